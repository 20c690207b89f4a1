// typical.h — the reference's `typical()` sampler over the 50277 logits, host side.
//
// API parity with the reference's include/rwkv/sampler/typical.h:20,60
//   int typical(float* logits, float temp = 0.9, float tau = 0.8);
//   std::vector<unsigned long long> typical(int batch, float* logits, float temp, float tau);
// restated on <random>/<algorithm> (the reference drags in the 58 kLoC NumCpp tree for this).
//
// What the reference actually computes (verified by running it: tests/golden/sampler_golden.json
// holds sequences drawn by the reference binary, and this header reproduces them exactly):
//   probs = exp(l) / sum(exp(l))        no max-shift; sequential double accumulation
//   the entropy / argsort / cumulative-tau cutoff is computed but then DISCARDED: the line
//       probs[shifted_logits > sorted_logits[cutoff]] = 0;          (typical.h:50)
//     assigns to a temporary, because NumCpp's NdArray::operator[](NdArray<bool>) returns a copy
//     (NumCpp/NdArray/NdArrayCore.hpp:778). `tau` therefore has no effect;
//   probs = probs ^ uint8(1/temp)       nc::power takes a uint8 exponent (NumCpp/Functions/power.hpp),
//                                       so 1/0.9 -> 1 and 1/0.8 -> 1: temperature is a no-op unless
//                                       temp <= 0.5; temp > 1 gives exponent 0 = a uniform draw;
//   token ~ std::discrete_distribution<int>(probs) on one process-wide std::mt19937_64 that is
//           default-seeded (NumCpp/Random/generator.hpp:35).
// `typical()` keeps exactly that behaviour (results identical to the reference on identical
// inputs). `typical_filtered()` is the algorithm the reference's own header comment describes
// (locally typical sampling with a real cutoff and a real temperature) for callers who want it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>

// The process-wide generator (default seed 5489); reseed for reproducible runs.
inline std::mt19937_64 &rwkv_sampler_generator() {
    static std::mt19937_64 gen;
    return gen;
}

// The draw itself restates what std::discrete_distribution<int>(probs)(gen) computes in libstdc++
// (bits/random.tcc: param_type::_M_initialize and operator()) without its two 400 KB vectors:
//   s = accumulate(probs); q_i = probs_i / s; cp_i = q_0 + ... + q_i (sequential), cp_last = 1.0;
//   u = generate_canonical<double, 53>(gen);  token = first i with cp_i >= u   (lower_bound).
// Same operations in the same order on the same doubles, hence the same tokens; tests/test_sampler.py
// checks the sequences against the ones drawn by the reference binary. (The uniform is drawn after the
// probabilities are built, as the reference does; it is the generator's only consumer.)
inline int rwkv_sampler_pick(const double *probs, int len, double u) {
    double s = 0.0;
    for (int i = 0; i < len; ++i) s += probs[i];
    double c = 0.0;
    for (int i = 0; i + 1 < len; ++i) {
        c += probs[i] / s;
        if (!(c < u)) return i;
    }
    return len - 1; // cp_last is forced to 1.0 and u < 1
}

// probs = exp(l) / sum(exp(l)), then ^ uint8(1/temp): the distribution `typical()` samples from
inline const std::vector<double> &rwkv_sampler_probs(const float *_logits, float _temp) {
    constexpr int len = 50277;
    static thread_local std::vector<double> probs(len);
    double total = 0.0;
    for (int i = 0; i < len; ++i) {
        probs[i] = std::exp((double)_logits[i]);
        total += probs[i];
    }
    const uint8_t exponent = _temp != 1.0 ? (uint8_t)(1.0 / _temp) : (uint8_t)1;
    if (exponent == 0) {
        for (int i = 0; i < len; ++i) probs[i] = 1.0;
    } else {
        for (int i = 0; i < len; ++i) {
            const double q = probs[i] / total;
            double v = q;
            for (uint8_t e = 1; e < exponent; ++e) v *= q;
            probs[i] = v;
        }
    }
    return probs;
}

// typical() with the uniform supplied by the caller (RWKV::sample uses it when the device sampler's
// answer is too close to an interval boundary to be trusted)
inline int typical_with_u(const float *_logits, float _temp, double u) {
    const std::vector<double> &probs = rwkv_sampler_probs(_logits, _temp);
    return rwkv_sampler_pick(probs.data(), (int)probs.size(), u);
}

inline int typical(float *_logits, float _temp = 0.9, float _tau = 0.8) {
    (void)_tau; // see the header comment: the reference's cutoff never reaches `probs`
    const std::vector<double> &probs = rwkv_sampler_probs(_logits, _temp);
    const double u = std::generate_canonical<double, 53>(rwkv_sampler_generator());
    return rwkv_sampler_pick(probs.data(), (int)probs.size(), u);
}

// Locally typical sampling as the reference's header comment (typical.h:1-18) specifies it:
// keep the tokens whose surprise is closest to the entropy until their mass reaches tau, apply
// the temperature as probs^(1/temp), sample. Not used by the reference-compatible entry points.
inline int typical_filtered(const float *_logits, float _temp = 0.9, float _tau = 0.8) {
    constexpr int len = 50277;
    std::vector<double> probs(len), shifted(len);
    double mx = _logits[0];
    for (int i = 1; i < len; ++i) mx = std::max(mx, (double)_logits[i]);
    double total = 0.0;
    for (int i = 0; i < len; ++i) {
        probs[i] = std::exp((double)_logits[i] - mx);
        total += probs[i];
    }
    double entropy = 0.0;
    for (int i = 0; i < len; ++i) {
        probs[i] /= total;
        if (probs[i] > 0.0) entropy -= probs[i] * std::log(probs[i]);
    }
    for (int i = 0; i < len; ++i) shifted[i] = probs[i] > 0.0 ? std::abs(-std::log(probs[i]) - entropy) : INFINITY;
    std::vector<uint32_t> order(len);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return shifted[a] < shifted[b]; });
    double running = 0.0;
    int cutoff = 0;
    for (int k = 0; k < len; ++k) {
        running += probs[order[k]];
        if (running < (double)_tau) ++cutoff;
    }
    const double threshold = shifted[order[std::min(cutoff, len - 1)]];
    for (int i = 0; i < len; ++i) {
        if (shifted[i] > threshold) probs[i] = 0.0;
        else if (_temp != 1.0f) probs[i] = std::pow(probs[i], 1.0 / (double)_temp);
    }
    std::discrete_distribution<int> dist(probs.begin(), probs.end());
    return dist(rwkv_sampler_generator());
}

inline std::vector<unsigned long long> typical(int batchsize, float *_logits, float _temp = 0.9, float _tau = 0.8) {
    std::vector<unsigned long long> out;
    out.reserve(batchsize > 0 ? batchsize : 0);
    for (int i = 0; i < batchsize; ++i) out.push_back(typical(&_logits[(size_t)i * 50277], _temp, _tau));
    return out;
}
