// typical.h — locally-typical sampling over the 50277 logits, host side.
//
// API parity with the reference's include/rwkv/sampler/typical.h:20,60
//   int typical(float* logits, float temp = 0.9, float tau = 0.8);
//   std::vector<unsigned long long> typical(int batch, float* logits, float temp, float tau);
// restated on <random>/<algorithm> (the reference drags in the 58 kLoC NumCpp tree for this).
// The arithmetic order is the one NumCpp executes, so a default-seeded run draws the same
// token sequence as the reference:
//   probs = exp(l) / sum(exp(l))             no max-shift; sequential double accumulation
//   s_i   = | -log p_i - sum_j(-log p_j * p_j) |      (NaN terms dropped from the sum)
//   order = stable argsort of s;  cutoff = #{k : cumsum(p[order])_k < tau}
//   p_i   = 0 where s_i > s[order[cutoff]]
//   p     = p ^ uint8(1/temp)                the reference's nc::power takes a uint8 exponent
//                                            (NumCpp/Functions/power.hpp), so 1/0.9 -> 1 and the
//                                            temperature is a no-op unless temp <= 0.5; temp > 1
//                                            gives exponent 0, i.e. a uniform draw. Kept as is.
//   token ~ std::discrete_distribution<int>(p) on one process-wide std::mt19937_64 that is
//           default-seeded (NumCpp/Random/generator.hpp:35).
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

inline int typical(float *_logits, float _temp = 0.9, float _tau = 0.8) {
    constexpr int len = 50277;
    std::vector<double> probs(len), surprise(len), shifted(len);

    double total = 0.0;
    for (int i = 0; i < len; ++i) {
        probs[i] = std::exp((double)_logits[i]);
        total += probs[i];
    }
    for (int i = 0; i < len; ++i) probs[i] /= total;

    double entropy = 0.0;
    for (int i = 0; i < len; ++i) {
        surprise[i] = -std::log(probs[i]);
        const double term = surprise[i] * probs[i];
        entropy += std::isnan(term) ? 0.0 : term;
    }
    for (int i = 0; i < len; ++i) shifted[i] = std::abs(surprise[i] - entropy);

    std::vector<uint32_t> order(len);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return shifted[a] < shifted[b]; });

    const double tau = (double)_tau;
    int cutoff = 0;
    double running = 0.0;
    for (int k = 0; k < len; ++k) {
        running = (k == 0) ? probs[order[0]] : running + probs[order[k]];
        if (running < tau) ++cutoff;
    }
    const double threshold = shifted[order[std::min(cutoff, len - 1)]];
    for (int i = 0; i < len; ++i)
        if (shifted[i] > threshold) probs[i] = 0.0;

    if (_temp != 1.0) {
        const uint8_t exponent = (uint8_t)(1.0 / _temp);
        for (int i = 0; i < len; ++i) {
            if (exponent == 0) {
                probs[i] = 1.0;
                continue;
            }
            double v = probs[i];
            for (uint8_t e = 1; e < exponent; ++e) v *= probs[i];
            probs[i] = v;
        }
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
