// q8.h — the reference's uint8 weight quantiser, restated for one input row.
//
// Follows converter/convert_model.py:108-119 (`quantize_matrix`): for a Linear weight
// W[out][in] the statistics are taken per *input* column j over all outputs:
//   mini_j = min_k W[k][j];  ran_j = (max_k W[k][j] - mini_j) / 255
//   q[k][j] = trunc((W[k][j] - mini_j) / ran_j)            (stored transposed: [in][out])
//   mini_j += mean_k(frac((W - mini)/ran)) * ran_j          (bias correction of truncation)
// and `ranges = ran` (f32), `zp = mini` (f32). The arithmetic is float64 like torch's
// promotion of (float32 tensor - float64 tensor).
//
// Because the stored layout is [in][out], one stored row j is self-contained: this
// helper takes the `out` real values of row j and emits the `out` bytes + (ran, zp).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace q8 {

struct RowParams {
    float range; // "r": dequant scale of this input row
    float zp;    // "o": dequant offset of this input row
};

inline RowParams quantize_row(const float *w, size_t n, uint8_t *q) {
    double lo = w[0], hi = w[0];
    for (size_t k = 1; k < n; ++k) {
        double v = w[k];
        if (v < lo) lo = v;
        if (v > hi) hi = v;
    }
    const double ran = (hi - lo) / 255.0;
    double frac_sum = 0.0;
    for (size_t k = 0; k < n; ++k) {
        const double t = ((double)w[k] - lo) / ran; // ran == 0 -> NaN, as in the reference
        const double fl = std::trunc(t);
        frac_sum += t - fl;
        // torch's .to(uint8) of a double truncates toward zero; values are in [0,255].
        q[k] = (uint8_t)(fl < 0.0 ? 0.0 : (fl > 255.0 ? 255.0 : fl));
    }
    const double zp = lo + (frac_sum / (double)n) * ran;
    return RowParams{(float)ran, (float)zp};
}

} // namespace q8
