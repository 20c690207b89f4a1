// sampler_cli.cpp — same contract as oracle/ref_sampler.cpp, against THIS repository's
// include/rwkv/sampler/typical.h.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rwkv/sampler/typical.h"

static inline float synth(uint64_t d, uint64_t i, float scale) {
    uint64_t z = (d * 0x9E3779B97F4A7C15ULL) ^ (i * 0xD6E8FEB86659FD93ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return scale * ((float)(z >> 40) * (1.0f / 8388608.0f) - 1.0f);
}
int main(int argc, char **argv) {
    if (argc < 5) return 1;
    const int n = atoi(argv[1]);
    const float temp = (float)atof(argv[2]), tau = (float)atof(argv[3]), scale = (float)atof(argv[4]);
    std::vector<float> logits(50277);
    for (int d = 0; d < n; ++d) {
        for (int i = 0; i < 50277; ++i) logits[i] = synth((uint64_t)d, (uint64_t)i, scale);
        logits[0] = -99.0f;
        printf("%d\n", typical(logits.data(), temp, tau));
    }
    return 0;
}
