// sample_gpu_test.cpp — RWKV::sample() (device sampler) must return the token typical(out, ...) returns for the
// same generator state, for every temperature class of the reference sampler. usage: sample_gpu_test model.bin
#include <cstdio>
#include <string>
#include "rwkv.h"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    setenv("RWKV_B200_QUIET", "1", 1);
    RWKV net;
    net.loadFile(argv[1]);
    auto &gen = rwkv_sampler_generator();
    const float temps[] = {0.9f, 0.5f, 1.0f, 0.3f, 2.0f}; // exponents 1, 2, (none), 3, 0 = uniform
    unsigned long long tok = 4118;
    int checked = 0;
    for (float temp : temps) {
        for (int step = 0; step < 40; ++step) {
            float *out = net.forward(tok);
            gen.seed(1000 + 97 * step + (unsigned)(temp * 10));
            const int host = typical(out, temp, 0.8f);
            gen.seed(1000 + 97 * step + (unsigned)(temp * 10));
            const int dev = net.sample(temp, 0.8f);
            if (host != dev) {
                printf("FAIL temp %.1f step %d: host %d device %d\n", temp, step, host, dev);
                return 1;
            }
            tok = (unsigned long long)host;
            ++checked;
        }
    }
    printf("ALL OK %d draws\n", checked);
    return 0;
}
