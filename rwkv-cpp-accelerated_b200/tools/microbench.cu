// microbench.cu — measures the two hardware rates the decode kernels are designed around:
//   (1) IDP.4A issue rate per SM (is 3 dp4a per 4 weight bytes comfortably below HBM speed?)
//   (2) a plain 128-bit streaming read of a buffer >> L2 (HBM read ceiling seen by SMs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_dp4a(int *out, int iters, unsigned a0, unsigned b0) {
    int acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    unsigned a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a), "r"(b));
    }
    int s = 0;
    for (int j = 0; j < 8; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float *out, int iters, float a0, float b0) {
    float acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    float a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[j]) : "f"(a), "f"(b));
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_read(const uint4 *__restrict__ p, size_t n, unsigned *out) {
    unsigned s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        s += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (s == 0x12345678u) out[0] = s;
}
int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms=%d clock=%d MHz l2=%d MB\n", pr.name, pr.multiProcessorCount, clk_khz / 1000, pr.l2CacheSize >> 20);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    const int iters = 1 << 14;
    for (int warps = 4; warps <= 32; warps *= 2) {
        dim3 g(pr.multiProcessorCount), blk(warps * 32);
        k_dp4a<<<g, blk>>>(d, 16, 1, 2);
        cudaEventRecord(a); k_dp4a<<<g, blk>>>(d, iters, 1, 2); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double ops = (double)pr.multiProcessorCount * warps * 32 * iters * 8;
        k_ffma<<<g, blk>>>((float *)d, 16, 1, 2);
        cudaEventRecord(a); k_ffma<<<g, blk>>>((float *)d, iters, 1.f, 2.f); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms2; cudaEventElapsedTime(&ms2, a, b);
        printf("warps/SM=%2d  dp4a %.1f Gop/s/SM (%.2f Tdp4a/s chip)   ffma %.1f Gop/s/SM\n", warps,
               ops / ms / 1e6 / pr.multiProcessorCount, ops / ms / 1e9, ops / ms2 / 1e6 / pr.multiProcessorCount);
    }
    size_t bytes = (size_t)4 << 30;
    uint4 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    unsigned *o; cudaMalloc(&o, 4);
    for (int bpsm = 1; bpsm <= 8; bpsm *= 2) {
        k_read<<<pr.multiProcessorCount * bpsm, 512>>>(buf, bytes / 16, o);
        cudaEventRecord(a); k_read<<<pr.multiProcessorCount * bpsm, 512>>>(buf, bytes / 16, o); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("stream read 4 GiB, %d CTA/SM x 512 thr: %.0f GB/s\n", bpsm, bytes / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
