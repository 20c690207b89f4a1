// corebench.cu — which GEMV core keeps up? Variants of the per-CTA int8-limb dot-product loop run on
// data that is already RESIDENT in shared memory (no TMA, no HBM), 8 or 16 warps per CTA, one CTA per
// SM. Reports shared-memory weight bytes consumed per clock per SM; the token kernel needs 22.7 B/clk/SM
// to match HBM and 2-3x that to absorb phase-boundary stalls.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o corebench corebench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int dp(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int N = 4096;     // row bytes
constexpr int ROWS = 40;    // rows resident in smem (160 KB)

// V0: one warp per row, limbs in registers (CPL=8), 6 accumulator chains, 3 REDUX, lane-0 tail.
// V1: same but NO cross-lane reduction (each lane just keeps its sums) - upper bound without REDUX.
// V2: same, per-lane partial sums stored to smem (STS) instead of REDUX.
// V3: V0 with 12 accumulator chains.
template <int V>
__global__ void __launch_bounds__(256) k_row(int iters, long long *out, int *sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t *tile = sm;                      // ROWS x N
    int *part = (int *)(sm + ROWS * N);      // [8 warps][32][4]
    long long *res = (long long *)(sm + ROWS * N + 8 * 32 * 16);
    for (int i = threadIdx.x; i < ROWS * N / 4; i += blockDim.x) ((uint32_t *)tile)[i] = i * 2654435761u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 a0[8], a1[8], a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a0[i] = make_uint4(lane + i, lane * 3 + i, 7 * i, 11);
        a1[i] = make_uint4(lane - i, lane * 5 + i, 9 * i, 13);
        a2[i] = make_uint4(lane ^ i, lane * 7 + i, 3 * i, 17);
    }
    const uint32_t base = s32(tile) + lane * 16;
    int acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int r = warp; r < ROWS; r += 8) {
            const uint32_t row = base + r * N;
            uint4 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = lds128(row + i * 512);
            int s[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (V == 3) {
                    s[0] = dp(w[i].x, a0[i].x, s[0]); s[1] = dp(w[i].x, a1[i].x, s[1]); s[2] = dp(w[i].x, a2[i].x, s[2]);
                    s[3] = dp(w[i].y, a0[i].y, s[3]); s[4] = dp(w[i].y, a1[i].y, s[4]); s[5] = dp(w[i].y, a2[i].y, s[5]);
                    s[6] = dp(w[i].z, a0[i].z, s[6]); s[7] = dp(w[i].z, a1[i].z, s[7]); s[8] = dp(w[i].z, a2[i].z, s[8]);
                    s[9] = dp(w[i].w, a0[i].w, s[9]); s[10] = dp(w[i].w, a1[i].w, s[10]); s[11] = dp(w[i].w, a2[i].w, s[11]);
                } else {
                    s[0] = dp(w[i].x, a0[i].x, s[0]); s[1] = dp(w[i].x, a1[i].x, s[1]); s[2] = dp(w[i].x, a2[i].x, s[2]);
                    s[3] = dp(w[i].y, a0[i].y, s[3]); s[4] = dp(w[i].y, a1[i].y, s[4]); s[5] = dp(w[i].y, a2[i].y, s[5]);
                    s[0] = dp(w[i].z, a0[i].z, s[0]); s[1] = dp(w[i].z, a1[i].z, s[1]); s[2] = dp(w[i].z, a2[i].z, s[2]);
                    s[3] = dp(w[i].w, a0[i].w, s[3]); s[4] = dp(w[i].w, a1[i].w, s[4]); s[5] = dp(w[i].w, a2[i].w, s[5]);
                }
            }
            int u0, u1, u2;
            if (V == 3) { u0 = s[0] + s[3] + s[6] + s[9]; u1 = s[1] + s[4] + s[7] + s[10]; u2 = s[2] + s[5] + s[8] + s[11]; }
            else { u0 = s[0] + s[3]; u1 = s[1] + s[4]; u2 = s[2] + s[5]; }
            if (V == 0 || V == 3) {
                const int t0r = __reduce_add_sync(0xffffffffu, u0);
                const int t1r = __reduce_add_sync(0xffffffffu, u1);
                const int t2r = __reduce_add_sync(0xffffffffu, u2);
                if (lane == 0) res[r] = (((long long)t2r << 8) + t1r) * 256 + t0r;
            } else if (V == 1) {
                acc += u0 ^ u1 ^ u2;
            } else {
                *reinterpret_cast<int4 *>(part + (warp * 32 + lane) * 4) = make_int4(u0, u1, u2, 0);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x7fffffff) sink[0] = acc + (int)res[0] + part[0];
}

// V4: K-split: warp w owns chunk-columns (1 chunk per lane: 512 B of every row), limbs for that slice
// in 12 registers, ALL rows of the tile per warp, NR rows interleaved (independent chains), per-lane
// partials stored to smem (no reduction instruction in the loop).
template <int NR>
__global__ void __launch_bounds__(256) k_ksplit(int iters, long long *out, int *sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t *tile = sm;
    int *part = (int *)(sm + ROWS * N); // [ROWS][8 warps] x int4 : 40*8*16 = 5 KB ... lanes reduced later
    for (int i = threadIdx.x; i < ROWS * N / 4; i += blockDim.x) ((uint32_t *)tile)[i] = i * 2654435761u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4 a0 = make_uint4(lane, lane * 3, 7, 11), a1 = make_uint4(lane + 1, lane * 5, 9, 13), a2 = make_uint4(lane ^ 5, lane * 7, 3, 17);
    const uint32_t base = s32(tile) + warp * 512 + lane * 16;
    int acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int r0 = 0; r0 < ROWS; r0 += NR) {
            uint4 w[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) w[k] = lds128(base + (r0 + k) * N);
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                int u0 = dp(w[k].x, a0.x, 0), u1 = dp(w[k].x, a1.x, 0), u2 = dp(w[k].x, a2.x, 0);
                u0 = dp(w[k].y, a0.y, u0); u1 = dp(w[k].y, a1.y, u1); u2 = dp(w[k].y, a2.y, u2);
                u0 = dp(w[k].z, a0.z, u0); u1 = dp(w[k].z, a1.z, u1); u2 = dp(w[k].z, a2.z, u2);
                u0 = dp(w[k].w, a0.w, u0); u1 = dp(w[k].w, a1.w, u1); u2 = dp(w[k].w, a2.w, u2);
                // cross-lane: REDUX per row-slice (24 per row over the CTA)
                const int t0r = __reduce_add_sync(0xffffffffu, u0);
                const int t1r = __reduce_add_sync(0xffffffffu, u1);
                const int t2r = __reduce_add_sync(0xffffffffu, u2);
                if (lane == 0) *reinterpret_cast<int4 *>(part + ((r0 + k) * 8 + warp) * 4) = make_int4(t0r, t1r, t2r, 0);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x7fffffff) sink[0] = acc + part[0];
}

// V5: lane = row (28 rows), chunk-major tile, activations broadcast from smem, G row-groups share
// one activation load (G = 1: what the token kernel does now; G = 4: ffn-K style amortisation).
template <int G>
__global__ void __launch_bounds__(256) k_lane(int iters, long long *out, int *sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t *tile = sm;                 // [chunk][G*32 rows][16 B], 256 chunks ... use 160 KB worth
    uint8_t *planes = sm + ROWS * N;    // 3 x 4096
    for (int i = threadIdx.x; i < (ROWS * N + 3 * 4096) / 4; i += blockDim.x) ((uint32_t *)sm)[i] = i * 2654435761u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = ROWS * N / (G * 32 * 16); // chunks resident
    const uint32_t wb = s32(tile) + lane * 16, ab = s32(planes);
    int s0 = 0, s1 = 0, s2 = 0, t0_ = 0, t1_ = 0, t2_ = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int c = warp; c < nchunk; c += 8) {
            const uint4 a0 = lds128(ab + (c & 255) * 16), a1 = lds128(ab + 4096 + (c & 255) * 16), a2 = lds128(ab + 8192 + (c & 255) * 16);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint4 w = lds128(wb + (c * G + g) * 512);
                s0 = dp(w.x, a0.x, s0); s1 = dp(w.x, a1.x, s1); s2 = dp(w.x, a2.x, s2);
                t0_ = dp(w.y, a0.y, t0_); t1_ = dp(w.y, a1.y, t1_); t2_ = dp(w.y, a2.y, t2_);
                s0 = dp(w.z, a0.z, s0); s1 = dp(w.z, a1.z, s1); s2 = dp(w.z, a2.z, s2);
                t0_ = dp(w.w, a0.w, t0_); t1_ = dp(w.w, a1.w, t1_); t2_ = dp(w.w, a2.w, t2_);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if ((s0 ^ s1 ^ s2 ^ t0_ ^ t1_ ^ t2_) == 0x7fffffff) sink[0] = 1;
}

template <class K> void run(const char *name, K kern, int iters, double bytes_per_iter) {
    long long *d;
    int *sink;
    cudaMalloc(&d, 148 * 8);
    cudaMalloc(&sink, 4);
    const size_t smem = ROWS * N + 3 * 4096 + 16384;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<148, 256, smem>>>(2, d, sink);
    kern<<<148, 256, smem>>>(iters, d, sink);
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-46s %8.1f B/clk/SM  (%.0f clk per 4 KB row-equivalent)  %s\n", name, bytes_per_iter * iters / avg, avg / iters / (bytes_per_iter / 4096.0),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
    cudaFree(sink);
}

int main() {
    const int iters = 200;
    const double bytes = (double)ROWS * N; // per iteration per CTA
    run("V0 warp-per-row, 6 chains, 3 REDUX, tail", k_row<0>, iters, bytes);
    run("V3 warp-per-row, 12 chains, 3 REDUX, tail", k_row<3>, iters, bytes);
    run("V1 warp-per-row, no cross-lane reduction", k_row<1>, iters, bytes);
    run("V2 warp-per-row, per-lane partials -> STS", k_row<2>, iters, bytes);
    run("V4 K-split (1 chunk/lane), 1 row at a time", k_ksplit<1>, iters, bytes);
    run("V4 K-split, 4 rows interleaved", k_ksplit<4>, iters, bytes);
    run("V4 K-split, 8 rows interleaved", k_ksplit<8>, iters, bytes);
    run("V5 lane=row, act broadcast, G=1", k_lane<1>, iters, bytes);
    run("V5 lane=row, act shared by 4 row groups", k_lane<4>, iters, bytes);
    return 0;
}
