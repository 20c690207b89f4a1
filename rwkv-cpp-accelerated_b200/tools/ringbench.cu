// ringbench.cu — the token kernel's producer/ring/consumer structure in isolation: a TMA producer
// warp streams tiles from an L2-resident buffer (so HBM is not the limit) through the mbarrier ring,
// consumer warps run a GEMV core variant. Reports weight bytes consumed per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../csrc -o ringbench ringbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../csrc/kernels.cuh"
using namespace rk;

__device__ __forceinline__ int dp_su(uint32_t a, uint32_t b, int c) { int d; asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

struct RP { uint32_t stage, phase; };

// VARIANT 0: lane=row core exactly as in token_kernel.cuh (8 warps split the chunks of a tile).
// VARIANT 1: lane=row, but every warp owns whole tiles (tile t -> warp t % 8): one wait/arrive per
//            tile per OWNER warp only (empty barrier count 1), long inner loop (all chunks of the tile).
// VARIANT 2: warp-per-row (limbs in registers), rows dealt round-robin, tile = 5 rows (old design).
// VARIANT 3: warp-per-row with one-row tiles owned by a single warp (per-warp rings interleaved).
template <int VAR, int NPROD>
__global__ void __launch_bounds__(384, 1) k_ring(const int8_t *src, size_t src_bytes, int ntiles, int tile_bytes, int stages,
                                                 long long *out, int *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ring = smem;
    uint8_t *planes = smem + (size_t)stages * tile_bytes;
    uint64_t *full = (uint64_t *)(planes + 3 * 4096);
    uint64_t *empty = full + 64;
    int *res = (int *)(empty + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = 4096, gr = 28;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(smem_u32(&full[i]), 1);
            mbar_init(smem_u32(&empty[i]), (VAR == 1 || VAR == 3) ? 1 : 8);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) ((uint32_t *)planes)[i] = i * 2654435761u;
    __syncthreads();
    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0 && warp - 8 < NPROD) {
            const uint64_t pol = policy_evict_normal();
            RP rp{0, 0};
            size_t off = (size_t)blockIdx.x * 1048576 % src_bytes;
            for (int t = 0; t < ntiles; ++t) {
                if (t % NPROD == warp - 8) {
                mbar_wait(smem_u32(&empty[rp.stage]), rp.phase ^ 1);
                const uint32_t fb = smem_u32(&full[rp.stage]);
                mbar_expect_tx(fb, tile_bytes);
                bulk_g2s(smem_u32(ring) + rp.stage * tile_bytes, src + off, tile_bytes, fb, pol);
                }
                off += tile_bytes;
                if (off + tile_bytes > src_bytes) off = 0;
                if (++rp.stage == (uint32_t)stages) { rp.stage = 0; rp.phase ^= 1; }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const uint32_t ring_a = smem_u32(ring), full0 = smem_u32(full), empty0 = smem_u32(empty), pl = smem_u32(planes);
    RP rp{0, 0};
    int acc = 0;
    const long long t0 = clock64();
    if (VAR == 0 || VAR == 1) {
        const bool on = lane < gr;
        int s0a = 0, s0b = 0, s1a = 0, s1b = 0, s2a = 0, s2b = 0;
        const int tc = tile_bytes / (gr * 16);
        for (int t = 0; t < ntiles; ++t) {
            const bool mine = VAR == 0 || (t & 7) == warp;
            if (mine) {
                mbar_wait(full0 + 8 * rp.stage, rp.phase);
                const uint32_t wbase = ring_a + rp.stage * tile_bytes + lane * 16;
                const uint32_t abase = pl + ((t * tc) & 255) * 16;
                const int c_beg = VAR == 0 ? warp : 0, c_step = VAR == 0 ? 8 : 1;
#pragma unroll 5
                for (int cc = c_beg; cc < tc; cc += c_step) {
                    uint4 w = make_uint4(0, 0, 0, 0);
                    if (on) w = lds128(wbase + cc * gr * 16);
                    const uint4 a0 = lds128(abase + cc * 16), a1 = lds128(abase + N + cc * 16), a2 = lds128(abase + 2 * N + cc * 16);
                    s0a = dp_su(w.x, a0.x, s0a); s1a = dp_su(w.x, a1.x, s1a); s2a = dp4a_ss(w.x, a2.x, s2a);
                    s0b = dp_su(w.y, a0.y, s0b); s1b = dp_su(w.y, a1.y, s1b); s2b = dp4a_ss(w.y, a2.y, s2b);
                    s0a = dp_su(w.z, a0.z, s0a); s1a = dp_su(w.z, a1.z, s1a); s2a = dp4a_ss(w.z, a2.z, s2a);
                    s0b = dp_su(w.w, a0.w, s0b); s1b = dp_su(w.w, a1.w, s1b); s2b = dp4a_ss(w.w, a2.w, s2b);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * rp.stage);
            }
            if (++rp.stage == (uint32_t)stages) { rp.stage = 0; rp.phase ^= 1; }
        }
        acc = s0a + s0b + s1a + s1b + s2a + s2b;
    } else {
        uint4 a0[8], a1[8], a2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[i] = lds128(pl + (lane + 32 * i) * 16); a1[i] = lds128(pl + N + (lane + 32 * i) * 16); a2[i] = lds128(pl + 2 * N + (lane + 32 * i) * 16); }
        const int rows = tile_bytes / N;
        int ubase = 0;
        for (int t = 0; t < ntiles; ++t) {
            const bool owner = VAR == 2 || (t & 7) == warp;
            if (owner) {
                mbar_wait(full0 + 8 * rp.stage, rp.phase);
                const uint32_t tile = ring_a + rp.stage * tile_bytes + lane * 16;
                for (int u = VAR == 2 ? ((warp - ubase) & 7) : 0; u < rows; u += VAR == 2 ? 8 : 1) {
                    const uint32_t row = tile + u * N;
                    uint4 w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = lds128(row + i * 512);
                    int s0a = 0, s0b = 0, s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        s0a = dp_su(w[i].x, a0[i].x, s0a); s1a = dp_su(w[i].x, a1[i].x, s1a); s2a = dp4a_ss(w[i].x, a2[i].x, s2a);
                        s0b = dp_su(w[i].y, a0[i].y, s0b); s1b = dp_su(w[i].y, a1[i].y, s1b); s2b = dp4a_ss(w[i].y, a2[i].y, s2b);
                        s0a = dp_su(w[i].z, a0[i].z, s0a); s1a = dp_su(w[i].z, a1[i].z, s1a); s2a = dp4a_ss(w[i].z, a2[i].z, s2a);
                        s0b = dp_su(w[i].w, a0[i].w, s0b); s1b = dp_su(w[i].w, a1[i].w, s1b); s2b = dp4a_ss(w[i].w, a2[i].w, s2b);
                    }
                    const int r0 = __reduce_add_sync(0xffffffffu, s0a + s0b), r1 = __reduce_add_sync(0xffffffffu, s1a + s1b),
                              r2 = __reduce_add_sync(0xffffffffu, s2a + s2b);
                    if (lane == 0) { res[(t & 63) * 3] = r0; res[(t & 63) * 3 + 1] = r1; res[(t & 63) * 3 + 2] = r2; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * rp.stage);
            }
            if (VAR == 2) ubase = (ubase + rows) & 7;
            if (++rp.stage == (uint32_t)stages) { rp.stage = 0; rp.phase ^= 1; }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x7fffffff) sink[0] = acc + res[0];
}

template <class K> void run(const char *name, K kern, const int8_t *src, size_t src_bytes, int tile_bytes, int stages) {
    long long *d; int *sink;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
    const int ntiles = 4000;
    const size_t smem = (size_t)stages * tile_bytes + 3 * 4096 + 2048;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<148, 384, smem>>>(src, src_bytes, 200, tile_bytes, stages, d, sink);
    kern<<<148, 384, smem>>>(src, src_bytes, ntiles, tile_bytes, stages, d, sink);
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-58s tile %6d x %2d: %6.1f B/clk/SM %s\n", name, tile_bytes, stages, (double)ntiles * tile_bytes / avg, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}

int main() {
    const size_t src_bytes = 24u << 20; // L2-resident after the first pass
    int8_t *src; cudaMalloc(&src, src_bytes + (1 << 20)); cudaMemset(src, 3, src_bytes + (1 << 20));
    run("V0 lane=row, 8 warps split each tile (token kernel now)", k_ring<0, 1>, src, src_bytes, 28 * 16 * 40, 8);
    run("V0 lane=row, 8 warps split each tile, big tiles", k_ring<0, 1>, src, src_bytes, 28 * 16 * 80, 4);
    run("V1 lane=row, one owner warp per tile", k_ring<1, 1>, src, src_bytes, 28 * 16 * 40, 8);
    run("V1 lane=row, one owner warp per tile, small tiles", k_ring<1, 1>, src, src_bytes, 28 * 16 * 16, 20);
    run("V2 warp-per-row, 5-row tiles shared by 8 warps (old)", k_ring<2, 1>, src, src_bytes, 5 * 4096, 8);
    run("V2 warp-per-row, 8-row tiles shared by 8 warps", k_ring<2, 1>, src, src_bytes, 8 * 4096, 5);
    run("V3 warp-per-row, one-row tiles, one owner warp per tile", k_ring<3, 1>, src, src_bytes, 4096, 40);
    run("V3 warp-per-row, two-row tiles, one owner warp per tile", k_ring<3, 1>, src, src_bytes, 8192, 20);
    run("V2 warp-per-row, 8-row tiles, 4 producers", k_ring<2, 4>, src, src_bytes, 8 * 4096, 5);
    run("V2 warp-per-row, 5-row tiles, 4 producers", k_ring<2, 4>, src, src_bytes, 5 * 4096, 8);
    run("V0 lane=row, 4 producers", k_ring<0, 4>, src, src_bytes, 28 * 16 * 40, 8);
    run("V2 warp-per-row, 8-row tiles, 4 producers, 6 stages", k_ring<2, 4>, src, src_bytes, 8 * 4096, 6);
    return 0;
}
