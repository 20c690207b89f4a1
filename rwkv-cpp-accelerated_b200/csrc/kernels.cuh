// kernels.cuh — sm_100a device code of the RWKV-v4 uint8 decode path.
//
// One token = embed_ln0 + n_layers x {att_kvr, att_out, ffn_rk, ffn_v} + head (+argmax).
// Every matrix phase is the same machine:
//
//   producer warp : one elected lane streams this CTA's weight rows HBM -> shared memory
//                   with 1-D bulk TMA (cp.async.bulk ... mbarrier::complete_tx) through a
//                   ring of STAGES tiles guarded by full/empty mbarriers. Weights do not
//                   depend on activations, so streaming starts before the prologue ends.
//   consumer warps: (1) prologue: build the activation vector(s) for this phase
//                   (layernorm / token shift / ...), pre-scale by the per-input-row dequant
//                   scale r_j, and quantise to three signed 7-bit "limbs"
//                   xs_j = S * (l2*2^14 + l1*2^7 + l0), kept in registers per lane;
//                   (2) main loop: one warp per weight row (or row segment): 128-bit LDS of
//                   the row bytes, 12 IDP.4A per 16 bytes against the limb registers,
//                   exact int32 accumulation, REDUX.SUM across the warp;
//                   (3) epilogue: y = S*total + sum_j x_j*oc_j, then the fused elementwise
//                   tail of the phase (WKV, residual add, sigmoid, relu^2 ...).
//
// Why integer limbs: the reference computes y_k = sum_j x_j*(w_jk*r_j + o_j) in fp32 with
// I2F + 2 FMA per weight byte (include/rwkv/cuda/rwkv.cu:279-294). Factorised as
// sum_j (x_j r_j) w'_jk + sum_j x_j (128 r_j + o_j) with w' = w-128 (stored as s8), the
// per-byte work is 3/4 of an IDP.4A, the accumulation is exact (so the result is
// deterministic and independent of the reduction order), and the only rounding is the
// 2^-21-of-max quantisation of the activation vector.
//
// HBM layout of a matrix: row-major [out][in] int8 (value ^ 0x80), so the rows one CTA
// owns are one contiguous byte range -> 1-D bulk copies, no tensor maps.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rk {

constexpr int kVocab = 50277;
constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32; // 256
constexpr int kThreads = kConsumers + 32;       // + producer warp
constexpr int kMaxStages = 12;
constexpr int kTileTraceMax = 4096;  // tiles per CTA recorded by the tile trace (debug)
constexpr int kTraceMax = 2048;      // trace stamps per CTA (token kernel, debug)
constexpr int kRedMax = 160;         // max contributors to a block reduction (grid size, rows of one CTA)
constexpr int kMaxSlice = 64;        // max elements of the residual stream one CTA owns (token kernel)
constexpr int kMaxRowsPerCta = 1024; // res64 capacity (rows x segments of one CTA)
constexpr int kMaxGrid = 1024;      // partials capacity
constexpr int kQMax = (1 << 20) - 1;

struct Ctrl {
    unsigned long long token; // input token of the current forward
    unsigned long long next;  // argmax of the last logits (forward_greedy)
    unsigned long long slot;  // state slot (PARRALEL mode)
    unsigned long long pos;   // cursor into a device-resident token stream (decode_timed)
    unsigned int bar_base;    // grid-barrier count at the start of the next token kernel
    unsigned int pad[3];
};

// Everything a kernel needs, passed by value (__grid_constant__).
struct Params {
    int L, E;
    int tile_bytes, stages;
    int plane_cap; // bytes reserved for limb planes in shared memory
    int tp_rank, tp_size;
    // repacked weights: int8 [rows_out][N_in]
    const int8_t *wk, *wv, *wr, *wo, *wfk, *wfv, *wfr, *whead;
    // per-input-row scale r and centred offset oc = 128*r + o
    const float *rk, *rv, *rr, *ro, *rfk, *rfv, *rfr, *rhead;
    const float *ock, *ocv, *ocr, *oco, *ocfk, *ocfv, *ocfr, *ochead;
    const double *ln;                                  // [4(L+1)][E]
    const double *mixk, *mixv, *mixr, *fmixk, *fmixr;  // [L][E]
    const double *decay, *bonus;                       // [L][E]
    const double *expdecay;                            // [L][E] exp(decay), tabulated at load
    const float *emb;                                  // [V][E]
    double *sxy, *saa, *sbb, *sdd;                     // [slots][L][E]
    double *x;                                         // [E] residual stream
    double *xy_new, *dd_new;                           // [E] token-shift state in flight
    float *xs_o;                                       // [E]  float(rwkv) * r_attout
    float *sr;                                         // [E]  sigmoid(ffn r)
    float *xs_v;                                       // [4E] relu(k)^2 * r_ffnv
    float *logits;                                     // [V]
    double *part_o;                                    // [2][kMaxGrid] max|xs|, sum x*oc  (att_kvr -> att_out)
    double *part_v;                                    // [2][kMaxGrid]                    (ffn_rk  -> ffn_v)
    Ctrl *ctrl;
    // ---- persistent token kernel (k_token) ------------------------------------------------
    int L_run;                         // layers to run (debug knob; normally == L)
    int issue_gap;                     // minimum SM cycles between two bulk-copy issues of the producer (0 = unpaced)
    int feed_mode;                     // 0: ctrl->token, 1: ctrl->next (free-running), 2: stream[ctrl->pos]
    int greedy;                        // 1: finish with an on-device argmax into ctrl->next
    const unsigned long long *stream;  // device-resident token stream (feed_mode 2)
    // The exchange block of this rank (one allocation, peer-mapped by the other ranks): gbar, acc,
    // vec and logits all point into xch[tp_rank]; xch[g] is rank g's block as seen from here.
    unsigned char *xch[8];
    unsigned int *gbar;                // grid barrier counter (monotonic; one rank: counts CTAs, several: counts ranks)
    unsigned int *lbar;                // rank-local arrival counter of the hierarchical barrier (tp_size > 1)
    unsigned long long *acc;           // [3 phases][16] integer accumulators (token_kernel.cuh)
    float *vec;                        // [2 parity][4E] next-phase activation vector(s), pre-scaled by r
    unsigned long long *trace;         // optional [grid][kTraceMax] globaltimer stamps (debug), or nullptr
    unsigned long long *ptrace;        // optional [2][grid][kTileTraceMax]: tile issue / tile ready times (debug)
};

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (error code on the host) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}
// 1-D bulk TMA: global -> shared, completion signalled as transaction bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void consumer_sync() { // named barrier 1: the 8 consumer warps only
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}
__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---------------------------------------------------------------------------------------
// Block-wide reductions over the 256 consumer threads (named barrier 1).
// `scratch` = 2 x 8 doubles of shared memory. All consumers get the result.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Sum of `a`, max of `b` in one pass. Deterministic (fixed tree).
__device__ __forceinline__ void cons_reduce(double &a, double &b, double *scratch, int ctid) {
    a = warp_sum(a);
    b = warp_max(b);
    const int w = ctid >> 5;
    consumer_sync(); // scratch free (previous reduction fully read)
    if ((ctid & 31) == 0) {
        scratch[w] = a;
        scratch[8 + w] = b;
    }
    consumer_sync();
    double s = 0.0, m = scratch[8];
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) {
        s += scratch[i];
        m = fmax(m, scratch[8 + i]);
    }
    a = s;
    b = m;
}

// ---------------------------------------------------------------------------------------
// Layernorm with the reference's rounding points (rwkv.cu:412-465, 40-57): the sum and
// the sum of squared deviations are rounded to f32 (they live in float accumulators
// there), variance is unbiased (E-1), no epsilon, sqrt in f32.
// Thread layout: consumer `ctid` owns groups g = ctid + 256*m of 4 consecutive elements.
// ---------------------------------------------------------------------------------------
// Register arrays in the prologues are sized by GROUPS = ceil(E / 1024), derived from the
// same template parameter CPL (16-byte chunks per lane) that sizes the limb registers:
//   CPL  2 -> E <= 1024, 4 -> E <= 2048, 8 -> E <= 4096, 10 -> E <= 5120.

template <class LoadX>
__device__ __forceinline__ void ln_stats(int E, int ctid, double *scratch, LoadX loadx, double &xmean,
                                         double &x2) {
    const int ng = E >> 2;
    double s = 0.0, dummy = 0.0;
    for (int g = ctid; g < ng; g += kConsumers) {
        double v[4];
        loadx(g, v);
        s += (v[0] + v[1]) + (v[2] + v[3]);
    }
    cons_reduce(s, dummy, scratch, ctid);
    const float mean_acc = (float)s;
    const double mean_f = (double)(mean_acc / (float)E); // variance kernel: float / float
    double q = 0.0;
    dummy = 0.0;
    for (int g = ctid; g < ng; g += kConsumers) {
        double v[4];
        loadx(g, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double d = v[e] - mean_f;
            q += d * d;
        }
    }
    cons_reduce(q, dummy, scratch, ctid);
    const float var_acc = (float)q;
    xmean = (double)mean_acc / (double)E;
    x2 = (double)sqrtf(var_acc / (float)(E - 1));
}

// ---------------------------------------------------------------------------------------
// Activation quantisation: 4 consecutive elements -> one 32-bit word in each limb plane.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void quantize4(const double (&xs)[4], double inv_s, uint8_t *planes,
                                          int plane_stride, int j) {
    uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int q = __double2int_rn(xs[e] * inv_s);
        q = max(-kQMax, min(kQMax, q));
        const int l0 = ((q + 64) & 127) - 64;
        const int q1 = (q - l0) >> 7;
        const int l1 = ((q1 + 64) & 127) - 64;
        const int l2 = (q1 - l1) >> 7;
        w0 |= (uint32_t)(l0 & 0xff) << (8 * e);
        w1 |= (uint32_t)(l1 & 0xff) << (8 * e);
        w2 |= (uint32_t)(l2 & 0xff) << (8 * e);
    }
    *reinterpret_cast<uint32_t *>(planes + j) = w0;
    *reinterpret_cast<uint32_t *>(planes + plane_stride + j) = w1;
    *reinterpret_cast<uint32_t *>(planes + 2 * plane_stride + j) = w2;
}

// ---------------------------------------------------------------------------------------
// Shared-memory carve-up (dynamic shared memory, 128-byte aligned base)
// ---------------------------------------------------------------------------------------
struct Smem {
    uint8_t *ring;       // stages * tile_bytes
    uint8_t *planes;     // limb planes
    long long *res64;    // [kMaxRowsPerCta] exact integer row totals
    double *scratch;     // [16] reductions
    double *scal;        // [12] S[0..2], off[0..2], [6..7] three floats 1/S, [8] trace counter
    uint64_t *full;      // [stages]
    uint64_t *empty;     // [stages]
    double *red;         // [6][kRedMax] + [8] block-reduction scratch (token kernel)
    double *xown;        // [kMaxSlice] this CTA's slice of the residual stream (token kernel)
    float *srown;        // [kMaxSlice] sigmoid(ffn r) of the slice
};

__device__ __forceinline__ Smem carve(uint8_t *base, const Params &p) {
    Smem s;
    s.ring = base;
    uint8_t *q = base + (size_t)p.stages * p.tile_bytes;
    s.planes = q;
    q += p.plane_cap;
    s.res64 = reinterpret_cast<long long *>(q);
    q += kMaxRowsPerCta * sizeof(long long);
    s.scratch = reinterpret_cast<double *>(q);
    q += 16 * sizeof(double);
    s.scal = reinterpret_cast<double *>(q);
    q += 12 * sizeof(double);
    s.full = reinterpret_cast<uint64_t *>(q);
    q += kMaxStages * sizeof(uint64_t);
    s.empty = reinterpret_cast<uint64_t *>(q);
    q += kMaxStages * sizeof(uint64_t);
    s.red = reinterpret_cast<double *>(q);
    q += (6 * kRedMax + 8) * sizeof(double);
    s.xown = reinterpret_cast<double *>(q);
    q += kMaxSlice * sizeof(double);
    s.srown = reinterpret_cast<float *>(q);
    return s;
}

__host__ __device__ inline size_t smem_bytes(int stages, int tile_bytes, int plane_cap) {
    return (size_t)stages * tile_bytes + plane_cap + kMaxRowsPerCta * 8 + 16 * 8 + 12 * 8 + 2 * kMaxStages * 8 +
           (6 * kRedMax + 8) * 8 + kMaxSlice * 12 + 128;
}

// One streamed sub-matrix of a phase: rows [r0, r1) of a row-major int8 matrix with N bytes
// per row; a row is cut into `nseg` segments of N/nseg bytes handled by different warps.
struct Sub {
    const int8_t *base;
    int N;
    int r0, r1;
    int nseg;
    int plane_off; // byte offset of this sub's limb planes inside Smem::planes
    int res_off;   // first res64 slot of this sub
};

__device__ __forceinline__ int tile_rows(const Params &p, const Sub &s) {
    int tr = p.tile_bytes / s.N;
    if (tr < 1) tr = 1;
    if (s.nseg > 1 && tr > 1) tr &= ~1; // two rows per pass of the 8 warps when nseg == 4
    return tr;
}

__device__ __forceinline__ void split_rows(int M, int &r0, int &r1) {
    r0 = (int)(((long long)M * blockIdx.x) / gridDim.x);
    r1 = (int)(((long long)M * (blockIdx.x + 1)) / gridDim.x);
}

// Producer: one lane walks the subs and issues the bulk copies.
__device__ __forceinline__ void produce(const Params &p, const Smem &sm, const Sub *subs, int nsub,
                                        uint64_t policy) {
    uint32_t it = 0;
    const uint32_t ring = smem_u32(sm.ring);
    for (int s = 0; s < nsub; ++s) {
        const Sub &sb = subs[s];
        const int tr = tile_rows(p, sb);
        for (int r = sb.r0; r < sb.r1; r += tr, ++it) {
            const int rows = min(tr, sb.r1 - r);
            const uint32_t bytes = (uint32_t)rows * (uint32_t)sb.N;
            const uint32_t st = it % (uint32_t)p.stages;
            const uint32_t k = it / (uint32_t)p.stages;
            if (k > 0) mbar_wait(smem_u32(&sm.empty[st]), (k - 1) & 1);
            const uint32_t fb = smem_u32(&sm.full[st]);
            mbar_expect_tx(fb, bytes);
            bulk_g2s(ring + st * (uint32_t)p.tile_bytes, sb.base + (size_t)r * sb.N, bytes, fb, policy);
        }
    }
}

// Consumer main loop for one sub. `it` is the running tile counter shared with produce().
template <int CPL>
__device__ __forceinline__ void consume(const Params &p, const Smem &sm, const Sub &sb, uint32_t &it,
                                        int warp, int lane) {
    const int seg_len = sb.N / sb.nseg;
    const int nchunks = seg_len >> 4;
    const int seg = warp % sb.nseg;
    // limb registers of this lane for its segment
    uint4 a0[CPL], a1[CPL], a2[CPL];
    {
        const uint32_t pl = smem_u32(sm.planes + sb.plane_off) + (uint32_t)(seg * seg_len);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunks) {
                a0[i] = lds128(pl + c * 16);
                a1[i] = lds128(pl + sb.N + c * 16);
                a2[i] = lds128(pl + 2 * sb.N + c * 16);
            } else {
                a0[i] = a1[i] = a2[i] = make_uint4(0, 0, 0, 0);
            }
        }
    }
    const int tr = tile_rows(p, sb);
    const uint32_t ring = smem_u32(sm.ring);
    for (int r = sb.r0; r < sb.r1; r += tr, ++it) {
        const int rows = min(tr, sb.r1 - r);
        const uint32_t st = it % (uint32_t)p.stages;
        const uint32_t k = it / (uint32_t)p.stages;
        mbar_wait(smem_u32(&sm.full[st]), k & 1);
        const uint32_t tile = ring + st * (uint32_t)p.tile_bytes;
        const int units = rows * sb.nseg;
        for (int u = warp; u < units; u += kConsumerWarps) {
            const int rl = u / sb.nseg;
            const uint32_t row = tile + (uint32_t)(rl * sb.N + seg * seg_len);
            int s0a = 0, s0b = 0, s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                if (c < nchunks) {
                    const uint4 w = lds128(row + c * 16);
                    s0a = dp4a_ss(w.x, a0[i].x, s0a);
                    s1a = dp4a_ss(w.x, a1[i].x, s1a);
                    s2a = dp4a_ss(w.x, a2[i].x, s2a);
                    s0b = dp4a_ss(w.y, a0[i].y, s0b);
                    s1b = dp4a_ss(w.y, a1[i].y, s1b);
                    s2b = dp4a_ss(w.y, a2[i].y, s2b);
                    s0a = dp4a_ss(w.z, a0[i].z, s0a);
                    s1a = dp4a_ss(w.z, a1[i].z, s1a);
                    s2a = dp4a_ss(w.z, a2[i].z, s2a);
                    s0b = dp4a_ss(w.w, a0[i].w, s0b);
                    s1b = dp4a_ss(w.w, a1[i].w, s1b);
                    s2b = dp4a_ss(w.w, a2[i].w, s2b);
                }
            }
            const int t0 = __reduce_add_sync(0xffffffffu, s0a + s0b);
            const int t1 = __reduce_add_sync(0xffffffffu, s1a + s1b);
            const int t2 = __reduce_add_sync(0xffffffffu, s2a + s2b);
            if (lane == 0) {
                const long long tot = (((long long)t2 << 7) + (long long)t1) * 128 + (long long)t0;
                long long *dst = &sm.res64[sb.res_off + (r - sb.r0) + rl];
                if (sb.nseg == 1) *dst = tot;
                else atomicAdd(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)tot);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.empty[st]));
    }
}

// Common kernel preamble: barrier init. Returns after a CTA-wide sync.
__device__ __forceinline__ void init_barriers(const Params &p, const Smem &sm) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&sm.full[i]), 1);
            mbar_init(smem_u32(&sm.empty[i]), kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// Zero the res64 slots [0, n) (needed where segments accumulate with atomics).
__device__ __forceinline__ void zero_res(const Smem &sm, int n, int ctid) {
    for (int i = ctid; i < n; i += kConsumers) sm.res64[i] = 0;
}

// Reduce the per-CTA partials {max|xs|, sum x*oc} the previous kernel left in `part`.
__device__ __forceinline__ void reduce_partials(const double *part, int nparts, int ctid, double *scratch,
                                                double &vmax, double &off) {
    double s = 0.0, m = 0.0;
    for (int i = ctid; i < nparts; i += kConsumers) {
        m = fmax(m, part[i]);
        s += part[kMaxGrid + i];
    }
    cons_reduce(s, m, scratch, ctid);
    vmax = m;
    off = s;
}

// =======================================================================================
// Kernel 0: x = LN0(double(emb[token]))          (rwkv.cu:513-524)
// =======================================================================================
__global__ void __launch_bounds__(kConsumers) k_embed_ln0(const __grid_constant__ Params p) {
    __shared__ double scratch[16];
    const int ctid = threadIdx.x;
    const int E = p.E;
    const float *row = p.emb + (size_t)p.ctrl->token * E;
    auto loadx = [&](int g, double (&v)[4]) {
        const float4 f = *reinterpret_cast<const float4 *>(row + 4 * g);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    };
    double xmean, x2;
    ln_stats(E, ctid, scratch, loadx, xmean, x2);
    const double *w = p.ln, *b = p.ln + E;
    for (int g = ctid; g < (E >> 2); g += kConsumers) {
        double v[4];
        loadx(g, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * g + e;
            p.x[j] = w[j] * ((v[e] - xmean) / x2) + b[j];
        }
    }
}

// =======================================================================================
// Kernel 1 (per layer): LN1 + token shift -> K,V,R GEMVs -> WKV -> xs_o, partials
//   rwkv.cu:535-545 (meanvar, cuda_layernorm, mixatt, kernel_mm8_threec, kernel_wkvc_forward)
// =======================================================================================
template <int CPL>
__global__ void __launch_bounds__(kThreads, 1) k_att_kvr(const __grid_constant__ Params p, int layer) {
    constexpr int GROUPS = (CPL + 1) / 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, p);
    init_barriers(p, sm);
    const int E = p.E;
    const size_t lo = (size_t)layer * E;
    int c0, c1;
    split_rows(E, c0, c1);
    const int nch = c1 - c0;
    Sub subs[3];
    const int8_t *mats[3] = {p.wk, p.wv, p.wr};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        subs[s].base = mats[s] + (size_t)layer * E * E;
        subs[s].N = E;
        subs[s].r0 = c0;
        subs[s].r1 = c1;
        subs[s].nseg = 1;
        subs[s].plane_off = s * 3 * E;
        subs[s].res_off = s * nch;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) { // producer warp (last warp)
        if (lane == 0) produce(p, sm, subs, 3, policy_evict_first());
        return;
    }
    const int ctid = threadIdx.x;
    // ---- prologue: LN1(x), token shift, scale, quantise --------------------------------
    const double *x = p.x;
    auto loadx = [&](int g, double (&v)[4]) {
        const double2 a = *reinterpret_cast<const double2 *>(x + 4 * g);
        const double2 b = *reinterpret_cast<const double2 *>(x + 4 * g + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    };
    double xmean, x2;
    ln_stats(E, ctid, sm.scratch, loadx, xmean, x2);
    const double *lw = p.ln + (size_t)(4 * layer + 2) * E, *lb = lw + E;
    const double *sxy = p.sxy + (size_t)p.ctrl->slot * p.L * E + lo;
    const double *mk = p.mixk + lo, *mv = p.mixv + lo, *mr = p.mixr + lo;
    const float *rk = p.rk + lo, *rv = p.rv + lo, *rr = p.rr + lo;
    const float *ok = p.ock + lo, *ov = p.ocv + lo, *orr = p.ocr + lo;
    const int ng = E >> 2;
    double xs[GROUPS][3][4];
    double mx[3] = {0, 0, 0}, of[3] = {0, 0, 0};
    int sl0, sl1; // slice of xy_new this CTA publishes
    split_rows(E, sl0, sl1);
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
        const int g = ctid + kConsumers * m;
        if (g < ng) {
            double v[4];
            loadx(g, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * g + e;
                const double ln = lw[j] * ((v[e] - xmean) / x2) + lb[j];
                const double st = sxy[j];
                const float fk = (float)(mk[j] * ln + (1.0 - mk[j]) * st);
                const float fv = (float)(mv[j] * ln + (1.0 - mv[j]) * st);
                const float fr = (float)(mr[j] * ln + (1.0 - mr[j]) * st);
                xs[m][0][e] = (double)fk * (double)rk[j];
                xs[m][1][e] = (double)fv * (double)rv[j];
                xs[m][2][e] = (double)fr * (double)rr[j];
                mx[0] = fmax(mx[0], fabs(xs[m][0][e]));
                mx[1] = fmax(mx[1], fabs(xs[m][1][e]));
                mx[2] = fmax(mx[2], fabs(xs[m][2][e]));
                of[0] += (double)fk * (double)ok[j];
                of[1] += (double)fv * (double)ov[j];
                of[2] += (double)fr * (double)orr[j];
                if (j >= sl0 && j < sl1) p.xy_new[j] = ln;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        cons_reduce(of[s], mx[s], sm.scratch, ctid);
        if (ctid == 0) {
            sm.scal[s] = mx[s] / (double)kQMax;
            sm.scal[3 + s] = of[s];
        }
    }
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
        const int g = ctid + kConsumers * m;
        if (g < ng) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const double inv = mx[s] > 0.0 ? (double)kQMax / mx[s] : 0.0;
                quantize4(xs[m][s], inv, sm.planes + s * 3 * E, E, 4 * g);
            }
        }
    }
    // WKV operands of this CTA's channels, fetched early so their latency hides under the GEMVs
    double aa = 0, bb = 0, wd = 0, ub = 0;
    float ro = 0, oco = 0;
    double *paa = p.saa + (size_t)p.ctrl->slot * p.L * E + lo, *pbb = p.sbb + (size_t)p.ctrl->slot * p.L * E + lo;
    if (ctid < nch) {
        const int c = c0 + ctid;
        aa = paa[c];
        bb = pbb[c];
        wd = p.decay[lo + c];
        ub = p.bonus[lo + c];
        ro = p.ro[lo + c];
        oco = p.oco[lo + c];
    }
    consumer_sync(); // planes + scal visible
    // ---- main loop -----------------------------------------------------------------------
    uint32_t it = 0;
#pragma unroll 1
    for (int s = 0; s < 3; ++s) consume<CPL>(p, sm, subs[s], it, warp, lane);
    consumer_sync(); // res64 complete
    // ---- epilogue: WKV per channel (rwkv.cu:221-259) -------------------------------------
    double pmax = 0.0, poff = 0.0;
    if (ctid < nch) {
        const int c = c0 + ctid;
        const float kf = (float)(sm.scal[0] * (double)sm.res64[ctid] + sm.scal[3]);
        const float vf = (float)(sm.scal[1] * (double)sm.res64[nch + ctid] + sm.scal[4]);
        const float rf = (float)(sm.scal[2] * (double)sm.res64[2 * nch + ctid] + sm.scal[5]);
        const double vv = (double)vf;
        const double e1 = exp(ub + wd + (double)kf);
        double y = (aa + e1 * vv) / (bb + e1);
        y = (1.0 / (1.0 + (double)expf(-rf))) * y;
        const double ek = exp((double)kf), ew = exp(wd);
        paa[c] = (aa + ek * vv) * ew;
        pbb[c] = (bb + ek) * ew;
        const float rw = (float)y; // the out-projection reads float(rwkv) (rwkv.cu:290)
        const double xo = (double)rw * (double)ro;
        p.xs_o[c] = (float)xo;
        pmax = fabs((double)(float)xo);
        poff = (double)rw * (double)oco;
    }
    cons_reduce(poff, pmax, sm.scratch, ctid);
    if (ctid == 0) {
        p.part_o[blockIdx.x] = pmax;
        p.part_o[kMaxGrid + blockIdx.x] = poff;
    }
}

// =======================================================================================
// Kernel 2 (per layer): out-projection GEMV + residual          (rwkv.cu:548-553)
// =======================================================================================
template <int CPL>
__global__ void __launch_bounds__(kThreads, 1) k_att_out(const __grid_constant__ Params p, int layer) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, p);
    init_barriers(p, sm);
    const int E = p.E;
    const size_t lo = (size_t)layer * E;
    Sub sb;
    sb.base = p.wo + (size_t)layer * E * E;
    sb.N = E;
    split_rows(E, sb.r0, sb.r1);
    sb.nseg = 1;
    sb.plane_off = 0;
    sb.res_off = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) {
        if (lane == 0) produce(p, sm, &sb, 1, policy_evict_first());
        return;
    }
    const int ctid = threadIdx.x;
    // publish the token-shift state computed by k_att_kvr (slice owned by this CTA)
    {
        double *sxy = p.sxy + (size_t)p.ctrl->slot * p.L * E + lo;
        for (int j = sb.r0 + ctid; j < sb.r1; j += kConsumers) sxy[j] = p.xy_new[j];
    }
    double vmax, off;
    reduce_partials(p.part_o, gridDim.x, ctid, sm.scratch, vmax, off);
    const double inv = vmax > 0.0 ? (double)kQMax / vmax : 0.0;
    const int ng = E >> 2;
    for (int g = ctid; g < ng; g += kConsumers) {
        const float4 f = *reinterpret_cast<const float4 *>(p.xs_o + 4 * g);
        const double xs[4] = {(double)f.x, (double)f.y, (double)f.z, (double)f.w};
        quantize4(xs, inv, sm.planes, E, 4 * g);
    }
    const int nrows = sb.r1 - sb.r0;
    double xold = 0.0;
    if (ctid < nrows) xold = p.x[sb.r0 + ctid];
    consumer_sync();
    uint32_t it = 0;
    consume<CPL>(p, sm, sb, it, warp, lane);
    consumer_sync();
    if (ctid < nrows) {
        const double s = vmax / (double)kQMax;
        const float y = (float)(s * (double)sm.res64[ctid] + off);
        const float xf = (float)xold + y; // residual round-trips through f32 (rwkv.cu:548-553)
        p.x[sb.r0 + ctid] = (double)xf;
    }
}

// =======================================================================================
// Kernel 3 (per layer): LN2 + token shift -> ffn R (E rows) and ffn K (4E rows) GEMVs ->
//   sigmoid / relu^2 -> sr, xs_v, partials            (rwkv.cu:557-573)
// =======================================================================================
template <int CPL>
__global__ void __launch_bounds__(kThreads, 1) k_ffn_rk(const __grid_constant__ Params p, int layer) {
    constexpr int GROUPS = (CPL + 1) / 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, p);
    init_barriers(p, sm);
    const int E = p.E;
    const size_t lo = (size_t)layer * E;
    Sub subs[2];
    subs[0].base = p.wfr + (size_t)layer * E * E;
    subs[0].N = E;
    split_rows(E, subs[0].r0, subs[0].r1);
    subs[0].nseg = 1;
    subs[0].plane_off = 0;
    subs[0].res_off = 0;
    subs[1].base = p.wfk + (size_t)layer * 4 * E * E;
    subs[1].N = E;
    split_rows(4 * E, subs[1].r0, subs[1].r1);
    subs[1].nseg = 1;
    subs[1].plane_off = 3 * E;
    subs[1].res_off = subs[0].r1 - subs[0].r0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) {
        if (lane == 0) produce(p, sm, subs, 2, policy_evict_first());
        return;
    }
    const int ctid = threadIdx.x;
    const double *x = p.x;
    auto loadx = [&](int g, double (&v)[4]) {
        const double2 a = *reinterpret_cast<const double2 *>(x + 4 * g);
        const double2 b = *reinterpret_cast<const double2 *>(x + 4 * g + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    };
    double xmean, x2;
    ln_stats(E, ctid, sm.scratch, loadx, xmean, x2);
    const double *lw = p.ln + (size_t)(4 * (layer + 1)) * E, *lb = lw + E;
    const double *sdd = p.sdd + (size_t)p.ctrl->slot * p.L * E + lo;
    const double *mk = p.fmixk + lo, *mr = p.fmixr + lo;
    const float *rr = p.rfr + lo, *rk = p.rfk + lo, *orr = p.ocfr + lo, *ok = p.ocfk + lo;
    const int ng = E >> 2;
    double xs[GROUPS][2][4];
    double mx[2] = {0, 0}, of[2] = {0, 0};
    int sl0, sl1;
    split_rows(E, sl0, sl1);
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
        const int g = ctid + kConsumers * m;
        if (g < ng) {
            double v[4];
            loadx(g, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * g + e;
                const double ln = lw[j] * ((v[e] - xmean) / x2) + lb[j];
                const double st = sdd[j];
                // mixffn keeps f64; the GEMV casts to float per element (rwkv.cu:341-342, 290)
                const float fr = (float)(mr[j] * ln + (1.0 - mr[j]) * st);
                const float fk = (float)(mk[j] * ln + (1.0 - mk[j]) * st);
                xs[m][0][e] = (double)fr * (double)rr[j];
                xs[m][1][e] = (double)fk * (double)rk[j];
                mx[0] = fmax(mx[0], fabs(xs[m][0][e]));
                mx[1] = fmax(mx[1], fabs(xs[m][1][e]));
                of[0] += (double)fr * (double)orr[j];
                of[1] += (double)fk * (double)ok[j];
                if (j >= sl0 && j < sl1) p.dd_new[j] = ln;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        cons_reduce(of[s], mx[s], sm.scratch, ctid);
        if (ctid == 0) {
            sm.scal[s] = mx[s] / (double)kQMax;
            sm.scal[3 + s] = of[s];
        }
    }
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
        const int g = ctid + kConsumers * m;
        if (g < ng) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const double inv = mx[s] > 0.0 ? (double)kQMax / mx[s] : 0.0;
                quantize4(xs[m][s], inv, sm.planes + s * 3 * E, E, 4 * g);
            }
        }
    }
    consumer_sync();
    uint32_t it = 0;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) consume<CPL>(p, sm, subs[s], it, warp, lane);
    consumer_sync();
    // ---- epilogue ------------------------------------------------------------------------
    const int nr = subs[0].r1 - subs[0].r0, nk = subs[1].r1 - subs[1].r0;
    for (int i = ctid; i < nr; i += kConsumers) { // sigmoid, rwkv.cu:199-219
        const float y = (float)(sm.scal[0] * (double)sm.res64[i] + sm.scal[3]);
        p.sr[subs[0].r0 + i] = (float)(1.0 / (1.0 + exp(-(double)y)));
    }
    double pmax = 0.0, poff = 0.0;
    const float *rv = p.rfv + (size_t)layer * 4 * E, *ov = p.ocfv + (size_t)layer * 4 * E;
    for (int i = ctid; i < nk; i += kConsumers) { // relu^2, rwkv.cu:177-197
        const int k = subs[1].r0 + i;
        float a = (float)(sm.scal[1] * (double)sm.res64[nr + i] + sm.scal[4]);
        a = a > 0.0f ? a : 0.0f;
        a = a * a;
        const float xv = (float)((double)a * (double)rv[k]);
        p.xs_v[k] = xv;
        pmax = fmax(pmax, (double)xv);
        poff += (double)a * (double)ov[k];
    }
    cons_reduce(poff, pmax, sm.scratch, ctid);
    if (ctid == 0) {
        p.part_v[blockIdx.x] = pmax;
        p.part_v[kMaxGrid + blockIdx.x] = poff;
    }
}

// =======================================================================================
// Kernel 4 (per layer): ffn V GEMV (4E -> E) + x += kv * sigmoid(r)     (rwkv.cu:574-577)
// Rows are 4E bytes long: four warps share a row (one E-byte segment each).
// =======================================================================================
template <int CPL>
__global__ void __launch_bounds__(kThreads, 1) k_ffn_v(const __grid_constant__ Params p, int layer) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, p);
    init_barriers(p, sm);
    const int E = p.E;
    const size_t lo = (size_t)layer * E;
    Sub sb;
    sb.base = p.wfv + (size_t)layer * 4 * E * E;
    sb.N = 4 * E;
    split_rows(E, sb.r0, sb.r1);
    sb.nseg = 4;
    sb.plane_off = 0;
    sb.res_off = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) {
        if (lane == 0) produce(p, sm, &sb, 1, policy_evict_first());
        return;
    }
    const int ctid = threadIdx.x;
    {
        double *sdd = p.sdd + (size_t)p.ctrl->slot * p.L * E + lo;
        for (int j = sb.r0 + ctid; j < sb.r1; j += kConsumers) sdd[j] = p.dd_new[j];
    }
    const int nrows = sb.r1 - sb.r0;
    zero_res(sm, nrows, ctid);
    double vmax, off;
    reduce_partials(p.part_v, gridDim.x, ctid, sm.scratch, vmax, off);
    const double inv = vmax > 0.0 ? (double)kQMax / vmax : 0.0;
    const int ng = E; // 4E / 4 groups
    for (int g = ctid; g < ng; g += kConsumers) {
        const float4 f = *reinterpret_cast<const float4 *>(p.xs_v + 4 * g);
        const double xs[4] = {(double)f.x, (double)f.y, (double)f.z, (double)f.w};
        quantize4(xs, inv, sm.planes, 4 * E, 4 * g);
    }
    double xold = 0.0;
    float srv = 0.0f;
    if (ctid < nrows) {
        xold = p.x[sb.r0 + ctid];
        srv = p.sr[sb.r0 + ctid];
    }
    consumer_sync();
    uint32_t it = 0;
    consume<CPL>(p, sm, sb, it, warp, lane);
    consumer_sync();
    if (ctid < nrows) {
        const double s = vmax / (double)kQMax;
        const float kv = (float)(s * (double)sm.res64[ctid] + off);
        p.x[sb.r0 + ctid] = xold + (double)(kv * srv); // blockout, rwkv.cu:407
    }
}

// =======================================================================================
// Kernel 5: LN_out + head GEMV (E -> V)                                  (rwkv.cu:585-589)
// =======================================================================================
template <int CPL>
__global__ void __launch_bounds__(kThreads, 1) k_head(const __grid_constant__ Params p) {
    constexpr int GROUPS = (CPL + 1) / 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, p);
    init_barriers(p, sm);
    const int E = p.E;
    Sub sb;
    sb.base = p.whead;
    sb.N = E;
    split_rows(kVocab, sb.r0, sb.r1);
    sb.nseg = 1;
    sb.plane_off = 0;
    sb.res_off = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) {
        if (lane == 0) produce(p, sm, &sb, 1, policy_evict_first());
        return;
    }
    const int ctid = threadIdx.x;
    const double *x = p.x;
    auto loadx = [&](int g, double (&v)[4]) {
        const double2 a = *reinterpret_cast<const double2 *>(x + 4 * g);
        const double2 b = *reinterpret_cast<const double2 *>(x + 4 * g + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    };
    double xmean, x2;
    ln_stats(E, ctid, sm.scratch, loadx, xmean, x2);
    const double *lw = p.ln + (size_t)(4 * p.L + 2) * E, *lb = lw + E;
    const int ng = E >> 2;
    double xs[GROUPS][4];
    double mx = 0.0, of = 0.0;
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
        const int g = ctid + kConsumers * m;
        if (g < ng) {
            double v[4];
            loadx(g, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * g + e;
                const float f = (float)(lw[j] * ((v[e] - xmean) / x2) + lb[j]);
                xs[m][e] = (double)f * (double)p.rhead[j];
                mx = fmax(mx, fabs(xs[m][e]));
                of += (double)f * (double)p.ochead[j];
            }
        }
    }
    cons_reduce(of, mx, sm.scratch, ctid);
    const double inv = mx > 0.0 ? (double)kQMax / mx : 0.0;
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
        const int g = ctid + kConsumers * m;
        if (g < ng) quantize4(xs[m], inv, sm.planes, E, 4 * g);
    }
    consumer_sync();
    uint32_t it = 0;
    consume<CPL>(p, sm, sb, it, warp, lane);
    consumer_sync();
    const int nrows = sb.r1 - sb.r0;
    const double s = mx / (double)kQMax;
    for (int i = ctid; i < nrows; i += kConsumers)
        p.logits[sb.r0 + i] = (float)(s * (double)sm.res64[i] + of);
}

// =======================================================================================
// argmax over the logits (first maximum wins, like a sequential `>` scan)
// =======================================================================================
__global__ void __launch_bounds__(1024) k_argmax(const __grid_constant__ Params p) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < kVocab; i += blockDim.x) {
        const float v = p.logits[i];
        if (v > best || (v == best && i < idx)) {
            best = v;
            idx = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > best || (ov == best && oi < idx)) {
            best = ov;
            idx = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        bv[threadIdx.x >> 5] = best;
        bi[threadIdx.x >> 5] = idx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < idx)) {
                best = bv[w];
                idx = bi[w];
            }
        p.ctrl->next = (unsigned long long)(idx == 0x7fffffff ? 0 : idx);
    }
}

// Next forward reads its token from the previous argmax (free-running greedy decode).
__global__ void k_feed_next(const __grid_constant__ Params p) {
    if (threadIdx.x == 0) p.ctrl->token = p.ctrl->next;
}
// Teacher-forced decode: token i of a device-resident stream.
__global__ void k_feed_stream(const __grid_constant__ Params p, const unsigned long long *stream) {
    if (threadIdx.x == 0) p.ctrl->token = stream[p.ctrl->pos++];
}

// =======================================================================================
// Load-time repack: u8 [R][C] (leading dim ldin) -> s8 out[c][r0 + r] (leading dim ldout),
// value ^ 0x80. 64x64-byte tiles through shared memory, 32-bit accesses on both sides.
// =======================================================================================
__global__ void __launch_bounds__(256) k_transpose_xor(const uint8_t *__restrict__ in, size_t ldin, int R, int C,
                                                       int8_t *__restrict__ out, size_t ldout, size_t r0) {
    __shared__ uint8_t t[64][68];
    const int bx = blockIdx.x * 64, by = blockIdx.y * 64; // bx: column block, by: row block
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; // 16 x 16
    for (int rr = ty; rr < 64; rr += 16) {
        const int r = by + rr;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = bx + tx * 4 + e;
            t[rr][tx * 4 + e] = (r < R && c < C) ? in[(size_t)r * ldin + c] : 0;
        }
    }
    __syncthreads();
    for (int cc = ty; cc < 64; cc += 16) {
        const int c = bx + cc;
        if (c >= C) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = by + tx * 4 + e;
            if (r < R) out[(size_t)c * ldout + r0 + r] = (int8_t)(t[tx * 4 + e][cc] ^ 0x80);
        }
    }
}

__global__ void k_exp_table(const double *__restrict__ in, double *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp(in[i]);
}

// oc[j] = 128*r[j] + o[j]
__global__ void k_centre_offsets(const float *__restrict__ r, const float *__restrict__ o, float *__restrict__ oc, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) oc[i] = (float)(128.0 * (double)r[i] + (double)o[i]);
}

// ---------------------------------------------------------------------------------------
// Device-side restatement of the reference sampler (include/rwkv/sampler/typical.h = what the
// reference's typical.h:20-58 actually computes): probs = exp(l)/sum, probs^e with e = uint8(1/temp),
// renormalise, cumulative sums, first index whose cumulative probability reaches the uniform `u`
// drawn on the host. One CTA, every thread owns a contiguous run of the vocabulary. Sums are block
// reductions, so cumulative values can differ from the host's sequential ones by ~1e-13; the kernel
// therefore also returns how far `u` is from the nearest interval boundary, and the caller falls
// back to the host path when that margin is below 1e-9 (probability ~1e-9 per draw): identical
// tokens by construction. out[0] = token, out[1] = margin.
// ---------------------------------------------------------------------------------------
constexpr int kSampleThreads = 1024;
__device__ __forceinline__ double sample_prob(float logit, double total, int exponent) {
    if (exponent == 0) return 1.0;
    const double q = exp((double)logit) / total;
    double v = q;
    for (int e = 1; e < exponent; ++e) v *= q;
    return v;
}
__device__ __forceinline__ double block_sum_scan(double v, double *sh, double &prefix_excl) {
    // inclusive scan over the 1024 threads; returns the block total, prefix_excl = sum of lower threads
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        double t = sh[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        sh[32 + lane] = t; // inclusive totals of the warps
    }
    __syncthreads();
    const double warp_off = w ? sh[32 + w - 1] : 0.0;
    const double total = sh[63];
    prefix_excl = warp_off + x - v;
    __syncthreads();
    return total;
}
__global__ void __launch_bounds__(kSampleThreads) k_sample_typical(const float *logits, int len, int exponent, double u,
                                                                  double *out) {
    __shared__ double sh[64];
    __shared__ double starts[kSampleThreads + 1];
    const int per = (len + kSampleThreads - 1) / kSampleThreads;
    const int i0 = min(len, (int)threadIdx.x * per), i1 = min(len, i0 + per);
    double dummy;
    double part = 0.0;
    for (int i = i0; i < i1; ++i) part += exp((double)logits[i]);
    const double total = block_sum_scan(part, sh, dummy);
    part = 0.0;
    for (int i = i0; i < i1; ++i) part += sample_prob(logits[i], total, exponent);
    double before;
    const double s = block_sum_scan(part, sh, before);
    // thread t's run covers cumulative probability (starts[t], starts[t+1]]: the runs tile [0, inf) exactly
    starts[threadIdx.x] = before / s;
    if (threadIdx.x == 0) {
        starts[kSampleThreads] = 2.0; // cp_last is forced to 1.0: everything above goes to the end
        out[0] = 0.0;                 // u <= 0 (probability 2^-53): nobody claims it; margin 0 sends the caller
        out[1] = 0.0;                 // to the host path
    }
    __syncthreads();
    const double lo = starts[threadIdx.x], hi = starts[threadIdx.x + 1];
    if (i0 < i1 && lo < u && !(hi < u)) {
        double c = lo, margin = 0.0;
        int tok = i1 - 1; // rounding left u just above this run's own running sum: margin 0 -> host path decides
        for (int i = i0; i < i1; ++i) {
            const double prev = c;
            c += sample_prob(logits[i], total, exponent) / s;
            if (i == len - 1 || !(c < u)) {
                tok = i;
                margin = i == len - 1 ? u - prev : fmin(u - prev, c - u);
                break;
            }
        }
        out[0] = (double)tok;
        out[1] = margin;
    }
}

} // namespace rk
