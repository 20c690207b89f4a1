// token_kernel.cuh — the persistent one-token kernel: one CTA per SM, launched cooperatively,
// the whole forward of one token in ONE launch.
//
// Why (profiles/r01a_staged_ffn_rk_ncu.md): with one kernel per phase, ~45 % of each kernel was the
// layernorm / token-shift prologue that all 148 CTAs repeated on the same vectors, and the weight
// stream stopped at every kernel boundary. Here
//   * each CTA owns a fixed slice of the residual stream (E/grid elements, kept in shared memory
//     for the whole token) and does the elementwise work for that slice only;
//   * CTAs exchange the small activation vectors and per-CTA partial reductions through L2, separated
//     by grid barriers (one monotonic counter, release/acquire at gpu scope);
//   * the producer warp never waits for a barrier: weights do not depend on activations, so it keeps
//     filling the shared-memory ring with the NEXT phase's tiles while the consumer warps sit in a
//     barrier, which keeps HBM busy across phase boundaries.
//
// Phase structure per layer (B = grid barrier):
//   [stats -> LN1 + token shift for own slice] B [gather xk,xv,xr; GEMV K,V,R rows of own channels;
//   WKV for own channels] B [gather rwkv; GEMV out-proj rows; residual for own slice] B
//   [stats -> LN2 + token shift for own slice] B [gather xr,xk; GEMV ffn-R rows (own slice) and
//   ffn-K rows; sigmoid / relu^2] B [gather k4; GEMV ffn-V rows; residual for own slice] B
// then [stats -> LN_out for own slice] B [gather; head GEMV; logits (+ local argmax)] (B [argmax]).
//
// Reference mapping is the same as for the staged kernels in kernels.cuh (rwkv.cu:493-593).
#pragma once
#include "kernels.cuh"

namespace rk {

// ---- thread layout of the token kernel -------------------------------------------------------
// Eight consumer warps (two per scheduler, each holding the limbs of one n_embed-byte row segment
// in registers) + one producer warpgroup: its first lane streams the weights, the rest of it only
// donates registers through setmaxnreg. (Sixteen consumer warps with two warps per row segment
// and a 104-register budget measured slower: spills.)
#ifndef RK_CORE_INLINE
#define RK_CORE_INLINE __forceinline__
#endif
#ifndef RK_GATHER_INLINE
#define RK_GATHER_INLINE __noinline__
#endif
constexpr int kTokWarps = 8;                   // one warp per unit of a tile, 232 registers each
constexpr int kTokConsumers = kTokWarps * 32;
#ifndef RK_PRODUCER_THREADS
#define RK_PRODUCER_THREADS 128
#endif
// The producer is a whole warpgroup (one lane of it works) so that setmaxnreg can hand its
// registers to the consumers: 384 threads compile to a budget of 168, consumers then take 232.
// (A 288-thread CTA does not help: 9 warps put 3 on one scheduler, 16384/96 = 170 registers.)
// ptxas still makes its pre-allocation choices against 168: left alone it re-derives thread ids,
// shared-window bases and kernel parameters inside every tile iteration and keeps one shared
// load in flight (ncu r01d: 45 % of the core's samples). The hot loop therefore takes its
// operands through opaque() - values the optimiser cannot rematerialise.
constexpr int kProducerThreads = RK_PRODUCER_THREADS;
constexpr int kTokThreads = kTokConsumers + kProducerThreads;
constexpr int kProducerRegs = 40;
// setmaxnreg budget: the consumers may only take what the producer warpgroup gives back
// (256 x (232 - 168) = 16384 = 128 x (168 - 40)).
constexpr int kConsumerRegs = 232;

__device__ __forceinline__ void tok_sync() { // named barrier 1: the eight consumer warps
    asm volatile("bar.sync 1, %0;" ::"n"(kTokConsumers) : "memory");
}

// Identity the optimiser cannot see through. The layer loop reads ~35 per-layer arrays at
// base + l*E + j; left alone, strength reduction turns every one of them into its own 64-bit
// induction pointer that stays live through the GEMV core (measured: 115 registers live across
// the core, which then has none left to keep more than one shared-memory load in flight).
__device__ __forceinline__ int opaque(int v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ size_t opaque(size_t v) {
    asm volatile("" : "+l"(v));
    return v;
}

// weights (signed bytes) x activation digits (unsigned bytes)
__device__ __forceinline__ int dp4a_su(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- exchange block, peers and scopes -----------------------------------------------------------
// Everything CTAs exchange lives in ONE allocation per GPU (the "exchange block": barrier counter,
// accumulators, activation vectors, logits). With tensor parallelism (tp_size ranks, one GPU each)
// the G GPUs simply form one grid of G*148 CTAs: slices and rows are split over the global CTA
// index, every publish is stored into the exchange block of EVERY rank (peer-mapped over NVLink,
// Params::xch[g]), every read is local, and the grid barrier counts the CTAs of all ranks at system
// scope. tp_size == 1 is the same code with one "peer" (itself) and gpu scope.
template <class T> __device__ __forceinline__ T *peer_ptr(const Params &p, T *local, int g) {
    return reinterpret_cast<T *>(p.xch[g] + (reinterpret_cast<unsigned char *>(local) - p.xch[p.tp_rank]));
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *ptr, bool sys) {
    unsigned int v;
    if (sys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_u64(unsigned long long *ptr, unsigned long long v, bool sys) {
    if (sys) asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
    else asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ void red_max_u64(unsigned long long *ptr, unsigned long long v, bool sys) {
    if (sys) asm volatile("red.relaxed.sys.global.max.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
    else asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}

// Accumulators: the per-CTA partial sums / maxima of a phase are combined with integer atomics into
// a few 64-bit words per rank instead of being published as per-CTA records (G*148 records would
// cost every reader several L2 round trips). Integer adds commute, so the result is still
// bit-deterministic. Sums are fixed point (2^-32; sum x^2: 2^-20), maxima are the bit patterns of
// non-negative doubles (order-preserving as unsigned integers), the arg-max key packs an
// order-preserving image of the logit above ~index (largest logit, then smallest index, wins).
// Three buffers rotate with the phase number: written in phase n, read in phase n+1, cleared in
// phase n+2 by CTA 0 of the owning rank.
constexpr int kAccSlots = 16;
constexpr int kAccS1 = 0, kAccS2 = 1, kAccMax = 2, kAccSum = 5, kAccArg = 8;
__device__ __forceinline__ unsigned long long fx32(double v) {
    return (unsigned long long)__double2ll_rn(v * 4294967296.0);
}
__device__ __forceinline__ double unfx32(unsigned long long u) { return (double)(long long)u * (1.0 / 4294967296.0); }
__device__ __forceinline__ unsigned long long fx20(double v) {
    return (unsigned long long)__double2ll_rn(v * 1048576.0);
}
__device__ __forceinline__ double unfx20(unsigned long long u) { return (double)(long long)u * (1.0 / 1048576.0); }
__device__ __forceinline__ unsigned long long *acc_buf(const Params &p, unsigned int phase) {
    return p.acc + (size_t)(phase % 3u) * kAccSlots;
}

// Grid barrier over the consumer threads of all CTAs of all ranks (the producer warps do not take
// part). `phase` = number of barriers completed so far (monotonic across launches, Ctrl::bar_base);
// it also selects the accumulator buffer. One rank: every CTA adds 1 to gbar and waits for
// phase*grid. Several ranks, hierarchical: every CTA arrives on the rank-local counter lbar (an
// acq_rel RMW chain, at system scope so that the CTA's own peer stores are acknowledged first);
// the last local arriver adds 1 to gbar of EVERY rank; everybody waits for phase*ranks on the own
// gbar. (Flat all-to-all increments made each counter take ranks*148 remote atomics per barrier:
// 6.6 us per barrier on 2 GPUs, 14.7 us on 4.)
// Afterwards CTA 0 clears the accumulator buffer that the phase after next will write.
__device__ __forceinline__ void grid_sync(const Params &p, unsigned int &phase, int ctid) {
    tok_sync();
    ++phase;
    if (ctid == 0) {
        unsigned int spins = 0;
        if (p.tp_size == 1) {
            // release: everything this CTA wrote (ordered before by the bar.sync above) is visible to
            // any thread that observes the increment with an acquire load.
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.gbar) : "memory");
            const unsigned int target = phase * gridDim.x;
            // (polling with relaxed loads and one fence.acq_rel after the loop measured 3 % slower)
            while ((int)(ld_acquire_u32(p.gbar, false) - target) < 0) {
                if (++spins > (1u << 25)) __trap();
            }
        } else {
            unsigned int old;
            asm volatile("atom.acq_rel.sys.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.lbar) : "memory");
            if (old + 1u == phase * gridDim.x) {
                for (int g = 0; g < p.tp_size; ++g)
                    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(peer_ptr(p, p.gbar, g)) : "memory");
            }
            const unsigned int target = phase * (unsigned int)p.tp_size;
            while ((int)(ld_acquire_u32(p.gbar, true) - target) < 0) {
                if (++spins > (1u << 25)) __trap();
            }
        }
    }
    tok_sync();
    if (blockIdx.x == 0 && ctid < kAccSlots) acc_buf(p, phase + 1)[ctid] = 0ull;
}

// Reductions in the phase-boundary code. Only a few threads ever hold data there (the <= 64 slice
// owners, the <= 160 row owners, or one lane per CTA partial), and warp shuffles are scarce (one
// warp-wide SHFL per clock per SM), so nothing here involves all eight warps:
//   owners_reduce : sum / max over the first `nact` threads; warps without data return at once;
//                   the participating warps shuffle-reduce and meet at named barrier 2.
//                   The result is valid in thread 0 only (which publishes it).
// Deterministic (fixed trees). NS sums then NM maxes (maxes are of non-negative values).
struct Red {
    double *buf; // [2][6][8] alternating scratch
    int par;
};
template <int NS, int NM>
__device__ __forceinline__ void owners_reduce(double *s, double *m, Red &rd, int ctid, int nact) {
    static_assert(NS + NM <= 6, "reduction scratch holds six values");
    const int nw = (nact + 31) >> 5;
    const int w = ctid >> 5;
    double *b = rd.buf + rd.par * 48;
    rd.par ^= 1; // every thread toggles on every call, whether or not its warp takes part
    if (w >= nw) return;
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] = warp_sum(s[k]);
#pragma unroll
    for (int k = 0; k < NM; ++k) m[k] = warp_max(m[k]);
    if (nw > 1) {
        if ((ctid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < NS; ++k) b[k * 8 + w] = s[k];
#pragma unroll
            for (int k = 0; k < NM; ++k) b[(NS + k) * 8 + w] = m[k];
        }
        asm volatile("bar.sync 2, %0;" ::"r"(nw * 32) : "memory");
        if (ctid == 0) {
            for (int i = 1; i < nw; ++i) {
#pragma unroll
                for (int k = 0; k < NS; ++k) s[k] += b[k * 8 + i];
#pragma unroll
                for (int k = 0; k < NM; ++k) m[k] = fmax(m[k], b[(NS + k) * 8 + i]);
            }
        }
    }
}

// Position in the shared-memory ring: stage index + parity of the current pass over the ring.
struct RingPos {
    uint32_t stage, phase;
    __device__ __forceinline__ void advance(uint32_t stages) {
        if (++stage == stages) {
            stage = 0;
            phase ^= 1;
        }
    }
};

// ---- row-major streaming ---------------------------------------------------------------------
// Weights are row-major [out][in] int8. A tile is exactly eight work units: a unit is one row
// segment of SEG = n_embed bytes (rows of E bytes: one unit per row, eight rows per tile; rows of
// 4E bytes: four units per row, two rows per tile), so consumer warp w always takes unit w of
// every tile - no dealing logic, no divisions in the loop. The activation limbs of a warp's
// segment live in its registers (CPL 16-byte chunks per lane x 3 planes).
// Measured in tools/ringbench.cu (same loop, L2-resident source): 49-52 B/clk/SM, 2.2x the HBM
// rate; tools/corebench.cu: 64 B/clk/SM for the bare loop. The "lane = row" alternative
// (activations broadcast from shared memory) is capped at 35-42 B/clk/SM by shared-memory
// bandwidth - a broadcast LDS.128 still writes 512 B of registers.
// One issuing lane. Several lanes taking tiles round-robin measured no faster (tools/ringbench.cu)
// and are unsafe: a lane two ring passes ahead aliases the parity of the empty barrier.
constexpr int kProducers = 1;

__device__ __forceinline__ void produce_sub(const Params &p, const Smem &sm, const int8_t *base, int N, int r0, int r1,
                                            RingPos &rp, uint64_t policy, int &tcount, int pw, unsigned long long *ptrace,
                                            long long &last_issue) {
    const uint32_t ring = smem_u32(sm.ring);
    const uint32_t full0 = smem_u32(sm.full), empty0 = smem_u32(sm.empty);
    const int tr = (8 * p.E) / N; // rows per tile: 8 (N = E) or 2 (N = 4E)
    for (int r = r0; r < r1; r += tr) {
        if ((tcount & (kProducers - 1)) == pw) {
            const uint32_t bytes = (uint32_t)(min(tr, r1 - r) * N);
            // first pass over the ring: a fresh mbarrier reports the "previous" phase as complete
            mbar_wait(empty0 + 8 * rp.stage, rp.phase ^ 1);
            if (p.issue_gap > 0) {
                // optional pacing of the bulk copies (set_option "issue_gap", SM cycles): a burst of five
                // 32 KB tiles at a phase boundary queues ahead of the consumers' latency-critical gather
                // loads. Measured: 1000 cycles +0.9 % (518 -> 523 tok/s), i.e. not the main effect; off by default.
                while (clock64() - last_issue < (long long)p.issue_gap) __nanosleep(32);
                last_issue = clock64();
            }
            const uint32_t fb = full0 + 8 * rp.stage;
            mbar_expect_tx(fb, bytes);
            bulk_g2s(ring + rp.stage * (uint32_t)p.tile_bytes, base + (size_t)r * N, bytes, fb, policy);
            if (ptrace != nullptr && tcount < kTileTraceMax) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                ptrace[(size_t)blockIdx.x * kTileTraceMax + tcount] = t;
            }
        }
        ++tcount;
        rp.advance((uint32_t)p.stages);
    }
}

// Half of a unit's weight bytes -> registers: chunks [HALF*H, HALF*H + H) of the lane.
template <int H, int HALF, bool FULL>
__device__ __forceinline__ void load_half(uint4 (&w)[H], uint32_t row, int lane, int nchunks) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
        if (FULL || lane + 32 * (HALF * H + i) < nchunks) w[i] = lds128(row + (HALF * H + i) * 512);
        else w[i] = make_uint4(0, 0, 0, 0);
    }
}

// Exact int32 dot products of half a unit against the three limb planes (two chains per plane).
template <int CPL, int H, int HALF>
__device__ __forceinline__ void dot_half(const uint4 (&w)[H], const uint4 (&a0)[CPL], const uint4 (&a1)[CPL],
                                         const uint4 (&a2)[CPL], int (&acc)[6]) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
        constexpr int o = HALF * H;
        acc[0] = dp4a_su(w[i].x, a0[o + i].x, acc[0]);
        acc[2] = dp4a_su(w[i].x, a1[o + i].x, acc[2]);
        acc[4] = dp4a_ss(w[i].x, a2[o + i].x, acc[4]);
        acc[1] = dp4a_su(w[i].y, a0[o + i].y, acc[1]);
        acc[3] = dp4a_su(w[i].y, a1[o + i].y, acc[3]);
        acc[5] = dp4a_ss(w[i].y, a2[o + i].y, acc[5]);
        acc[0] = dp4a_su(w[i].z, a0[o + i].z, acc[0]);
        acc[2] = dp4a_su(w[i].z, a1[o + i].z, acc[2]);
        acc[4] = dp4a_ss(w[i].z, a2[o + i].z, acc[4]);
        acc[1] = dp4a_su(w[i].w, a0[o + i].w, acc[1]);
        acc[3] = dp4a_su(w[i].w, a1[o + i].w, acc[3]);
        acc[5] = dp4a_ss(w[i].w, a2[o + i].w, acc[5]);
    }
}

// Consumer side of one streamed sub-matrix: warp w takes unit w of every tile. All arguments are
// plain values in registers (the hot loop takes them through opaque(), see above).
// NSEG = N / E (1 or 4). planes: shared address of limb plane 0 of this sub's activation vector
// (planes 1, 2 at +N, +2N); res: shared address of this sub's int64 [row][NSEG] partial totals.
// Variants measured and dropped (gpurun_out A/B of 2026-09-24, 7B shape, tokens/s): this loop 525;
// all eight loads pinned ahead of the arithmetic with a warp barrier 512; the same as a noinline
// function 497; software-pipelined in half units (loads of tile t+1 under the arithmetic of tile t,
// 0.50 instead of 0.64 us per 32 KB tile) as a noinline function 430 - the faster consumer let the
// producer put 24 MB of bulk copies in flight at every phase boundary and the latency-critical
// gather loads queued behind them (gather 2.9 -> 6.9 us); inlined it exceeds the register budget.
template <int CPL, bool FULL, int NSEG>
__device__ RK_CORE_INLINE RingPos consume_sub(uint32_t ring, uint32_t full0, uint32_t empty0, uint32_t tile_bytes,
                                            uint32_t stages, uint32_t planes, uint32_t res, int N, int nr, RingPos rp,
                                            int warp, int lane, unsigned long long *ptrace, int *tile_cnt) {
    static_assert(CPL % 2 == 0, "chunks per lane must be even");
    constexpr int TR = 8 / NSEG; // rows per tile
    constexpr int H = CPL / 2;
    const int seg_len = N / NSEG;
    const int nchunks = seg_len >> 4;
    const int seg = warp % NSEG, rl = warp / NSEG; // this warp's unit inside every tile
    const uint32_t unit_off = (uint32_t)(rl * N + seg * seg_len + lane * 16);
    const int ntiles = (nr + TR - 1) / TR;
    if (ntiles <= 0) return rp;
    uint4 a0[CPL], a1[CPL], a2[CPL];
    {
        const uint32_t pl = planes + (uint32_t)(seg * seg_len);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            if (FULL || c < nchunks) {
                a0[i] = lds128(pl + c * 16);
                a1[i] = lds128(pl + N + c * 16);
                a2[i] = lds128(pl + 2 * N + c * 16);
            } else {
                a0[i] = a1[i] = a2[i] = make_uint4(0, 0, 0, 0);
            }
        }
    }
    uint32_t dst = res + (uint32_t)((rl * NSEG + seg) * 8);
    int row = rl;
    for (int t = 0; t < ntiles; ++t) {
        mbar_wait(full0 + 8 * rp.stage, rp.phase);
        if (ptrace != nullptr && threadIdx.x == 0) {
            const int c = *tile_cnt;
            if (c < kTileTraceMax) {
                unsigned long long tm;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm));
                ptrace[((size_t)gridDim.x + blockIdx.x) * kTileTraceMax + c] = tm;
            }
            *tile_cnt = c + 1;
        }
        if (row < nr) {
            const uint32_t wrow = ring + rp.stage * tile_bytes + unit_off;
            uint4 h0[H], h1[H];
            load_half<H, 0, FULL>(h0, wrow, lane, nchunks);
            load_half<H, 1, FULL>(h1, wrow, lane, nchunks);
            int acc[6] = {0, 0, 0, 0, 0, 0};
            dot_half<CPL, H, 0>(h0, a0, a1, a2, acc);
            dot_half<CPL, H, 1>(h1, a0, a1, a2, acc);
            const int t0 = __reduce_add_sync(0xffffffffu, acc[0] + acc[1]);
            const int t1 = __reduce_add_sync(0xffffffffu, acc[2] + acc[3]);
            const int t2 = __reduce_add_sync(0xffffffffu, acc[4] + acc[5]);
            if (lane == 0) {
                const long long tot = (((long long)t2 << 8) + (long long)t1) * 256 + (long long)t0;
                asm volatile("st.shared.u64 [%0], %1;" ::"r"(dst), "l"(tot) : "memory");
            }
        }
        dst += 8 * 8;
        row += TR;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * rp.stage);
        rp.advance(stages);
    }
    return rp;
}

// Slice ownership of this CTA.
struct Slices {
    int e0, e1, ne; // residual-stream elements / att channels / rows of every E-row matrix
    int k0, k1, nk; // rows of the 4E-row ffn key matrix
    int v0, v1, nv; // rows of the head
};
__device__ __forceinline__ void split_rows_g(int M, int gb, int gn, int &r0, int &r1) {
    r0 = (int)(((long long)M * gb) / gn);
    r1 = (int)(((long long)M * (gb + 1)) / gn);
}
// gb / gn: index of this CTA in, and size of, the grid formed by all ranks
__device__ __forceinline__ Slices make_slices(int E, int gb, int gn) {
    Slices s;
    split_rows_g(E, gb, gn, s.e0, s.e1);
    s.ne = s.e1 - s.e0;
    split_rows_g(4 * E, gb, gn, s.k0, s.k1);
    s.nk = s.k1 - s.k0;
    split_rows_g(kVocab, gb, gn, s.v0, s.v1);
    s.nv = s.v1 - s.v0;
    return s;
}

// The producer's whole-token schedule. MUST enumerate subs in exactly the consumers' order.
// TRACE: the debug time stamps are compiled in (a separate instantiation: the branches alone cost 2 %).
template <bool TRACE>
__device__ __forceinline__ void produce_token(const Params &p, const Smem &sm, const Slices &sl, int pw) {
    unsigned long long *const ptrace = TRACE ? p.ptrace : nullptr;
    // evict_first keeps the 7 GB/token weight stream from displacing the exchange vectors and the per-layer
    // parameters in L2: with evict_normal the same kernel runs at 441 instead of 516 tok/s.
    const uint64_t pol = policy_evict_first();
    const int E = p.E;
    RingPos rp{0, 0};
    int tcount = 0;
    long long last_issue = 0;
    for (int l = 0; l < p.L_run; ++l) {
        const size_t mo = (size_t)l * E * E;
        produce_sub(p, sm, p.wk + mo, E, sl.e0, sl.e1, rp, pol, tcount, pw, ptrace, last_issue);
        produce_sub(p, sm, p.wv + mo, E, sl.e0, sl.e1, rp, pol, tcount, pw, ptrace, last_issue);
        produce_sub(p, sm, p.wr + mo, E, sl.e0, sl.e1, rp, pol, tcount, pw, ptrace, last_issue);
        produce_sub(p, sm, p.wo + mo, E, sl.e0, sl.e1, rp, pol, tcount, pw, ptrace, last_issue);
        produce_sub(p, sm, p.wfr + mo, E, sl.e0, sl.e1, rp, pol, tcount, pw, ptrace, last_issue);
        produce_sub(p, sm, p.wfk + 4 * mo, E, sl.k0, sl.k1, rp, pol, tcount, pw, ptrace, last_issue);
        produce_sub(p, sm, p.wfv + 4 * mo, 4 * E, sl.e0, sl.e1, rp, pol, tcount, pw, ptrace, last_issue);
    }
    produce_sub(p, sm, p.whead, E, sl.v0, sl.v1, rp, pol, tcount, pw, ptrace, last_issue);
}

// mean / std of the full residual stream from the accumulated sum(x), sum(x^2), with the
// reference's f32 rounding of the two accumulators (rwkv.cu:412-465, 43-44).
// sum((x-m)^2) = s2 - 2 m s1 + E m^2. Every calling lane loads the two words itself (broadcast).
__device__ __forceinline__ void stats_from_acc(const Params &p, const unsigned long long *acc, double &xmean, double &x2) {
    const double s1 = unfx32(__ldcg(acc + kAccS1)), s2 = unfx20(__ldcg(acc + kAccS2));
    const double E = (double)p.E;
    const float mean_acc = (float)s1;
    const double mean_f = (double)(mean_acc / (float)p.E);
    double var = s2 - 2.0 * mean_f * s1 + E * mean_f * mean_f;
    if (var < 0.0) var = 0.0;
    const float var_acc = (float)var;
    xmean = (double)mean_acc / E;
    x2 = (double)sqrtf(var_acc / (float)(p.E - 1));
}

// Accumulate this CTA's {sum x, sum x^2} of its slice on every rank.
__device__ __forceinline__ void publish_stats(const Params &p, const Smem &sm, unsigned int phase, int ne, Red &rd, int ctid) {
    double s[2] = {0.0, 0.0};
    if (ctid < ne) {
        const double v = sm.xown[ctid];
        s[0] = v;
        s[1] = v * v;
    }
    owners_reduce<2, 0>(s, nullptr, rd, ctid, ne);
    if (ctid == 0 && ne > 0) {
        unsigned long long *acc = acc_buf(p, phase);
        const bool sys = p.tp_size > 1;
        for (int g = 0; g < p.tp_size; ++g) {
            unsigned long long *a = peer_ptr(p, acc, g);
            red_add_u64(a + kAccS1, fx32(s[0]), sys);
            red_add_u64(a + kAccS2, fx20(s[1]), sys);
        }
    }
}

// Accumulate per-vector {max |xs|, sum x*oc} of this CTA (data in the first `nact` threads).
template <int NVEC>
__device__ __forceinline__ void publish_vparts(const Params &p, unsigned int phase, double *mx, double *of, Red &rd,
                                               int ctid, int nact) {
    owners_reduce<NVEC, NVEC>(of, mx, rd, ctid, nact);
    if (ctid == 0 && nact > 0) {
        unsigned long long *acc = acc_buf(p, phase);
        const bool sys = p.tp_size > 1;
        for (int g = 0; g < p.tp_size; ++g) {
            unsigned long long *a = peer_ptr(p, acc, g);
#pragma unroll
            for (int v = 0; v < NVEC; ++v) {
                red_max_u64(a + kAccMax + v, (unsigned long long)__double_as_longlong(mx[v]), sys);
                red_add_u64(a + kAccSum + v, fx32(of[v]), sys);
            }
        }
    }
}

// Activation quantiser of the token kernel: q = round(xs * inv) as a 23-bit two's complement integer
// (|q| <= 2^22 - 1), and the three limb planes are simply its three low BYTES: bytes 0 and 1 are
// unsigned digits, byte 2 is the signed top digit, q = b2*65536 + b1*256 + b0. The GEMV uses
// dp4a.s32.u32 for the two unsigned planes and dp4a.s32.s32 for the signed one (all exact int32).
// Rounding goes through the float adder (1.5*2^23 + x has ulp 1; the low mantissa bits are the
// integer) - no F2I, no per-digit bit surgery; four elements are transposed with seven PRMTs.
constexpr int kQMaxTok = 4194303; // 2^22 - 1
__device__ __forceinline__ uint32_t round_q(float xs, float inv) {
    return __float_as_uint(fmaf(xs, inv, 12582912.0f)) - 0x4B400000u; // two's complement q
}
__device__ __forceinline__ void quantize4f(const float4 f, float inv, uint8_t *planes, int stride, int j) {
    const uint32_t t0 = round_q(f.x, inv), t1 = round_q(f.y, inv), t2 = round_q(f.z, inv), t3 = round_q(f.w, inv);
    const uint32_t lo01 = __byte_perm(t0, t1, 0x5140), lo23 = __byte_perm(t2, t3, 0x5140); // [a.b0,b.b0,a.b1,b.b1]
    const uint32_t hi01 = __byte_perm(t0, t1, 0x0062), hi23 = __byte_perm(t2, t3, 0x0062); // [a.b2,b.b2,..]
    *reinterpret_cast<uint32_t *>(planes + j) = __byte_perm(lo01, lo23, 0x5410);
    *reinterpret_cast<uint32_t *>(planes + stride + j) = __byte_perm(lo01, lo23, 0x7632);
    *reinterpret_cast<uint32_t *>(planes + 2 * stride + j) = __byte_perm(hi01, hi23, 0x5410);
}

// Debug tracing: thread 0 of each CTA appends %globaltimer to its row of `trace`.
__device__ __forceinline__ void trace_stamp(unsigned long long *trace, const Smem &sm, int ctid) {
    if (trace != nullptr && ctid == 0) {
        int *cnt = reinterpret_cast<int *>(sm.scal + 8);
        const int c = *cnt;
        if (c < kTraceMax) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            trace[(size_t)blockIdx.x * kTraceMax + c] = t;
            *cnt = c + 1;
        }
    }
}

__device__ __forceinline__ void trace_stamp2(unsigned long long *trace, double *scal, int ctid) {
    if (trace != nullptr && ctid == 0) {
        int *cnt = reinterpret_cast<int *>(scal + 8);
        const int c = *cnt;
        if (c < kTraceMax) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            trace[(size_t)blockIdx.x * kTraceMax + c] = t;
            *cnt = c + 1;
        }
    }
}

// After a barrier: fetch the `nvec` activation vectors of length N (vector v -> limb planes at offset
// v*3*N) and the accumulated {max |xs|, sum x*oc}, then quantise from registers. All loads (up to
// 24 x 16 B per thread) are issued before anything is consumed, so the whole gather costs ONE L2
// round trip; every CTA walks the vector from a different starting offset so that the CTAs do not
// hit the same L2 lines at the same moment.
constexpr int kGatherBatches = 6; // x 4 float4 groups x 256 threads x 4 elements >= 4*5120
__device__ RK_GATHER_INLINE void gather_quantise(uint8_t *planes, double *scal, const float *vec,
                                             const unsigned long long *acc, int nvec, int N, int ctid, int rot_num,
                                             int rot_den, unsigned long long *trace) {
    const int ng = N >> 2;                                             // float4 groups per vector
    const int nb = (ng + 4 * kTokConsumers - 1) / (4 * kTokConsumers); // batches of 4 groups per thread per vector
    const int total = nvec * nb;                                       // <= kGatherBatches
    const int rot = (int)(((long long)ng * rot_num) / rot_den);
    float4 f[kGatherBatches][4];
#pragma unroll
    for (int t = 0; t < kGatherBatches; ++t) {
        if (t < total) {
            const int v = t / nb, b = t - v * nb;
            const float4 *src = reinterpret_cast<const float4 *>(vec + (size_t)v * N);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int g = ctid + kTokConsumers * (4 * b + k);
                int gg = g + rot;
                if (gg >= ng) gg -= ng;
                f[t][k] = g < ng ? __ldcg(src + gg) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    if (ctid < 3) {
        double mm = 0.0, ss = 0.0;
        if (ctid < nvec) {
            mm = __longlong_as_double((long long)__ldcg(acc + kAccMax + ctid));
            ss = unfx32(__ldcg(acc + kAccSum + ctid));
        }
        scal[ctid] = mm / (double)kQMaxTok;
        scal[3 + ctid] = ss;
        reinterpret_cast<float *>(scal + 6)[ctid] = mm > 0.0 ? (float)((double)kQMaxTok / mm) : 0.0f;
    }
    trace_stamp2(trace, scal, ctid); // loads issued
    tok_sync();
    trace_stamp2(trace, scal, ctid); // scales known
    float inv[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) inv[v] = reinterpret_cast<const float *>(scal + 6)[v];
#pragma unroll
    for (int t = 0; t < kGatherBatches; ++t) {
        if (t < total) {
            const int v = t / nb, b = t - v * nb;
            const float iv = v == 0 ? inv[0] : v == 1 ? inv[1] : inv[2];
            uint8_t *pl = planes + (size_t)v * 3 * N;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int g = ctid + kTokConsumers * (4 * b + k);
                int gg = g + rot;
                if (gg >= ng) gg -= ng;
                if (g < ng) quantize4f(f[t][k], iv, pl, N, 4 * gg);
            }
        }
    }
    trace_stamp2(trace, scal, ctid); // own quantisation done
    tok_sync();
}

// CPL: 16-byte chunks per lane of one n_embed-byte row segment; FULL: n_embed == CPL*512;
// TRACE: with the %globaltimer stamps of tools/trace_token.py (set_option("trace", 1)).
template <int CPL, bool FULL, bool TRACE>
__global__ void __launch_bounds__(kTokThreads, 1) k_token(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, p);
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&sm.full[i]), 1);
            mbar_init(smem_u32(&sm.empty[i]), kTokWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int E = p.E;
    const int gn = (int)gridDim.x * p.tp_size, gb = p.tp_rank * (int)gridDim.x + (int)blockIdx.x;
    const Slices sl = make_slices(E, gb, gn);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= kTokWarps) {
        if (kProducerThreads == 128) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        if (lane == 0 && warp - kTokWarps < kProducers) produce_token<TRACE>(p, sm, sl, warp - kTokWarps);
        return;
    }
    if (kProducerThreads == 128) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    const int ctid = threadIdx.x;
    const int ne = sl.ne, nk = sl.nk;
    const int nwe = (ne + 31) >> 5; // warps that own residual elements
    Red rd{sm.red, 0};
    if (ctid == 0) {
        *reinterpret_cast<int *>(sm.scal + 8) = 0;
        *reinterpret_cast<int *>(sm.scal + 9) = 0;
    }
    unsigned long long *const c_trace = TRACE ? p.trace : nullptr;
    auto stamp = [&]() {
        if (TRACE) trace_stamp(c_trace, sm, ctid);
    };
    stamp();
    const bool mine = ctid < ne;      // this thread owns residual element j
    const int j0 = sl.e0 + (mine ? ctid : 0);
    const int j = j0;
    Ctrl *ctrl = p.ctrl;
    unsigned int phase = ctrl->bar_base;             // barriers completed so far: phase number, selects the accumulator buffer
    unsigned long long token = ctrl->token;
    if (p.feed_mode == 1) token = ctrl->next;
    else if (p.feed_mode == 2) token = p.stream[ctrl->pos];
    const size_t so0 = (size_t)ctrl->slot * p.L * E; // state slot offset
    const size_t so = so0;
    unsigned int q = 0;                              // exchange-buffer parity counter (one per barrier)
    const bool sys = p.tp_size > 1;
    RingPos rp{0, 0};
    const uint32_t c_ring = opaque(smem_u32(sm.ring)), c_full = opaque(smem_u32(sm.full)),
                   c_empty = opaque(smem_u32(sm.empty));
    const uint32_t c_planes = opaque(smem_u32(sm.planes)), c_res = opaque(smem_u32(sm.res64));
    const uint32_t c_tile = opaque((uint32_t)p.tile_bytes), c_stages = opaque((uint32_t)p.stages);
    const int c_warp = opaque(warp), c_lane = opaque(lane);
    unsigned long long *const c_ptrace = TRACE ? reinterpret_cast<unsigned long long *>(opaque((size_t)p.ptrace)) : nullptr;
    int *const c_tcnt = reinterpret_cast<int *>(sm.scal + 9);
    // exact integer total of row `i` of the sub whose partials start at res64[off] (nseg per row)
    auto row_total = [&](int off, int i, int nseg) {
        long long t = 0;
        for (int sgm = 0; sgm < nseg; ++sgm) t += sm.res64[off + i * nseg + sgm];
        return (double)t;
    };
    auto vecp = [&](unsigned int qq) { return p.vec + (size_t)(qq & 1) * 4 * E; };
    // store one value of a published vector into the exchange block of every rank
    auto put = [&](float *local, float v) {
        if (!sys) { *local = v; return; }
        for (int g = 0; g < p.tp_size; ++g) *peer_ptr(p, local, g) = v;
    };

    // ---- x = LN0(emb[token]) for the own slice (rwkv.cu:513-524) --------------------------------
    {
        const float *row = p.emb + (size_t)token * E;
        auto loadx = [&](int g, double (&v)[4]) {
            const float4 f = *reinterpret_cast<const float4 *>(row + 4 * g);
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        };
        // two-pass statistics of the embedding row with the reference's f32 rounding (rwkv.cu:412-465)
        double *sc = sm.red; // [32] scratch: 16 warp sums
        double sacc = 0.0;
        for (int g = ctid; g < (E >> 2); g += kTokConsumers) {
            double v[4];
            loadx(g, v);
            sacc += (v[0] + v[1]) + (v[2] + v[3]);
        }
        sacc = warp_sum(sacc);
        if (lane == 0) sc[warp] = sacc;
        tok_sync();
        double tot = 0.0;
        for (int w = 0; w < kTokWarps; ++w) tot += sc[w];
        const float mean_acc = (float)tot;
        const double mean_f = (double)(mean_acc / (float)E);
        double qacc = 0.0;
        for (int g = ctid; g < (E >> 2); g += kTokConsumers) {
            double v[4];
            loadx(g, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) qacc += (v[e] - mean_f) * (v[e] - mean_f);
        }
        qacc = warp_sum(qacc);
        if (lane == 0) sc[16 + warp] = qacc;
        tok_sync();
        double qtot = 0.0;
        for (int w = 0; w < kTokWarps; ++w) qtot += sc[16 + w];
        const double xmean = (double)mean_acc / (double)E;
        const double x2 = (double)sqrtf((float)qtot / (float)(E - 1));
        tok_sync(); // scratch is reused by owners_reduce below
        if (mine) sm.xown[ctid] = p.ln[j] * (((double)row[j] - xmean) / x2) + p.ln[E + j];
        publish_stats(p, sm, phase, ne, rd, ctid);
    }
    // parameters of the first LN1 / token-shift slice computation
    double lw = 0, lb = 0, mk = 0, mv = 0, mr = 0, st = 0;
    float rk = 0, rv = 0, rr = 0, ok = 0, ov = 0, orr = 0;
    auto prefetch_att = [&](int l, int j, size_t so) {
        {   // unconditional (j is clamped to an owned element): a guarded assignment would keep the old
            // values live through the whole layer and push them into local memory
            const size_t lo = (size_t)l * E + j;
            lw = p.ln[(size_t)(4 * l + 2) * E + j];
            lb = p.ln[(size_t)(4 * l + 3) * E + j];
            mk = p.mixk[lo]; mv = p.mixv[lo]; mr = p.mixr[lo];
            rk = p.rk[lo]; rv = p.rv[lo]; rr = p.rr[lo];
            ok = p.ock[lo]; ov = p.ocv[lo]; orr = p.ocr[lo];
            st = p.sxy[so + lo];
        }
    };
    if (p.L_run > 0) prefetch_att(0, j0, so0);
    stamp();
    grid_sync(p, phase, ctid);
    stamp();
    ++q;

    for (int l = 0; l < p.L_run; ++l) {
        const int j = opaque(j0);
        const size_t so = opaque(so0);
        const int lq = opaque(l);
        const size_t lo = (size_t)lq * E;
        // ======== LN1 + token shift for the own slice (rwkv.cu:535-540) ==========================
        {
            double xmean = 0.0, x2 = 1.0;
            if (warp < nwe) stats_from_acc(p, acc_buf(p, phase - 1), xmean, x2);
            double mx[3] = {0, 0, 0}, of[3] = {0, 0, 0};
            if (mine) {
                const double ln = lw * ((sm.xown[ctid] - xmean) * (1.0 / x2)) + lb;
                const float fk = (float)(mk * ln + (1.0 - mk) * st);
                const float fv = (float)(mv * ln + (1.0 - mv) * st);
                const float fr = (float)(mr * ln + (1.0 - mr) * st);
                const float xk = (float)((double)fk * (double)rk);
                const float xv = (float)((double)fv * (double)rv);
                const float xr = (float)((double)fr * (double)rr);
                float *vec = vecp(q);
                put(vec + j, xk);
                put(vec + E + j, xv);
                put(vec + 2 * E + j, xr);
                mx[0] = fabs((double)xk); mx[1] = fabs((double)xv); mx[2] = fabs((double)xr);
                of[0] = (double)fk * (double)ok; of[1] = (double)fv * (double)ov; of[2] = (double)fr * (double)orr;
                p.sxy[so + lo + j] = ln; // only the owner ever reads or writes this element
            }
            publish_vparts<3>(p, phase, mx, of, rd, ctid, ne);
        }
        stamp();
    grid_sync(p, phase, ctid);
    stamp();
        ++q;
        // ======== K, V, R GEMVs for the own channels + WKV (rwkv.cu:542-545) =====================
        gather_quantise(sm.planes, sm.scal, vecp(q - 1), acc_buf(p, phase - 1), 3, E, ctid, gb, gn, c_trace);
        stamp();
        {
            double aa = 0, bb = 0, wd = 0, ub = 0, ewd = 0;
            float ro = 0, oco = 0;
            if (mine) {
                aa = p.saa[so + lo + j];
                bb = p.sbb[so + lo + j];
                wd = p.decay[lo + j];
                ub = p.bonus[lo + j];
                ewd = p.expdecay[lo + j];
                ro = p.ro[lo + j];
                oco = p.oco[lo + j];
            }
            const size_t mo = (size_t)lq * E * E;
            (void)mo;
            rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(0), c_res + (uint32_t)(0) * 8u, E, sl.e1 - sl.e0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
            rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(3 * E), c_res + (uint32_t)(ne * 1) * 8u, E, sl.e1 - sl.e0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
            rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(6 * E), c_res + (uint32_t)(2 * ne * 1) * 8u, E, sl.e1 - sl.e0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
            tok_sync();
            stamp();
            double mx[3] = {0, 0, 0}, of[3] = {0, 0, 0};
            if (mine) {
                const float kf = (float)(sm.scal[0] * row_total(0, ctid, 1) + sm.scal[3]);
                const float vf = (float)(sm.scal[1] * row_total(ne * 1, ctid, 1) + sm.scal[4]);
                const float rf = (float)(sm.scal[2] * row_total(2 * ne * 1, ctid, 1) + sm.scal[5]);
                const double vv = (double)vf;
                const double e1 = exp(ub + wd + (double)kf);
                double y = (aa + e1 * vv) / (bb + e1);
                y = (1.0 / (1.0 + (double)expf(-rf))) * y;
                const double ek = exp((double)kf), ew = ewd; // exp(decay) is static: tabulated at load
                p.saa[so + lo + j] = (aa + ek * vv) * ew;
                p.sbb[so + lo + j] = (bb + ek) * ew;
                const float rw = (float)y;
                const float xo = (float)((double)rw * (double)ro);
                put(vecp(q) + j, xo);
                mx[0] = fabs((double)xo);
                of[0] = (double)rw * (double)oco;
            }
            publish_vparts<1>(p, phase, mx, of, rd, ctid, ne);
        }
        stamp();
    grid_sync(p, phase, ctid);
    stamp();
        ++q;
        // ======== out-projection + residual (rwkv.cu:548-553) =====================================
        gather_quantise(sm.planes, sm.scal, vecp(q - 1), acc_buf(p, phase - 1), 1, E, ctid, gb, gn, c_trace);
        stamp();
        // parameters of the LN2 / ffn token-shift slice computation (used two barriers later)
        double flw = 0, flb = 0, fmk = 0, fmr = 0, fst = 0;
        float frr = 0, frk = 0, forr = 0, fok = 0;
        if (mine) {
            flw = p.ln[(size_t)(4 * (lq + 1)) * E + j];
            flb = p.ln[(size_t)(4 * (lq + 1) + 1) * E + j];
            fmk = p.fmixk[lo + j]; fmr = p.fmixr[lo + j];
            frr = p.rfr[lo + j]; frk = p.rfk[lo + j];
            forr = p.ocfr[lo + j]; fok = p.ocfk[lo + j];
            fst = p.sdd[so + lo + j];
        }
        rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(0), c_res + (uint32_t)(0) * 8u, E, sl.e1 - sl.e0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
        tok_sync();
        stamp();
        if (mine) {
            const float y = (float)(sm.scal[0] * row_total(0, ctid, 1) + sm.scal[3]);
            const float xf = (float)sm.xown[ctid] + y;
            sm.xown[ctid] = (double)xf;
        }
        publish_stats(p, sm, phase, ne, rd, ctid);
        stamp();
    grid_sync(p, phase, ctid);
    stamp();
        ++q;
        // ======== LN2 + token shift for the own slice (rwkv.cu:557-562) ===========================
        {
            double xmean = 0.0, x2 = 1.0;
            if (warp < nwe) stats_from_acc(p, acc_buf(p, phase - 1), xmean, x2);
            double mx[3] = {0, 0, 0}, of[3] = {0, 0, 0};
            if (mine) {
                const double ln = flw * ((sm.xown[ctid] - xmean) * (1.0 / x2)) + flb;
                const float fr = (float)(fmr * ln + (1.0 - fmr) * fst);
                const float fk = (float)(fmk * ln + (1.0 - fmk) * fst);
                const float xr = (float)((double)fr * (double)frr);
                const float xk = (float)((double)fk * (double)frk);
                float *vec = vecp(q);
                put(vec + j, xr);
                put(vec + E + j, xk);
                mx[0] = fabs((double)xr); mx[1] = fabs((double)xk);
                of[0] = (double)fr * (double)forr; of[1] = (double)fk * (double)fok;
                p.sdd[so + lo + j] = ln;
            }
            publish_vparts<2>(p, phase, mx, of, rd, ctid, ne);
        }
        stamp();
    grid_sync(p, phase, ctid);
    stamp();
        ++q;
        // ======== ffn R (own slice rows) and ffn K (4E rows) + sigmoid / relu^2 (rwkv.cu:566-573) ==
        gather_quantise(sm.planes, sm.scal, vecp(q - 1), acc_buf(p, phase - 1), 2, E, ctid, gb, gn, c_trace);
        stamp();
        {
            float rvk[2] = {0, 0}, ovk[2] = {0, 0}; // ffn-V scale / offset of the own K rows (<= 2 per thread)
            const float *rvp = p.rfv + (size_t)lq * 4 * E, *ovp = p.ocfv + (size_t)lq * 4 * E;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int i = ctid + t * kTokConsumers;
                if (i < nk) {
                    rvk[t] = rvp[sl.k0 + i];
                    ovk[t] = ovp[sl.k0 + i];
                }
            }
            rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(0), c_res + (uint32_t)(0) * 8u, E, sl.e1 - sl.e0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
            rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(3 * E), c_res + (uint32_t)(ne * 1) * 8u, E, sl.k1 - sl.k0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
            tok_sync();
            stamp();
            if (mine) {
                const float y = (float)(sm.scal[0] * row_total(0, ctid, 1) + sm.scal[3]);
                sm.srown[ctid] = (float)(1.0 / (1.0 + exp(-(double)y)));
            }
            double mx[3] = {0, 0, 0}, of[3] = {0, 0, 0};
            float *vec = vecp(q);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int i = ctid + t * kTokConsumers;
                if (i < nk) {
                    float a = (float)(sm.scal[1] * row_total(ne * 1, i, 1) + sm.scal[4]);
                    a = a > 0.0f ? a : 0.0f;
                    a = a * a;
                    const float xv = (float)((double)a * (double)rvk[t]);
                    put(vec + sl.k0 + i, xv);
                    mx[0] = fmax(mx[0], (double)xv);
                    of[0] += (double)a * (double)ovk[t];
                }
            }
            publish_vparts<1>(p, phase, mx, of, rd, ctid, nk < kRedMax ? nk : kRedMax);
        }
        stamp();
    grid_sync(p, phase, ctid);
    stamp();
        ++q;
        // ======== ffn V (rows of 4E bytes, four warps per row) + residual (rwkv.cu:574-577) =========
        gather_quantise(sm.planes, sm.scal, vecp(q - 1), acc_buf(p, phase - 1), 1, 4 * E, ctid, gb, gn, c_trace);
        stamp();
        rp = consume_sub<CPL, FULL, 4>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(0), c_res + (uint32_t)(0) * 8u, 4 * E, sl.e1 - sl.e0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
        // issued after the register-hungry core; the loads land during the epilogue + grid barrier
        if (l + 1 < p.L_run) prefetch_att(lq + 1, j, so);
        tok_sync();
        stamp();
        if (mine) {
            const float kv = (float)(sm.scal[0] * row_total(0, ctid, 4 * 1) + sm.scal[3]);
            sm.xown[ctid] = sm.xown[ctid] + (double)(kv * sm.srown[ctid]);
        }
        publish_stats(p, sm, phase, ne, rd, ctid);
        stamp();
    grid_sync(p, phase, ctid);
    stamp();
        ++q;
    }

    // ======== LN_out for the own slice, head GEMV (rwkv.cu:585-589) ================================
    {
        double xmean = 0.0, x2 = 1.0;
        if (warp < nwe) stats_from_acc(p, acc_buf(p, phase - 1), xmean, x2);
        double mx[3] = {0, 0, 0}, of[3] = {0, 0, 0};
        if (mine) {
            const double *lwp = p.ln + (size_t)(4 * p.L + 2) * E;
            const float f = (float)(lwp[j] * ((sm.xown[ctid] - xmean) * (1.0 / x2)) + lwp[E + j]);
            const float xh = (float)((double)f * (double)p.rhead[j]);
            put(vecp(q) + j, xh);
            mx[0] = fabs((double)xh);
            of[0] = (double)f * (double)p.ochead[j];
            p.x[j] = sm.xown[ctid]; // residual stream after the last layer (debug / tests)
        }
        publish_vparts<1>(p, phase, mx, of, rd, ctid, ne);
    }
    stamp();
    grid_sync(p, phase, ctid);
    stamp();
    ++q;
    gather_quantise(sm.planes, sm.scal, vecp(q - 1), acc_buf(p, phase - 1), 1, E, ctid, gb, gn, c_trace);
    stamp();
    rp = consume_sub<CPL, FULL, 1>(c_ring, c_full, c_empty, c_tile, c_stages, c_planes + (uint32_t)(0), c_res + (uint32_t)(0) * 8u, E, sl.v1 - sl.v0, rp, c_warp, c_lane, c_ptrace, c_tcnt);
    tok_sync();
    stamp();
    {
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        for (int i = ctid; i < sl.nv; i += kTokConsumers) {
            const float y = (float)(sm.scal[0] * row_total(0, i, 1) + sm.scal[3]);
            put(p.logits + sl.v0 + i, y);
            if (y > best) { // i ascending per thread: first maximum kept
                best = y;
                bidx = sl.v0 + i;
            }
        }
        if (p.greedy) {
            // block arg-max, first index wins ties
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov2 = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ov2 > best || (ov2 == best && oi < bidx)) {
                    best = ov2;
                    bidx = oi;
                }
            }
            float *bv = reinterpret_cast<float *>(sm.red);
            int *bi = reinterpret_cast<int *>(sm.red + 16);
            tok_sync();
            if (lane == 0) {
                bv[warp] = best;
                bi[warp] = bidx;
            }
            tok_sync();
            if (ctid == 0) {
                for (int w = 1; w < kTokWarps; ++w)
                    if (bv[w] > best || (bv[w] == best && bi[w] < bidx)) {
                        best = bv[w];
                        bidx = bi[w];
                    }
                if (bidx != 0x7fffffff) {
                    // order-preserving image of the float above ~index: max = largest logit, then smallest index
                    unsigned int u = __float_as_uint(best);
                    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                    const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned int)bidx);
                    unsigned long long *acc = acc_buf(p, phase);
                    for (int g = 0; g < p.tp_size; ++g) red_max_u64(peer_ptr(p, acc, g) + kAccArg, key, sys);
                }
            }
            stamp();
    grid_sync(p, phase, ctid);
    stamp();
            if (blockIdx.x == 0 && ctid == 0) {
                const unsigned long long key = __ldcg(acc_buf(p, phase - 1) + kAccArg);
                const unsigned int ix = 0xffffffffu - (unsigned int)(key & 0xffffffffull);
                ctrl->next = (unsigned long long)(key == 0ull ? 0u : ix);
            }
        }
    }
    if (blockIdx.x == 0 && ctid == 0) {
        ctrl->bar_base = phase;
        if (p.feed_mode == 2) ctrl->pos = ctrl->pos + 1;
    }
}

} // namespace rk
