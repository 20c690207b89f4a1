// engine.cu — host side of the B200 RWKV-v4 uint8 decode engine + the C ABI (include/rwkv_b200.h).
//
// Responsibilities:
//   * load a reference-format .bin (include/rwkv/cuda/rwkv.cu:638-717 semantics), stage it
//     through pinned memory, and repack on the device (transpose to [out][in], centre to s8,
//     fold 128*r + o into one offset vector);
//   * keep state, embedding table, weights and logits resident in HBM;
//   * issue one token as 2 + 4*L kernels, normally replayed as a single CUDA graph;
//   * measurement hooks used by bench.py (device-timed decode, per-kernel event profile).
//
// There is deliberately no CPU code path: every entry point that computes fails with an
// error when no sm_100 device is present.
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/rwkv/enums/enum.h"
#include "../../include/rwkv_b200.h"
#include "binfmt.h"
#include "kernels.cuh"
#include "token_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(100 + (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                        __FILE__, __LINE__);                                                       \
    } while (0)

enum KernelClass { K_EMBED = 0, K_ATT_KVR, K_ATT_OUT, K_FFN_RK, K_FFN_V, K_HEAD, K_ARGMAX, K_TOKEN, K_COUNT };
const char *kKernelNames[K_COUNT] = {"embed_ln0", "att_kvr", "att_out", "ffn_rk", "ffn_v", "head", "argmax", "token"};

} // namespace

struct rwkv_b200_model {
    int device = 0;
    int sms = 0;
    int grid = 0;
    int cpl = 0;
    double *d_sample = nullptr, *h_sample = nullptr; // device sampler result {token, margin}
    size_t xch_bytes = 0;   // exchange block (peer-visible with tensor parallelism)
    bool tp_wired = false;  // peers' exchange blocks imported
    std::vector<void *> ipc_opened;
    unsigned long long L = 0, E = 0, max_gpt = 1;
    cudaStream_t stream = nullptr;
    rk::Params p{};
    size_t smem = 0;
    std::vector<void *> allocs;
    void *tensors[RWKV_B200_NUM_TENSORS] = {};
    double *spp = nullptr; // state_pp lives on the device only to honour the tensor table
    rk::Ctrl *h_ctrl = nullptr; // pinned [max_gpt]
    float *h_logits = nullptr;  // pinned [max_gpt][V]
    unsigned long long *h_next = nullptr;
    // graphs
    bool use_graph = true;
    bool token_mode = true; // one persistent cooperative kernel per token (default) vs one kernel per phase
    int max_layers = -1; // debug: run only the first n layers
    cudaGraphExec_t g_plain = nullptr, g_greedy = nullptr, g_free = nullptr, g_stream = nullptr;
    const unsigned long long *g_stream_src = nullptr;
    unsigned long long launches = 0;
    int tp_rank = 0, tp_size = 1;
};

namespace {

using M = rwkv_b200_model;

template <class T> int dmalloc(M *m, T **out, size_t count) {
    void *p = nullptr;
    CK(cudaMalloc(&p, count * sizeof(T) + 256));
    m->allocs.push_back(p);
    *out = reinterpret_cast<T *>(p);
    return 0;
}

int layers_to_run(const M *m) { return m->max_layers >= 0 && m->max_layers < (int)m->L ? m->max_layers : (int)m->L; }

unsigned long long kernels_per_token(const M *m, bool greedy) {
    if (m->token_mode) return 1ull;
    return 2ull + 4ull * layers_to_run(m) + (greedy ? 1 : 0);
}

// ---- kernel dispatch on the model width -------------------------------------------------
template <int CPL> int set_attrs(size_t smem) {
    CK(cudaFuncSetAttribute(rk::k_att_kvr<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_att_out<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_ffn_rk<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_ffn_v<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_head<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}
template <int CPL> int set_attrs_tok(size_t smem) {
    CK(cudaFuncSetAttribute(rk::k_token<CPL, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_token<CPL, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_token<CPL, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(rk::k_token<CPL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}
template <int CPL> const void *token_entry(bool full, bool trace) {
    if (trace) return full ? (const void *)rk::k_token<CPL, true, true> : (const void *)rk::k_token<CPL, false, true>;
    return full ? (const void *)rk::k_token<CPL, true, false> : (const void *)rk::k_token<CPL, false, false>;
}

struct EventPair {
    cudaEvent_t a, b;
};

// Optional per-launch timing (profile mode): records an event pair around each launch.
struct Prof {
    std::vector<std::pair<int, EventPair>> ev;
};

template <int CPL> int launch_class(M *m, int cls, int layer, cudaStream_t s) {
    const dim3 g(m->grid), b(rk::kThreads);
    switch (cls) {
    case K_ATT_KVR: rk::k_att_kvr<CPL><<<g, b, m->smem, s>>>(m->p, layer); break;
    case K_ATT_OUT: rk::k_att_out<CPL><<<g, b, m->smem, s>>>(m->p, layer); break;
    case K_FFN_RK: rk::k_ffn_rk<CPL><<<g, b, m->smem, s>>>(m->p, layer); break;
    case K_FFN_V: rk::k_ffn_v<CPL><<<g, b, m->smem, s>>>(m->p, layer); break;
    case K_HEAD: rk::k_head<CPL><<<g, b, m->smem, s>>>(m->p); break;
    default: return fail(3, "bad kernel class %d", cls);
    }
    return 0;
}

int launch_one(M *m, int cls, int layer, cudaStream_t s, Prof *prof) {
    EventPair ep{};
    if (prof) {
        CK(cudaEventCreate(&ep.a));
        CK(cudaEventCreate(&ep.b));
        CK(cudaEventRecord(ep.a, s));
    }
    int rc = 0;
    if (cls == K_EMBED) rk::k_embed_ln0<<<1, rk::kConsumers, 0, s>>>(m->p);
    else if (cls == K_ARGMAX) rk::k_argmax<<<1, 1024, 0, s>>>(m->p);
    else {
        switch (m->cpl) {
        case 2: rc = launch_class<2>(m, cls, layer, s); break;
        case 4: rc = launch_class<4>(m, cls, layer, s); break;
        case 8: rc = launch_class<8>(m, cls, layer, s); break;
        case 10: rc = launch_class<10>(m, cls, layer, s); break;
        default: rc = fail(3, "unsupported chunks-per-lane %d", m->cpl);
        }
    }
    if (rc) return rc;
    CK(cudaGetLastError());
    if (prof) {
        CK(cudaEventRecord(ep.b, s));
        prof->ev.push_back({cls, ep});
    }
    return 0;
}

// Ring geometry per mode. Token kernel: a tile is eight row segments of n_embed bytes (one per
// consumer warp); staged kernels: 20 KB tiles. As many stages as fit beside the limb planes.
void configure_mode(M *m) {
    rk::Params &p = m->p;
    p.tile_bytes = m->token_mode ? (int)(8 * m->E) : 20480;
    if (p.tile_bytes < (int)(4 * m->E)) p.tile_bytes = (int)(4 * m->E);
    p.stages = (int)std::min<size_t>(rk::kMaxStages, (232448 - rk::smem_bytes(0, 0, p.plane_cap)) / p.tile_bytes);
    m->smem = rk::smem_bytes(p.stages, p.tile_bytes, p.plane_cap);
}

int launch_token(M *m, int feed, bool greedy, const unsigned long long *stream, cudaStream_t s) {
    if (m->tp_size > 1 && !m->tp_wired)
        return fail(7, "tensor parallelism: call rwkv_b200_tp_import with every rank's handle before the first forward");
    rk::Params prm = m->p;
    prm.L_run = layers_to_run(m);
    prm.feed_mode = feed;
    prm.greedy = greedy ? 1 : 0;
    prm.stream = stream;
    void *args[] = {&prm};
    const bool full = m->E == (unsigned long long)m->cpl * 512ull;
    const bool trace = m->p.trace != nullptr;
    const void *fn = m->cpl == 2 ? token_entry<2>(full, trace) : m->cpl == 4 ? token_entry<4>(full, trace)
                   : m->cpl == 8 ? token_entry<8>(full, trace) : token_entry<10>(full, trace);
    CK(cudaLaunchCooperativeKernel(fn, dim3(m->grid), dim3(rk::kTokThreads), args, m->smem, s));
    return 0;
}

// The kernel sequence of one token.
int enqueue_token(M *m, cudaStream_t s, bool greedy, Prof *prof) {
    int rc;
    if ((rc = launch_one(m, K_EMBED, 0, s, prof))) return rc;
    const int nl = layers_to_run(m);
    for (int l = 0; l < nl; ++l) {
        if ((rc = launch_one(m, K_ATT_KVR, l, s, prof))) return rc;
        if ((rc = launch_one(m, K_ATT_OUT, l, s, prof))) return rc;
        if ((rc = launch_one(m, K_FFN_RK, l, s, prof))) return rc;
        if ((rc = launch_one(m, K_FFN_V, l, s, prof))) return rc;
    }
    if ((rc = launch_one(m, K_HEAD, 0, s, prof))) return rc;
    if (greedy && (rc = launch_one(m, K_ARGMAX, 0, s, prof))) return rc;
    return 0;
}

enum GraphKind { G_PLAIN, G_GREEDY, G_FREE, G_STREAM };

int build_graph(M *m, GraphKind kind, const unsigned long long *stream_src, cudaGraphExec_t *out) {
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (kind == G_FREE) rk::k_feed_next<<<1, 32, 0, m->stream>>>(m->p);
    if (kind == G_STREAM) rk::k_feed_stream<<<1, 32, 0, m->stream>>>(m->p, stream_src);
    rc = enqueue_token(m, m->stream, kind == G_GREEDY || kind == G_FREE, nullptr);
    cudaError_t e = cudaStreamEndCapture(m->stream, &g);
    if (rc) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) return fail(100 + (int)e, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(100 + (int)e, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return 0;
}

void drop_graphs(M *m) {
    for (cudaGraphExec_t *g : {&m->g_plain, &m->g_greedy, &m->g_free, &m->g_stream}) {
        if (*g) cudaGraphExecDestroy(*g);
        *g = nullptr;
    }
    m->g_stream_src = nullptr;
}

int run_token(M *m, bool greedy) {
    if (m->token_mode) {
        int rc = launch_token(m, 0, greedy, nullptr, m->stream);
        if (rc) return rc;
    } else if (!m->use_graph) {
        int rc = enqueue_token(m, m->stream, greedy, nullptr);
        if (rc) return rc;
    } else {
        cudaGraphExec_t *g = greedy ? &m->g_greedy : &m->g_plain;
        if (!*g) {
            int rc = build_graph(m, greedy ? G_GREEDY : G_PLAIN, nullptr, g);
            if (rc) return rc;
        }
        CK(cudaGraphLaunch(*g, m->stream));
    }
    m->launches += kernels_per_token(m, greedy);
    return 0;
}

// ---- loader --------------------------------------------------------------------------------
struct FileReader {
    int fd = -1;
    uint8_t *pin = nullptr;
    size_t pin_bytes = 0;
    ~FileReader() {
        if (fd >= 0) close(fd);
        if (pin) cudaFreeHost(pin);
    }
};

int read_exact(int fd, void *dst, size_t n, uint64_t off) {
    uint8_t *d = (uint8_t *)dst;
    while (n) {
        ssize_t got = pread(fd, d, n, (off_t)off);
        if (got < 0) {
            if (errno == EINTR) continue;
            return fail(4, "read error: %s", strerror(errno));
        }
        if (got == 0) return fail(4, "model file truncated");
        d += got;
        off += (uint64_t)got;
        n -= (size_t)got;
    }
    return 0;
}

// file[off, off+n) -> device dst, through the pinned staging buffer.
int upload(M *m, FileReader &fr, uint64_t off, size_t n, void *dst) {
    uint8_t *d = (uint8_t *)dst;
    while (n) {
        const size_t c = n < fr.pin_bytes ? n : fr.pin_bytes;
        int rc = read_exact(fr.fd, fr.pin, c, off);
        if (rc) return rc;
        CK(cudaMemcpyAsync(d, fr.pin, c, cudaMemcpyHostToDevice, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        d += c;
        off += c;
        n -= c;
    }
    return 0;
}

template <class T> int upload_tensor(M *m, FileReader &fr, int tid, T **out) {
    const size_t n = binfmt::elems(tid, m->L, m->E);
    int rc = dmalloc(m, out, n);
    if (rc) return rc;
    return upload(m, fr, binfmt::offset(tid, m->L, m->E), n * sizeof(T), *out);
}

// uint8 matrix family `tid`: `mats` matrices of [rows_in][cols_out] -> int8 [cols_out][rows_in]
int upload_matrix(M *m, FileReader &fr, int tid, size_t mats, size_t rows_in, size_t cols_out, uint8_t *d_raw,
                  int8_t **out) {
    int rc = dmalloc(m, out, mats * rows_in * cols_out);
    if (rc) return rc;
    const uint64_t base = binfmt::offset(tid, m->L, m->E);
    for (size_t i = 0; i < mats; ++i) {
        rc = upload(m, fr, base + i * rows_in * cols_out, rows_in * cols_out, d_raw);
        if (rc) return rc;
        dim3 g((unsigned)((cols_out + 63) / 64), (unsigned)((rows_in + 63) / 64));
        rk::k_transpose_xor<<<g, 256, 0, m->stream>>>(d_raw, cols_out, (int)rows_in, (int)cols_out,
                                                      *out + i * rows_in * cols_out, rows_in, 0);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(m->stream));
    }
    return 0;
}

int centre(M *m, const float *r, const float *o, size_t n, const float **out) {
    float *oc = nullptr;
    int rc = dmalloc(m, &oc, n);
    if (rc) return rc;
    rk::k_centre_offsets<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(r, o, oc, n);
    CK(cudaGetLastError());
    *out = oc;
    return 0;
}

int do_load(M *m, const char *path, int quiet) {
    FileReader fr;
    fr.fd = open(path, O_RDONLY);
    if (fr.fd < 0) return fail(2, "Error opening file %s", path);
    int64_t hdr[2];
    int rc = read_exact(fr.fd, hdr, sizeof(hdr), 0);
    if (rc) return rc;
    m->L = (unsigned long long)hdr[0];
    m->E = (unsigned long long)hdr[1];
    const unsigned long long L = m->L, E = m->E;
    if (!quiet) {
        printf("n_layers: %llu\nn_embed: %llu\n", L, E);
        fflush(stdout);
    }
    if (L == 0 || L > 4096 || E == 0 || E % 16 != 0 || E > 5120)
        return fail(5, "unsupported model shape: n_layers=%llu n_embed=%llu (need n_embed %% 16 == 0, <= 5120)", L, E);
    struct stat st;
    if (fstat(fr.fd, &st) != 0 || (uint64_t)st.st_size < binfmt::file_bytes(L, E))
        return fail(4, "model file too short: %lld bytes, need %llu", (long long)st.st_size,
                    (unsigned long long)binfmt::file_bytes(L, E));

    CK(cudaSetDevice(m->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, m->device));
    if (prop.major < 10) return fail(6, "device %d is sm_%d%d; this engine is built for sm_100a only", m->device, prop.major, prop.minor);
    m->sms = prop.multiProcessorCount;
    m->grid = m->sms;
    CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));

    m->cpl = E <= 1024 ? 2 : E <= 2048 ? 4 : E <= 4096 ? 8 : 10;
    if (const char *e = getenv("RWKV_B200_MODE")) {
        if (std::string(e) == "staged" && m->tp_size == 1) m->token_mode = false;
    }
    rk::Params &p = m->p;
    p.L = (int)L;
    p.E = (int)E;
    p.plane_cap = (int)(12 * E);
    configure_mode(m);
    p.tp_rank = m->tp_rank;
    p.tp_size = m->tp_size;
    if ((4 * E + m->grid - 1) / m->grid + (E + m->grid - 1) / m->grid + 2 > (unsigned long long)rk::kMaxRowsPerCta ||
        (E + m->grid - 1) / m->grid + 1 > (unsigned long long)rk::kMaxSlice ||
        (4 * E + m->grid - 1) / m->grid + 1 > 2ull * rk::kConsumers ||
        (binfmt::kVocab + m->grid - 1) / m->grid + 1 > (unsigned long long)rk::kMaxRowsPerCta ||
        m->grid > rk::kRedMax || (4 * E + m->grid - 1) / m->grid + 1 > (unsigned long long)rk::kRedMax)
        return fail(5, "grid of %d CTAs is too small for n_embed=%llu", m->grid, E);
    switch (m->cpl) {
    case 2: rc = set_attrs<2>(232448), rc = rc ? rc : set_attrs_tok<2>(232448); break;
    case 4: rc = set_attrs<4>(232448), rc = rc ? rc : set_attrs_tok<4>(232448); break;
    case 8: rc = set_attrs<8>(232448), rc = rc ? rc : set_attrs_tok<8>(232448); break;
    default: rc = set_attrs<10>(232448), rc = rc ? rc : set_attrs_tok<10>(232448); break;
    }
    if (rc) return rc;

    fr.pin_bytes = 64u << 20;
    CK(cudaMallocHost((void **)&fr.pin, fr.pin_bytes));

    auto say = [&](int tid) {
        if (!quiet) {
            printf("loading: %s\n", binfmt::name(tid));
            fflush(stdout);
        }
    };
    // The reference prints every tensor in file order (rwkv.cu:679); keep that UX.
    for (int t = 0; t < binfmt::kNumTensors; ++t) say(t);

    // ---- small parameter tensors, reference dtype and shape ---------------------------------
    float *emb, *kr, *vr, *rr, *o1, *o2, *o3, *aor, *aoo, *fkr, *fvr, *frr, *fko, *fvo, *fro, *hr, *ho;
    double *ln, *mixk, *mixv, *mixr, *fmk, *fmr, *decay, *bonus;
#define UP(tid, var)                                                                               \
    if ((rc = upload_tensor(m, fr, tid, &var))) return rc;                                          \
    m->tensors[tid] = var;
    UP(EMBED, emb) UP(LAYERNORMS, ln) UP(MIXK, mixk) UP(MIXV, mixv) UP(MIXR, mixr)
    UP(KR, kr) UP(VR, vr) UP(RR, rr) UP(O1, o1) UP(O2, o2) UP(O3, o3)
    UP(ATTOUTR, aor) UP(ATTOUTO, aoo) UP(FFNMIXK, fmk) UP(FFNMIXV, fmr)
    UP(FFNKR, fkr) UP(FFNVR, fvr) UP(FFNRR, frr) UP(FFNKO, fko) UP(FFNVO, fvo) UP(FFNRO, fro)
    UP(DECAY, decay) UP(BONUS, bonus) UP(HEADR, hr) UP(HEADO, ho)
#undef UP
    p.emb = emb; p.ln = ln; p.mixk = mixk; p.mixv = mixv; p.mixr = mixr; p.fmixk = fmk; p.fmixr = fmr;
    p.decay = decay; p.bonus = bonus;
    p.rk = kr; p.rv = vr; p.rr = rr; p.ro = aor; p.rfk = fkr; p.rfv = fvr; p.rfr = frr; p.rhead = hr;
    {
        double *ed = nullptr;
        if ((rc = dmalloc(m, &ed, (size_t)(L * E)))) return rc;
        rk::k_exp_table<<<(unsigned)((L * E + 255) / 256), 256, 0, m->stream>>>(decay, ed, (size_t)(L * E));
        CK(cudaGetLastError());
        p.expdecay = ed;
    }
    if ((rc = centre(m, kr, o1, L * E, &p.ock))) return rc;
    if ((rc = centre(m, vr, o2, L * E, &p.ocv))) return rc;
    if ((rc = centre(m, rr, o3, L * E, &p.ocr))) return rc;
    if ((rc = centre(m, aor, aoo, L * E, &p.oco))) return rc;
    if ((rc = centre(m, fkr, fko, L * E, &p.ocfk))) return rc;
    if ((rc = centre(m, fvr, fvo, L * 4 * E, &p.ocfv))) return rc;
    if ((rc = centre(m, frr, fro, L * E, &p.ocfr))) return rc;
    if ((rc = centre(m, hr, ho, E, &p.ochead))) return rc;

    // ---- uint8 matrices: stage raw, transpose + centre on the device --------------------------
    uint8_t *d_raw = nullptr;
    const size_t raw_bytes = std::max<size_t>(4 * E * E, binfmt::kVocab * E);
    CK(cudaMalloc((void **)&d_raw, raw_bytes));
    int8_t *wk, *wv, *wr, *wo, *wfk, *wfv, *wfr, *whead;
    rc = upload_matrix(m, fr, KM, L, E, E, d_raw, &wk);
    if (!rc) rc = upload_matrix(m, fr, VM, L, E, E, d_raw, &wv);
    if (!rc) rc = upload_matrix(m, fr, RM, L, E, E, d_raw, &wr);
    if (!rc) rc = upload_matrix(m, fr, ATTOUT, L, E, E, d_raw, &wo);
    if (!rc) rc = upload_matrix(m, fr, FFNK, L, E, 4 * E, d_raw, &wfk);
    if (!rc) rc = upload_matrix(m, fr, FFNV, L, 4 * E, E, d_raw, &wfv);
    if (!rc) rc = upload_matrix(m, fr, FFNR, L, E, E, d_raw, &wfr);
    if (!rc) rc = upload_matrix(m, fr, HEAD, 1, E, binfmt::kVocab, d_raw, &whead);
    cudaFree(d_raw);
    if (rc) return rc;
    p.wk = wk; p.wv = wv; p.wr = wr; p.wo = wo; p.wfk = wfk; p.wfv = wfv; p.wfr = wfr; p.whead = whead;
    m->tensors[KM] = wk; m->tensors[VM] = wv; m->tensors[RM] = wr; m->tensors[ATTOUT] = wo;
    m->tensors[FFNK] = wfk; m->tensors[FFNV] = wfv; m->tensors[FFNR] = wfr; m->tensors[HEAD] = whead;

    // ---- state, activations, control ------------------------------------------------------------
    const size_t sn = (size_t)(L * E * m->max_gpt);
    if ((rc = dmalloc(m, &p.sxy, sn)) || (rc = dmalloc(m, &p.saa, sn)) || (rc = dmalloc(m, &p.sbb, sn)) ||
        (rc = dmalloc(m, &p.sdd, sn)) || (rc = dmalloc(m, &m->spp, sn)))
        return rc;
    for (double *s : {p.sxy, p.saa, p.sbb, p.sdd, m->spp}) CK(cudaMemsetAsync(s, 0, sn * sizeof(double), m->stream));
    double *b1, *fkb, *fvb;
    float *b3;
    if ((rc = dmalloc(m, &p.x, E)) || (rc = dmalloc(m, &p.xy_new, E)) || (rc = dmalloc(m, &p.dd_new, E)) ||
        (rc = dmalloc(m, &p.xs_o, E)) || (rc = dmalloc(m, &p.sr, E)) || (rc = dmalloc(m, &p.xs_v, 4 * E)) ||
        (rc = dmalloc(m, &p.part_o, 2 * rk::kMaxGrid)) ||
        (rc = dmalloc(m, &p.part_v, 2 * rk::kMaxGrid)) || (rc = dmalloc(m, &p.ctrl, 1)) ||
        (rc = dmalloc(m, &b1, E)) || (rc = dmalloc(m, &fkb, E)) || (rc = dmalloc(m, &fvb, E)) ||
        (rc = dmalloc(m, &b3, E)))
        return rc;
    CK(cudaMemsetAsync(p.ctrl, 0, sizeof(rk::Ctrl), m->stream));
    // exchange block: [0,64) barrier counter, [64,128) rank-local arrival counter | [128,512) accumulators | [1024, +32E) vec | logits
    {
        unsigned char *x = nullptr;
        m->xch_bytes = (1024 + 32 * (size_t)E + 4 * (size_t)binfmt::kVocab + 255) & ~(size_t)255;
        if ((rc = dmalloc(m, &x, m->xch_bytes))) return rc;
        CK(cudaMemsetAsync(x, 0, m->xch_bytes, m->stream));
        for (int g = 0; g < 8; ++g) p.xch[g] = x; // peers are wired by rwkv_b200_tp_import
        p.gbar = reinterpret_cast<unsigned int *>(x);
        p.lbar = reinterpret_cast<unsigned int *>(x + 64);
        p.acc = reinterpret_cast<unsigned long long *>(x + 128);
        p.vec = reinterpret_cast<float *>(x + 1024);
        p.logits = reinterpret_cast<float *>(x + 1024 + 32 * (size_t)E);
    }
    {
        int coop = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, m->device));
        if (!coop && m->token_mode) return fail(6, "device does not support cooperative launch");
    }
    CK(cudaMemsetAsync(p.x, 0, E * sizeof(double), m->stream));
    m->tensors[X] = p.x;
    m->tensors[STATEXY] = p.sxy; m->tensors[STATEAA] = p.saa; m->tensors[STATEBB] = p.sbb;
    m->tensors[STATEPP] = m->spp; m->tensors[STATEDD] = p.sdd;
    m->tensors[BUFFER1] = b1; m->tensors[BUFFER2] = p.logits; m->tensors[BUFFER3] = b3; m->tensors[BUFFER4] = p.sr;
    m->tensors[FFNKBUFFER] = fkb; m->tensors[FFNVBUFFER] = fvb; m->tensors[FFNRBUFFER] = p.xs_v;

    CK(cudaMallocHost((void **)&m->h_ctrl, sizeof(rk::Ctrl) * m->max_gpt));
    CK(cudaMallocHost((void **)&m->h_logits, sizeof(float) * binfmt::kVocab * m->max_gpt));
    CK(cudaMallocHost((void **)&m->h_next, sizeof(unsigned long long)));
    memset(m->h_logits, 0, sizeof(float) * binfmt::kVocab * m->max_gpt);
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

int check_model(const M *m) {
    if (!m) return fail(1, "null model handle");
    return 0;
}

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char *rwkv_b200_last_error(void) { return g_err.c_str(); }
int rwkv_b200_abi_version(void) { return 1; }

int rwkv_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rwkv_b200_load_tp(const char *path, unsigned long long max_gpt, int device, int quiet, int tp_rank, int tp_size,
                      rwkv_b200_model **out, unsigned long long *n_layers, unsigned long long *n_embed) {
    if (!path || !out) return fail(1, "null argument");
    *out = nullptr;
    if (tp_size < 1 || tp_size > 8 || tp_rank < 0 || tp_rank >= tp_size)
        return fail(7, "tensor parallelism: rank %d of %d is not supported (1..8 ranks)", tp_rank, tp_size);
    if (rwkv_b200_device_count() <= device || device < 0)
        return fail(6, "CUDA device %d not available (no CPU fallback exists)", device);
    M *m = new M;
    m->device = device;
    m->max_gpt = max_gpt ? max_gpt : 1;
    m->tp_rank = tp_rank;
    m->tp_size = tp_size;
    int rc = do_load(m, path, quiet);
    if (rc) {
        std::string keep = g_err;
        rwkv_b200_free(m);
        g_err = keep;
        return rc;
    }
    *out = m;
    if (n_layers) *n_layers = m->L;
    if (n_embed) *n_embed = m->E;
    return 0;
}

int rwkv_b200_load(const char *path, unsigned long long max_gpt, int device, int quiet, rwkv_b200_model **out,
                   unsigned long long *n_layers, unsigned long long *n_embed) {
    return rwkv_b200_load_tp(path, max_gpt, device, quiet, 0, 1, out, n_layers, n_embed);
}

void rwkv_b200_free(rwkv_b200_model *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    drop_graphs(m);
    for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : m->allocs) cudaFree(p);
    if (m->h_ctrl) cudaFreeHost(m->h_ctrl);
    if (m->h_logits) cudaFreeHost(m->h_logits);
    if (m->h_next) cudaFreeHost(m->h_next);
    if (m->h_sample) cudaFreeHost(m->h_sample);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

void *rwkv_b200_tensor(rwkv_b200_model *m, int index) {
    if (!m || index < 0 || index >= RWKV_B200_NUM_TENSORS) return nullptr;
    return m->tensors[index];
}
unsigned long long rwkv_b200_n_layers(const rwkv_b200_model *m) { return m ? m->L : 0; }
unsigned long long rwkv_b200_n_embed(const rwkv_b200_model *m) { return m ? m->E : 0; }
unsigned long long rwkv_b200_max_gpt(const rwkv_b200_model *m) { return m ? m->max_gpt : 0; }

void *rwkv_b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 1;
    if (cudaMallocHost(&p, bytes) == cudaSuccess) return p;
    cudaGetLastError();
    // No driver (tokenizer-only use): tag the block so host_free knows it came from malloc.
    uint64_t *raw = (uint64_t *)malloc(bytes + 16);
    if (!raw) return nullptr;
    raw[0] = 0x6d616c6c6f636564ULL;
    return raw + 2;
}
void rwkv_b200_host_free(void *p) {
    if (!p) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost) {
        cudaFreeHost(p);
        return;
    }
    cudaGetLastError();
    uint64_t *raw = (uint64_t *)p - 2;
    if (raw[0] == 0x6d616c6c6f636564ULL) free(raw);
}

int rwkv_b200_state_upload(rwkv_b200_model *m, const double *xy, const double *aa, const double *bb,
                           const double *pp, const double *dd, unsigned long long slots) {
    int rc = check_model(m);
    if (rc) return rc;
    if (slots > m->max_gpt) return fail(1, "state_upload: %llu slots > max_gpt %llu", slots, m->max_gpt);
    CK(cudaSetDevice(m->device));
    const size_t n = (size_t)(m->L * m->E * slots) * sizeof(double);
    const double *src[5] = {xy, aa, bb, pp, dd};
    double *dst[5] = {m->p.sxy, m->p.saa, m->p.sbb, m->spp, m->p.sdd};
    for (int i = 0; i < 5; ++i)
        if (src[i]) CK(cudaMemcpyAsync(dst[i], src[i], n, cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

int rwkv_b200_state_download(rwkv_b200_model *m, double *xy, double *aa, double *bb, double *pp, double *dd,
                             unsigned long long slots) {
    int rc = check_model(m);
    if (rc) return rc;
    if (slots > m->max_gpt) return fail(1, "state_download: %llu slots > max_gpt %llu", slots, m->max_gpt);
    CK(cudaSetDevice(m->device));
    const size_t n = (size_t)(m->L * m->E * slots) * sizeof(double);
    double *dst[5] = {xy, aa, bb, pp, dd};
    const double *src[5] = {m->p.sxy, m->p.saa, m->p.sbb, m->spp, m->p.sdd};
    for (int i = 0; i < 5; ++i)
        if (dst[i]) CK(cudaMemcpyAsync(dst[i], src[i], n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

int rwkv_b200_state_zero(rwkv_b200_model *m) {
    int rc = check_model(m);
    if (rc) return rc;
    CK(cudaSetDevice(m->device));
    const size_t n = (size_t)(m->L * m->E * m->max_gpt) * sizeof(double);
    for (double *s : {m->p.sxy, m->p.saa, m->p.sbb, m->p.sdd, m->spp}) CK(cudaMemsetAsync(s, 0, n, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

int rwkv_b200_forward(rwkv_b200_model *m, const unsigned long long *tokens, unsigned long long n_tokens, int mode,
                      float *logits_out) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!tokens || n_tokens == 0) return fail(1, "forward: no tokens");
    if (n_tokens > m->max_gpt) return fail(1, "Context too large, max context is %llu", m->max_gpt);
    CK(cudaSetDevice(m->device));
    const size_t V = binfmt::kVocab;
    for (unsigned long long t = 0; t < n_tokens; ++t) {
        if (tokens[t] >= V) return fail(1, "token id %llu out of range", tokens[t]);
        rk::Ctrl &c = m->h_ctrl[t];
        c.token = tokens[t];
        c.next = 0;
        c.slot = (mode == RWKV_B200_MODE_PARRALEL) ? t : 0;
        c.pos = 0;
        CK(cudaMemcpyAsync(m->p.ctrl, &c, 32, cudaMemcpyHostToDevice, m->stream));
        if ((rc = run_token(m, false))) return rc;
        if (logits_out)
            CK(cudaMemcpyAsync(m->h_logits + t * V, m->p.logits, V * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    }
    CK(cudaStreamSynchronize(m->stream));
    if (logits_out && logits_out != m->h_logits) memcpy(logits_out, m->h_logits, n_tokens * V * sizeof(float));
    return 0;
}

int rwkv_b200_forward_greedy(rwkv_b200_model *m, unsigned long long token, unsigned long long *next, float *logits_out) {
    int rc = check_model(m);
    if (rc) return rc;
    const size_t V = binfmt::kVocab;
    if (token >= V) return fail(1, "token id %llu out of range", token);
    CK(cudaSetDevice(m->device));
    rk::Ctrl &c = m->h_ctrl[0];
    c.token = token;
    c.next = 0;
    c.slot = 0;
    c.pos = 0;
    CK(cudaMemcpyAsync(m->p.ctrl, &c, 32, cudaMemcpyHostToDevice, m->stream));
    if ((rc = run_token(m, true))) return rc;
    CK(cudaMemcpyAsync(m->h_next, &m->p.ctrl->next, sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
    if (logits_out) CK(cudaMemcpyAsync(m->h_logits, m->p.logits, V * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (next) *next = *m->h_next;
    if (logits_out && logits_out != m->h_logits) memcpy(logits_out, m->h_logits, V * sizeof(float));
    return 0;
}

float *rwkv_b200_logits_host(rwkv_b200_model *m) { return m ? m->h_logits : nullptr; }

int rwkv_b200_sample_typical(rwkv_b200_model *m, float temp, double u, unsigned long long *token, double *margin) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!token) return fail(1, "null argument");
    CK(cudaSetDevice(m->device));
    if (!m->d_sample) {
        if ((rc = dmalloc(m, &m->d_sample, 2))) return rc;
        CK(cudaMallocHost((void **)&m->h_sample, 2 * sizeof(double)));
    }
    // the reference applies the temperature as probs ^ uint8(1 / temp) (include/rwkv/sampler/typical.h)
    const int exponent = temp != 1.0f ? (int)(unsigned char)(1.0 / (double)temp) : 1;
    rk::k_sample_typical<<<1, rk::kSampleThreads, 0, m->stream>>>(m->p.logits, (int)binfmt::kVocab, exponent, u, m->d_sample);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(m->h_sample, m->d_sample, 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    *token = (unsigned long long)m->h_sample[0];
    if (margin) *margin = m->h_sample[1];
    m->launches += 1;
    return 0;
}

int rwkv_b200_decode_timed(rwkv_b200_model *m, const unsigned long long *tokens, unsigned long long n,
                           int teacher_forced, float *ms) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!tokens || n == 0 || !ms) return fail(1, "decode_timed: bad arguments");
    CK(cudaSetDevice(m->device));
    unsigned long long *d_tok = nullptr;
    const unsigned long long cnt = teacher_forced ? n : 1;
    for (unsigned long long i = 0; i < cnt; ++i)
        if (tokens[i] >= binfmt::kVocab) return fail(1, "token id %llu out of range", tokens[i]);
    CK(cudaMalloc((void **)&d_tok, cnt * sizeof(unsigned long long)));
    CK(cudaMemcpy(d_tok, tokens, cnt * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    rk::Ctrl c{tokens[0], tokens[0], 0, 0, 0, {0, 0, 0}};
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaMemcpy(m->p.ctrl, &c, 32, cudaMemcpyHostToDevice));
    cudaGraphExec_t g = nullptr;
    if (!m->token_mode) {
        rc = build_graph(m, teacher_forced ? G_STREAM : G_FREE, d_tok, &g);
        if (rc) {
            cudaFree(d_tok);
            return rc;
        }
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaStreamSynchronize(m->stream);
    cudaEventRecord(a, m->stream);
    cudaError_t e = cudaSuccess;
    for (unsigned long long i = 0; i < n && e == cudaSuccess; ++i) {
        if (m->token_mode) {
            if (launch_token(m, teacher_forced ? 2 : 1, !teacher_forced, d_tok, m->stream)) e = cudaErrorLaunchFailure;
        } else {
            e = cudaGraphLaunch(g, m->stream);
        }
    }
    cudaEventRecord(b, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e == cudaSuccess) cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (g) cudaGraphExecDestroy(g);
    cudaFree(d_tok);
    if (e != cudaSuccess) return fail(100 + (int)e, "decode_timed failed: %s", cudaGetErrorString(e));
    m->launches += n * (kernels_per_token(m, !teacher_forced) + (m->token_mode ? 0 : 1));
    return 0;
}

int rwkv_b200_kernel_count(void) { return K_COUNT; }
const char *rwkv_b200_kernel_name(int k) { return (k >= 0 && k < K_COUNT) ? kKernelNames[k] : ""; }

int rwkv_b200_profile(rwkv_b200_model *m, const unsigned long long *tokens, unsigned long long n, float *ms_sum,
                      unsigned long long *launches, double *bytes) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!tokens || !ms_sum || !launches || !bytes) return fail(1, "profile: bad arguments");
    CK(cudaSetDevice(m->device));
    for (int k = 0; k < K_COUNT; ++k) {
        ms_sum[k] = 0.f;
        launches[k] = 0;
    }
    const double E = (double)m->E, V = (double)binfmt::kVocab;
    // algorithmic HBM bytes of one launch: weight bytes + the vectors the phase must touch
    bytes[K_EMBED] = 4 * E + 16 * E + 8 * E;
    bytes[K_ATT_KVR] = 3 * E * E + E * (8 + 16 + 8 + 24 + 12 + 12) + E * (8 * 4 + 8 + 4);
    bytes[K_ATT_OUT] = E * E + E * (4 + 4 + 16 + 16);
    bytes[K_FFN_RK] = 5 * E * E + E * (8 + 16 + 8 + 16 + 8 + 8) + E * 8 + E * 4 + 4 * E * (4 + 4 + 4);
    bytes[K_FFN_V] = 4 * E * E + 4 * E * 4 + E * (4 + 16 + 16);
    bytes[K_HEAD] = V * E + E * (8 + 16 + 8) + 4 * V;
    bytes[K_ARGMAX] = 4 * V;
    bytes[K_TOKEN] = (double)binfmt::algorithmic_bytes_per_token(m->L, m->E);
    for (unsigned long long t = 0; t < n; ++t) {
        if (tokens[t] >= binfmt::kVocab) return fail(1, "token id out of range");
        rk::Ctrl c{tokens[t], 0, 0, 0, 0, {0, 0, 0}};
        if (m->token_mode) {
            CK(cudaMemcpyAsync(m->p.ctrl, &c, 32, cudaMemcpyHostToDevice, m->stream));
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            CK(cudaStreamSynchronize(m->stream));
            CK(cudaEventRecord(a, m->stream));
            if ((rc = launch_token(m, 0, true, nullptr, m->stream))) return rc;
            CK(cudaEventRecord(b, m->stream));
            CK(cudaStreamSynchronize(m->stream));
            float t_ms = 0.f;
            cudaEventElapsedTime(&t_ms, a, b);
            cudaEventDestroy(a);
            cudaEventDestroy(b);
            ms_sum[K_TOKEN] += t_ms;
            launches[K_TOKEN] += 1;
            m->launches += 1;
            continue;
        }
        CK(cudaMemcpyAsync(m->p.ctrl, &c, 32, cudaMemcpyHostToDevice, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        Prof prof;
        rc = enqueue_token(m, m->stream, true, &prof);
        if (rc) return rc;
        CK(cudaStreamSynchronize(m->stream));
        for (auto &pe : prof.ev) {
            float t_ms = 0.f;
            cudaEventElapsedTime(&t_ms, pe.second.a, pe.second.b);
            ms_sum[pe.first] += t_ms;
            launches[pe.first] += 1;
            cudaEventDestroy(pe.second.a);
            cudaEventDestroy(pe.second.b);
        }
        m->launches += kernels_per_token(m, true);
    }
    return 0;
}

unsigned long long rwkv_b200_launch_count(const rwkv_b200_model *m) { return m ? m->launches : 0; }

int rwkv_b200_set_option(rwkv_b200_model *m, const char *key, const char *value) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!key || !value) return fail(1, "set_option: null");
    const std::string k = key;
    const int v = atoi(value);
    if (k == "graph") m->use_graph = v != 0;
    else if (k == "trace") {
        if (v && !m->p.trace) {
            unsigned long long *t = nullptr;
            if (dmalloc(m, &t, (size_t)rk::kMaxGrid * rk::kTraceMax)) return fail(1, "trace alloc failed");
            cudaMemset(t, 0, (size_t)rk::kMaxGrid * rk::kTraceMax * 8);
            m->p.trace = t;
            unsigned long long *pt = nullptr;
            if (dmalloc(m, &pt, (size_t)3 * rk::kRedMax * rk::kTileTraceMax)) return fail(1, "trace alloc failed");
            cudaMemset(pt, 0, (size_t)3 * rk::kRedMax * rk::kTileTraceMax * 8);
            m->p.ptrace = pt;
        } else if (!v) {
            m->p.trace = nullptr;
            m->p.ptrace = nullptr;
        }
    } else if (k == "mode") {
        const bool want_token = std::string(value) == "token";
        if (!want_token && std::string(value) != "staged") return fail(1, "mode must be 'token' or 'staged'");
        if (!want_token && m->tp_size > 1) return fail(1, "the staged kernels are single-GPU only");
        m->token_mode = want_token;
        configure_mode(m);
    }
    else if (k == "max_layers") m->max_layers = v;
    else if (k == "issue_gap") {
        if (v < 0 || v > 100000) return fail(1, "issue_gap is a cycle count in 0..100000");
        m->p.issue_gap = v;
    }
    else if (k == "stages") {
        if (v < 2 || v > rk::kMaxStages) return fail(1, "stages must be 2..%d", rk::kMaxStages);
        const size_t smem = rk::smem_bytes(v, m->p.tile_bytes, m->p.plane_cap);
        if (smem > 232448) return fail(1, "stages=%d needs %zu bytes of shared memory", v, smem);
        m->p.stages = v;
        m->smem = smem;
    } else if (k == "tile_bytes") {
        if (m->token_mode) return fail(1, "the token kernel's tile is fixed at 8*n_embed bytes");
        if (v < (int)(4 * m->E) || v % 16) return fail(1, "tile_bytes must be a multiple of 16 and >= 4*n_embed");
        int st = m->p.stages;
        while (st > 2 && rk::smem_bytes(st, v, m->p.plane_cap) > 232448) --st;
        const size_t smem = rk::smem_bytes(st, v, m->p.plane_cap);
        if (smem > 232448) return fail(1, "tile_bytes=%d needs %zu bytes of shared memory", v, smem);
        m->p.tile_bytes = v;
        m->p.stages = st;
        m->smem = smem;
    } else if (k == "grid") {
        if (v < 1 || v > rk::kMaxGrid) return fail(1, "grid out of range");
        m->grid = v;
    } else return fail(1, "unknown option '%s'", key);
    drop_graphs(m);
    return 0;
}

// Debug/test hook: copy a named device vector to the host. Returns the element count.
long long rwkv_b200_debug_read(rwkv_b200_model *m, const char *name, void *dst, size_t dst_bytes) {
    if (check_model(m) || !name || !dst) return -1;
    cudaSetDevice(m->device);
    const std::string k = name;
    const void *src = nullptr;
    size_t bytes = 0, count = 0;
    const size_t E = m->E;
    if (k == "x") src = m->p.x, count = E, bytes = E * 8;
    else if (k == "xy_new") src = m->p.xy_new, count = E, bytes = E * 8;
    else if (k == "dd_new") src = m->p.dd_new, count = E, bytes = E * 8;
    else if (k == "xs_o") src = m->p.xs_o, count = E, bytes = E * 4;
    else if (k == "sr") src = m->p.sr, count = E, bytes = E * 4;
    else if (k == "xs_v") src = m->p.xs_v, count = 4 * E, bytes = 16 * E;
    else if (k == "logits") src = m->p.logits, count = binfmt::kVocab, bytes = 4 * binfmt::kVocab;
    else if (k == "trace" && m->p.trace) src = m->p.trace, count = (size_t)m->grid * rk::kTraceMax, bytes = count * 8;
    else if (k == "ptrace" && m->p.ptrace) src = m->p.ptrace, count = (size_t)3 * m->grid * rk::kTileTraceMax, bytes = count * 8;
    else return -1;
    if (dst_bytes < bytes) return -1;
    if (cudaStreamSynchronize(m->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)count;
}

size_t rwkv_b200_tp_buffer_bytes(const rwkv_b200_model *m) { return m ? m->xch_bytes : 0; }

int rwkv_b200_tp_export(rwkv_b200_model *m, void *ipc_handle_64) {
    if (check_model(m) || !ipc_handle_64) return fail(1, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaSetDevice(m->device);
    CK(cudaStreamSynchronize(m->stream)); // the block is zero-filled before anybody maps it
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, m->p.xch[m->tp_rank]));
    memcpy(ipc_handle_64, &h, 64);
    return 0;
}

int rwkv_b200_tp_import(rwkv_b200_model *m, const void *ipc_handles) {
    if (check_model(m) || !ipc_handles) return fail(1, "null argument");
    if (m->tp_wired) return fail(7, "peer exchange blocks already imported");
    cudaSetDevice(m->device);
    const unsigned char *hs = static_cast<const unsigned char *>(ipc_handles);
    for (int g = 0; g < m->tp_size; ++g) {
        if (g == m->tp_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + 64 * (size_t)g, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        m->ipc_opened.push_back(ptr);
        m->p.xch[g] = static_cast<unsigned char *>(ptr);
    }
    m->tp_wired = true;
    drop_graphs(m);
    return 0;
}

} // extern "C"
