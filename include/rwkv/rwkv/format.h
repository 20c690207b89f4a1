// format.h — the reference ".bin" model layout as data: the single source of truth for the
// host API (rwkv.h: names/types/getSize), the engine loader, the synthetic-model generator,
// the converter and the CPU oracle.
//
// Follows: file header + tensor order   converter/cpp_save_tensor.cpp:75-93
//          element counts               include/rwkv/rwkv/rwkv.h:124-128 (getSize)
//          element sizes                include/rwkv/rwkv/rwkv.h:84      (types)
//          names                        include/rwkv/rwkv/rwkv.h:10-56
// Written as a shape table (dtype + three multipliers) instead of literal arrays.
#pragma once
#include <cstdint>
#include <cstddef>

namespace rwkv_format {

constexpr uint64_t kVocab = 50277;
constexpr int kNumTensors = 46;

enum DType : uint8_t { F64 = 8, F32 = 4, U8 = 1 };

// elems = vmul * (per-layer ? L : 1) * E^epow * emul   (+ special cases below)
struct Spec {
    const char *name;
    DType dtype;
    uint8_t per_layer; // multiply by L
    uint8_t epow;      // power of E (0,1,2)
    uint8_t emul;      // constant multiplier (1 or 4)
    uint8_t kind;      // 0 regular, 1 = V*E (embed/head), 2 = 4(L+1)*E (layernorms), 3 = V (buffer2)
};

constexpr Spec kSpecs[kNumTensors] = {
    {"xbuf", F64, 0, 1, 1, 0},         {"embed", F32, 0, 1, 1, 1},
    {"layernorms", F64, 0, 1, 1, 2},   {"state_xy", F64, 1, 1, 1, 0},
    {"state_aa", F64, 1, 1, 1, 0},     {"state_bb", F64, 1, 1, 1, 0},
    {"state_pp", F64, 1, 1, 1, 0},     {"state_dd", F64, 1, 1, 1, 0},
    {"buffer1", F64, 0, 1, 1, 0},      {"buffer2", F32, 0, 0, 1, 3},
    {"buffer3", F32, 0, 1, 1, 0},      {"buffer4", F32, 0, 1, 1, 0},
    {"mix_k", F64, 1, 1, 1, 0},        {"mix_v", F64, 1, 1, 1, 0},
    {"mix_r", F64, 1, 1, 1, 0},        {"km", U8, 1, 2, 1, 0},
    {"vm", U8, 1, 2, 1, 0},            {"rm", U8, 1, 2, 1, 0},
    {"kr", F32, 1, 1, 1, 0},           {"vr", F32, 1, 1, 1, 0},
    {"rr", F32, 1, 1, 1, 0},           {"o1", F32, 1, 1, 1, 0},
    {"o2", F32, 1, 1, 1, 0},           {"o3", F32, 1, 1, 1, 0},
    {"att_out", U8, 1, 2, 1, 0},       {"att_out_r", F32, 1, 1, 1, 0},
    {"att_out_o", F32, 1, 1, 1, 0},    {"ffn_mix_k", F64, 1, 1, 1, 0},
    {"ffn_mix_v", F64, 1, 1, 1, 0},    {"ffn_k", U8, 1, 2, 4, 0},
    {"ffn_v", U8, 1, 2, 4, 0},         {"ffn_r", U8, 1, 2, 1, 0},
    {"ffn_kr", F32, 1, 1, 1, 0},       {"ffn_vr", F32, 1, 1, 4, 0},
    {"ffn_rr", F32, 1, 1, 1, 0},       {"ffn_ko", F32, 1, 1, 1, 0},
    {"ffn_vo", F32, 1, 1, 4, 0},       {"ffn_ro", F32, 1, 1, 1, 0},
    {"ffn_k_buffer", F64, 0, 1, 1, 0}, {"ffn_v_buffer", F64, 0, 1, 1, 0},
    {"ffn_r_buffer", F32, 0, 1, 4, 0}, {"decay", F64, 1, 1, 1, 0},
    {"bonus", F64, 1, 1, 1, 0},        {"head", U8, 0, 1, 1, 1},
    {"head_r", F32, 0, 1, 1, 0},       {"head_o", F32, 0, 1, 1, 0},
};

inline uint64_t elems(int i, uint64_t L, uint64_t E) {
    const Spec &s = kSpecs[i];
    switch (s.kind) {
    case 1: return kVocab * E;
    case 2: return 4 * (L + 1) * E;
    case 3: return kVocab;
    default: break;
    }
    uint64_t n = s.emul;
    if (s.per_layer) n *= L;
    for (int p = 0; p < s.epow; ++p) n *= E;
    return n;
}
inline uint64_t elsize(int i) { return (uint64_t)kSpecs[i].dtype; }
inline uint64_t bytes(int i, uint64_t L, uint64_t E) { return elems(i, L, E) * elsize(i); }
inline const char *name(int i) { return kSpecs[i].name; }

constexpr uint64_t kHeaderBytes = 16; // two little-endian int64: n_layers, n_embed

// Byte offset of tensor i inside the file.
inline uint64_t offset(int i, uint64_t L, uint64_t E) {
    uint64_t off = kHeaderBytes;
    for (int t = 0; t < i; ++t) off += bytes(t, L, E);
    return off;
}
inline uint64_t file_bytes(uint64_t L, uint64_t E) { return offset(kNumTensors, L, E); }

// uint8 GEMV weight bytes touched once per decoded token: 13*L*E^2 + V*E.
inline uint64_t weight_bytes_per_token(uint64_t L, uint64_t E) {
    return 13 * L * E * E + kVocab * E;
}
// Full algorithmic HBM bytes per token (BASELINE.md section 2 "small terms").
inline uint64_t algorithmic_bytes_per_token(uint64_t L, uint64_t E) {
    return weight_bytes_per_token(L, E) + 4 * (20 * L * E + 2 * E) +
           8 * (7 * L * E + 4 * (L + 1) * E) + 2 * 8 * 4 * L * E + 4 * E + 4 * kVocab;
}

} // namespace rwkv_format
