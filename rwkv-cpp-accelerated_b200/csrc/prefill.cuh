// prefill.cuh — batched forward over T tokens. Placeholder until the int8 tensor-core
// GEMM path lands: reports "not enabled" so that rwkv_b200_forward runs token by token.
#pragma once
#include "common.cuh"

namespace rk {

constexpr int kPrefillMinTokens = 1 << 30;
struct PrefillState {
    bool disabled = true;
};
inline bool prefill_enabled(const PrefillState &) { return false; }
inline int prefill_forward(PrefillState &, const Params &, cudaStream_t, const unsigned long long *, int, bool, float *) { return 3; }
inline const char *prefill_error() { return "batched prefill is not built"; }
inline unsigned long long prefill_launches(const PrefillState &) { return 0; }
inline void prefill_free(PrefillState &) {}

} // namespace rk
