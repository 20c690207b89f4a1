// host_api_test.cpp — exercises the host classes of include/rwkv.h that need no GPU:
// error messages, RWKVState value semantics, tensor table. Prints "ok <name>" lines.
#include <cmath>
#include <cstdio>
#include <string>
#include "rwkv.h"

#define CHECK(cond, name)                                 \
    do {                                                  \
        if (!(cond)) {                                    \
            printf("FAIL %s (line %d)\n", name, __LINE__); \
            return 1;                                     \
        }                                                 \
        printf("ok %s\n", name);                          \
    } while (0)

int main() {
    // tensor table (R.h:10-56, 84, 124-138)
    CHECK(std::string(getName(0)) == "xbuf" && std::string(getName(45)) == "head_o" && names[24] == "att_out", "names");
    CHECK(Mtypes(X) == 8 && Mtypes(EMBED) == 4 && Mtypes(KM) == 1 && types[HEAD] == 1 && types[FFNRBUFFER] == 4, "types");
    CHECK(getSize(EMBED, 12, 768) == 50277ull * 768 && getSize(LAYERNORMS, 12, 768) == 4 * 13 * 768 &&
              getSize(FFNK, 12, 768) == 12ull * 768 * 768 * 4 && getSize(BUFFER2, 12, 768) == 50277 &&
              getSize(FFNVR, 2, 64) == 2 * 64 * 4,
          "getSize");
    unsigned long long total = 16;
    for (int i = 0; i < 46; ++i) total += getSize(i, 32, 4096) * Mtypes(i);
    CHECK(total == 8036852132ull, "file_bytes_7b");

    // errors (R.h:285,344,349)
    RWKV net;
    try {
        net.forward(1);
        CHECK(false, "not_loaded_throws");
    } catch (const std::runtime_error &e) {
        CHECK(std::string(e.what()) == "RWKV not loaded", "not_loaded_message");
    }

    // RWKVState: zero init, deep copy, substate bounds, slot stride
    RWKVState s(2, 8, 3);
    bool zero = true;
    for (int i = 0; i < 2 * 8 * 3; ++i) zero = zero && s.statexy[i] == 0 && s.statepp[i] == 0;
    CHECK(zero && s.num_layers == 2 && s.num_embed == 8 && s.stateSize == 3, "state_zero_init");
    for (int i = 0; i < 2 * 8 * 3; ++i) {
        s.stateaa[i] = i;
        s.statedd[i] = -i;
    }
    RWKVState c(s);
    c.stateaa[0] = 99;
    CHECK(s.stateaa[0] == 0 && c.stateaa[5] == 5 && c.stateSize == 3, "state_deep_copy");
    RWKVState sub = s.getSubState(2);
    CHECK(sub.stateSize == 1 && sub.stateaa[0] == 32 && sub.statedd[15] == -47, "substate_slot_stride");
    try {
        s.getSubState(3);
        CHECK(false, "substate_bounds_throws");
    } catch (const std::runtime_error &e) {
        CHECK(std::string(e.what()) == "State get offset out of bounds, max offset is 3", "substate_bounds_message");
    }
    RWKVState one(2, 8, 1);
    for (int i = 0; i < 16; ++i) one.statebb[i] = 1000 + i;
    s.setSubState(one, 1);
    CHECK(s.statebb[16] == 1000 && s.statebb[31] == 1015 && s.statebb[15] == 0 && s.statebb[32] == 0 && s.hostAhead, "set_substate");
    RWKVState e = net.emptyState();
    CHECK(e.stateSize == 1, "empty_state");
    printf("ALL OK\n");
    return 0;
}
