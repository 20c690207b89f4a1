// ref_harness.cpp — driver around the UNMODIFIED reference (rwkv.h + cuda/rwkv.cu).
// TEST / BASELINE INFRASTRUCTURE: built only by oracle/Makefile into oracle/_ref/, from
// the sources where they lie under /root/reference; nothing from the reference is
// copied into this repository. It gives the two things the reference itself never
// shipped: a golden forward (logits + state dumps) and a tokens/s number measured
// through its own public API (RWKV::loadFile + RWKV::forward, rwkv.h:281,378).
//
//   ref_harness <model.bin> <tokens.txt> <dump.bin> [--warmup W] [--dump-every K] [--greedy N]
//
// tokens.txt: whitespace separated token ids (teacher forced). With --greedy N only the
// first id is used and the next N-1 inputs are the argmax of the reference's own logits;
// the ids actually fed are written back as "<dump.bin>.tokens".
// dump.bin: u64 magic, u64 n_dumped, u64 V, u64 L, u64 E, then n_dumped x {u64 step,
// f32 logits[V]}, then the five state arrays (xy, aa, bb, pp, dd), L*E f64 each.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "rwkv/rwkv/rwkv.h" // the reference's host API (resolved through -I/root/reference/include)

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s model.bin tokens.txt dump.bin [--warmup W] [--dump-every K] [--greedy N]\n", argv[0]);
        return 1;
    }
    std::string model = argv[1], tokfile = argv[2], dump = argv[3];
    unsigned long long warmup = 0, dump_every = 1, greedy = 0;
    for (int i = 4; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--warmup")) warmup = strtoull(argv[i + 1], nullptr, 10);
        else if (!strcmp(argv[i], "--dump-every")) dump_every = strtoull(argv[i + 1], nullptr, 10);
        else if (!strcmp(argv[i], "--greedy")) greedy = strtoull(argv[i + 1], nullptr, 10);
    }
    std::vector<unsigned long long> tokens;
    {
        std::ifstream tf(tokfile);
        unsigned long long t;
        while (tf >> t) tokens.push_back(t);
    }
    if (tokens.empty()) {
        fprintf(stderr, "ref_harness: no tokens\n");
        return 1;
    }
    if (greedy) tokens.resize(1);
    const unsigned long long total = greedy ? greedy : tokens.size();

    RWKV net;
    net.loadFile(model, 1);
    const unsigned long long V = 50277, L = net.num_layers, E = net.num_embed;

    FILE *out = fopen(dump.c_str(), "wb");
    if (!out) {
        perror("ref_harness: dump");
        return 2;
    }
    const unsigned long long magic = 0x524546484152ULL; // "REFHAR"
    unsigned long long n_dumped = 0;
    unsigned long long hdr[5] = {magic, 0, V, L, E};
    fwrite(hdr, sizeof(hdr), 1, out);

    double timed_s = 0.0;
    unsigned long long timed_n = 0;
    for (unsigned long long step = 0; step < total; ++step) {
        const unsigned long long tok = tokens[step];
        auto t0 = std::chrono::steady_clock::now();
        float *logits = net.forward(tok);
        auto t1 = std::chrono::steady_clock::now();
        if (step >= warmup) {
            timed_s += std::chrono::duration<double>(t1 - t0).count();
            ++timed_n;
        }
        if (dump_every && (step % dump_every == 0 || step + 1 == total)) {
            fwrite(&step, sizeof(step), 1, out);
            fwrite(logits, sizeof(float), V, out);
            ++n_dumped;
        }
        if (greedy && step + 1 < total) {
            unsigned long long best = 0;
            for (unsigned long long i = 1; i < V; ++i)
                if (logits[i] > logits[best]) best = i;
            tokens.push_back(best);
        }
    }
    const size_t n = (size_t)(L * E);
    fwrite(net.state->statexy, sizeof(double), n, out);
    fwrite(net.state->stateaa, sizeof(double), n, out);
    fwrite(net.state->statebb, sizeof(double), n, out);
    fwrite(net.state->statepp, sizeof(double), n, out);
    fwrite(net.state->statedd, sizeof(double), n, out);
    hdr[1] = n_dumped;
    fseek(out, 0, SEEK_SET);
    fwrite(hdr, sizeof(hdr), 1, out);
    fclose(out);

    {
        std::ofstream tf(dump + ".tokens");
        for (auto t : tokens) tf << t << "\n";
    }
    printf("\nREF_RESULT {\"tokens\": %llu, \"seconds\": %.6f, \"tokens_per_s\": %.3f, \"n_layers\": %llu, \"n_embed\": %llu}\n",
           timed_n, timed_s, timed_n ? (double)timed_n / timed_s : 0.0, L, E);
    return 0;
}
