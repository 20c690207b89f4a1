// ref_tokenizer.cpp — golden-vector generator: runs the UNMODIFIED reference tokenizer
// (include/rwkv/tokenizer/tokenizer.h, resolved through -I/root/reference/include) on
// hex-encoded inputs read from stdin, one per line. TEST INFRASTRUCTURE, CPU only, built
// by `make -C oracle ref-tools` into oracle/_ref/.
//   usage: ref_tokenizer <vocab.json> <merges.txt> < cases.hex > ids.txt
// Output per input line: the token ids separated by spaces, then " | " and the hex of
// decode(ids).
#include <cstdio>
#include <iostream>
#include <string>
#include "rwkv/tokenizer/tokenizer.h"

static std::string unhex(const std::string &h) {
    std::string s;
    for (size_t i = 0; i + 1 < h.size(); i += 2) s += (char)std::stoi(h.substr(i, 2), nullptr, 16);
    return s;
}
int main(int argc, char **argv) {
    if (argc < 3) return 1;
    auto t = GPT2Tokenizer::load(argv[1], argv[2]);
    if (!t.has_value()) return 2;
    GPT2Tokenizer tok = t.value();
    std::cerr << "vocab_size " << tok.vocab_size() << "\n";
    std::string line;
    while (std::getline(std::cin, line)) {
        const std::string text = unhex(line);
        auto ids = tok.encode(text);
        for (size_t i = 0; i < ids.size(); ++i) printf(i ? " %lld" : "%lld", ids[i]);
        const std::string back = tok.decode(ids);
        printf(" | ");
        for (unsigned char c : back) printf("%02x", c);
        printf("\n");
    }
    return 0;
}
