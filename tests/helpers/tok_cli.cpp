// tok_cli.cpp — same command-line contract as oracle/ref_tokenizer.cpp, but compiled
// against THIS repository's include/rwkv/tokenizer/tokenizer.h.
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
