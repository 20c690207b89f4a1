// tokenizer.h — GPT-NeoX-20B byte-level BPE tokenizer with the reference's exact behaviour.
//
// API parity with include/rwkv/tokenizer/tokenizer.h of the reference (class GPT2Tokenizer:
// load / encode / decode / tokenize / vocab_size, tokenizer.h:53,105,127,139,160). The token
// ids must be bit-exact, so the reference's observable quirks are kept on purpose:
//   Q1  load() throws away the first line of merges.txt as a "#version" header even though the
//       shipped file has none, so the first real merge ("Ġ Ġ") is lost and every rank shifts
//       by one (tokenizer.h:66-74);
//   Q2  the pre-tokeniser is the ECMAScript pattern
//         's|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+
//       evaluated byte-wise in the "C" locale (tokenizer.h:50), i.e. bytes >= 0x80 are never
//       alpha/digit/space;
//   Q3  unknown symbols encode to id 0, unknown ids decode to "", unknown code points decode
//       to a NUL byte (operator[] default-insertions at tokenizer.h:111,130,133);
//   Q4  BPE merges the lowest-rank adjacent pair everywhere, left to right, until no ranked
//       pair is left (tokenizer.h:172-247).
// The implementation is new: a hand-written scanner instead of <regex>, a generated
// byte<->code-point table instead of two 256-entry literals, and a ~100-line flat JSON
// reader instead of the vendored simdjson (the vocab is one flat {string: int} object).
#pragma once
#include <cstdint>
#include <fstream>
#include <iostream>
#include <memory>
#include <optional>
#include <sstream>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

namespace rwkv_tok_detail {

inline void append_utf8(std::string &s, uint32_t cp) {
    if (cp < 0x80) {
        s += (char)cp;
    } else if (cp < 0x800) {
        s += (char)(0xC0 | (cp >> 6));
        s += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
        s += (char)(0xE0 | (cp >> 12));
        s += (char)(0x80 | ((cp >> 6) & 0x3F));
        s += (char)(0x80 | (cp & 0x3F));
    } else {
        s += (char)(0xF0 | (cp >> 18));
        s += (char)(0x80 | ((cp >> 12) & 0x3F));
        s += (char)(0x80 | ((cp >> 6) & 0x3F));
        s += (char)(0x80 | (cp & 0x3F));
    }
}

// GPT-2 "bytes_to_unicode": printable Latin-1 bytes map to themselves, the other 68 bytes
// to U+0100.. in ascending byte order. Equals the literal tables at tokenizer.h:23-33.
struct ByteTable {
    std::string enc[256];                        // byte -> UTF-8 of its code point
    std::unordered_map<std::string, char> dec;   // UTF-8 of code point -> byte
    ByteTable() {
        uint32_t next = 256;
        for (int b = 0; b < 256; ++b) {
            const bool keep = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
            const uint32_t cp = keep ? (uint32_t)b : next++;
            append_utf8(enc[b], cp);
            dec.emplace(enc[b], (char)b);
        }
    }
};
inline const ByteTable &byte_table() {
    static const ByteTable t;
    return t;
}

inline size_t utf8_len(unsigned char c) {
    if ((c & 0xf8) == 0xf0) return 4;
    if ((c & 0xf0) == 0xe0) return 3;
    if ((c & 0xe0) == 0xc0) return 2;
    return 1;
}

// "C"-locale character classes on raw bytes.
inline bool is_alpha(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_other(unsigned char c) { return !is_space(c) && !is_alpha(c) && !is_digit(c); }

// Length of the match of the pre-tokeniser pattern anchored at text[i] (always >= 1 for i < n).
inline size_t pretoken_len(const std::string &t, size_t i) {
    const size_t n = t.size();
    const unsigned char c = (unsigned char)t[i];
    // 's|'t|'re|'ve|'m|'ll|'d   (ordered alternation: first that matches)
    if (c == '\'' && i + 1 < n) {
        const char a = t[i + 1];
        const char b = i + 2 < n ? t[i + 2] : '\0';
        if (a == 's' || a == 't') return 2;
        if (a == 'r' && b == 'e') return 3;
        if (a == 'v' && b == 'e') return 3;
        if (a == 'm') return 2;
        if (a == 'l' && b == 'l') return 3;
        if (a == 'd') return 2;
    }
    // " ?[[:alpha:]]+" | " ?[[:digit:]]+" | " ?[^\s[:alpha:][:digit:]]+"
    {
        const size_t j = (c == ' ' && i + 1 < n) ? i + 1 : i;
        const unsigned char d = (unsigned char)t[j];
        bool (*cls)(unsigned char) = nullptr;
        if (is_alpha(d)) cls = is_alpha;
        else if (is_digit(d)) cls = is_digit;
        else if (is_other(d)) cls = is_other;
        if (cls) {
            size_t k = j + 1;
            while (k < n && cls((unsigned char)t[k])) ++k;
            return k - i;
        }
    }
    // "\s+(?!\S)": the whole whitespace run if it reaches the end of the text, otherwise all
    // but its last character (needs a run of at least two); else "\s+": the whole run.
    size_t k = i;
    while (k < n && is_space((unsigned char)t[k])) ++k;
    if (k == n) return k - i;
    if (k - i >= 2) return k - i - 1;
    return k - i;
}

struct PairHash {
    size_t operator()(const std::pair<std::string, std::string> &p) const noexcept {
        const std::hash<std::string> h;
        size_t a = h(p.first), b = h(p.second);
        return a ^ (b + 0x9e3779b97f4a7c15ULL + (a << 6) + (a >> 2));
    }
};

// Minimal reader for a flat JSON object {"key": integer, ...} with string escapes.
inline bool parse_flat_json(const std::string &s, std::vector<std::pair<std::string, int64_t>> &out) {
    size_t i = 0;
    const size_t n = s.size();
    auto ws = [&]() {
        while (i < n && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t')) ++i;
    };
    auto hex4 = [&](uint32_t &v) -> bool {
        if (i + 4 > n) return false;
        v = 0;
        for (int k = 0; k < 4; ++k) {
            const char c = s[i++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else return false;
        }
        return true;
    };
    ws();
    if (i >= n || s[i] != '{') return false;
    ++i;
    ws();
    if (i < n && s[i] == '}') return true;
    while (i < n) {
        ws();
        if (i >= n || s[i] != '"') return false;
        ++i;
        std::string key;
        while (i < n && s[i] != '"') {
            if (s[i] != '\\') {
                key += s[i++];
                continue;
            }
            if (++i >= n) return false;
            const char e = s[i++];
            switch (e) {
            case '"': key += '"'; break;
            case '\\': key += '\\'; break;
            case '/': key += '/'; break;
            case 'b': key += '\b'; break;
            case 'f': key += '\f'; break;
            case 'n': key += '\n'; break;
            case 'r': key += '\r'; break;
            case 't': key += '\t'; break;
            case 'u': {
                uint32_t cp;
                if (!hex4(cp)) return false;
                if (cp >= 0xD800 && cp <= 0xDBFF && i + 6 <= n && s[i] == '\\' && s[i + 1] == 'u') {
                    i += 2;
                    uint32_t lo;
                    if (!hex4(lo)) return false;
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                append_utf8(key, cp);
                break;
            }
            default: return false;
            }
        }
        if (i >= n) return false;
        ++i; // closing quote
        ws();
        if (i >= n || s[i] != ':') return false;
        ++i;
        ws();
        bool neg = false;
        if (i < n && s[i] == '-') {
            neg = true;
            ++i;
        }
        if (i >= n || s[i] < '0' || s[i] > '9') return false;
        int64_t v = 0;
        while (i < n && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
        out.emplace_back(std::move(key), neg ? -v : v);
        ws();
        if (i < n && s[i] == ',') {
            ++i;
            continue;
        }
        if (i < n && s[i] == '}') return true;
        return false;
    }
    return false;
}

} // namespace rwkv_tok_detail

class GPT2Tokenizer {
    using BPE = std::pair<std::string, std::string>;
    using BPERanks = std::unordered_map<BPE, size_t, rwkv_tok_detail::PairHash>;
    using Encoder = std::unordered_map<std::string, int64_t>;
    using Decoder = std::unordered_map<int64_t, std::string>;

  public:
    static std::optional<GPT2Tokenizer> load(std::string_view vocab_file, std::string_view merges_file) {
        std::ifstream merges(std::string(merges_file).c_str());
        if (!merges.good()) {
            std::cerr << "Error: could not open merges file " << merges_file << std::endl;
            return std::nullopt;
        }
        GPT2Tokenizer tok;
        std::string line;
        std::getline(merges, line); // Q1: first line is discarded unconditionally
        for (size_t rank = 0; std::getline(merges, line); ++rank) {
            const size_t sp = line.find(' ');
            if (sp == std::string::npos) { // the reference would build a garbage pair here; no such line exists
                continue;
            }
            tok.m_bpe_ranks.emplace(BPE{line.substr(0, sp), line.substr(sp + 1)}, rank);
        }

        std::ifstream vocab(std::string(vocab_file).c_str(), std::ios::binary);
        if (!vocab.good()) {
            std::cerr << "Error: could not open vocab file " << vocab_file << std::endl;
            return std::nullopt;
        }
        std::stringstream buf;
        buf << vocab.rdbuf();
        std::vector<std::pair<std::string, int64_t>> entries;
        if (!rwkv_tok_detail::parse_flat_json(buf.str(), entries)) {
            std::cerr << "Error: " << vocab_file << " is not a flat JSON object of string -> integer" << std::endl;
            return std::nullopt;
        }
        for (auto &kv : entries) {
            tok.m_encoder.emplace(kv.first, kv.second);
            tok.m_decoder.emplace(kv.second, kv.first);
        }
        return tok;
    }

    std::vector<long long> encode(const std::string &text) {
        const std::vector<std::string> pieces = tokenize(text);
        std::vector<long long> ids;
        ids.reserve(pieces.size());
        for (const std::string &p : pieces) {
            const auto it = m_encoder.find(p);
            ids.push_back(it == m_encoder.end() ? 0 : (long long)it->second); // Q3
        }
        return ids;
    }

    std::string decode(const std::vector<long long> &token_ids) {
        const auto &bt = rwkv_tok_detail::byte_table();
        std::string text;
        for (const long long id : token_ids) {
            const auto it = m_decoder.find((int64_t)id);
            if (it == m_decoder.end()) continue; // Q3: unknown id contributes nothing
            const std::string &sym = it->second;
            for (size_t i = 0; i < sym.size();) {
                const size_t len = rwkv_tok_detail::utf8_len((unsigned char)sym[i]);
                const auto b = bt.dec.find(sym.substr(i, len));
                text += (b == bt.dec.end()) ? '\0' : b->second; // Q3
                i += len;
            }
        }
        return text;
    }

    std::vector<std::string> tokenize(const std::string &text) {
        const auto &bt = rwkv_tok_detail::byte_table();
        std::vector<std::string> result;
        for (size_t i = 0; i < text.size();) {
            const size_t len = rwkv_tok_detail::pretoken_len(text, i); // Q2
            std::vector<std::string> word;
            word.reserve(len);
            for (size_t k = i; k < i + len; ++k) word.push_back(bt.enc[(unsigned char)text[k]]);
            bpe(word);
            for (std::string &w : word) result.push_back(std::move(w));
            i += len;
        }
        return result;
    }

    size_t vocab_size() const noexcept { return m_encoder.size(); }

  protected:
    GPT2Tokenizer() = default;

    BPERanks m_bpe_ranks;
    Encoder m_encoder;
    Decoder m_decoder;

  private:
    // Q4: in-place byte-pair merging of one pre-token given as a list of symbols.
    void bpe(std::vector<std::string> &word) const {
        const size_t none = (size_t)-1;
        while (word.size() > 1) {
            size_t best_rank = none, best_pos = 0;
            for (size_t i = 0; i + 1 < word.size(); ++i) {
                const auto it = m_bpe_ranks.find(BPE{word[i], word[i + 1]});
                if (it != m_bpe_ranks.end() && it->second < best_rank) {
                    best_rank = it->second;
                    best_pos = i;
                }
            }
            if (best_rank == none) break;
            const std::string first = word[best_pos], second = word[best_pos + 1];
            std::vector<std::string> merged;
            merged.reserve(word.size());
            for (size_t i = 0; i < word.size();) {
                if (i + 1 < word.size() && word[i] == first && word[i + 1] == second) {
                    merged.push_back(first + second);
                    i += 2;
                } else {
                    merged.push_back(word[i]);
                    i += 1;
                }
            }
            word.swap(merged);
        }
    }
};
