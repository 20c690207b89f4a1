"""Regenerates tests/golden/*.json by running the UNMODIFIED reference code in this
container (it cannot travel to the GPU box, so the vectors are committed):

  tokenizer_golden.json  ids + decode round trip from the reference GPT2Tokenizer
                         (oracle/_ref/ref_tokenizer, built from /root/reference by
                         `make -C oracle ref-tools`) over fixed + seeded fuzz inputs
  sampler_golden.json    token sequences drawn by the reference typical() (NumCpp) from
                         synthetic logits, default-seeded generator

Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
VOCAB_DIR = "/root/reference/include/rwkv/tokenizer/vocab"

FIXED = [
    "To see the world in a grain of",  # tests/test_pybind.py:21
    "### Instruction: Write a story/book using the themes and details provided\n\n### Input:",  # storygen.cpp:5-7
    "Bob: Hello Alice, how are you doing?\n\nAlice:",
    "\n\n### Response:",
    "a  b", "    indented", "Hello, world! It's 2023.", "café 日本", "x\n\n\ny", "I'll we've don't",
    " ", "A", "", "'", "''s", "'re're'r", "it's'sit", "'T 'S 'LL", "don't stop", "x'y'z'", "a'll b'd c'm d've e're f't g's",
    "tabs\tand\ttabs\t\t", "trailing spaces   ", "   ", "\n", "\n\n", "\r\n\r\n", " \n \n", "a \nb", "a\n b", "1 2  3   4",
    "3.14159 2,000,000 1e-9", "foo_bar-baz+qux=42;", "C++ && C# || F#", "<|endoftext|>", "http://example.com/a?b=c&d=e",
    "éèê üß ñ", "中文测试，标点。", "\U0001F600 emoji \U0001F680",
    "mixed nbsp emspace", "The quick brown fox jumps over the lazy dog. " * 3,
    "def f(x):\n    return x ** 2  # square\n", "{\"k\": [1, 2, {\"z\": null}]}", b"\x00\x01\x7f\x80\xff",
]


def fuzz(n, seed):
    rnd = random.Random(seed)
    alphabets = [
        "abcdefghijklmnopqrstuvwxyz", "ABCDEFGHIJKLMNOPQRSTUVWXYZ", "0123456789", " \t\n\r\x0b\x0c", "    ", "'", "'stredvml",
        ".,;:!?-_()[]{}<>/\\|@#$%^&*+=~`\"", "éüñåø", "中文日本語", "\U0001F600\U0001F680",
    ]
    words = ["the", "of", "and", " to", "in", "is", "you", "that", "it's", "we've", "I'll", "they're", "don't", "2023", "3.5",
             "\n\n", "  ", " ", "\t", "Alice:", "Bob:", "###", "Response", "story", "hello", "world", "GPU", "RWKV"]
    out = []
    for _ in range(n):
        parts = []
        for _ in range(rnd.randint(1, 12)):
            if rnd.random() < 0.5:
                parts.append(rnd.choice(words))
            else:
                a = rnd.choice(alphabets)
                parts.append("".join(rnd.choice(a) for _ in range(rnd.randint(1, 6))))
            if rnd.random() < 0.4:
                parts.append(rnd.choice([" ", "  ", "\n", ""]))
        out.append("".join(parts))
    # raw random bytes (invalid UTF-8 included): the tokenizer is byte-level
    for _ in range(n // 8):
        out.append(bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 24))))
    return out


def tokenizer_golden():
    cases = [c if isinstance(c, bytes) else c.encode("utf-8", "surrogatepass") for c in FIXED + fuzz(400, 1234)]
    hexes = "\n".join(c.hex() for c in cases) + "\n"
    r = subprocess.run([os.path.join(REF, "ref_tokenizer"), VOCAB_DIR + "/vocab.json", VOCAB_DIR + "/merges.txt"],
                       input=hexes, capture_output=True, text=True, check=True)
    lines = r.stdout.split("\n")[:len(cases)]
    rows = []
    for c, line in zip(cases, lines):
        ids_s, back = line.split(" | ") if " | " in line else (line.replace(" |", ""), "")
        rows.append({"hex": c.hex(), "ids": [int(x) for x in ids_s.split()], "decoded_hex": back.strip()})
    vocab_size = int(r.stderr.split()[-1])
    with open(os.path.join(HERE, "tokenizer_golden.json"), "w") as f:
        json.dump({"source": "reference tokenizer.h @ /root/reference, g++ 13, this script", "vocab_size": vocab_size,
                   "cases": rows}, f, separators=(",", ":"))
    print("tokenizer: %d cases, vocab_size %d" % (len(rows), vocab_size))


def sampler_golden():
    runs = []
    for n, temp, tau, scale in [(12, 0.9, 0.8, 6.0), (12, 0.8, 0.7, 9.0), (8, 1.0, 0.95, 4.0), (8, 0.5, 0.8, 6.0),
                                (6, 0.3, 0.6, 3.0), (6, 1.5, 0.8, 6.0)]:
        r = subprocess.run([os.path.join(REF, "ref_sampler"), str(n), repr(temp), repr(tau), repr(scale)],
                           capture_output=True, text=True, check=True)
        runs.append({"n": n, "temp": temp, "tau": tau, "scale": scale, "tokens": [int(x) for x in r.stdout.split()]})
    with open(os.path.join(HERE, "sampler_golden.json"), "w") as f:
        json.dump({"source": "reference typical.h + NumCpp @ /root/reference, default-seeded mt19937_64", "runs": runs}, f)
    print("sampler:", [len(x["tokens"]) for x in runs])


if __name__ == "__main__":
    tokenizer_golden()
    sampler_golden()
