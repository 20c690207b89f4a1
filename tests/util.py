import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
PKG_DIR = os.path.join(ROOT, "rwkv-cpp-accelerated_b200")
VOCAB_DIR = os.path.join(INCLUDE, "rwkv", "tokenizer", "vocab")


def compile_cpp(src, out, link_engine=False, extra=()):
    cmd = ["g++", "-O1", "-std=c++17", "-I" + INCLUDE, src, "-o", out] + list(extra)
    if link_engine:
        cmd += ["-L" + PKG_DIR, "-lrwkv_b200", "-Wl,-rpath," + PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "g++ failed:\n" + r.stderr[-4000:]
    return out
