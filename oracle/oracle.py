"""ctypes view of oracle/librwkv_oracle.so — TEST INFRASTRUCTURE (see rwkv_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the engine package never does.
"""
import ctypes
import os

import numpy as np

VOCAB = 50277
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librwkv_oracle.so")
REF_HARNESS = os.path.join(_HERE, "_ref", "ref_harness")
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_load.restype = ctypes.c_void_p
        lib.oracle_load.argtypes = [ctypes.c_char_p]
        lib.oracle_free.argtypes = [ctypes.c_void_p]
        lib.oracle_n_layers.restype = ctypes.c_ulonglong
        lib.oracle_n_layers.argtypes = [ctypes.c_void_p]
        lib.oracle_n_embed.restype = ctypes.c_ulonglong
        lib.oracle_n_embed.argtypes = [ctypes.c_void_p]
        lib.oracle_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.oracle_max_threads.restype = ctypes.c_int
        lib.oracle_forward.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, dp, dp, dp, dp, dp,
                                       ctypes.POINTER(ctypes.c_float)]
        lib.oracle_last_x.restype = dp
        lib.oracle_last_x.argtypes = [ctypes.c_void_p]
        _LIB = lib
    return _LIB


class Oracle:
    """CPU restatement of the reference forward on a reference-format .bin."""

    def __init__(self, path, threads=None):
        lib = _lib()
        self.h = lib.oracle_load(path.encode())
        if not self.h:
            raise RuntimeError("oracle: cannot load %s" % path)
        self.n_layers = lib.oracle_n_layers(self.h)
        self.n_embed = lib.oracle_n_embed(self.h)
        self.threads = threads or lib.oracle_max_threads()
        lib.oracle_set_threads(self.h, self.threads)
        self.reset()

    def reset(self):
        n = self.n_layers * self.n_embed
        self.state = {k: np.zeros(n, np.float64) for k in ("xy", "aa", "bb", "pp", "dd")}

    def forward(self, token):
        dp = ctypes.POINTER(ctypes.c_double)
        out = np.empty(VOCAB, np.float32)
        s = self.state
        _lib().oracle_forward(self.h, int(token), *[s[k].ctypes.data_as(dp) for k in ("xy", "aa", "bb", "pp", "dd")],
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return out

    def last_x(self):
        p = _lib().oracle_last_x(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n_embed,)).copy()

    def close(self):
        if self.h:
            _lib().oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_ref_dump(path):
    """Parse the dump written by oracle/ref_harness.cpp."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:40].view(np.uint64)
    assert hdr[0] == 0x524546484152, "bad ref dump magic"
    n, V, L, E = int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[4])
    off = 40
    steps, logits = [], []
    for _ in range(n):
        steps.append(int(raw[off:off + 8].view(np.uint64)[0]))
        off += 8
        logits.append(raw[off:off + 4 * V].view(np.float32).copy())
        off += 4 * V
    st = {}
    for k in ("xy", "aa", "bb", "pp", "dd"):
        st[k] = raw[off:off + 8 * L * E].view(np.float64).copy()
        off += 8 * L * E
    return {"steps": steps, "logits": logits, "state": st, "L": L, "E": E}
