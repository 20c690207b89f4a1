"""Cycle counters of the slice statistics inside the real token kernel (set_option dbg=4): no global stores in between."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
pkg = importlib.import_module("rwkv-cpp-accelerated_b200")
eng = pkg.Engine(bench.model_path(sys.argv[1] if len(sys.argv) > 1 else "7b", pkg))
eng.set_option("trace", 1)
eng.set_option("dbg", int(os.environ.get("DBG", "4")))
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    eng.set_option(k, v)
tok = bench.SEED_TOKEN
for _ in range(3):
    tok = eng.forward_greedy(tok)
tr = eng.read_trace().astype(np.int64)[:, :16]
n = tr[:, 4].clip(1)
for i, nm in enumerate(["first trees", "publish+records", "syncwarp", "second trees"]):
    per = tr[:, i] / n
    print("%-16s cycles per call: mean %.0f min %.0f max %.0f" % (nm, per.mean(), per.min(), per.max()))
print("calls per token", n.mean())
if int(os.environ.get("DBG", "4")) & 8:
    tr = eng.read_trace().astype(np.int64)[:, :16]
    n = tr[:, 13].clip(1)
    for i, nm in enumerate(["scales (div)", "quantise loop", "offset trees", "barrier", "offset chain"]):
        per = tr[:, 8 + i] / n
        print("gather: %-16s cycles per call: mean %.0f min %.0f max %.0f" % (nm, per.mean(), per.min(), per.max()))
    print("active lanes with thread 0: after the barrier %.2f, after the quantise loop %.2f" % ((tr[:, 14] / n).mean(), (tr[:, 15] / n).mean()))
    print("gathers per token", tr[:, 13].mean())
    eng.close()
    sys.exit(0)
tr = eng.read_trace().astype(np.int64)[:, :16]
n = tr[:, 12].clip(1)
for i, nm in enumerate(["first trees", "publish+records", "syncwarp", "second trees"]):
    per = tr[:, 8 + i] / n
    print("cold call: %-16s cycles per call: mean %.0f min %.0f max %.0f" % (nm, per.mean(), per.min(), per.max()))
print("cold calls per token", tr[:, 12].mean())
eng.close()
