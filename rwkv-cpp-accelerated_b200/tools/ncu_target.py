"""Short decode for ncu (one launch of the token kernel per token): N tokens of the bench workload.
usage: python ncu_target.py [workload=7b] [tokens=3]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module("rwkv-cpp-accelerated_b200")
workload = sys.argv[1] if len(sys.argv) > 1 else "7b"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = pkg.Engine(bench.model_path(workload, pkg))
tok = bench.SEED_TOKEN
for _ in range(n):
    tok = eng.forward_greedy(tok)
print("ncu_target done", tok)
