"""Tokens/s of ONE stream decoded by a tensor-parallel group (launched under torchrun, one rank per GPU).
usage: torchrun --nproc-per-node G tp_bench.py [workload=7b] [steps=128] [opt=value ...]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    workload = sys.argv[1] if len(sys.argv) > 1 else "7b"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("rwkv-cpp-accelerated_b200")
    if rank == 0:
        bench.model_path(workload, pkg)
    dist.barrier()
    path = bench.model_path(workload, pkg)
    eng = pkg.Engine(path, device=local, tp_rank=rank, tp_size=world)
    pkg.tp.connect(eng)
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        eng.set_option(k, v)
    eng.state_zero()
    eng.decode_timed([bench.SEED_TOKEN] * 8, teacher_forced=False)
    eng.state_zero()
    dist.barrier()
    torch.cuda.synchronize()
    ms = eng.decode_timed([bench.SEED_TOKEN] * steps, teacher_forced=False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print("TP%d %s: %.1f tok/s (%.3f ms/token, max over ranks)" % (world, workload, steps / (ms / 1e3), ms / steps), flush=True)
    if any(a == "trace=1" for a in sys.argv[3:]):
        trace_report(eng, rank, workload)
    dist.barrier()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def trace_report(eng, rank, workload):
    """Per-layer segment times of rank 0's CTAs from the %globaltimer stamps (same layout as tools/trace_token.py)."""
    import numpy as np
    L, E = bench.SHAPES[workload]
    tok = bench.SEED_TOKEN
    for _ in range(4):
        tok = eng.forward_greedy(tok)
    if rank != 0:
        return
    tr = eng.read_trace().astype(np.int64)
    n = int((tr[0] > 0).sum())
    tr = tr[:, :n] - tr[:, 0].min()
    ST = ["st.sync", "st.pub", "st.first", "st.all", "st.calc", "st.ret"]
    GA = ["g.meet", "g.words", "g.max", "g.sync", "g.quant"]
    names = ["ln0." + x for x in ST] + ["ln0.pub"]
    layer = ["kvr." + x for x in GA] + ["kvr.gemv", "kvr.epi"] + ["out." + x for x in GA] + ["out.gemv", "out.resid"] + \
            ["out." + x for x in ST] + ["out.ln2pub"] + ["rk." + x for x in GA] + ["rk.gemv", "rk.epi", "fr.gemv", "fr.epi"] + \
            ["v." + x for x in GA] + ["v.gemv", "v.resid"] + ["v." + x for x in ST] + ["v.ln1pub", "v.end"]
    for _ in range(L):
        names += layer
    names += ["head." + x for x in GA] + ["head.gemv", "head.epi", "done"]
    d = np.diff(tr, axis=1)
    print("stamps per CTA", n, "segments", d.shape[1], "expected", len(names), "token time us", tr[:, -1].max() / 1e3)
    agg = {}
    for i, nm in enumerate(names[:d.shape[1]]):
        agg.setdefault(nm, []).append(d[:, i])
    tot = 0.0
    for nm, lst in agg.items():
        a = np.stack(lst, 1) / 1e3
        if nm.split(".")[0] in ("kvr", "out", "rk", "fr", "v"):
            tot += a.mean()
        print("%-12s %8.2f %8.2f %8.2f   x%d" % (nm, a.mean(), a.mean(1).min(), a.mean(1).max(), a.shape[1]))
    print("per layer total (us):", round(tot, 2))


if __name__ == "__main__":
    main()
