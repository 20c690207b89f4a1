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
    dist.barrier()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
