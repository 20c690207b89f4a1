"""Host side of the tensor-parallel path on CPU: the row partition over the grid formed by all ranks and
the handle exchange over a world-size-2 gloo group (no GPU, no compute calls)."""
import importlib
import os
import sys

import pytest

from util import ROOT


@pytest.fixture(scope="module")
def tp():
    sys.path.insert(0, ROOT)
    return importlib.import_module("rwkv-cpp-accelerated_b200").tp


@pytest.mark.parametrize("rows", [768, 4096, 4 * 5120, 50277])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_covers_every_row_once(tp, rows, world):
    grid = 148
    parts = tp.partition(rows, grid, world)
    assert len(parts) == grid * world
    assert parts[0][0] == 0 and parts[-1][1] == rows
    for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
        assert a1 == b0 and a0 <= a1            # contiguous, ascending, no overlap
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1         # balanced to one row
    # a rank's CTAs own one contiguous block: that is what it streams from HBM per token
    for r in range(world):
        r0, r1 = tp.rank_rows(rows, grid, world, r)
        assert r1 - r0 == sum(sizes[r * grid:(r + 1) * grid])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    tp = importlib.import_module("rwkv-cpp-accelerated_b200").tp
    got = tp.exchange_handles(bytes([rank]) * 64, group=dist.group.WORLD)
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


def test_handle_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r] == [bytes([0]) * 64, bytes([1]) * 64]   # rank order, identical on every rank


def test_tp_load_needs_a_gpu_and_valid_rank():
    pkg = importlib.import_module("rwkv-cpp-accelerated_b200")
    lib = pkg.load_library()
    import ctypes
    h = ctypes.c_void_p()
    L, E = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    rc = lib.rwkv_b200_load_tp(b"/nonexistent.bin", 1, 0, 1, 3, 2, ctypes.byref(h), ctypes.byref(L), ctypes.byref(E))
    assert rc != 0 and b"rank 3 of 2" in lib.rwkv_b200_last_error()
    assert lib.rwkv_b200_tp_buffer_bytes(None) == 0
