"""Tensor-parallel plumbing: one process per GPU, `torch.distributed` only to swap the 64-byte CUDA IPC
handles of the ranks' exchange blocks at start-up. The data path has no collective: the token kernel
stores into the peers' blocks over NVLink and synchronises with a system-scope grid barrier
(csrc/token_kernel.cuh)."""


def partition(rows, grid, world):
    """Row range [r0, r1) of every CTA of the grid formed by `world` ranks x `grid` CTAs - the same
    arithmetic as split_rows_g() in csrc/token_kernel.cuh. Returns a list indexed by rank*grid + cta."""
    n = grid * world
    return [((rows * b) // n, (rows * (b + 1)) // n) for b in range(n)]


def rank_rows(rows, grid, world, rank):
    """Rows [r0, r1) of a matrix that rank `rank` streams per token (union of its CTAs' ranges)."""
    parts = partition(rows, grid, world)
    return parts[rank * grid][0], parts[(rank + 1) * grid - 1][1]


def exchange_handles(handle, group=None):
    """All-gather one bytes object per rank, in rank order, over `group` (default: a gloo group created
    next to the default one - object collectives on NCCL need a CUDA context per pickle)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if group is None:
        group = dist.new_group(backend="gloo")
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(handle), group=group)
    return out


def connect(engine, group=None):
    """Wire a tensor-parallel Engine with its peers: export, all-gather, import."""
    import torch.distributed as dist
    handles = exchange_handles(engine.tp_export(), group)
    if len(handles) != engine.tp_size:
        raise RuntimeError("process group has %d ranks, engine expects %d" % (len(handles), engine.tp_size))
    engine.tp_import(handles)
    dist.barrier(group=group) if group is not None else dist.barrier()
