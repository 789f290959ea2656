"""Sharding of a batch of independent ciphertexts over the GPUs of one box.

The units of the hot path are independent and every per-key constant is
replicated, so the only multi-GPU machinery is a static contiguous block
partition (the shape of the reference's CPU/QAT split, ipcl/mod_exp.cpp:
702-731: prefix to one back-end, suffix to the other) plus, when the batch
starts on one rank, a scatter before and a gather after.  torch.distributed is
the plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(count, world, rank):
    """[start, stop) of rank's block: count//world each, remainder to the last
    ranks one element apiece (so no rank differs by more than one)."""
    base, rem = divmod(count, world)
    start = rank * base + max(0, rank - (world - rem))
    stop = start + base + (1 if rank >= world - rem else 0)
    return start, stop


def shard_sizes(count, world):
    return [shard_range(count, world, r)[1] - shard_range(count, world, r)[0]
            for r in range(world)]


def _world_rank(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def scatter_rows(full, count, cols, dtype, device, src=0, group=None):
    """Rank `src` holds `full` (count x cols); every rank gets its block.
    Grouped point-to-point sends: NCCL has no native scatter with ragged
    sizes, batched isend/irecv is the idiom."""
    world, rank = _world_rank(group)
    s, e = shard_range(count, world, rank)
    local = torch.empty((e - s, cols), dtype=dtype, device=device)
    if world == 1:
        local.copy_(full)
        return local
    ops = []
    if rank == src:
        for r in range(world):
            rs, re = shard_range(count, world, r)
            if r == src:
                local.copy_(full[rs:re])
            elif re > rs:
                ops.append(dist.P2POp(dist.isend, full[rs:re].contiguous(), r, group))
    elif e > s:
        ops.append(dist.P2POp(dist.irecv, local, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return local


def gather_rows(local, count, dst=0, group=None):
    """Inverse of scatter_rows: rank `dst` returns the (count x cols) result,
    the others return None."""
    world, rank = _world_rank(group)
    if world == 1:
        return local
    cols = local.shape[1]
    full = None
    ops = []
    if rank == dst:
        full = torch.empty((count, cols), dtype=local.dtype, device=local.device)
        for r in range(world):
            rs, re = shard_range(count, world, r)
            if r == dst:
                full[rs:re].copy_(local)
            elif re > rs:
                ops.append(dist.P2POp(dist.irecv, full[rs:re], r, group))
    elif local.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, local.contiguous(), dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full


def max_over_ranks(values, device, group=None):
    """element-wise max of a list of floats over all ranks (timing rule: a
    multi-GPU number is the max over ranks)"""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t.tolist()]
