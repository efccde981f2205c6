"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The hot path shards by RAYS with the voxel set replicated: forward needs no communication.
Collectives exist only where the path has a real exchange step:
  * training: gradient all-reduce (mean) of values.weight.grad + MLP grads — DDP, or allreduce_grads() below;
  * pruning : voxels are sharded, each rank scores its slice, the uint8 keep mask is all-gathered
              (new design; the reference recomputes the full mask on every rank, encoder.py:605-618);
  * pruning with train stats: all-reduce MAX of max_voxel_probs (encoder.py:613-614, kept in encoder.pruning).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world, align=1):
    """Contiguous slice [lo, hi) of n items owned by `rank`: sizes differ by at most one block of `align` items, and
    every boundary except the last is a multiple of `align`."""
    blocks = (n + align - 1) // align
    base, rem = divmod(blocks, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return min(lo * align, n), min(hi * align, n)


def shard_rays(ray_start, ray_dir, rank, world, dim=1):
    """Per-rank slice of the rays along `dim` (images first: pass dim of the view axis when V >= world)."""
    lo, hi = shard_range(ray_dir.size(dim), rank, world)
    sl = [slice(None)] * ray_dir.dim()
    sl[dim] = slice(lo, hi)
    rs = ray_start[tuple(sl)] if ray_start.size(dim) == ray_dir.size(dim) else ray_start
    return rs, ray_dir[tuple(sl)]


def allgather_keep_mask(local_keep, n, rank, world, group=None, align=1):
    """All-gather of the per-shard keep masks (uint8, n bytes in total) -> bool [n] on every rank."""
    if world == 1 or not dist.is_initialized():
        return local_keep.bool()
    sizes = [shard_range(n, r, world, align)[1] - shard_range(n, r, world, align)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.uint8, device=local_keep.device)
    buf[: local_keep.numel()] = local_keep.to(torch.uint8)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)]).bool()


def allreduce_grads(params, world, group=None):
    """Mean all-reduce of the gradients (what DDP does for the reference; explicit variant for manual loops).
    The embedding gradient is reduced on its own so it can be launched as soon as the scatter-add kernel ends."""
    if world == 1 or not dist.is_initialized():
        return
    handles = []
    for p in params:
        if p.grad is not None:
            handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    for p in params:
        if p.grad is not None:
            p.grad.div_(world)


def gather_frames(local_frames, rank, world, dst=0, group=None):
    """Rendering: each rank renders a contiguous range of frames (render_multigpu.py:95-104); gather on dst."""
    if world == 1 or not dist.is_initialized():
        return local_frames
    out = [None] * world if rank == dst else None
    dist.gather_object(local_frames, out, dst=dst, group=group)
    return [f for part in out for f in part] if rank == dst else None
