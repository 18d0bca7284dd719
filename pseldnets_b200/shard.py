"""Sharding of a clip batch across the GPUs of one box, and the cross-rank result check.

The hot path has no exchange step: every 10-s clip is independent, so rank r of W
simply owns a contiguous slice of the global batch -- the same partition the reference's
`UserDistributedBatchSampler` produces (/root/reference/src/data/components/sampler.py:11-14,44:
global batch = batch_size x world, each rank takes its slice).  No collective touches features;
the one all_gather below moves 16 bytes per clip (sum and sum of squares of each clip's feature
map, in fp64) and exists only so a benchmark or test can verify every rank's output from rank 0.
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU test-suite).
"""
import torch
import torch.distributed as dist


def clip_shard(n_clips, rank, world):
    """[start, stop) of the clips rank `rank` owns; sizes differ by at most one, order preserved."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def clip_checksums(y):
    """(n, 2) fp64 tensor: per clip [sum, sum of squares] of its feature map."""
    yd = y.reshape(y.shape[0], -1).double()
    return torch.stack([yd.sum(dim=1), (yd * yd).sum(dim=1)], dim=1)


def gather_clip_checksums(y, n_clips_global, group=None):
    """All ranks call this with their local feature maps (shard order = clip_shard); every rank
    gets the (n_clips_global, 2) checksum table in global clip order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local = clip_checksums(y)
    if world == 1:
        return local
    sizes = [clip_shard(n_clips_global, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    padded = torch.zeros((mx, 2), dtype=torch.float64, device=y.device)
    padded[:local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    assert sizes[rank][1] - sizes[rank][0] == local.shape[0], 'local batch does not match clip_shard()'
    return torch.cat([bufs[r][:b - a] for r, (a, b) in enumerate(sizes)], dim=0)
