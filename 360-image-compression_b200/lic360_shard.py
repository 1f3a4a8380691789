"""Image-level sharding of the codec path across the GPUs of one box (every image has its own two
bitstreams, coder state and wavefront, so the path partitions by image with NO data-path collective).

One process per GPU (torchrun / mp.spawn, like the reference's trainers, train/trainDDP_IMP_ENT.py:105-108); the only
cross-rank traffic is bookkeeping: the max-over-ranks step time and the gathered stream sizes.  Used by bench.py
(NCCL on the GPU box) and covered by a world_size-2 gloo test on CPU (tests/test_shard_gloo.py).
"""
import torch
import torch.distributed as dist


def rank_images(n_images, world, rank):
    """Indices of the images rank `rank` codes: contiguous blocks, remainder to the first ranks (16 images over
    1/2/4/8 GPUs -> 16/8/4/2 per GPU, BASELINE.json configs[2])."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank_images: bad world/rank %r/%r" % (world, rank))
    base, rem = divmod(int(n_images), world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def max_over_ranks(value, device="cpu"):
    """Step time as the driver wants it: the slowest rank's."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_stream_sizes(local, n_images, device="cpu"):
    """local: {image index: (imp_bytes, code_bytes)} of this rank -> the full table on every rank.
    Each image must be reported by exactly one rank (checked)."""
    tab = torch.zeros((n_images, 3), dtype=torch.int64, device=device)
    for i, (a, b) in local.items():
        tab[i, 0] = int(a)
        tab[i, 1] = int(b)
        tab[i, 2] = 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tab, op=dist.ReduceOp.SUM)
    owners = tab[:, 2].tolist()
    if any(o != 1 for o in owners):
        raise RuntimeError("gather_stream_sizes: images coded by %s ranks (expected exactly one each)" % owners)
    return [(int(a), int(b)) for a, b, _ in tab.tolist()]
