"""Clip sharding across the GPUs of one box.

Denoising is independent per clip (eval-mode norms use running statistics, diffusion_trainer.py:862), so the loop
runs with no collective: rank r owns a contiguous block of ceil(N / world) clips -- the same partition
DistributedSampler gives the reference's test loaders (datasets/prepare_data.py:87-103) -- and the predicted maps are
gathered once after the loop (344 KB per clip).  torch.distributed is the plumbing (NCCL over NVLink on the GPUs,
gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_clips, rank, world):
    """[start, stop) of the clips owned by ``rank``; blocks of ceil(n/world), the tail ranks may be short/empty."""
    per = (n_clips + world - 1) // world
    start = min(n_clips, rank * per)
    return start, min(n_clips, start + per)


def gather_maps(local_maps, n_clips, group=None):
    """All ranks receive the [n_clips, 1, H, W] maps in clip order.  ``local_maps`` is this rank's
    [n_local, 1, H, W] block (n_local may be smaller than ceil(n/world) on the last ranks: it is zero-padded
    for the equal-size all_gather and trimmed afterwards)."""
    if not dist.is_available() or not dist.is_initialized():
        return local_maps[:n_clips]
    world = dist.get_world_size(group)
    per = (n_clips + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    pad[: local_maps.shape[0]] = local_maps
    out = torch.empty((world * per,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_clips]
