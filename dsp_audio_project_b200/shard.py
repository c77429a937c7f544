"""Channel sharding across the GPUs of one box.

Every channel/clip is independent end to end, so the batch is cut into
contiguous blocks of ``ceil(C / world)`` channels, one process per GPU, with NO
collective on the data path.  The only exchange offered is an optional gather
of per-clip spectra to every rank (``torch.distributed`` all-gather: NCCL on
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def channel_block(n_channels: int, world: int, rank: int):
    """[start, stop) of the contiguous channel block rank owns."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    if n_channels < 0:
        raise ValueError("negative channel count")
    per = -(-n_channels // world) if n_channels else 0
    start = min(rank * per, n_channels)
    stop = min(start + per, n_channels)
    return start, stop


def block_sizes(n_channels: int, world: int):
    return [channel_block(n_channels, world, r)[1] - channel_block(n_channels, world, r)[0]
            for r in range(world)]


def gather_spectra(local, n_channels: int, group=None):
    """All-gather per-rank spectra blocks [c_local, ...] into [n_channels, ...]
    on every rank.  Blocks are padded to the common block size for the
    collective and trimmed afterwards (the last rank may own fewer channels)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = -(-n_channels // world) if n_channels else 0
    pad = per - local.shape[0]
    if pad < 0:
        raise ValueError("local block larger than ceil(n_channels / world)")
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
    local = local.contiguous()
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:n_channels]
