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


class SpectraGather:
    """The optional exchange of SURVEY.md 8e off the data path: the all-gather of a wave's spectra runs on a side
    stream, so the next wave's kernels (main stream) overlap the NVLink transfer.

        g = SpectraGather(n_channels)           # once; n_channels = clips of a wave over ALL ranks
        chain.run(x, z=z, mag=mag)              # wave k on the current stream
        g.start(mag)                            # returns at once; the collective waits for wave k on the device
        chain.run(x2, z=z2, mag=mag2)           # wave k+1 overlaps the gather
        full = g.wait()                         # [n_channels, frames, bins] on every rank, current stream ordered after it

    ``mag`` must not be overwritten before wait() (use two spectra buffers, as above).  NCCL on GPUs; with the gloo
    backend (CPU tests) the same calls run synchronously."""

    def __init__(self, n_channels: int, group=None):
        import torch
        import torch.distributed as dist

        self.n_channels = int(n_channels)
        self.group = group
        self.world = dist.get_world_size(group)
        self.per = -(-self.n_channels // self.world) if self.n_channels else 0
        self._cuda = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        self._side = torch.cuda.Stream() if self._cuda else None
        self._out = None
        self._pad = None
        self._done = None

    def start(self, local):
        import torch
        import torch.distributed as dist

        if local.shape[0] > self.per:
            raise ValueError("local block larger than ceil(n_channels / world)")
        shape = (self.world * self.per,) + tuple(local.shape[1:])
        if self._out is None or tuple(self._out.shape) != shape or self._out.dtype != local.dtype:
            self._out = local.new_empty(shape)
        if not self._cuda:
            src = local
            if local.shape[0] < self.per:
                src = torch.cat([local, local.new_zeros((self.per - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
            dist.all_gather_into_tensor(self._out, src.contiguous(), group=self.group)
            return self
        ready = torch.cuda.Event()
        ready.record()                                   # wave k's kernels, on the caller's stream
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            src = local
            if local.shape[0] < self.per:                # the last rank may own fewer clips: pad on the side stream
                if self._pad is None or tuple(self._pad.shape) != (self.per,) + tuple(local.shape[1:]):
                    self._pad = local.new_zeros((self.per,) + tuple(local.shape[1:]))
                self._pad[:local.shape[0]].copy_(local, non_blocking=True)
                src = self._pad
            local.record_stream(self._side)
            dist.all_gather_into_tensor(self._out, src.contiguous(), group=self.group)
            self._done = torch.cuda.Event()
            self._done.record()
        return self

    def wait(self):
        import torch

        if self._cuda and self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None
        return self._out[:self.n_channels]
