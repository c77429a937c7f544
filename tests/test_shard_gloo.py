"""Sharding logic on CPU: index math plus a world_size-2 gloo run of the
optional spectra gather (the only collective the path offers)."""
import os
import socket

import numpy as np
import pytest

from dsp_audio_project_b200.shard import block_sizes, channel_block


def test_channel_blocks_cover_everything():
    for n in (0, 1, 7, 8, 9, 1024, 262144, 262145):
        for world in (1, 2, 3, 4, 8):
            spans = [channel_block(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(block_sizes(n, world)) == n
    assert channel_block(262144, 8, 3) == (98304, 131072)
    with pytest.raises(ValueError):
        channel_block(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_channels, out_dir):
    import torch
    import torch.distributed as dist
    from dsp_audio_project_b200.shard import channel_block, gather_spectra

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = channel_block(n_channels, world, rank)
        # stand-in spectra: value encodes (channel, frame, bin)
        c = torch.arange(start, stop, dtype=torch.float32).view(-1, 1, 1)
        local = c * 100 + torch.arange(3).view(1, 3, 1) * 10 + torch.arange(5).view(1, 1, 5)
        full = gather_spectra(local, n_channels)
        from dsp_audio_project_b200.shard import SpectraGather
        g = SpectraGather(n_channels)
        assert torch.equal(g.start(local).wait(), full)          # the side-stream form (synchronous under gloo)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_channels", [7, 8])
def test_gather_spectra_world2_gloo(tmp_path, n_channels):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_channels, str(tmp_path)), nprocs=2, join=True)
    c = np.arange(n_channels, dtype=np.float32).reshape(-1, 1, 1)
    want = c * 100 + np.arange(3).reshape(1, 3, 1) * 10 + np.arange(5).reshape(1, 1, 5)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == want.shape and np.array_equal(got, want)
