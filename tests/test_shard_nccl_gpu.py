"""The optional spectra gather on hardware (north star "optional NCCL gather of per-clip spectra"):
two ranks, one GPU each, NCCL; every rank runs the chain on its own clip block, the gather of wave k runs on a side
stream while wave k+1's kernels run, and both ranks end up with the spectra a single GPU computes for all clips.
Skipped on a box with one GPU (the driver's N=2 runs and `gpurun --gpus 2` see two)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GAINS = {"Sub-Bass": 6, "Bass": -3, "Low Mids": 4, "High Mids": -6, "Presence": 3, "Brilliance": -9}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, n_in, out_dir):
    import torch
    import torch.distributed as dist

    import dsp_audio_project_b200 as pk
    from dsp_audio_project_b200.shard import SpectraGather, channel_block, gather_spectra

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        c0, c1 = channel_block(n_clips, world, rank)
        chain = pk.Chain(3, 2, 44100, GAINS, n_fft=1024, dtype=np.float32)
        waves = []
        for w in range(3):                                            # three waves, seeds 100 + w
            x = torch.empty((c1 - c0, n_in), dtype=torch.float32, device=dev)
            pk.generate_uniform(x, 100 + w, -0.5, 0.5, first_channel=c0)
            waves.append(x)
        g = SpectraGather(n_clips)
        mags = [None, None]
        fulls = []
        for w, x in enumerate(waves):
            _, _, mags[w % 2] = chain.run(x)                          # wave w; the gather of wave w-1 is in flight
            if w > 0:
                fulls.append(g.wait().clone())
            g.start(mags[w % 2])
        fulls.append(g.wait().clone())
        # the blocking form on the same data
        blocking = gather_spectra(mags[(len(waves) - 1) % 2], n_clips)
        torch.cuda.synchronize()
        assert torch.equal(blocking, fulls[-1])
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), torch.stack(fulls).cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7])
def test_gather_spectra_world2_nccl_side_stream(tmp_path, n_clips):
    import torch
    import torch.multiprocessing as mp

    if not torch.cuda.is_available():
        pytest.fail("a CUDA device is required")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import dsp_audio_project_b200 as pk

    n_in = 6000
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_clips, n_in, str(tmp_path)), nprocs=2, join=True)
    # one GPU, all clips: what every rank must hold after the gather
    chain = pk.Chain(3, 2, 44100, GAINS, n_fft=1024, dtype=np.float32)
    want = []
    for w in range(3):
        x = torch.empty((n_clips, n_in), dtype=torch.float32, device="cuda:0")
        pk.generate_uniform(x, 100 + w, -0.5, 0.5)
        want.append(chain.run(x)[2].cpu().numpy())
    want = np.stack(want)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == want.shape and np.array_equal(got, want), r
