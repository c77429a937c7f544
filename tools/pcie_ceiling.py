"""Copy-only ceiling of the host-buffer chain (bench.py's `e2e`): pinned host <-> device copies of the
chain's per-call byte counts with no kernels at all, per rank and in aggregate.

  python tools/pcie_ceiling.py                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29511 tools/pcie_ceiling.py   # N ranks, one per GPU, copying at the same time

Prints one JSON line (rank 0): GB/s for H2D alone, D2H alone, and both directions at once on two streams
(the pattern dspb200_chain_host_* produces), each as the slowest rank's rate and the sum over ranks; and
`chain_floor_ms`: the time the both-directions pattern needs for one 1024-clip call (1.806 GB in, 2.925 GB
out), i.e. the floor under `e2e` whatever the kernels do."""
from __future__ import annotations

import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        from dsp_audio_project_b200 import shard
        shard.pin_to_gpu_numa(local)
    except Exception:
        pass
    clips, n_in, n_out, frames, bins = 1024, 441000, 480000, 117, 2049
    in_bytes = clips * n_in * 4
    out_bytes = clips * (n_out + frames * bins) * 4
    hx = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    hz = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    dx = torch.empty(in_bytes, dtype=torch.uint8, device="cuda")
    dz = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    pieces = 32      # the chain copies slab by slab; same granularity here

    def run(h2d: bool, d2h: bool, reps: int = 3) -> float:
        best = 1e30
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for p in range(pieces):
                if h2d:
                    a, b = p * in_bytes // pieces, (p + 1) * in_bytes // pieces
                    with torch.cuda.stream(s1):
                        dx[a:b].copy_(hx[a:b], non_blocking=True)
                if d2h:
                    a, b = p * out_bytes // pieces, (p + 1) * out_bytes // pieces
                    with torch.cuda.stream(s2):
                        hz[a:b].copy_(dz[a:b], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = min(best, dt)
        return best

    t_in = run(True, False)
    t_out = run(False, True)
    t_both = run(True, True)
    if rank == 0:
        gb = 1e-9
        print(json.dumps({
            "n_gpus": world, "pinned_bytes_in": in_bytes, "pinned_bytes_out": out_bytes,
            "h2d_gbs_per_rank": in_bytes * gb / t_in, "h2d_gbs_total": world * in_bytes * gb / t_in,
            "d2h_gbs_per_rank": out_bytes * gb / t_out, "d2h_gbs_total": world * out_bytes * gb / t_out,
            "both_gbs_per_rank": (in_bytes + out_bytes) * gb / t_both,
            "both_gbs_total": world * (in_bytes + out_bytes) * gb / t_both,
            "chain_floor_ms": t_both * 1e3,
            "chain_floor_gsamples_s": world * clips * n_in / t_both * 1e-9,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
