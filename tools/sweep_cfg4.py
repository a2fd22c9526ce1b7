"""BASELINE config 4: pendulum iLQR batch sweep, B in {1k, 10k, 100k, 1M} problems in TOTAL, split across the ranks of
the job (strong scaling; one process per GPU, no collective inside the iteration):
    python tools/sweep_cfg4.py                      (1 GPU)
    torchrun --nproc-per-node N tools/sweep_cfg4.py (N GPUs)
Rank 0 prints one JSON line per batch size: ms per pass (max over ranks, CUDA events), trajectory-steps/s, the
backward kernel's time and achieved HBM bandwidth on rank 0 (the "backward-pass bandwidth study")."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch.distributed as dist
    from pddp_b200 import _lib, sharding
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, _, src = bench.measured_peaks()
    for total in (1000, 10000, 100000, 1000000):
        lo, hi = sharding.shard_bounds(total, world, rank)
        rec = bench.measure("pendulum_known_b1m", torch.float32, 10, 5, dev, rank, world, dist, B=hi - lo)
        if rank == 0:
            k = rec["kernels"]
            nbytes = bench.stage_elements(2, 1)["backward"] * 4 * (hi - lo) * 100
            bw_ms = k["backward"][0]
            print(json.dumps({"workload": "pendulum_known (cfg 4)", "problems_total": total, "n_gpus": world,
                              "problems_per_gpu": hi - lo, "ms_per_pass": rec["ms_per_step"],
                              "trajectory_steps_per_s": total * 100 / (rec["ms_per_step"] * 1e-3),
                              "backward_ms": bw_ms, "backward_gbs": nbytes / (bw_ms * 1e-3) / 1e9,
                              "backward_frac_of_hbm": nbytes / (bw_ms * 1e-3) / 1e9 / hbm_peak, "hbm_peak_gbs": hbm_peak,
                              "peak_source": src, "kernels_ms": {n: v[0] for n, v in k.items()},
                              "sm_mhz": rec["clocks"]["sm_mhz"] if rec["clocks"] else None}), flush=True)
        del rec
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
