#!/usr/bin/env python
"""Benchmark of the PDDP/iLQR iteration hot path (BASELINE.json metric: trajectory-steps/s).

A "step" is ONE pass of the hot path over the whole batch of problems:
    linearise (dynamics + cost derivatives) -> backward Riccati -> rollout of all line-search
    alphas + trajectory cost -> per-problem accept/reject.
One trajectory-step = one time step of one problem through one such pass, so
    value = n_gpus * B * N * steps / seconds.

Workload (N=1 and per rank for N>1, weak scaling): BASELINE.json configs[1] --
"Cartpole PDDP with BNN dynamics (MC-dropout, 50 particles), horizon 100, batch 4096 initial
states": UT-Cholesky encoding (nz=14), MLP 6->200->200->8, 10 line-search alphas, u in [-10, 10].
Synthetic data: random-init BNN weights (reference initialiser, fc_out x0.02 so 100-step rollouts
stay finite -- SURVEY.md 6), CDropout eval masks and standardised eps_in[0] drawn once.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, no collective in
                                                                 the iteration, final NCCL gather)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (geo name, encoding, B, N, P, A, hidden, u bound)
    "cartpole_bnn_b4096": dict(problem="cartpole", enc=1, B=4096, N=100, P=50, A=10, hidden=200, umax=10.0),
    "double_cartpole_bnn_fullcov_b1024": dict(problem="double_cartpole", enc=0, B=1024, N=100, P=50, A=10,
                                              hidden=200, umax=20.0),
    "cartpole_bnn_mpc_b8192": dict(problem="cartpole", enc=1, B=8192, N=50, P=50, A=16, hidden=200, umax=10.0),
    "pendulum_known_b1m": dict(problem="pendulum", enc=4, B=1 << 20, N=100, P=0, A=10, hidden=0, umax=2.5),
    "cartpole_bnn_small": dict(problem="cartpole", enc=1, B=64, N=10, P=50, A=10, hidden=200, umax=10.0),
    # action_size 4 (Jacobi eigen-clipping + 4-dimensional box QP in the backward pass)
    "rendezvous_known_b64k": dict(problem="rendezvous", enc=4, B=1 << 16, N=100, P=0, A=10, hidden=0, umax=1.0),
}
GEOMETRY = {"pendulum": (0, 2, (0,), (1,)), "cartpole": (1, 4, (2,), (0, 1, 3)),
            "double_cartpole": (2, 6, (2, 4), (0, 1, 3, 5)), "rendezvous": (3, 8, (), tuple(range(8)))}
KNOWN_PARAMS = {"pendulum": [0.1, 1.0, 1.0, 0.1, 9.80665], "rendezvous": [0.1, 1.0, 0.1]}
ACTION_SIZE = {"rendezvous": 4}


def cost_constants(problem, dtype=torch.float64):
    """Q, R, Q_term, x_goal of the reference's example costs (pddp/examples/*/cost.py)."""
    if problem == "pendulum":
        Q = torch.tensor([[1, .5, 0], [.5, .25, 0], [0, 0, .25]], dtype=dtype)
        return Q, 0.1 * torch.eye(1, dtype=dtype), 100 * torch.eye(3, dtype=dtype), torch.tensor(
            [0.0, math.sin(math.pi), math.cos(math.pi)], dtype=dtype)
    if problem == "cartpole":
        Q = torch.zeros(5, 5, dtype=dtype)
        Q[0, 0] = 1
        Q[0, 3] = Q[3, 0] = .5
        Q[3, 3] = Q[4, 4] = .25
        return Q, 0.1 * torch.eye(1, dtype=dtype), torch.eye(5, dtype=dtype), torch.tensor(
            [0, 0, 0, math.sin(math.pi), math.cos(math.pi)], dtype=dtype)
    if problem == "rendezvous":            # pddp/examples/rendezvous/cost.py:29-43
        Q = torch.eye(8, dtype=dtype)
        Q[0, 2] = Q[2, 0] = Q[1, 3] = Q[3, 1] = -1
        return Q, 0.1 * torch.eye(4, dtype=dtype), Q.clone(), torch.zeros(8, dtype=dtype)
    C = torch.tensor([[1, -.6, 0, -.6, 0], [0, 0, .6, 0, .6]], dtype=dtype)
    Q = torch.zeros(8, 8, dtype=dtype)
    dims = [0, 4, 5, 6, 7]
    Q[torch.tensor(dims)[:, None], torch.tensor(dims)[None, :]] = C.T @ C
    goal = torch.tensor([0, 0, 0, 0, 0, 1, 0, 1.0], dtype=dtype)
    return Q, 0.1 * torch.eye(1, dtype=dtype), 100 * torch.eye(8, dtype=dtype), goal


def synth_bnn(problem, P, H, seed):
    """Random-init network exactly as pddp/models/bnn/modules.py:797-801,838-849 builds it
    (Xavier-normal with ReLU gain, biases U(-0.1,0.1)); CDropout eval masks (modules.py:540-548)
    and standardised eps_in[0] (modules.py:321-329)."""
    g = torch.Generator().manual_seed(seed)
    _, D, ang, _ = GEOMETRY[problem]
    dims = [D + len(ang) + 1, H, H, 2 * D]
    W, b = [], []
    for din, dout in zip(dims[:-1], dims[1:]):
        std = math.sqrt(2.0) * math.sqrt(2.0 / (din + dout))
        W.append(torch.randn(dout, din, generator=g) * std)
        b.append(torch.rand(dout, generator=g) * 0.2 - 0.1)
    W[-1] *= 0.02
    b[-1] *= 0.02
    masks = []
    for _ in range(2):
        r = torch.rand(P, H, generator=g).clamp(1e-6, 1 - 1e-6)
        masks.append(torch.sigmoid((r.log() - (1 - r).log()) / 0.1))
    eps = torch.randn(P, D, generator=g)
    eps0 = (eps - eps.mean(0)) / eps.std(0)
    return W, b, masks, eps0


def synth_inputs(w, seed, dtype):
    """z0: env reset distribution (mean 1e-2 randn around the start state, variance 1e-2 per dim,
    gym_env.py:75-85) in the workload's encoding; U = 0.1 randn (examples/animation.py:27)."""
    g = torch.Generator().manual_seed(seed)
    _, D, _, _ = GEOMETRY[w["problem"]]
    start = torch.zeros(D)
    if w["problem"] == "double_cartpole":
        start[2] = start[4] = math.pi
    if w["problem"] == "rendezvous":       # pddp/examples/rendezvous/env.py:106-108
        start = torch.tensor([-10.0, -10.0, 10.0, 10.0, 0.0, -5.0, 5.0, 0.0])
    mean = start + 1e-2 * torch.randn(w["B"], D, generator=g)
    if w["enc"] == 4:
        z0 = mean
    elif w["enc"] == 1:
        iu = torch.triu_indices(D, D)
        z0 = torch.cat([mean, (0.1 * torch.eye(D))[iu[0], iu[1]].expand(w["B"], -1)], -1)
    else:
        z0 = torch.cat([mean, (1e-2 * torch.eye(D)).reshape(-1).expand(w["B"], -1)], -1)
    U = 0.1 * torch.randn(w["B"], w["N"], ACTION_SIZE.get(w["problem"], 1), generator=g)
    return z0.to(dtype).contiguous(), U.to(dtype).contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], None, set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v == "Active":
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
def oracle_problem(w, dtype, n_problems, seed):
    """The same synthetic workload expressed for the CPU oracle (checker / cpu baseline only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pddp_oracle as O
    geo, D, ang, nonang = GEOMETRY[w["problem"]]
    Q, R, Qt, goal = cost_constants(w["problem"], dtype)
    cost = O.QRCostSpec(Q, R, Qt, goal, torch.zeros(ACTION_SIZE.get(w["problem"], 1), dtype=dtype), D, ang, nonang)
    if w["P"]:
        W, b, masks, eps0 = synth_bnn(w["problem"], w["P"], w["hidden"], seed)
        dyn = O.BNNSpec(list(zip(W, b)), masks, eps0, D, 1, ang, nonang).to(dtype)
    else:
        dyn = {"pendulum": O.pendulum_spec, "rendezvous": O.rendezvous_spec}[w["problem"]](*KNOWN_PARAMS[w["problem"]])
    z0, U = synth_inputs(dict(w, B=max(n_problems, 1)), seed + 1, dtype)
    return O, dyn, cost, z0, U


def time_oracle(w, dtype, n_problems, seed=0):
    """Seconds for `n_problems` sequential problem-iterations of the reference algorithm on the
    host cores (the reference optimises one problem at a time, SURVEY.md section 2 note)."""
    O, dyn, cost, z0, U = oracle_problem(w, dtype, n_problems, seed)
    nu = ACTION_SIZE.get(w["problem"], 1)
    lo, hi = torch.full((nu,), -w["umax"], dtype=dtype), torch.full((nu,), w["umax"], dtype=dtype)
    alphas = O.fit_alphas(dtype, w["A"])
    t0 = time.perf_counter()
    for i in range(n_problems):
        lin = O.linearize(z0[i], U[i], dyn, cost, w["enc"], lo, hi)
        try:
            k, K = O.backward_pass(*lin, reg=1.0, u_min=lo, u_max=hi, U=U[i])
            Zb, Ub = O.rollout(dyn, lin[0], U[i], k, K, alphas, w["enc"], lo, hi)
            O.trajectory_cost(cost, Zb, Ub, w["enc"]).argmin()
        except RuntimeError:
            pass
    return time.perf_counter() - t0


def run_reference(args, w, rank):
    """--impl reference: the reference's algorithm on the host CPU (oracle port; the reference is
    pure Python and cannot travel to the GPU box)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dtype = torch.float32
    sample = 1
    for _ in range(args.warmup and 1):
        time_oracle(dict(w, N=min(w["N"], 5)), dtype, 1)
    times = [time_oracle(w, dtype, sample, seed=i) for i in range(args.steps)]
    sec = sum(times)
    value = sample * w["N"] * args.steps / sec
    line = {"impl": "reference", "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, w),
            "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
                             "sample": "%d problem(s) x 1 iteration per step (the reference handles one problem "
                                       "at a time; whole-batch time = B x this)" % sample},
            "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(name, w):
    nz = {4: GEOMETRY[w["problem"]][1], 1: (3 * GEOMETRY[w["problem"]][1] + GEOMETRY[w["problem"]][1] ** 2) // 2,
          0: GEOMETRY[w["problem"]][1] * (1 + GEOMETRY[w["problem"]][1])}[w["enc"]]
    return {"workload": name, "problems_per_gpu": w["B"], "horizon": w["N"], "particles": w["P"],
            "alphas": w["A"], "nz": nz, "hidden": [w["hidden"]] * 2 if w["hidden"] else None,
            "encoding": {0: "FULL_COVARIANCE_MATRIX", 1: "UPPER_TRIANGULAR_CHOLESKY", 4: "IGNORE_UNCERTAINTY"}[w["enc"]],
            "bounded": True, "mu": 1.0, "l2": "inputs_exceed_l2 (per-step working set >> 126 MB)",
            "parallelism": "independent problems sharded per GPU, no collective in the iteration"}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cartpole_bnn_b4096", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from pddp_b200 import _lib
    from pddp_b200.solver import BatchedSolver, BNNDynamics, KnownDynamics, QRCostConstants

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    lib = _lib.load()
    geo, D, ang, nonang = GEOMETRY[w["problem"]]
    Q, R, Qt, goal = cost_constants(w["problem"])
    cost = QRCostConstants(Q, R, Qt, goal)
    if w["P"]:
        W, b, masks, eps0 = synth_bnn(w["problem"], w["P"], w["hidden"], seed=0)
        dyn = BNNDynamics(geo, W, b, masks, eps0)
    else:
        dyn = KnownDynamics(geo, KNOWN_PARAMS[w["problem"]])
    solver = BatchedSolver(dyn, cost, w["enc"], w["B"], w["N"], dtype=dtype, device=dev, max_alphas=w["A"])
    z0_h, U_h = synth_inputs(w, seed=1 + rank, dtype=dtype)
    z0_h, U_h = z0_h.pin_memory(), U_h.pin_memory()
    nu = ACTION_SIZE.get(w["problem"], 1)
    lo, hi = [-w["umax"]] * nu, [w["umax"]] * nu
    alphas = (1.025 ** (-torch.arange(float(w["A"]), dtype=torch.float64) ** 2)).to(dtype)
    solver.set_problem(z0_h.to(dev), U_h.to(dev), lo, hi, alphas=alphas, iterations=1 << 30)

    def step():
        # every problem stays in play with a fixed regularisation so the work per step is constant
        solver.active.fill_(1)
        solver.mu.fill_(1.0)
        solver.iterate()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()
    not_pd = int((solver.bw_status != 0).sum().item())
    launches0 = int(lib.pddp_launch_count())
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = int(lib.pddp_launch_count()) - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    not_pd = max(not_pd, int((solver.bw_status != 0).sum().item()))
    total_steps = world * w["B"] * w["N"] * args.steps
    value = total_steps / (ms * 1e-3)

    # ---- end to end: host buffers in, host buffers out, every step ------------------------------
    Z_h = torch.empty(solver.view("Z").shape, dtype=dtype).pin_memory()
    Uo_h = torch.empty(solver.view("U").shape, dtype=dtype).pin_memory()
    J_h = torch.empty(w["B"], dtype=dtype).pin_memory()
    st_h = torch.empty(w["B"], dtype=torch.int32).pin_memory()
    h2d = z0_h.numel() * z0_h.element_size() + U_h.numel() * U_h.element_size()
    d2h = sum(t.numel() * t.element_size() for t in (Z_h, Uo_h, J_h, st_h))

    def e2e_step():
        solver.set_problem(z0_h.to(dev, non_blocking=True), U_h.to(dev, non_blocking=True), lo, hi,
                           alphas=None, iterations=1)
        solver.mu.fill_(1.0)
        solver.iterate()
        Z_h.copy_(solver.view("Z"), non_blocking=True)
        Uo_h.copy_(solver.view("U"), non_blocking=True)
        J_h.copy_(solver.J_opt, non_blocking=True)
        st_h.copy_(solver.state, non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    sync_all()
    ms_e2e = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev,
                          dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = total_steps / (float(ms_e2e.item()) * 1e-3)

    # ---- dominant kernel: per-launch duration from CUDA events on the launch stream ---------------
    roofline = None
    hbm_peak, tensor_peak, peak_src = measured_peaks()
    if w["P"]:
        lib.pddp_profile_enable(1)
        step()
        import ctypes as C
        pms, pcnt = (C.c_double * 4)(), (C.c_int64 * 4)()
        _lib.check(lib.pddp_profile_read(pms, pcnt), "profile_read")
        lib.pddp_profile_enable(0)
        macs_per_row = sum(a * b_ for a, b_ in zip([D + len(ang) + 1, w["hidden"], w["hidden"]],
                                                  [w["hidden"], w["hidden"], 2 * D]))
        rows = {0: w["B"] * w["P"] * (1 + D + 1), 1: w["B"] * w["A"] * w["P"]}
        kinds = {}
        for kind, name in ((0, "mlp_linearise"), (1, "mlp_rollout"), (2, "moment_linearise"), (3, "rollout_step")):
            if pcnt[kind]:
                kinds[name] = {"launches_per_step": int(pcnt[kind]), "avg_ms": pms[kind] / pcnt[kind],
                               "ms_per_step": pms[kind]}
        dom = 1 if pms[1] >= pms[0] else 0
        flops = 2.0 * macs_per_row * rows[dom]
        achieved = flops / (pms[dom] / pcnt[dom] * 1e-3) / 1e12
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        # (profiles/r1_summary.md, default workload only; null for the others)
        traffic = {1: 36.7e6 + 3.0e6, 0: 4.6e6 + 0.2e6}[dom] if args.workload == "cartpole_bnn_b4096" else None
        roofline = {"bound": "tensor", "kernel": ["bnn_mlp (linearise rows)", "bnn_mlp (rollout rows)"][dom],
                    "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                    "traffic": traffic, "peak_source": "bf16 dense sustained, " + peak_src,
                    "tensor_passes_per_product": 3,
                    "frac_of_fp32_accurate_bound": 3.0 * achieved / tensor_peak,
                    "note": "fp32-accurate products on the tensor core cost 3 FP16 passes (a0*b0 + a0*b1 + a1*b0), "
                            "so `frac` of the 16-bit dense peak cannot exceed 1/3; tensor pipe 63 % active, shared-memory "
                            "data pipe 98 %, and the tile's layer-0 -> mid-stage -> layer-1 -> epilogue dependency chain "
                            "plus the 1 kW power cap set the time (profiles/r1_summary.md section 5)",
                    "flops_per_launch": flops, "avg_launch_ms": pms[dom] / pcnt[dom],
                    "mlp_share_of_step": (pms[0] + pms[1]) / (ms / args.steps), "kernels": kinds}
    else:
        nz, nu = solver.nz, solver.nu      # linearise writes + backward reads/writes + rollout reads/writes
        elems = (2 * nz * nz + 2 * nz * nu + 2 * nz + nu + nu * nu + 1 + nu) + (
            2 * nz * nz + 2 * nz * nu + nz + nu + nu * nu + nu + nu * nz) + (nz + 2 * nu + nu * nz) + (nz + nu)
        bytes_per_step = elems * (4 if dtype == torch.float32 else 8) * w["B"] * w["N"]
        achieved = bytes_per_step / (ms / args.steps * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole pass (linearise+backward+rollout+accept)", "achieved": achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                    "peak_source": peak_src}

    # ---- final gather of the results over NCCL (outside the timed iteration) ---------------------
    gather_ms = None
    if world > 1:
        out = [torch.empty_like(solver.view("U").contiguous()) for _ in range(world)] if rank == 0 else None
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        dist.gather(solver.view("U").contiguous(), out, dst=0)
        torch.cuda.synchronize(dev)
        gather_ms = (time.perf_counter() - t0) * 1e3

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sec = time_oracle(w, torch.float32, 1)
        cpu = {"value": w["N"] / sec, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
               "sample": "1 problem x 1 iteration of the same workload (%.1f s); the reference handles one "
                         "problem at a time" % sec}
    if rank == 0:
        line = {"metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": workload_config(args.workload, w), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "not_pd_problems": not_pd, "final_gather_ms": gather_ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
