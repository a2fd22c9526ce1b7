"""The tcgen05 (3xTF32) particle-MLP kernel against the fp64 SIMT path on the same synthetic
problem: cartpole (UT-Cholesky, 1+5 rows per particle) and double cartpole (full covariance,
1+7 rows per particle), H = 200, P = 50, longer horizons than the golden fixtures.

fp32 tolerance from the north star: 1e-3 relative (to each tensor's own scale here).

ReLU kinks: a hidden unit whose pre-activation is within fp32 rounding of zero can switch on in one
arithmetic and off in another; its tangent then changes by O(1) for that single (problem, step,
particle).  This happens between ANY two fp32 implementations (the SIMT fp32 kernel shows the same
isolated outliers against fp64, tools/tc_stats.py) and it contaminates that one problem's gains and
candidate controls.  So derivative-like outputs are held to 1e-3 for at least 85 % of the problems
(median problem: 1e-4) and to 5e-2 for the worst one; values, costs and trajectories are held to
1e-3 everywhere."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def scale_err(a, b, per_problem=False):
    a, b = a.double().cpu(), b.double().cpu()
    e = (a - b).abs() / b.abs().max().clamp_min(1e-30)
    return e.reshape(e.shape[0], -1).max(1).values if per_problem else e.max().item()


EXACT = ("Z", "L", "L_z", "L_zz", "Z_new", "J")          # values / costs / trajectories
KINKY = ("F_z", "F_u", "k", "K", "U_new")                # carry ReLU-derivative outliers


@pytest.mark.parametrize("workload,N", [("cartpole_bnn_b4096", 30), ("double_cartpole_bnn_fullcov_b1024", 12)])
def test_tc_matches_fp64(workload, N):
    import bench
    from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
    w = dict(bench.WORKLOADS[workload], B=37, N=N)
    geo, D, ang, nonang = bench.GEOMETRY[w["problem"]]
    W, b, masks, eps0 = bench.synth_bnn(w["problem"], w["P"], w["hidden"], seed=3)
    cost = QRCostConstants(*bench.cost_constants(w["problem"]))
    z0, U = bench.synth_inputs(w, seed=4, dtype=torch.float64)
    out = {}
    for dtype in (torch.float64, torch.float32):
        s = BatchedSolver(BNNDynamics(geo, W, b, masks, eps0), cost, w["enc"], w["B"], w["N"], dtype=dtype)
        s.set_problem(z0.to(dtype).cuda(), U.to(dtype).cuda(), [-w["umax"]], [w["umax"]])
        s.mu.fill_(1.0)
        s.linearize(); s.backward(); s.rollout()
        torch.cuda.synchronize()
        assert s.lin_status.cpu().abs().sum() == 0 and s.bw_status.cpu().abs().sum() == 0
        out[dtype] = {n: s.matrices(n).clone() for n in ("Z", "F_z", "F_u", "L", "L_z", "L_zz", "k", "K", "Z_new", "U_new")}
        out[dtype]["J"] = s.J_all.clone()
    errs = {n: scale_err(out[torch.float32][n], out[torch.float64][n]) for n in out[torch.float64]}
    print(workload, {k: "%.1e" % v for k, v in errs.items()})
    for n in EXACT:
        assert errs[n] < 1e-3, (n, errs[n])
    for n in KINKY:
        per = scale_err(out[torch.float32][n], out[torch.float64][n], per_problem=True)
        assert per.median() < 1e-4, (n, per.median())
        assert (per < 1e-3).float().mean() >= 0.85, (n, per)
        assert errs[n] < 5e-2, (n, errs[n])
