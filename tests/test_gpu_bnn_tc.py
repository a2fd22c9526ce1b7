"""The tcgen05 particle-MLP kernel (bnn_mlp_tc.cuh: TF32 x TF32 + one BF16 cross-term pass) against
the fp64 SIMT path on the same synthetic problem: cartpole (UT-Cholesky, 1+5 passes per super-tile)
and double cartpole (full covariance, 1+7 passes), H = 200, P = 50, longer horizons than the
golden fixtures; plus batch sizes that give every CTA several tiles on both tracks and particle
changes inside a CTA's range, and hidden widths that are not 200 (other K-block counts).

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


def _run_both(w, W, b, masks, eps0, N, seed=4):
    import bench
    from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
    geo = bench.GEOMETRY[w["problem"]][0]
    cost = QRCostConstants(*bench.cost_constants(w["problem"]))
    z0, U = bench.synth_inputs(w, seed=seed, dtype=torch.float64)
    out = {}
    for dtype in (torch.float64, torch.float32):
        s = BatchedSolver(BNNDynamics(geo, W, b, masks, eps0), cost, w["enc"], w["B"], N, dtype=dtype)
        s.set_problem(z0.to(dtype).cuda(), U.to(dtype).cuda(), [-w["umax"]], [w["umax"]])
        s.mu.fill_(1.0)
        s.linearize(); s.backward(); s.rollout()
        torch.cuda.synchronize()
        assert s.lin_status.cpu().abs().sum() == 0 and s.bw_status.cpu().abs().sum() == 0
        out[dtype] = {n: s.matrices(n).clone() for n in ("Z", "F_z", "F_u", "L", "k", "K", "Z_new", "U_new")}
        out[dtype]["J"] = s.J_all.clone()
    return out


@pytest.mark.parametrize("B", [300, 1111])
def test_tc_many_tiles_per_cta(B):
    """B = 300: 3 super-tiles per particle (the last one ragged); B = 1111: every CTA owns tiles of
    two particles and both tracks run several tiles (rollout: 87 tiles per particle)."""
    import bench
    w = dict(bench.WORKLOADS["cartpole_bnn_b4096"], B=B, N=3)
    W, b, masks, eps0 = bench.synth_bnn(w["problem"], w["P"], w["hidden"], seed=5)
    out = _run_both(w, W, b, masks, eps0, 3)
    for n in out[torch.float64]:
        per = scale_err(out[torch.float32][n], out[torch.float64][n], per_problem=True)
        assert per.median() < 1e-5, (n, per.median())
        assert (per < 1e-3).float().mean() >= 0.97, (n, per.max())


@pytest.mark.parametrize("H0,H1", [(96, 144), (207, 72)])
def test_tc_other_hidden_widths(H0, H1):
    """Hidden widths other than 200: 7 / 13 K-blocks of 16 (H0 + 1 bias column), N < 208."""
    import bench
    w = dict(bench.WORKLOADS["cartpole_bnn_b4096"], B=150, N=3)
    g = torch.Generator().manual_seed(11)
    D, P = 4, w["P"]
    dims = [(6, H0), (H0, H1), (H1, 2 * D)]
    W = [torch.randn(o, i, generator=g) * math.sqrt(2.0 / (i + o)) * math.sqrt(2.0) for i, o in dims]
    W[-1] *= 0.02
    b = [torch.rand(o, generator=g) * 0.2 - 0.1 for _, o in dims]
    b[-1] *= 0.02
    masks = []
    for h in (H0, H1):
        r = torch.rand(P, h, generator=g).clamp(1e-6, 1 - 1e-6)
        masks.append(torch.sigmoid((r.log() - (1 - r).log()) / 0.1))
    eps = torch.randn(P, D, generator=g)
    eps0 = (eps - eps.mean(0)) / eps.std(0)
    out = _run_both(w, W, b, masks, eps0, 3)
    for n in out[torch.float64]:
        per = scale_err(out[torch.float32][n], out[torch.float64][n], per_problem=True)
        assert per.median() < 1e-5, (n, per.median())
        assert (per < 1e-3).float().mean() >= 0.97, (n, per.max())
