"""Edge cases and size-independent properties of the CUDA path (through the C ABI).

* degenerate shapes against the CPU oracle on the same seeded inputs: horizon 1, a single
  line-search alpha, one problem, a particle count that is not a multiple of anything (7), a
  ragged last tile;
* inactive problems are left untouched;
* at BASELINE.json's full batch sizes (4096 BNN problems / 2^17 known-dynamics problems): a batch
  made of ONE problem repeated must give the same row wherever the problem sits in the batch (tile,
  track, CTA, warp and lane placement must not change a result: bit-identical on the known-dynamics
  path; to fp32 rounding on the BNN path, where the two tile tracks of the MLP kernel sum the
  K-blocks in a different cyclic order), and the line-search outputs must be self-consistent (J_new = J_all[amin], a re-rollout with the winning alpha alone
  reproduces Z_new / U_new)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

F64 = torch.float64


def _cartpole_cost(O):
    Q = torch.zeros(5, 5, dtype=F64)
    Q[0, 0] = 1
    Q[0, 3] = Q[3, 0] = .5
    Q[3, 3] = Q[4, 4] = .25
    return O.QRCostSpec(Q, 0.1 * torch.eye(1, dtype=F64), torch.eye(5, dtype=F64),
                        torch.tensor([0, 0, 0, 0.0, -1.0], dtype=F64), torch.zeros(1, dtype=F64), 4, (2,), (0, 1, 3))


def _pendulum_cost(O):
    return O.QRCostSpec(torch.tensor([[1, .5, 0], [.5, .25, 0], [0, 0, .25]], dtype=F64), 0.1 * torch.eye(1, dtype=F64),
                        100 * torch.eye(3, dtype=F64), torch.tensor([0.0, 0.0, -1.0], dtype=F64),
                        torch.zeros(1, dtype=F64), 2, (0,), (1,))


def _small_bnn(O, P, H, seed):
    g = torch.Generator().manual_seed(seed)
    D = 4
    W = [torch.randn(H, 6, generator=g, dtype=F64) * 0.5, torch.randn(H, H, generator=g, dtype=F64) * 0.25,
         torch.randn(2 * D, H, generator=g, dtype=F64) * 0.02]
    b = [0.1 * torch.randn(H, generator=g, dtype=F64), 0.1 * torch.randn(H, generator=g, dtype=F64),
         0.01 * torch.randn(2 * D, generator=g, dtype=F64)]
    masks = [torch.rand(P, H, generator=g, dtype=F64), torch.rand(P, H, generator=g, dtype=F64)]
    eps = torch.randn(P, D, generator=g, dtype=F64)
    eps0 = (eps - eps.mean(0)) / eps.std(0)
    return W, b, masks, eps0


def _check_against_oracle(O, solver, odyn, ocost, enc, z0, U, alphas, reg, tol=1e-6):
    B = z0.shape[0]
    solver.set_problem(z0.to(solver.dtype).cuda(), U.to(solver.dtype).cuda(), alphas=alphas.to(solver.dtype))
    solver.mu.fill_(reg)
    solver.linearize(); solver.backward(); solver.rollout()
    torch.cuda.synchronize()
    assert solver.lin_status.cpu().abs().sum() == 0 and solver.bw_status.cpu().abs().sum() == 0
    worst = 0.0
    for b in range(B):
        lin = O.linearize(z0[b], U[b], odyn, ocost, enc)
        k, K = O.backward_pass(*lin, reg=reg)
        Zb, Ub = O.rollout(odyn, lin[0], U[b], k, K, alphas, enc)
        J = O.trajectory_cost(ocost, Zb, Ub, enc)
        pairs = [(solver.matrices(n).cpu()[b], w) for n, w in zip("Z F_z F_u L L_z L_u L_zz L_uz L_uu".split(), lin)]
        # the kernel's winner must be the oracle's minimum -- up to the tolerance: in fp32 two step sizes whose costs
        # agree to rounding may swap
        amin = int(solver.amin.cpu()[b])
        assert float(J[amin]) <= float(J.min()) + tol * max(1.0, float(J.abs().max()))
        if tol <= 1e-6:
            assert amin == int(J.argmin())
        pairs += [(solver.matrices("k").cpu()[b], k), (solver.matrices("K").cpu()[b], K), (solver.J_all.cpu()[b], J),
                  (solver.view("Z_new").cpu()[b], Zb[:, amin]), (solver.view("U_new").cpu()[b], Ub[:, amin])]
        for got, want in pairs:
            err = (got.double().reshape(want.shape) - want).abs().max().item() / max(1.0, want.abs().max().item())
            worst = max(worst, err)
    assert worst < tol, worst


@pytest.mark.parametrize("B,N,A", [(1, 1, 1), (3, 1, 4), (1, 7, 1), (5, 3, 10)])
def test_degenerate_shapes_known(B, N, A):
    import pddp_oracle as O
    from pddp_b200 import _lib
    from pddp_b200.solver import BatchedSolver, KnownDynamics, QRCostConstants
    ocost = _pendulum_cost(O)
    g = torch.Generator().manual_seed(B * 100 + N * 10 + A)
    z0 = 1e-2 * torch.randn(B, 2, generator=g, dtype=F64)
    U = 0.1 * torch.randn(B, N, 1, generator=g, dtype=F64)
    s = BatchedSolver(KnownDynamics(_lib.GEO_PENDULUM, [0.1, 1.0, 1.0, 0.1, 9.80665]),
                      QRCostConstants(ocost.Q, ocost.R, ocost.Q_term, ocost.x_goal), O.IGNORE_UNCERTAINTY, B, N,
                      dtype=F64, max_alphas=max(A, 1))
    _check_against_oracle(O, s, O.pendulum_spec(0.1), ocost, O.IGNORE_UNCERTAINTY, z0, U, O.fit_alphas(F64, A), 50.0)


# (P = 210 in fp64: the particles of 32 (problem, alpha) pairs no longer fit shared memory -> the roll step reads them from
# global memory instead of staging them)
@pytest.mark.parametrize("B,N,A,P", [(1, 1, 1, 7), (2, 2, 3, 7), (3, 1, 10, 1 + 12), (1, 2, 3, 210)])
def test_degenerate_shapes_bnn(B, N, A, P):
    import pddp_oracle as O
    from pddp_b200 import _lib
    from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
    ocost = _cartpole_cost(O)
    W, b, masks, eps0 = _small_bnn(O, P, 32, seed=P + A)
    odyn = O.BNNSpec(list(zip(W, b)), masks, eps0, 4, 1, (2,), (0, 1, 3))
    g = torch.Generator().manual_seed(B * 100 + N * 10 + A)
    mean = 1e-2 * torch.randn(B, 4, generator=g, dtype=F64)
    z0 = torch.stack([O.encode(m, V=1e-2 * torch.ones(4, dtype=F64), enc=O.DEFAULT) for m in mean])
    U = 0.1 * torch.randn(B, N, 1, generator=g, dtype=F64)
    s = BatchedSolver(BNNDynamics(_lib.GEO_CARTPOLE, W, b, masks, eps0),
                      QRCostConstants(ocost.Q, ocost.R, ocost.Q_term, ocost.x_goal), O.DEFAULT, B, N, dtype=F64,
                      max_alphas=max(A, 1))
    _check_against_oracle(O, s, odyn, ocost, O.DEFAULT, z0, U, O.fit_alphas(F64, A), 1.0)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-6), (torch.float32, 1e-3)])
@pytest.mark.parametrize("B,N,A,P", [(3, 2, 5, 7), (1, 3, 3, 5)])
def test_pendulum_bnn_odd_particle_and_pair_counts(B, N, A, P, dtype, tol):
    """Pendulum BNN (D = 2): a (problem, alpha) pair's particle block is 8 P bytes in fp32, so with an odd P and an odd
    number of pairs the roll-step kernel's bulk copy is rounded up to 16 bytes (it reads into the workspace padding), the
    staged rows are 8-byte rows, and the last CTA is ragged.  Against the CPU oracle on the same inputs."""
    import pddp_oracle as O
    from pddp_b200 import _lib
    from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
    ocost = _pendulum_cost(O)
    g = torch.Generator().manual_seed(17 * P + A)
    D, H = 2, 32
    W = [torch.randn(H, 4, generator=g, dtype=F64) * 0.5, torch.randn(H, H, generator=g, dtype=F64) * 0.25,
         torch.randn(2 * D, H, generator=g, dtype=F64) * 0.02]
    b = [0.1 * torch.randn(H, generator=g, dtype=F64), 0.1 * torch.randn(H, generator=g, dtype=F64),
         0.01 * torch.randn(2 * D, generator=g, dtype=F64)]
    masks = [torch.rand(P, H, generator=g, dtype=F64), torch.rand(P, H, generator=g, dtype=F64)]
    eps = torch.randn(P, D, generator=g, dtype=F64)
    eps0 = (eps - eps.mean(0)) / eps.std(0)
    odyn = O.BNNSpec(list(zip(W, b)), masks, eps0, D, 1, (0,), (1,))
    mean = 1e-2 * torch.randn(B, D, generator=g, dtype=F64)
    z0 = torch.stack([O.encode(m, V=1e-2 * torch.ones(D, dtype=F64), enc=O.DEFAULT) for m in mean])
    U = 0.1 * torch.randn(B, N, 1, generator=g, dtype=F64)
    s = BatchedSolver(BNNDynamics(_lib.GEO_PENDULUM, W, b, masks, eps0),
                      QRCostConstants(ocost.Q, ocost.R, ocost.Q_term, ocost.x_goal), O.DEFAULT, B, N, dtype=dtype,
                      max_alphas=A)
    _check_against_oracle(O, s, odyn, ocost, O.DEFAULT, z0, U, O.fit_alphas(F64, A), 1.0, tol=tol)


def test_inactive_problems_are_untouched():
    import bench
    from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
    w = dict(bench.WORKLOADS["cartpole_bnn_b4096"], B=9, N=3)
    W, b, masks, eps0 = bench.synth_bnn(w["problem"], w["P"], w["hidden"], seed=2)
    s = BatchedSolver(BNNDynamics(bench.GEOMETRY[w["problem"]][0], W, b, masks, eps0),
                      QRCostConstants(*bench.cost_constants(w["problem"])), w["enc"], w["B"], w["N"], dtype=torch.float32)
    z0, U = bench.synth_inputs(w, seed=3, dtype=torch.float32)
    s.set_problem(z0.cuda(), U.cuda(), [-10.0], [10.0])
    s.mu.fill_(1.0)
    names = ("Z", "F_z", "F_u", "L", "L_z", "L_zz", "k", "K", "Z_new", "U_new")
    for n in names:
        s.view(n).fill_(-7.0)
    s.active[torch.tensor([1, 4, 8])] = 0
    s.linearize(); s.backward(); s.rollout()
    torch.cuda.synchronize()
    for n in names:
        v = s.view(n).cpu()
        for bb in (1, 4, 8):
            assert (v[bb] == -7.0).all(), (n, bb)
        assert not (v[0] == -7.0).all(), n


def _row_identical(t):
    t = t.reshape(t.shape[0], -1)
    return bool((t == t[0:1]).all())


def _row_spread(t):
    """per row: max |row - row 0| relative to the tensor's scale"""
    t = t.reshape(t.shape[0], -1).double()
    return (t - t[0:1]).abs().max(1).values / t.abs().max().clamp_min(1e-30)


def test_full_batch_placement_invariance_bnn():
    """4096 copies of one cartpole problem (BASELINE cfg 2 batch size, short horizon): every row of
    every output is the same to fp32 rounding, whichever tile / track / CTA / lane computed it."""
    import bench
    from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
    w = dict(bench.WORKLOADS["cartpole_bnn_b4096"], N=2)
    W, b, masks, eps0 = bench.synth_bnn(w["problem"], w["P"], w["hidden"], seed=0)
    s = BatchedSolver(BNNDynamics(bench.GEOMETRY[w["problem"]][0], W, b, masks, eps0),
                      QRCostConstants(*bench.cost_constants(w["problem"])), w["enc"], w["B"], w["N"], dtype=torch.float32)
    z0, U = bench.synth_inputs(dict(w, B=1), seed=5, dtype=torch.float32)
    s.set_problem(z0.expand(w["B"], -1).contiguous().cuda(), U.expand(w["B"], -1, -1).contiguous().cuda(), [-10.0], [10.0])
    s.mu.fill_(1.0)
    s.linearize(); s.backward(); s.rollout()
    torch.cuda.synchronize()
    assert s.lin_status.cpu().abs().sum() == 0 and s.bw_status.cpu().abs().sum() == 0
    # values: every row equals row 0 to fp32 rounding; derivative-like outputs may differ on isolated
    # rows where a ReLU sits within rounding of its kink (tests/test_gpu_bnn_tc.py)
    for n in ("Z", "L", "Z_new"):
        assert _row_spread(s.matrices(n)).max() < 1e-5, n
    assert _row_spread(s.J_all).max() < 1e-5
    for n in ("F_z", "F_u", "L_z", "L_zz", "k", "K", "U_new"):
        spread = _row_spread(s.matrices(n))
        assert spread.median() < 1e-6 and (spread < 1e-4).float().mean() > 0.99, (n, spread.max())
    # line-search self-consistency
    J_all, amin = s.J_all.clone(), s.amin.clone().long()
    assert torch.equal(s.J_new, J_all.gather(1, amin[:, None])[:, 0])
    Z_new, U_new = s.view("Z_new").clone(), s.view("U_new").clone()
    s.alphas = s.alphas[amin[0]:amin[0] + 1].contiguous()
    s.J_all = torch.zeros(w["B"], 1, dtype=torch.float32, device="cuda")
    s.rollout()
    torch.cuda.synchronize()
    # (a different batch shape moves rows between tiles / tracks: equal to fp32 rounding, not bitwise)
    for got, want in ((s.view("Z_new"), Z_new), (s.view("U_new"), U_new)):
        assert ((got - want).abs().max() / want.abs().max()).item() < 1e-5


@pytest.mark.parametrize("layout", [0, 1])
def test_full_batch_placement_invariance_known(layout):
    """2^17 copies of one pendulum problem through the thread-per-problem scans (both layouts)."""
    import pddp_oracle as O
    from pddp_b200 import _lib
    from pddp_b200.solver import BatchedSolver, KnownDynamics, QRCostConstants
    ocost = _pendulum_cost(O)
    B, N = 1 << 17, 20
    g = torch.Generator().manual_seed(1)
    z0 = (1e-2 * torch.randn(1, 2, generator=g)).expand(B, -1).contiguous()
    U = (0.1 * torch.randn(1, N, 1, generator=g)).expand(B, -1, -1).contiguous()
    s = BatchedSolver(KnownDynamics(_lib.GEO_PENDULUM, [0.1, 1.0, 1.0, 0.1, 9.80665]),
                      QRCostConstants(ocost.Q, ocost.R, ocost.Q_term, ocost.x_goal), O.IGNORE_UNCERTAINTY, B, N,
                      dtype=torch.float32, layout=layout)
    s.set_problem(z0.cuda(), U.cuda(), [-2.5], [2.5])
    s.mu.fill_(1.0)
    s.linearize(); s.backward(); s.rollout()
    torch.cuda.synchronize()
    assert s.bw_status.cpu().abs().sum() == 0
    for n in ("Z", "F_z", "L_zz", "k", "K", "Z_new", "U_new"):
        assert _row_identical(s.matrices(n)), n
    assert torch.equal(s.J_new, s.J_all.gather(1, s.amin.long()[:, None])[:, 0])
    # against the oracle in fp64 on that one problem
    lin = O.linearize(z0[0].double(), U[0].double(), O.pendulum_spec(0.1), ocost, O.IGNORE_UNCERTAINTY,
                      torch.tensor([-2.5], dtype=F64), torch.tensor([2.5], dtype=F64))
    k, K = O.backward_pass(*lin, reg=1.0, u_min=torch.tensor([-2.5], dtype=F64), u_max=torch.tensor([2.5], dtype=F64),
                           U=U[0].double())
    assert (s.matrices("K").cpu()[0].double().reshape(K.shape) - K).abs().max() <= 1e-3 * K.abs().max()
    assert (s.matrices("k").cpu()[0].double().reshape(k.shape) - k).abs().max() <= 1e-3 * max(1.0, k.abs().max())
