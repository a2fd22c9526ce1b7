"""fp32 (tcgen05 or, with PDDP_FORCE_SIMT=1, SIMT) against fp64 on the same synthetic problems:
error quantiles per output.  usage: python tools/tc_stats.py [B] [N_cartpole] [N_double]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 37
NS = (int(sys.argv[2]) if len(sys.argv) > 2 else 30, int(sys.argv[3]) if len(sys.argv) > 3 else 12)
from pddp_b200.solver import BatchedSolver, BNNDynamics, QRCostConstants
for workload, N in (("cartpole_bnn_b4096", NS[0]), ("double_cartpole_bnn_fullcov_b1024", NS[1])):
    w = dict(bench.WORKLOADS[workload], B=B, N=N)
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
        out[dtype] = {n: s.matrices(n).double().cpu() for n in ("Z", "F_z", "F_u", "k", "K", "Z_new", "U_new")}
        out[dtype]["J"] = s.J_all.double().cpu()
    for n in out[torch.float64]:
        a, bb = out[torch.float32][n], out[torch.float64][n]
        e = (a - bb).abs().flatten(); sc = bb.abs().max()
        q = torch.quantile(e[torch.randperm(e.numel())[:1000000]], torch.tensor([0.5, 0.99, 0.9999], dtype=torch.float64))
        print("%s %-4s scale %.2e  med %.1e  p99 %.1e  p99.99 %.1e  max %.1e" % (workload[:8], n, sc, q[0], q[1], q[2], e.max()))
