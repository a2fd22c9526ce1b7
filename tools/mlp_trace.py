"""Timeline experiment for the tcgen05 particle MLP (library built with -DPDDP_EXP_TRACE, loaded through
PDDP_B200_LIB): runs passes of the default workload, then reads CTA 0's clock64() trace of the LAST MLP launch
(the rollout MLP of the last step) and prints per-tile phase durations for both tracks.
  python tools/mlp_trace.py [linearise|rollout] > gpurun_out/trace.txt"""
import ctypes as C
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pddp_b200 import _lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "rollout"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
lib = _lib.load()
w, solver, *_ = bench.build_solver("cartpole_bnn_b4096", torch.float32, dev, 0)
for _ in range(2):
    solver.active.fill_(1); solver.mu.fill_(1.0)
    solver.iterate()
if which == "linearise":
    solver.active.fill_(1); solver.linearize()
torch.cuda.synchronize()
raw = C.CDLL(_lib.LIB_PATH)
n = 8 * 8192
buf = (C.c_longlong * n)()
raw.pddp_debug_trace.argtypes = [C.POINTER(C.c_longlong), C.c_size_t]
rc = raw.pddp_debug_trace(buf, n)
assert rc == 0, rc
tr = np.frombuffer(buf, dtype=np.int64).reshape(8, 8192)
t0 = min(x for x in [tr[0, 0], tr[2, 0], tr[4, 0]] if x > 0)
print("which", which)
for t in range(2):
    print("== track %d: L1 issuer [tile: wait_acc1_empty_start, mma_start, last_issue | sums: wait_b, wait_a, issue]" % t)
    iss = tr[t].reshape(-1, 8)
    epi = tr[2 + t].reshape(-1, 4)
    mid = tr[4 + t].reshape(-1, 4)
    K = int((iss[:, 2] > 0).sum())
    for k in range(min(K, 60)):
        i, e, m = iss[k], epi[k], mid[k]
        print("tile %2d | L1: wait_acc1 %6d start %7d mma_phase %6d (wait_b %5d wait_a %5d issue %5d) | EPI: wait_full %6d start %7d colloop %6d tail %5d | MID: start %7d dur %6d (wait_acc0 %5d wait_a1e %5d)" % (
            k, i[1] - i[0], i[1] - t0, i[2] - i[1], i[3], i[4], i[5],
            e[1] - e[0], e[1] - t0, e[2] - e[1], e[3] - e[2],
            m[0] - t0, m[1] - m[0], m[2], m[3]))
    if K > 4:
        per = (iss[K - 2, 1] - iss[2, 1]) / (K - 4)
        print("track %d: mean period %.0f cycles over tiles 2..%d" % (t, per, K - 2))
        sel = slice(2, K - 2)
        print("   means: L1 wait_acc1 %.0f mma_phase %.0f (wait_b %.0f wait_a %.0f issue %.0f) | EPI wait_full %.0f colloop %.0f tail %.0f | MID dur %.0f wait_acc0 %.0f wait_a1e %.0f" % (
            (iss[sel, 1] - iss[sel, 0]).mean(), (iss[sel, 2] - iss[sel, 1]).mean(), iss[sel, 3].mean(), iss[sel, 4].mean(), iss[sel, 5].mean(),
            (epi[sel, 1] - epi[sel, 0]).mean(), (epi[sel, 2] - epi[sel, 1]).mean(), (epi[sel, 3] - epi[sel, 2]).mean(),
            (mid[sel, 1] - mid[sel, 0]).mean(), mid[sel, 2].mean(), mid[sel, 3].mean()))
