"""One-line digest of bench.py's JSON line (stdin): python bench.py ... | python tools/bench_brief.py"""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    k = d.get("roofline", {}).get("kernels", {})
    print("%s: %.2f ms/step, %.3f M/s, e2e %.3f M/s, clocks %s %s, mlp lin %.4f roll %.4f ms, mom %.4f step %.4f, not_pd %s" % (
        d["config"]["workload"], d["ms_per_step"], d["value"] / 1e6, (d.get("e2e") or {}).get("value", 0) / 1e6,
        d["clocks"]["sm_mhz"], d["clocks"]["reasons"], k.get("mlp_linearise", {}).get("avg_ms", 0),
        k.get("mlp_rollout", {}).get("avg_ms", 0), k.get("moment_linearise", {}).get("avg_ms", 0),
        k.get("rollout_step", {}).get("avg_ms", 0), d.get("not_pd_problems")))
