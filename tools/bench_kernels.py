"""One line per bench JSON (stdin or files): workload, ms per step, per-kernel avg ms.  python tools/bench_kernels.py f.json ..."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        k = (d.get("roofline") or {}).get("kernels") or {}
        print(f, d["config"]["workload"], d.get("dtype"), "%.3f ms/step" % d["ms_per_step"],
              {n: round(v["avg_ms"], 4) for n, v in k.items()}, "not_pd", d.get("not_pd_problems"))
