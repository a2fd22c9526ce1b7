#!/bin/bash
# On the GPU box: bench the default workload with each variant library given (names under pddp_b200/lib/variants).
#   tools/run_variants.sh base nb7 ...     ->  gpurun_out/var_<name>.json + a one-line digest each
mkdir -p gpurun_out
for v in "$@"; do
  PDDP_B200_LIB=$PWD/pddp_b200/lib/variants/libpddp_$v.so timeout -k 5 ${VTIMEOUT:-150} python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline \
      > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  echo -n "$v: "; python tools/bench_brief.py < gpurun_out/var_$v.json
done
