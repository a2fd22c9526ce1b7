#!/bin/bash
# compute-sanitizer over the GPU tests of the kernels with hand-written synchronisation (mbarrier pipelines,
# sub-warp teams, tcgen05 / TMEM): racecheck (shared-memory hazards), synccheck (barrier misuse), memcheck
# (out-of-bounds / misaligned accesses).  Run on the GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize_*.log
# (summaries are committed under profiles/).  The sanitizer slows kernels down 10-100x: the selection is small.
set -u
SEL=${SAN_SEL:-'tests/test_gpu_bnn.py::test_linearize tests/test_gpu_bnn.py::test_backward_and_rollout tests/test_gpu_bnn_tc.py::test_tc_many_tiles_per_cta tests/test_gpu_backward_nu.py tests/test_gpu_train.py::test_training_matches_the_reference'}
KNOWN='tests/test_gpu_known.py -k "pendulum_ign_bounded or cartpole_ut_f64 or double_cartpole_full or rendezvous_ign_bounded"'
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool" 
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest $SEL -m gpu -v -x -p no:cacheprovider > gpurun_out/sanitize${SAN_TAG:-}_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize${SAN_TAG:-}_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|exit " gpurun_out/sanitize${SAN_TAG:-}_$tool.log | tail -5
done
