#!/bin/bash
# compute-sanitizer passes over the small parity cases of the sweep kernels (run under gpurun)
OUT=gpurun_out
R=${ROUND:-r2}
PY="python -m pytest -x -q -p no:cacheprovider"
run() {  # name tool timeout pytest-args...
  local name=$1 tool=$2 t=$3; shift 3
  timeout $t compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 $PY "$@" > $OUT/${R}_sanitize_$name.log 2>&1
  echo "rc=$?" >> $OUT/${R}_sanitize_$name.log
}
run memcheck_pair memcheck 600 tests/test_star7_pair_gpu.py -k "fp32_matches_oracle or equal_coefficients"
run memcheck_sweeps memcheck 600 tests/test_sweeps_gpu.py -k "(matches_oracle or fused_two_sweep or residual_reduced) and not variants"
run memcheck_selfcheck memcheck 600 tests/test_selfcheck_suite.py -k "b200"
run racecheck_pair racecheck 600 tests/test_star7_pair_gpu.py -k "fp32_matches_oracle and (shape1 or shape4 or shape6 or shape12 or shape14)"
run racecheck_himeno_pair racecheck 600 tests/test_sweeps_gpu.py -k "fused_two_sweep and (dims0 or dims2 or dims6)"
run synccheck_pair synccheck 300 tests/test_star7_pair_gpu.py -k "fp32_matches_oracle and (shape1 or shape4 or shape12)"
run synccheck_himeno_pair synccheck 300 tests/test_sweeps_gpu.py -k "fused_two_sweep and (dims0 or dims2)"
run memcheck_autotune memcheck 600 tests/test_autotune_gpu.py
run initcheck_sweeps initcheck 600 tests/test_sweeps_gpu.py -k "(matches_oracle or fused_two_sweep or residual_reduced) and not variants"
# two ranks on one GPU: the in-kernel neighbour ordering (named barriers, per-CTA reporting)
timeout 500 compute-sanitizer --tool synccheck --target-processes all --error-exitcode 9 --print-limit 20 $PY tests/test_multigpu.py -k "tail_chunk or (tune_on_their_own_iterations and 2)" > $OUT/${R}_sanitize_synccheck_two_ranks.log 2>&1
echo "rc=$?" >> $OUT/${R}_sanitize_synccheck_two_ranks.log
for f in $OUT/${R}_sanitize_*.log; do echo "== $(basename $f .log)"; grep -E "passed|failed|SUMMARY|rc=" $f; done > $OUT/${R}_sanitizer.txt
cat $OUT/${R}_sanitizer.txt
