#!/bin/bash
# compute-sanitizer passes over the small parity cases of the sweep kernels (run under gpurun)
OUT=gpurun_out
PY="python -m pytest -x -q -p no:cacheprovider"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 $PY tests/test_star7_pair_gpu.py -k "fp32_matches_oracle or equal_coefficients" > $OUT/sanitize_memcheck_pair.log 2>&1; echo "rc=$?" >> $OUT/sanitize_memcheck_pair.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 $PY tests/test_sweeps_gpu.py -k "matches_oracle and not variants" > $OUT/sanitize_memcheck_sweeps.log 2>&1; echo "rc=$?" >> $OUT/sanitize_memcheck_sweeps.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 $PY tests/test_star7_pair_gpu.py -k "fp32_matches_oracle and (shape1 or shape4 or shape6)" > $OUT/sanitize_racecheck_pair.log 2>&1; echo "rc=$?" >> $OUT/sanitize_racecheck_pair.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 $PY tests/test_star7_pair_gpu.py -k "fp32_matches_oracle and (shape1 or shape4)" > $OUT/sanitize_synccheck_pair.log 2>&1; echo "rc=$?" >> $OUT/sanitize_synccheck_pair.log
