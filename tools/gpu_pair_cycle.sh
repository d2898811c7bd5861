#!/bin/bash
# one develop-measure cycle for the fused two-sweep kernel (run under gpurun)
OUT=gpurun_out
timeout 300 python -m pytest tests/test_star7_pair_gpu.py -x -q > $OUT/pytest_pair.log 2>&1; echo rc=$? >> $OUT/pytest_pair.log
EXP_CONFIGS=${EXP_CONFIGS:-short} timeout 300 python tools/exp_pair.py > $OUT/exp_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:Star7Pair -s 4 -c 1 -f -o $OUT/prof_pair python bench.py --count 8 --steps 1 --warmup 3 --no-himeno --no-cpu --no-strong > $OUT/prof_pair.log 2>&1
