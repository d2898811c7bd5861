#!/bin/bash
# ncu captures of the dominant kernels on one GPU (run under gpurun); reports land in gpurun_out/
set -x
OUT=gpurun_out
B="python bench.py --count 8 --steps 1 --warmup 3 --himeno-nn 2 --pstag-count 4 --no-cpu --no-strong"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:Star7Pair -s 4 -c 1 -f -o $OUT/prof_pair $B > $OUT/prof_pair.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:Star7KernelV2 -s 4 -c 1 -f -o $OUT/prof_star7 $B --opt star7_fuse=0 > $OUT/prof_star7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:Himeno -s 4 -c 1 -f -o $OUT/prof_himeno $B > $OUT/prof_himeno.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:Pstag -s 4 -c 1 -f -o $OUT/prof_pstag $B > $OUT/prof_pstag.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ReduceStage1 -c 1 -f -o $OUT/prof_reduce $B > $OUT/prof_reduce.log 2>&1
# launch list of the default bench command at reduced counts (shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --count 40 --steps 2 --warmup 3 --himeno-nn 4 --pstag-count 4 --no-cpu --no-strong > $OUT/launches_bench.log 2>&1
