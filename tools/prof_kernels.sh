#!/bin/bash
# ncu captures of the dominant kernels on one GPU (run under gpurun); reports land in gpurun_out/
set -x
OUT=gpurun_out
R=${ROUND:-r2}
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --count 8 --steps 1 --warmup 3 --himeno-nn 8 --pstag-count 4 --no-cpu --no-small --no-parity --strong-count 8"
timeout 300 $NCU -k regex:Star7Pair -s 4 -c 1 -o $OUT/${R}_prof_pair $B --no-strong > $OUT/${R}_prof_pair.log 2>&1
# the z-slab form of the same kernel on one GPU (nothing to exchange: the form's code only)
timeout 300 $NCU -k regex:Star7Pair -s 4 -c 1 -o $OUT/${R}_prof_pair_slabform $B --no-strong --opt debug_slab=4 > $OUT/${R}_prof_pair_slabform.log 2>&1
# x-tiled fused pass on 1024-wide rows (BASELINE config 4's grid on one GPU)
timeout 300 $NCU --kernel-name-base demangled -k "regex:Star7PairKernel<float, \(int\)3" -s 2 -c 1 -o $OUT/${R}_prof_pair_xtile $B > $OUT/${R}_prof_pair_xtile.log 2>&1
# single 7-point sweep (the tail of the schedule, and everything a program calling PSStencilRun(..., 1) runs)
timeout 300 $NCU -k regex:Star7KernelV2 -s 2 -c 1 -o $OUT/${R}_prof_star7 $B --no-strong > $OUT/${R}_prof_star7.log 2>&1
timeout 300 $NCU -k regex:HimenoPair -s 2 -c 1 -o $OUT/${R}_prof_himeno_pair $B --no-strong > $OUT/${R}_prof_himeno_pair.log 2>&1
# residual form of the single Himeno sweep (the with_residual bench leg)
timeout 300 $NCU --kernel-name-base demangled -k "regex:HimenoKernel<\(int\)15, \(bool\)1" -s 2 -c 1 -o $OUT/${R}_prof_himeno_gosa $B --no-strong > $OUT/${R}_prof_himeno_gosa.log 2>&1
timeout 300 $NCU -k regex:Pstag -s 4 -c 1 -o $OUT/${R}_prof_pstag $B --no-strong > $OUT/${R}_prof_pstag.log 2>&1
# launch list of the default bench command at reduced counts (shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${R}_launches.csv python bench.py --count 40 --steps 2 --warmup 3 --himeno-nn 8 --pstag-count 4 --strong-count 8 --no-cpu > $OUT/${R}_launches_bench.log 2>&1
ls -la $OUT/${R}_prof_*.ncu-rep
