#!/bin/bash
# round-2 1-GPU session: failed tests, Himeno experiments + ncu of the fused Himeno pass, pstag tuning
OUT=gpurun_out
timeout 600 python -m pytest tests/test_sweeps_gpu.py tests/test_multigpu.py tests/test_baseline_shapes_gpu.py -m gpu -q --timeout 400 -x > $OUT/r2_pytest4.log 2>&1; echo rc=$? >> $OUT/r2_pytest4.log
EXP_CONFIGS="|himeno_fuse=0|himeno_pair_pf=1|himeno_pair_pf=4|himeno_pair_zc=64|himeno_fuse=0+himeno_sthint=2|himeno_fuse=0+himeno_sthint=3|himeno_fuse=0+himeno_sthint=1" timeout 300 python tools/exp_himeno.py XL 20 > $OUT/r2_exp_himeno.log 2>&1
PSTAG_VARIANTS=13,9,12,4,8 PSTAG_STAGES=6 PSTAG_OCCS=0,4,5,6 timeout 300 python tools/tune_pstag.py 512 > $OUT/r2_tune_pstag.log 2>&1
B="python bench.py --count 8 --steps 1 --warmup 3 --himeno-nn 8 --pstag-count 4 --no-cpu --no-strong --no-small --no-parity"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:HimenoPair -s 2 -c 1 -f -o $OUT/r2_prof_himeno_pair $B > $OUT/r2_prof_himeno_pair.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:HimenoKernel -s 6 -c 1 -f -o $OUT/r2_prof_himeno_gosa $B > $OUT/r2_prof_himeno_gosa.log 2>&1
tail -3 $OUT/r2_pytest4.log; cat $OUT/r2_exp_himeno.log | tail -20; tail -8 $OUT/r2_tune_pstag.log
