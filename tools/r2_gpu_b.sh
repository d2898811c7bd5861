#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_sweeps_gpu.py -m gpu -q --timeout 400 -x -k "himeno" > $OUT/r2_pytest5.log 2>&1; echo rc=$? >> $OUT/r2_pytest5.log
EXP_CONFIGS="|himeno_pair_pf=0|himeno_pair_pf=1|himeno_pair_pf=3|himeno_pair_zc=64|himeno_pair_zc=32" timeout 300 python tools/exp_himeno.py XL 20 > $OUT/r2_exp_himeno2.log 2>&1
B="python bench.py --count 8 --steps 1 --warmup 3 --himeno-nn 8 --pstag-count 4 --no-cpu --no-strong --no-small --no-parity"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:HimenoPair -s 2 -c 1 -f -o $OUT/r2_prof_himeno_pair2 $B > $OUT/r2_prof_himeno_pair2.log 2>&1
tail -3 $OUT/r2_pytest5.log; cat $OUT/r2_exp_himeno2.log | grep sweeps
