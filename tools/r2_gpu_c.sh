#!/bin/bash
OUT=gpurun_out
EXP_COUNT=1000 EXP_CONFIGS="star7_fuse=1|star7_fuse=1+debug_slab=4|star7_fuse=1|star7_fuse=1+debug_slab=4" timeout 200 python tools/exp_pair_mgpu.py > $OUT/r2_exp_pair_n1b.log 2>&1
cat $OUT/r2_exp_pair_n1b.log | cut -c1-120
