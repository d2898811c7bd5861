#!/bin/bash
OUT=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
EXP_CONFIGS="slab_zbl=0|slab_zbl=0+early_signal=0|halo_push=0|halo_push=0+sync_mode=0|slab_zbl=0+pstag_variant=4|slab_zbl=0+early_signal=0+pstag_variant=4" timeout 300 $T tools/exp_pstag_mgpu.py > $OUT/r2_exp_pstag_mgpu3.log 2>&1
sed 's/rank \([01]\)\/2/\nrank \1\/2/g' $OUT/r2_exp_pstag_mgpu3.log | grep "^rank" | sort -k4 | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2c.json'))
print('n2 value',d['value'],'himeno',d['himeno']['sweep_only']['glups'],'pstag',d['periodic_staggered_fp64']['glups'],'strong',d['strong_scaling_1024']['glups'], d.get('parity_ok'))
PY
