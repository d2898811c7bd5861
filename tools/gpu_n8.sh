#!/bin/bash
# 8-GPU session: the bench line at N=8 (weak) and the multi-rank parity tests on real GPUs
OUT=gpurun_out
R=${ROUND:-r2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $T bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/${R}_bench_n8.json 2> $OUT/${R}_bench_n8.err; echo "bench rc=$?"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 400 > $OUT/${R}_pytest_mgpu8.log 2>&1; echo rc=$? >> $OUT/${R}_pytest_mgpu8.log
tail -3 $OUT/${R}_pytest_mgpu8.log; head -c 600 $OUT/${R}_bench_n8.json
