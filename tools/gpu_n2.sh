#!/bin/bash
# 2-GPU session: multi-rank parity tests on real GPUs, Himeno on z-slabs, bench at N=2
OUT=gpurun_out
R=${ROUND:-r2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 400 -x > $OUT/${R}_pytest_mgpu2.log 2>&1; echo rc=$? >> $OUT/${R}_pytest_mgpu2.log
timeout 600 $T bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/${R}_bench_n2.json 2> $OUT/${R}_bench_n2.err; echo "bench rc=$?"
tail -30 $OUT/${R}_pytest_mgpu2.log | cut -c1-220
