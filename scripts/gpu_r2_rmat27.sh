#!/bin/bash
# cfg5-scale graph (R-MAT scale 27: 134 M vertices, 2.1 G directed entries before merging) on ONE GPU with a 16-column block:
# the per-rank work of an 8-way feature split at d = 128.  nnz > 2^31 through the native builder and the hop kernels.
OUT=gpurun_out/r2_rmat27
mkdir -p $OUT
nvidia-smi --query-gpu=memory.total,memory.used --format=csv | tail -1
free -g | head -2
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 1 --split-one --workload rmat27 --feat-dim 16 --steps 3 --warmup 3 --no-e2e > $OUT/rmat27_d16.json 2> $OUT/rmat27_d16.err
echo "exit $?"
tail -c 1500 $OUT/rmat27_d16.json; echo; tail -5 $OUT/rmat27_d16.err | cut -c1-300
