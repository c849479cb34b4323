#!/bin/bash
# cfg5-scale graph on 8 GPUs: feature split of d = 128 into 8 blocks of 16 columns, A^ (2.26 G non-zeros) replicated
OUT=gpurun_out/r2_rmat27
mkdir -p $OUT
free -g | head -2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --workload rmat27 --steps 3 --warmup 3 --no-e2e > $OUT/rmat27_n8.json 2> $OUT/rmat27_n8.err
echo "exit $?"
tail -c 1200 $OUT/rmat27_n8.json; echo; tail -5 $OUT/rmat27_n8.err | cut -c1-300
