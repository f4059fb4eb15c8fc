#!/bin/bash
# usage: tools/gpu_multi7.sh TAG NGPU  (driver-like bench at N GPUs + reference arm)
TAG=${1:-multi}; N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== bench --gpus $N"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-3000 | tee $OUT/bench_c4_n${N}.txt
echo "== bench reference --gpus $N"; timeout 600 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-500
