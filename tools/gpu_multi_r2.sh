#!/bin/bash
# usage: tools/gpu_multi_r2.sh TAG NGPU : bit-identity check of the sharded path + bench at N GPUs (outputs kept)
TAG=${1:-r02multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT profiles
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 5000 3000 2>&1 | grep -E "multi-GPU check|identical|Error|error|Traceback|line |timeout" | head -40 | tee $OUT/multi_gpu_check_n$N.txt
echo "== bench --gpus $N"; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-4000 | tee $OUT/bench_c4_n${N}.txt
