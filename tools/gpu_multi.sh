#!/bin/bash
# usage: tools/gpu_multi.sh TAG NGPU
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 3000 5000 2>&1 | grep -E "multi-GPU check|identical|Error|error|Traceback|line " | head -20
echo "== bench --gpus $N (10k x 5k)"; timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload coex_10k_x_5k --no-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-1500 | tee $OUT/bench_c2_n$N.txt
echo "== bench --gpus $N (100k x 20k)"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-2500 | tee $OUT/bench_c4_n$N.txt
echo "== bench reference --gpus $N"; timeout 900 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2 | cut -c1-600
