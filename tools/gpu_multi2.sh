#!/bin/bash
# usage: tools/gpu_multi2.sh TAG NGPU   (pairs vs allgather schedule)
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
F='grep -v "^W\|^\*\*\*\|OMP_NUM"'
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 3000 5000 2>&1 | grep -E "multi-GPU check|identical|Error|error|Traceback|line " | head -30
echo "== multi_gpu_check odd sizes"; timeout 600 $TR tools/multi_gpu_check.py 1111 3001 2>&1 | grep -E "multi-GPU check|identical on all|Error|error|Traceback|line " | head -30
echo "== bench --gpus $N pairs"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-3000 | tee $OUT/bench_c4_n${N}_pairs.txt
echo "== bench --gpus $N allgather"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --schedule allgather --no-e2e 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-3000 | tee $OUT/bench_c4_n${N}_allgather.txt
