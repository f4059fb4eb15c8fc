#!/bin/bash
# usage: tools/gpu_multi6.sh TAG NGPU  (default transport = auto -> copy engines)
TAG=${1:-multi}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 5000 3000 2>&1 | grep -E "multi-GPU check|identical on all|Error|error|Traceback|line " | head -30
echo "== bench --gpus $N (auto transport)"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-3000 | tee $OUT/bench_c4_n${N}_pairs_ce.txt
echo "== bench --gpus $N nccl"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --transport nccl --no-e2e 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-1400 | tee $OUT/bench_c4_n${N}_pairs_nccl.txt
