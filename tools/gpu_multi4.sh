#!/bin/bash
# usage: tools/gpu_multi4.sh TAG NGPU
TAG=${1:-multi}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 5000 3000 2>&1 | grep -E "multi-GPU check|identical on all|Error|error|Traceback|line " | head -30
echo "== bench --gpus $N pairs"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-3000 | tee $OUT/bench_c4_n${N}_pairs.txt
echo "== bench --gpus $N allgather"; timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --schedule allgather --no-e2e 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-1400 | tee $OUT/bench_c4_n${N}_allgather.txt
echo "== DE sweep --gpus $N"; timeout 600 $TR bench.py --gpus $N --workload de_1m_x_20k_x_1000 --steps 3 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-500 | tee $OUT/bench_de_sweep_n${N}.txt
