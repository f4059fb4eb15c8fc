#!/bin/bash
# usage: tools/gpu_r2g.sh TAG NGPU : GPU tests, multi-process and one-process multi-GPU checks, timings, bench
TAG=${1:-r02g}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
df -h /dev/shm | tee $OUT/shm.txt; nproc; free -g | head -2; nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 5000 3000 2>&1 | grep -E "multi-GPU check|identical|Error|error|Traceback|line |timeout" | head -40 | tee $OUT/multi_gpu_check_n$N.txt
echo "== all_devices_check"; timeout 600 python tools/all_devices_check.py 5000 3000 2>&1 | grep -E "all-devices|Error|error|Traceback|line " | head -30 | tee $OUT/all_devices_check_n$N.txt
echo "== all_devices timing (full size)"; timeout 900 python tools/all_devices_check.py 20000 100000 --time 2>&1 | grep -E "all-devices|Error|error|Traceback|line " | head -30 | tee $OUT/all_devices_full_n$N.txt
echo "== bench --gpus $N"; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-6000 | tee $OUT/bench_c4_n${N}.txt
