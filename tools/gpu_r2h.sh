#!/bin/bash
# usage: tools/gpu_r2h.sh TAG NGPU : N-GPU box: concurrent PCIe probe, bit-identity checks (multi-process and one process), timings, bench
TAG=${1:-r02h}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
df -h /dev/shm | tail -1 | tee $OUT/shm.txt; nproc; free -g | head -2; nvidia-smi topo -m > $OUT/topo.txt 2>&1; numactl -H > $OUT/numa.txt 2>&1
echo "== pcie probe, $N ranks at once"; timeout 300 $TR tools/pcie_probe.py 2>&1 | grep -E "^\{|Error|error|Traceback" | tee $OUT/pcie_probe_n$N.json
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 5000 3000 2>&1 | grep -E "multi-GPU check|identical|Error|error|Traceback|line |timeout" | head -60 | tee $OUT/multi_gpu_check_n$N.txt
echo "== all_devices timing (full size)"; timeout 900 python tools/all_devices_check.py 20000 100000 --time --only-all 2>&1 | grep -E "all-devices|Error|error|Traceback|line " | head -30 | tee $OUT/all_devices_full_n$N.txt
echo "== bench --gpus $N"; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-8000 | tee $OUT/bench_c4_n${N}.txt
