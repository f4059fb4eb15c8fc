#!/bin/bash
# usage: tools/gpu_multi5.sh TAG NGPU   (NCCL vs copy-engine transport)
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== multi_gpu_check"; timeout 600 $TR tools/multi_gpu_check.py 3000 5000 2>&1 | grep -E "multi-GPU check|identical on all|Error|error|Traceback|line |rror" | head -40
for tr in ce nccl ce; do
echo "== bench --gpus $N pairs transport $tr"; timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --transport $tr --no-e2e 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('ms/step %.2f kernel_ms %.2f int8 %.0f mhz %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['executed_int8_tops'], d['clocks']['sm_mhz']))
except Exception as e:
    print('parse error', e)"
done
echo "== bench --gpus $N ce with e2e"; timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --transport ce 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-3000 | tee $OUT/bench_c4_n${N}_pairs_ce.txt
