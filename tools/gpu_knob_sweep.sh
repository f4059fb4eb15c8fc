#!/bin/bash
# sweep one nsr_set_option knob on the headline step, two interleaved rounds: tools/gpu_knob_sweep.sh TAG NAME v1 v2 ...
TAG=$1; NAME=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for round in 1 2; do
  for v in "$@"; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-de --opt $NAME=$v > $OUT/bench_${v}_$round.txt 2>&1
    tail -1 $OUT/bench_${v}_$round.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$NAME=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['kernel_ms'],2))"
  done
done
