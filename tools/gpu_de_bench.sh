#!/bin/bash
TAG=${1:-r02s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for wl in de_50k_x_10k_x_300 de_1m_x_20k_x_1000; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT/bench_$wl.txt 2>&1
  tail -1 $OUT/bench_$wl.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl', round(d['ms_per_step'], 3), d['roofline']['phase_ms'])"
done
