#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=3 2>&1 | tail -8
for dyn in 1 0 1 0; do
  echo "== bench dynamic=$dyn"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-de --opt umma_dynamic=$dyn > $OUT/bench_dyn${dyn}.txt 2>&1; tail -1 $OUT/bench_dyn${dyn}.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f kernel_ms %.2f %s int8 %.0f mhz %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_per_step'], d['roofline']['executed_int8_tops'], d['clocks']['sm_mhz']))"
done
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | cut -c1-4200
