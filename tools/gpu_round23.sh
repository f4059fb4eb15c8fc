#!/bin/bash
OUT=gpurun_out/${1:-r}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
for a in 8192 0; do
echo "== bench adaptive_min_cells=$a"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-de --opt adaptive_min_cells=$a 2>&1 | tail -1 | tee $OUT/bench_adapt$a.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step %.2f kernel %.2f frac %.3f refined %s products %.2f int8 %.0f mhz %s' % (d['ms_per_step'], r['kernel_ms'], r['frac'], r.get('tiles_refined'), r['digit_products_executed'], r['executed_int8_tops'], d['clocks']['sm_mhz']))"
done
