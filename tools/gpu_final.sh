#!/bin/bash
# What the driver runs at round end, on one GPU: GPU tests, smoke, both bench arms.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-700 | tee $OUT/bench_reference.txt
echo "== bench ours"; timeout 900 python bench.py > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f value %.4g e2e %.1f ms (%.4g) launches %d frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac']))
print('cpu', d['cpu_baseline']['value']); print('binnet', d['binnet']['ms'], d['binnet']['roofline']['frac']); print('de', {k: (round(v['ms'], 2)) for k, v in d['de'].items() if isinstance(v, dict)}); print('normvar', d['normvar']['ms'], d['normvar']['roofline']['frac'])"
