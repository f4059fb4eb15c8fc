#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -24
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f e2e %.1f binnet %s de %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['binnet'], d['de']))"
