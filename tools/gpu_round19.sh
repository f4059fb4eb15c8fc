#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -16
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f e2e %.1f launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])); print('binnet', d['binnet']); print('de', d['de']); print('normvar', d['normvar'])"
echo "== ncu full: binnet, normvar, group_stats (one launch each)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"binnet_rows" -s 1 -c 1 -f -o $OUT/prof_binnet python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normvar_stats|normvar_apply|sym_pinv" -c 3 -f -o $OUT/prof_normvar python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"group_stats" -s 1 -c 1 -f -o $OUT/prof_group python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
ls -la $OUT | tail -5
