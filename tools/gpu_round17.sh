#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -24
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f e2e %.1f launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])); print('binnet', d['binnet']); print('de', d['de']); print('normvar', d['normvar'])"
echo "== ncu launch list (headline step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"contract|residual|coef|stats_finalize|sumsq|cov_" -c 60 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_launches_stdout.txt 2>&1
echo "== ncu full: binnet, single=1, normvar kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"binnet_rows|group_stats|normvar_" -c 6 -f -o $OUT/prof_aux \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_aux_stdout.txt 2>&1
ls -la $OUT | tail -5
