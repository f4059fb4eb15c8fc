#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/gpu_bringup.py residual 2>&1 | tail -4 | cut -c1-330
for st in "umma_vs_simt default 128 700 5000 0" "umma_vs_simt default 128 1333 3001 0" "umma_vs_simt precise 128 700 5000 1" "umma_golden default 128" "perf 5000 10000 default 128 1 0" "perf 8192 65536 default 128 0 0"; do
  timeout 300 python tools/gpu_bringup.py $st 2>&1 | grep -E "identical|residualize_ms|rror|relP" | cut -c1-420
done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -14 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 12 --warmup 3 2>&1 | tail -1 | cut -c1-3500 | tee $OUT/bench_ours.txt
echo "== ncu projection"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"residual_mma|coef_mma" -s 3 -c 3 -f -o $OUT/prof_project \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_project_stdout.txt 2>&1
ls -la $OUT | tail -5
