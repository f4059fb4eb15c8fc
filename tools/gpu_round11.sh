#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/gpu_bringup.py residual 2>&1 | tail -2 | cut -c1-200
for st in "umma_vs_simt default 128 700 5000 0" "umma_golden default 128" "perf 8192 65536 default 128 0 0"; do
  timeout 300 python tools/gpu_bringup.py $st 2>&1 | grep -E "identical|residualize_ms|rror|relP" | cut -c1-420
done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=4 2>&1 | tail -12
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | cut -c1-3800
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-900 | tee $OUT/bench_reference.txt
