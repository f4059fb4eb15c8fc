#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bringup quick"; timeout 1500 python tools/gpu_bringup.py --quick > $OUT/bringup.txt 2>&1; grep -E "^===|identical|residualize_ms|Error|error|assert|timeout|nsr umma" $OUT/bringup.txt | cut -c1-600 | tail -60
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench ours"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-3000 | tee $OUT/bench_ours.txt
echo "== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 --workload coex_10k_x_5k --no-cpu 2>&1 | tail -1 | cut -c1-3000 | tee $OUT/bench_c2.txt
