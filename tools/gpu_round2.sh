#!/bin/bash
# quick bring-up + parity tests + bench (no ncu)
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bringup quick"; timeout 1500 python tools/gpu_bringup.py --quick > $OUT/bringup.txt 2>&1; grep -E "^===|identical|residualize_ms|Error|error|assert|quant_err" $OUT/bringup.txt | cut -c1-400 | tail -60
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench ours"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -4 | tee $OUT/bench_ours.txt
echo "== bench ours 10k x 5k"; timeout 900 python bench.py --steps 10 --warmup 3 --workload coex_10k_x_5k --no-cpu 2>&1 | tail -2 | tee $OUT/bench_c2.txt
