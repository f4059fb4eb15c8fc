#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bringup quick"; timeout 1500 python tools/gpu_bringup.py --quick > $OUT/bringup.txt 2>&1; grep -E "^===|identical|residualize_ms|Error|error|assert|timeout|nsr umma" $OUT/bringup.txt | cut -c1-600 | tail -60
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench ours"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-3000 | tee $OUT/bench_ours.txt
echo "== ncu full: projection"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"residual_kernel|coef_kernel" -s 3 -c 3 -f -o $OUT/prof_project \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_project_stdout.txt 2>&1
echo "== ncu full: contraction"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_umma -s 1 -c 1 -f -o $OUT/prof_contract \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_contract_stdout.txt 2>&1
ls -la $OUT
