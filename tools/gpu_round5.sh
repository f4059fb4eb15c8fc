#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bringup residual + perf"; timeout 900 python tools/gpu_bringup.py residual > $OUT/residual.txt 2>&1; tail -3 $OUT/residual.txt | cut -c1-300
for v in "default 0" "default 1" "fast 0" "fast 1"; do set -- $v
  echo "== bench $1 pair=$2 (12 steps)"; timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --precision $1 --umma-pair $2 2>&1 | tail -1 > $OUT/bench_$1_$2.txt
  python - <<PY
import json
d=json.loads(open("$OUT/bench_$1_$2.txt").read())
print("$1 pair=$2", "ms/step", round(d["ms_per_step"],1), "value %.3e"%d["value"], d["clocks"], "kernel_ms", round(d["roofline"]["kernel_ms"],1), d["roofline"]["kernel_ms_per_step"], "int8", round(d["roofline"]["executed_int8_tops"]))
PY
done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
