#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],1), "mhz", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "int8", round(d["roofline"]["executed_int8_tops"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2]).read()[-300:])
PY
}
for st in "umma_vs_simt default 128 700 5000 0" "umma_vs_simt fast 128 1333 3001 0" "umma_vs_simt precise 128 700 5000 0" "umma_golden default 128" "perf 5000 10000 default 128 1 0" "perf 8192 65536 default 128 0 0"; do
  timeout 300 python tools/gpu_bringup.py $st 2>&1 | grep -E "identical|residualize_ms|rror|relP" | cut -c1-420
done
for rep in 1 2; do
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --opt umma_stack=1 2>&1 | tail -1 > $OUT/stack1_$rep.txt; show stack1_$rep $OUT/stack1_$rep.txt
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --opt umma_stack=0 2>&1 | tail -1 > $OUT/stack0_$rep.txt; show stack0_$rep $OUT/stack0_$rep.txt
done
timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --opt umma_stack=1 --precision fast 2>&1 | tail -1 > $OUT/fast_stack1.txt; show fast_stack1 $OUT/fast_stack1.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-3500 | tee $OUT/bench_ours.txt
