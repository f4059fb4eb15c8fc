#!/bin/bash
OUT=gpurun_out/${1:-ab3}
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
timeout 600 python tools/gpu_bringup.py residual 2>&1 | tail -2 | cut -c1-200
for st in "umma_vs_simt default 128 700 5000 1" "umma_vs_simt default 128 1333 3001 0" "umma_vs_simt precise 128 700 5000 1" "perf 5000 10000 default 128 1 1" "perf 5000 10000 default 128 1 0" "perf 8192 65536 default 128 0 1" "perf 8192 65536 default 128 0 0"; do
  timeout 300 python tools/gpu_bringup.py $st 2>&1 | grep -E "identical|residualize_ms|rror" | cut -c1-420
done
for rep in 1 2; do
  (cd _old_b && timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de 2>&1 | tail -1 > ../$OUT/oldb_$rep.txt); show oldb_$rep $OUT/oldb_$rep.txt
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --umma-pair 0 2>&1 | tail -1 > $OUT/new_p0_$rep.txt; show new_p0_$rep $OUT/new_p0_$rep.txt
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --umma-pair 1 2>&1 | tail -1 > $OUT/new_p1_$rep.txt; show new_p1_$rep $OUT/new_p1_$rep.txt
done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
