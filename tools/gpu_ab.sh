#!/bin/bash
# A/B on the SAME box: old kernel (commit ea8d068, 4-warp serial epilogue) vs current
OUT=gpurun_out/${1:-ab}
mkdir -p $OUT
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],1), "mhz", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "int8", round(d["roofline"]["executed_int8_tops"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for rep in 1 2; do
  (cd _old && timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de 2>&1 | tail -1 > ../$OUT/old_$rep.txt); show old_$rep $OUT/old_$rep.txt
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --umma-pair 0 --opt epi_overlap=0 2>&1 | tail -1 > $OUT/new_p0ov0_$rep.txt; show new_p0ov0_$rep $OUT/new_p0ov0_$rep.txt
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --umma-pair 1 --opt epi_overlap=1 2>&1 | tail -1 > $OUT/new_p1ov1_$rep.txt; show new_p1ov1_$rep $OUT/new_p1ov1_$rep.txt
done
nvidia-smi --query-gpu=name,power.limit,power.max_limit,temperature.gpu,clocks.sm --format=csv
