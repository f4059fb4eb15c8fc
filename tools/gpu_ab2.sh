#!/bin/bash
OUT=gpurun_out/${1:-ab2}
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
for rep in 1 2; do
 for v in _old _old_a _old_b; do
  (cd $v && timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de 2>&1 | tail -1 > ../$OUT/${v}_$rep.txt); show ${v}_$rep $OUT/${v}_$rep.txt
 done
 timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --umma-pair 0 --opt epi_overlap=0 2>&1 | tail -1 > $OUT/new_p0ov0_$rep.txt; show new_p0ov0_$rep $OUT/new_p0ov0_$rep.txt
done
