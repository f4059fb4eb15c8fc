#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de "$@" 2>&1 | tail -1 > $OUT/bench_$name.txt
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.txt").read())
    print("$name", "ms/step", round(d["ms_per_step"],1), "value %.3e"%d["value"], "mhz", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_max"), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "int8", round(d["roofline"]["executed_int8_tops"]))
except Exception as e:
    print("$name FAILED", e, open("$OUT/bench_$name.txt").read()[-500:])
PY
}
run p0_ov1_s500 --umma-pair 0
run p0_ov1_s0 --umma-pair 0 --opt epi_sleep_ns=0
run p0_ov1_s4000 --umma-pair 0 --opt epi_sleep_ns=4000
run p0_ov0_s500 --umma-pair 0 --opt epi_overlap=0
run p1_ov0_s500 --umma-pair 1 --opt epi_overlap=0
run p1_ov1_s4000 --umma-pair 1 --opt epi_sleep_ns=4000
run fast_p1_ov1_s4000 --umma-pair 1 --opt epi_sleep_ns=4000 --precision fast
