#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for st in "umma_vs_simt default 128 700 5000 1" "umma_vs_simt default 128 1333 3001 0" "umma_vs_simt precise 128 700 5000 1" "perf 5000 10000 default 128 1 1" "perf 5000 10000 default 128 1 0" "perf 8192 65536 default 128 0 1" "perf 8192 65536 default 128 0 0"; do
  timeout 300 python tools/gpu_bringup.py $st 2>&1 | grep -E "identical|residualize_ms|rror" | cut -c1-420
done
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
run p0_ov1 --umma-pair 0
run p0_ov0 --umma-pair 0 --opt epi_overlap=0
run p1_ov1 --umma-pair 1
run p1_ov0 --umma-pair 1 --opt epi_overlap=0
run fast_p1_ov1 --umma-pair 1 --precision fast
run fast_p1_ov0 --umma-pair 1 --precision fast --opt epi_overlap=0
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
