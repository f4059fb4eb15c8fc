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
timeout 300 python tools/gpu_bringup.py umma_vs_simt default 128 700 5000 0 2>&1 | grep -E "identical|rror"; NSR_EPI_WARPS=16 timeout 300 python tools/gpu_bringup.py umma_vs_simt default 128 1333 3001 0 2>&1 | grep -E "identical|rror"
for rep in 1 2; do
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --opt epi_warps=8 2>&1 | tail -1 > $OUT/ew8_$rep.txt; show ew8_$rep $OUT/ew8_$rep.txt
  timeout 600 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-de --opt epi_warps=16 2>&1 | tail -1 > $OUT/ew16_$rep.txt; show ew16_$rep $OUT/ew16_$rep.txt
done
echo "== ncu full: contraction (ew8)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_umma -s 1 -c 1 -f -o $OUT/prof_contract \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_contract_stdout.txt 2>&1
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"contract|residual|coef|stats_finalize|sumsq" -c 40 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_launches_stdout.txt 2>&1
ls -la $OUT | tail -4
