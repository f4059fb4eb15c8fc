#!/bin/bash
# One GPU trip: parity tests, smoke, bench (both arms), ncu launch list + full captures.
# usage: tools/gpu_round.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_before.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench ours"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -4 | tee $OUT/bench_ours.txt
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_reference.txt
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"contract|residual|coef|stats_finalize" -c 60 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_launches_stdout.txt 2>&1
tail -2 $OUT/ncu_launches_stdout.txt
echo "== ncu full: contraction"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_umma -s 1 -c 1 -f -o $OUT/prof_contract \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_contract_stdout.txt 2>&1
echo "== ncu full: projection"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"residual_kernel|coef_kernel" -s 3 -c 3 -f -o $OUT/prof_project \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_project_stdout.txt 2>&1
ls -la $OUT
