#!/bin/bash
# Round-2 profiling pass on one GPU: launch list of the headline step + ncu --set full of the hot kernels (no source
# import: the reports must stay below the 64 MiB that travel back).
TAG=${1:-r02prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu launch list (headline step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_launches_stdout.txt 2>&1
tail -3 $OUT/launches.csv | cut -c1-200
echo "== ncu full: contraction"
timeout 900 ncu --set full --clock-control none -k regex:contract_umma -s 1 -c 1 -f -o $OUT/prof_contract \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_contract_stdout.txt 2>&1
echo "== ncu full: projection"
timeout 900 ncu --set full --clock-control none -k regex:"coef_mma_kernel|residual_mma_kernel" -s 3 -c 2 -f -o $OUT/prof_project \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_project_stdout.txt 2>&1
echo "== ncu launch list (de config 3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file $OUT/launches_de_c3.csv python bench.py --workload de_50k_x_10k_x_300 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_de_stdout.txt 2>&1
ls -la $OUT | tail -12
