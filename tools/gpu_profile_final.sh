#!/bin/bash
# Final profiling pass of a round: launch list of the headline step + ncu --set full of the hot kernels.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu launch list (headline step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_launches_stdout.txt 2>&1
tail -3 $OUT/launches.csv | cut -c1-200
echo "== ncu full: contraction"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_umma -s 1 -c 1 -f -o $OUT/prof_contract \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_contract_stdout.txt 2>&1
echo "== ncu full: binnet + single=1 kernels + basis"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"binnet_rows|group_stats|cov_gram_partial|cov_apply" -c 8 -f -o $OUT/prof_aux \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_aux_stdout.txt 2>&1
ls -la $OUT | tail -6
