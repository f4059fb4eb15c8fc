#!/bin/bash
# Launch lists (per-launch durations under ncu) of the library's own kernels: headline coex step and de config 3.
TAG=${1:-r02ll}; OUT=gpurun_out/$TAG; mkdir -p $OUT
K='regex:coef_|residual_|stats_|contract_|cov_|gram|de4|chol|trsm|tri_inverse|pvalue|group_|sym_pinv|single1|tiles_upload|sub_sym|diag_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_launches_stdout.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv \
   --log-file $OUT/launches_de_c3.csv python bench.py --workload de_50k_x_10k_x_300 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_de_stdout.txt 2>&1
wc -l $OUT/*.csv
