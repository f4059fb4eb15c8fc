#!/bin/bash
# usage: tools/gpu_r2k.sh TAG op... : launch lists (every kernel, torch's included) of the steps around the hot path
TAG=${1:-r02k}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for op in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_$op.csv python tools/aux_run.py $op > $OUT/ncu_$op.log 2>&1
  tail -1 $OUT/ncu_$op.log | cut -c1-300
done
