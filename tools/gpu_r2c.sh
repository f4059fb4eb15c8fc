#!/bin/bash
OUT=gpurun_out/r02c
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest.txt
bash tools/gpu_multi_r2.sh r02c 2
echo "== coex N=1 quick"; timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-de > $OUT/bench_coex.txt 2>&1; tail -1 $OUT/bench_coex.txt | cut -c1-1200
echo "== ncu launches c3"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c3.csv python tools/de_probe.py c3 2 2>&1 | tail -2
