#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -16
echo "== bench DE sweep (1M cells, 2500 genes/GPU)"; timeout 900 python bench.py --workload de_1m_x_20k_x_1000 --steps 3 --warmup 3 > $OUT/bench_de_sweep_n1.txt 2>&1; tail -3 $OUT/bench_de_sweep_n1.txt | cut -c1-1500
