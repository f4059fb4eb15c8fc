#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pcie"; timeout 300 python tools/pcie_probe.py 2>&1 | tail -1 | tee $OUT/pcie.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=4 2>&1 | tail -12
echo "== bench DE sweep"; timeout 900 python bench.py --workload de_1m_x_20k_x_1000 --steps 3 --warmup 3 > $OUT/bench_de_sweep_n1.txt 2>&1; tail -1 $OUT/bench_de_sweep_n1.txt | cut -c1-400
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_ours.txt 2>&1; tail -1 $OUT/bench_ours.txt | cut -c1-3800
