#!/bin/bash
# round 2, call B: new kernels (segments, exact planes, de4 solve) + DE benches
OUT=gpurun_out/r02b
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest.txt
echo "== bench de c3"; timeout 600 python bench.py --workload de_50k_x_10k_x_300 --steps 5 --warmup 3 > $OUT/bench_de_c3.txt 2>&1; tail -1 $OUT/bench_de_c3.txt | cut -c1-2500
echo "== bench de c5"; timeout 900 python bench.py --workload de_1m_x_20k_x_1000 --steps 3 --warmup 3 > $OUT/bench_de_c5.txt 2>&1; tail -1 $OUT/bench_de_c5.txt | cut -c1-2500
echo "== bench coex quick"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-de > $OUT/bench_coex.txt 2>&1; tail -1 $OUT/bench_coex.txt | cut -c1-1500
