#!/bin/bash
# usage: tools/gpu_r2i.sh TAG : GPU tests + the full default bench line (all blocks) on one GPU
TAG=${1:-r02i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest.txt
echo "== bench (default)"; timeout 1200 python bench.py > $OUT/bench_default.txt 2>&1; tail -1 $OUT/bench_default.txt | cut -c1-200
