#!/bin/bash
# What the driver runs at round end, on one GPU: GPU tests, smoke(), both bench arms.
TAG=${1:-r02final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > $OUT/bench_reference.txt 2>&1; tail -1 $OUT/bench_reference.txt | cut -c1-400
echo "== bench"; timeout 1200 python bench.py > $OUT/bench_default.txt 2>&1; tail -1 $OUT/bench_default.txt | cut -c1-300
