#!/bin/bash
# usage: tools/gpu_r2n.sh TAG : tests of the steps around the hot path + their timings
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest (normvar, lcpm)"; timeout 900 python -m pytest tests/test_normvar.py tests/test_lcpm.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.txt
echo "== aux timings"; timeout 600 python tools/aux_run.py lcpm compute_var normvar 2>&1 | tail -1 | tee $OUT/aux.json
