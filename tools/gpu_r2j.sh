#!/bin/bash
# usage: tools/gpu_r2j.sh TAG : GPU tests, the steps around the hot path timed, ncu --set full of their kernels
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest.txt
echo "== aux timings"; timeout 600 python tools/aux_run.py 2>&1 | tail -1 | tee $OUT/aux.json
echo "== ncu aux"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lcpm_|colvar|normvar_|coef_mma|sym_pinv" -c 40 -f -o $OUT/prof_aux python tools/aux_run.py lcpm compute_var normvar > $OUT/ncu_aux.log 2>&1; tail -2 $OUT/ncu_aux.log
ls -la $OUT
