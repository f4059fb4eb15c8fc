#!/bin/bash
# usage: tools/gpu_r2l.sh TAG : GPU tests of the steps around the hot path, their timings, one ncu --set full capture per kernel
TAG=${1:-r02l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest (normvar, lcpm)"; timeout 900 python -m pytest tests/test_normvar.py tests/test_lcpm.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.txt
echo "== aux timings"; timeout 600 python tools/aux_run.py lcpm compute_var normvar 2>&1 | tail -1 | tee $OUT/aux.json
for spec in "normvar:normvar_apply_kernel" "normvar:normvar_gemm_kernel" "compute_var:colvar_kernel" "lcpm:lcpm_colstats_kernel" "lcpm:lcpm_apply_kernel"; do
  op=${spec%%:*}; k=${spec##*:}
  timeout 600 ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -f -o $OUT/prof_$k python tools/aux_run.py $op > $OUT/ncu_$k.log 2>&1
  tail -1 $OUT/ncu_$k.log | cut -c1-200
done
ls -la $OUT
