#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (normvar, binnet)"; timeout 900 python -m pytest tests/test_normvar.py tests/test_binnet.py -q -m gpu 2>&1 | tail -8
echo "== normvar bench"; timeout 600 python - <<'PY'
import torch, json, bench
print(json.dumps(bench.bench_normvar(torch, torch.device('cuda', 0))))
PY
echo "== ncu normvar"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normvar_gemm|normvar_reduce|normvar_apply" -c 3 -f -o $OUT/prof_normvar python - > /dev/null 2>&1 <<'PY'
import torch, bench
bench.bench_normvar(torch, torch.device('cuda', 0), reps=1)
PY
ls -la $OUT | tail -3
