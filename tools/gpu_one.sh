#!/bin/bash
# usage: tools/gpu_one.sh bench_function   (one secondary bench block on its own)
timeout 200 python - "$1" <<'PY'
import sys, json, torch, bench
print(json.dumps(getattr(bench, sys.argv[1])(torch, torch.device('cuda', 0))))
PY
