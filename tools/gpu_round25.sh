#!/bin/bash
timeout 300 python - <<'PY'
import torch, json, bench
print(json.dumps(bench.bench_lcpm(torch, torch.device('cuda', 0))))
PY
