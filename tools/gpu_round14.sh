#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=4 2>&1 | tail -14
bash tools/gpu_multi3.sh $TAG 2
