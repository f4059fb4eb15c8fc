#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (binnet)"; timeout 900 python -m pytest tests/test_binnet.py -q -m gpu 2>&1 | tail -6
echo "== bench (binnet block)"; timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('binnet', d['binnet']); print('normvar', d['normvar']['ms'])"
echo "== ncu binnet"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"binnet_rows" -s 1 -c 1 -f -o $OUT/prof_binnet python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
ls -la $OUT | tail -2
