#!/bin/bash
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench (de block)"; timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('de', d['de']); print('binnet', d['binnet']['ms'], 'normvar', d['normvar']['ms'])"
