#!/bin/bash
# GPU test suite + smoke (what the driver runs first at round end)
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=3 2>&1 | tail -10
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
