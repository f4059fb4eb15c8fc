#!/bin/bash
# usage: tools/gpu_tests_only.sh TAG [pytest args] : the GPU test suite (or a selection), output kept
TAG=${1:-tests}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -40 | tee $OUT/pytest.txt
