#!/bin/bash
TAG=${1:-r02o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -u -m pytest tests/test_normvar.py tests/test_lcpm.py -x -v -m gpu --timeout 40 --timeout-method=thread 2>&1 | tail -60 | tee $OUT/pytest.txt
