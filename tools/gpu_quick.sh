#!/bin/bash
# usage: tools/gpu_quick.sh [pytest args]   (a subset of the GPU tests, for iterating)
timeout 600 python -m pytest "${@:-tests}" -x -q -m gpu 2>&1 | tail -14
