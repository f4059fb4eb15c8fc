#!/bin/bash
# usage: tools/grun.sh LOG [gpurun args...] -- 'command'
# rebuilds the in-tree library first so the snapshot never carries a stale .so
LOG=$1; shift
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 /tmp/build.log; exit 1; }
/usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
