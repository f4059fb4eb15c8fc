#!/bin/bash
# A/B of two builds of the library on the same box: alternate them, 10 timed steps each, two rounds.  The other build
# (e.g. the previous commit's libnsr_b200.so) goes to tools/ab/libnsr_old.so (git-ignored) or $NSR_AB_OLD.
OUT=gpurun_out/${1:-r02ab}; mkdir -p $OUT
cp normalisr_b200/libnsr_b200.so /tmp/lib_new.so
for round in 1 2; do
  for v in new old; do
    if [ $v = old ]; then cp ${NSR_AB_OLD:-tools/ab/libnsr_old.so} normalisr_b200/libnsr_b200.so; else cp /tmp/lib_new.so normalisr_b200/libnsr_b200.so; fi
    touch normalisr_b200/libnsr_b200.so
    timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-de > $OUT/bench_${v}_$round.txt 2>&1
    tail -1 $OUT/bench_${v}_$round.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['kernel_ms'],2), round(d['roofline_projection']['kernel_ms'],2))"
  done
done
cp /tmp/lib_new.so normalisr_b200/libnsr_b200.so
