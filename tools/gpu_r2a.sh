#!/bin/bash
# round 2, call A: new parity tests, prefetch A/B, DE launch lists
OUT=gpurun_out/r02a
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.txt
for pf in 0 1; do
  echo "== bench prefetch=$pf"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-de --opt prefetch=$pf > $OUT/bench_pf$pf.txt 2>&1
  tail -1 $OUT/bench_pf$pf.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f contract %.2f proj %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_projection']))"
done
echo "== ncu launches c3"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c3.csv python tools/de_probe.py c3 2 2>&1 | tail -2
echo "== ncu launches c5"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c5.csv python tools/de_probe.py c5 2 2>&1 | tail -2
