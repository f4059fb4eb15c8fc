#!/bin/bash
OUT=gpurun_out/r02d
mkdir -p $OUT
echo "== fused test"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_projection or residual_planes or host_pipeline" 2>&1 | tail -15
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.txt
for f in 0 1; do
  echo "== bench fused=$f"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-de --opt fused_projection=$f > $OUT/bench_f$f.txt 2>&1
  tail -1 $OUT/bench_f$f.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f contract %.2f proj ms %.3f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_projection']['kernel_ms'], d['roofline_projection']['frac']))"
done
echo "== ncu fused"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_fused -s 2 -c 1 -o $OUT/prof_fused python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu_fused.log 2>&1; tail -2 $OUT/ncu_fused.log
echo "== bench de c3"; timeout 600 python bench.py --workload de_50k_x_10k_x_300 --steps 5 --warmup 3 --no-cpu > $OUT/bench_de_c3.txt 2>&1; tail -1 $OUT/bench_de_c3.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['phase_ms'], d['e2e'])"
