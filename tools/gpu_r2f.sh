#!/bin/bash
OUT=gpurun_out/r02f
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.txt
echo "== bench coex"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-de > $OUT/bench_coex.txt 2>&1
tail -1 $OUT/bench_coex.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms/step %.2f contract %.2f proj ms %.3f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_projection']['kernel_ms'], d['roofline_projection']['frac']))"
echo "== bench de c3"; timeout 600 python bench.py --workload de_50k_x_10k_x_300 --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT/bench_de_c3.txt 2>&1; tail -1 $OUT/bench_de_c3.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['phase_ms'])"
echo "== bench de c5"; timeout 900 python bench.py --workload de_1m_x_20k_x_1000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/bench_de_c5.txt 2>&1; tail -1 $OUT/bench_de_c5.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['phase_ms'])"
echo "== ncu projection"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"residual_mma|coef_mma" -s 4 -c 2 -o $OUT/prof_project python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-de > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log; ls -la $OUT
