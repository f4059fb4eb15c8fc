#!/bin/bash
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (normvar, binnet)"; timeout 900 python -m pytest tests/test_normvar.py tests/test_binnet.py -q -m gpu 2>&1 | tail -8
echo "== ncu launch list: normvar + single1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"normvar_|sym_pinv|group_stats|coef_mma|index|gather" -c 60 --csv \
   --log-file $OUT/launches_aux.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_aux_stdout.txt 2>&1
python - $OUT/launches_aux.csv <<'PY'
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1]))); hdr=None; agg=collections.OrderedDict()
for r in rows:
    if len(r)>5 and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); name=d["Kernel Name"][:70]
        try: v=float(d["Metric Value"].replace(",",""))
        except: continue
        u=d["Metric Unit"]; v = v/1e3 if u in ("usecond","us") else v/1e6 if u in ("nsecond","ns") else v*1e3 if u in ("second","s") else v
        a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print("%-72s n=%3d %9.3f ms (%.3f each)"%(k,c,t,t/c))
PY
tail -1 $OUT/ncu_aux_stdout.txt | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('normvar (under ncu, not a bench value)', d['normvar'])
except Exception as e: print('no json', e)"
echo "== ncu full: normvar + group_stats"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normvar_stats|normvar_apply" -c 2 -f -o $OUT/prof_normvar \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"group_stats" -s 1 -c 1 -f -o $OUT/prof_group \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
echo "== normvar bench (plain)"; timeout 600 python - <<'PY'
import torch, json, bench
print(json.dumps(bench.bench_normvar(torch, torch.device('cuda', 0))))
PY
ls -la $OUT | tail -5
