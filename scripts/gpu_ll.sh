#!/bin/bash
# launch list of a short bench: bash scripts/gpu_ll.sh [bench args]
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --samples 1.07e9 --steps 1 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 "$@" > gpurun_out/bench_ncu.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(lines):
    k=r.get('Kernel Name','?'); v=float(r.get('Metric Value','0').replace(',',''))
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print('%-60s n=%4d  %10.1f us  %5.1f%%'%(k[:60],n,t/1e3,100*t/tot))
PY
