#!/bin/bash
# bench + launch list (ncu) on one B200.  Results -> gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 2 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --samples 4e8 --steps 1 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 > gpurun_out/bench_ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows=[]
try:
    with open('gpurun_out/launches.csv') as f:
        lines=[l for l in f if not l.startswith('==')]
    rd=csv.DictReader(lines)
    agg=collections.defaultdict(lambda:[0,0.0])
    for r in rd:
        k=r.get('Kernel Name','?'); v=float(r.get('Metric Value','0').replace(',',''))
        agg[k][0]+=1; agg[k][1]+=v
    tot=sum(v[1] for v in agg.values())
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        print('%-60s n=%4d  %10.1f us  %5.1f%%'%(k[:60],n,t/1e3,100*t/tot))
except Exception as e:
    print('launch summary failed',e)
PY
