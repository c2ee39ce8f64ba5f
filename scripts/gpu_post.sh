#!/bin/bash
# per-kernel time, DRAM and SM utilisation of everything but the slicer
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:'chunk_|run_|extract_|scan_|seam_' -c 60 --csv --log-file gpurun_out/post.csv \
  python bench.py --samples 1.07e9 --steps 1 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 > gpurun_out/post.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/post.csv') if not l.startswith('==')]
rd=csv.DictReader(lines)
agg=collections.OrderedDict()
for r in rd:
    k=(r['ID'], r['Kernel Name'][:40])
    agg.setdefault(k,{})[r['Metric Name']]=r['Metric Value']
for (i,k),m in list(agg.items())[:60]:
    print('%3s %-40s grid %7s  %9s us  dram %5s%%  sm %5s%%  inst %10s  rd %8s wr %8s'%(i,k,m.get('launch__grid_size'),m.get('gpu__time_duration.sum'),m.get('dram__throughput.avg.pct_of_peak_sustained_elapsed'),m.get('sm__throughput.avg.pct_of_peak_sustained_elapsed'),m.get('smsp__inst_executed.sum'),m.get('dram__bytes_read.sum'),m.get('dram__bytes_write.sum')))
PY
