#!/bin/bash
# correctness + a short bench + instruction count of the slicer launches
mkdir -p gpurun_out
bash scripts/gpu_check.sh
timeout 600 python bench.py --samples 2e9 --steps 2 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print(b['tiles'], 'slicer_ms', b['slicer_ms_per_step'], 'mism', b['seam_mismatches'], 'value', b['value'], 'dev_ms', b['device_ms_per_step'])"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:slicer_kernel -s 2 -c 2 --csv --log-file gpurun_out/inst.csv \
    python bench.py --samples 1e9 --steps 1 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 > /dev/null 2>&1
grep -v "^==" gpurun_out/inst.csv | cut -d, -f5,9,12- | tail -8
