#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_fast \
   python bench.py --samples 1e9 --steps 1 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 > gpurun_out/prof.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
echo "== no fade"
timeout 600 python bench.py --fade 0 --samples 2e9 --steps 2 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print(b['tiles'], 'slicer_ms', b['slicer_ms_per_step'], 'mism', b['seam_mismatches'], 'value', b['value'], 'dev_ms', b['device_ms_per_step'], 'wall_ms', b['ms_per_step'])"
