#!/bin/bash
# slabs per launch of the streaming slicer: bash scripts/gpu_sweep_super.sh "<NFC_SUPER_SLAB> <NFC_SUPER_BALANCE>" ...
mkdir -p gpurun_out
for c in "$@"; do
  set -- $c
  echo "== super_slab $1 balance $2"
  NFC_SUPER_SLAB=$1 NFC_SUPER_BALANCE=$2 timeout 300 python bench.py --samples 1e10 --steps 3 --warmup 2 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 \
    | python -c "import json,sys; b=json.loads(sys.stdin.read()); r=b['roofline']; print('value', round(b['value']), 'wall_ms', round(b['ms_per_step'],2), 'dev_ms', round(b['device_ms_per_step'],2), 'slicer_stage_ms', round(b['slicer_ms_per_step'],2), 'kernel_ms', round(r['avg_launch_ms']*r['launches_per_step'],2), 'frac', round(r['frac'],4), 'mism', b['seam_mismatches'], 'frames', b['frames_per_step'])"
done 2>&1 | tee gpurun_out/sweep_super.txt
