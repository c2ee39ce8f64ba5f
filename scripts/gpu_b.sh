#!/bin/bash
# short benches: bash scripts/gpu_b.sh "<bench args>" ...
for a in "$@"; do
  echo "== $a"
  timeout 600 python bench.py --samples 2e9 --steps 2 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 $a | python -c "import json,sys; b=json.loads(sys.stdin.read()); print(b['tiles'], 'slicer_ms', b['slicer_ms_per_step'], 'mism', b['seam_mismatches'], 'value', b['value'], 'dev_ms', b['device_ms_per_step'], 'wall_ms', b['ms_per_step'])"
done
