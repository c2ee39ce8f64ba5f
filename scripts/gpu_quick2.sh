#!/bin/bash
# correctness + short benches of the slicer variants
mkdir -p gpurun_out
bash scripts/gpu_check.sh
for mb in ${MINBS:-3 2 4}; do
  echo "== MINB=$mb"
  NFC_SLICER_MINB=$mb timeout 600 python bench.py --samples 2e9 --steps 2 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print(b['tiles'], 'slicer_ms', b['slicer_ms_per_step'], 'mism', b['seam_mismatches'], 'value', b['value'], 'dev_ms', b['device_ms_per_step'], 'wall_ms', b['ms_per_step'])"
done
