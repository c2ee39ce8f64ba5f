#!/bin/bash
# configs[3]: streams in flight per GPU
mkdir -p gpurun_out
for w in 4 8 16 32 64; do
  echo "== workers $w"
  timeout 300 python bench.py --batch 512 --batch-workers $w --steps 2 --warmup 1 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('value', round(b['value']), 'ms', round(b['ms_per_step'],1))"
done 2>&1 | tee gpurun_out/batch_sweep.txt
NFC_TIMING=1 timeout 200 python bench.py --batch 8 --batch-workers 1 --steps 1 --warmup 1 2>&1 | tail -40 > gpurun_out/batch_timing.txt
tail -30 gpurun_out/batch_timing.txt
