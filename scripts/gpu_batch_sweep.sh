#!/bin/bash
# configs[3]: streams in flight per GPU, blocking against spinning waits (two rounds: spinning results are bimodal)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_batch.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for r in 1 2; do for w in ${WORKERS:-4 8 16 32}; do for m in "" "--batch-spin"; do
  echo "== round $r workers $w $m"
  timeout 300 python bench.py --batch 512 --batch-workers $w $m --steps 2 --warmup 1 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('value', round(b['value']), 'ms', round(b['ms_per_step'],1))"
done; done; done 2>&1 | tee gpurun_out/batch_sweep.txt
