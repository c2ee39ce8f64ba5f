#!/bin/bash
# configs[3]: hardware queues (CUDA_DEVICE_MAX_CONNECTIONS: every nfc_stream owns three CUDA streams, the default is 8 queues)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_batch.py tests/test_gpu_blocks.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for c in 8 32; do for w in 4 8 16; do
  echo "== connections $c workers $w"
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 120 python scripts/t_batch_split.py $w 2>&1 | tail -1
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python bench.py --batch 512 --batch-workers $w --steps 2 --warmup 1 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('value', round(b['value']), 'ms', round(b['ms_per_step'],1))"
done; done 2>&1 | tee gpurun_out/batch_conn.txt
