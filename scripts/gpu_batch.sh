#!/bin/bash
# one-pass batch: parity per capture against the oracle, then the batch bench (one pass and the legacy per-stream path)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/pytest_batch.log 2>&1; rc=$?
echo "batch tests exit $rc"; tail -25 gpurun_out/pytest_batch.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu_quick.log 2>&1; echo "parity exit $?"; tail -2 gpurun_out/pytest_gpu_quick.log
timeout 300 python bench.py --batch 512 --steps 3 --warmup 2 > gpurun_out/bench_batch512.json 2> gpurun_out/bench_batch512.err; echo "batch bench exit $?"
python -c "import json; b=json.load(open('gpurun_out/bench_batch512.json')); print('one pass: value %.0f Msamples/s, %.2f ms/step, frames %d' % (b['value'], b['ms_per_step'], b['frames_per_step']))"; tail -3 gpurun_out/bench_batch512.err
timeout 300 python bench.py --batch 512 --batch-legacy --steps 2 --warmup 1 > gpurun_out/bench_batch512_legacy.json 2> gpurun_out/bench_batch512_legacy.err
python -c "import json; b=json.load(open('gpurun_out/bench_batch512_legacy.json')); print('legacy: value %.0f Msamples/s, %.2f ms/step, frames %d' % (b['value'], b['ms_per_step'], b['frames_per_step']))"
