#!/bin/bash
# end-of-round confirmation: GPU parity tests, smoke, the default bench line
bash scripts/gpu_check.sh
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "import json; b=json.load(open('gpurun_out/bench.json')); print(b['value'], b['ms_per_step'], b['clocks'], b['roofline']['frac'], b['selfcheck']['identical'], b['e2e']['value'])"
timeout 300 python bench.py --batch 512 --steps 2 --warmup 1 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err
echo "batch exit $?"; tail -c 400 gpurun_out/bench_batch.json
