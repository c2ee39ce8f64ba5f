#!/bin/bash
# bench lines: the default command, the reference arm, and the other configs' geometries
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "import json; b=json.load(open('gpurun_out/bench.json')); print(b['value'], b['ms_per_step'], b['clocks'], b['roofline']['frac'], b['roofline']['traffic']['dram_bytes_per_sample'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -c 900 gpurun_out/bench_ref.json
bash scripts/gpu_cfgs.sh
