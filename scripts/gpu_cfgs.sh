#!/bin/bash
# bench lines of the other BASELINE.json configs on one B200: configs[4] geometry (20 MS/s) and configs[3] (batch of captures, one GPU's share)
mkdir -p gpurun_out
timeout 400 python bench.py --rate 20e6 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_20ms.json 2> gpurun_out/bench_20ms.err
echo "20ms exit $?"; tail -c 1500 gpurun_out/bench_20ms.json
timeout 400 python bench.py --batch 512 --steps 2 --warmup 1 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err
echo "batch exit $?"; tail -c 1500 gpurun_out/bench_batch.json
