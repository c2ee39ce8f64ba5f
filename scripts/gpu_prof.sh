#!/bin/bash
# one ncu --set full capture of the slicer kernel + a short bench with path counters
mkdir -p gpurun_out
timeout 600 python bench.py --samples 1e9 --steps 2 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
tail -c 1500 gpurun_out/bench_small.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_kernel -s 2 -c 1 -f -o gpurun_out/prof_slicer \
   python bench.py --samples 2e8 --steps 1 --warmup 1 --no-cpu-baseline --e2e-samples 1e6 > gpurun_out/prof.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
