#!/bin/bash
# official artefact set: bench line, launch list, one ncu --set full capture of the streaming slicer
bash scripts/gpu_bench.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_fast \
   python bench.py --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
