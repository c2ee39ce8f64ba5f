#!/bin/bash
# confirmation after a rebuild: GPU parity tests, smoke, sweep of the slabs per launch, one ncu --set full capture of the streaming slicer
bash scripts/gpu_check.sh
bash scripts/gpu_sweep_super.sh "4 0" "4 1" "5 1" "8 1" "3 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_fast \
   python bench.py --samples 4.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
