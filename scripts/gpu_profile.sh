#!/bin/bash
# ncu evidence of one command: launch list (per-launch durations) of a whole step, then one --set full capture of the
# dominant kernel (second launch: the first one follows the allocations).  Summarise with scripts/summarize_profile.py.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-selfcheck --no-configs --e2e-samples 1e6 > gpurun_out/launches.log 2>&1
echo "ncu launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_fast \
   python bench.py --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --no-configs --e2e-samples 1e6 > gpurun_out/prof.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
