#!/bin/bash
# session 2 re-entry: full GPU suite (incl. hardening), default bench line with timing breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
cat gpurun_out/bench_default.json | cut -c1-3000
NFC_TIMING=1 timeout 150 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
tail -40 gpurun_out/bench_timing.err
