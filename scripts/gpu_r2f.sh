#!/bin/bash
# fail-fast parity, bench on default and calm traffic, per-slab stage timing, ncu launch list of one step
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu_quick.log 2>&1; rc=$?
echo "quick parity exit $rc"; tail -3 gpurun_out/pytest_gpu_quick.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu_quick.log | head -20; exit 1; fi
run() {
  timeout 150 python bench.py --fade $1 --steps 3 --warmup 3 --no-cpu-baseline $3 --e2e-samples 1e6 > gpurun_out/bench_$2.json 2> gpurun_out/bench_$2.err
  python - <<PY
import json
try:
    b = json.load(open("gpurun_out/bench_$2.json"))
    t = b["tiles"]
    print("%-10s value %.0f ms %.2f dev %.2f slicer_stage %.2f kern %.3f frac %.3f self %s | pipe %d runs %d aborts %d rep %d fix %d" % ("$2", b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["selfcheck"] and b["selfcheck"]["identical"], t["pipe_tiles"], t["pipe_runs"], t["pipe_aborts"], t["repeated_passes"], t["fixpoint_tiles"]))
except Exception as e:
    print("$2: no line", e)
PY
  tail -2 gpurun_out/bench_$2.err
}
run 0 calm --no-selfcheck
run 0.05 fade " "
NFC_TIMING=1 timeout 150 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
tail -25 gpurun_out/bench_timing.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/launches.log 2>&1
echo "ncu launches exit $?"
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
