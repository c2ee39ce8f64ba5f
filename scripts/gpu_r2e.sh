#!/bin/bash
# parity tests (fail fast), then the default bench traffic with / without the in-pipeline precise pass, and calm traffic
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu_quick.log 2>&1; rc=$?
echo "quick parity exit $rc"; tail -3 gpurun_out/pytest_gpu_quick.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu_quick.log | head -20; exit 1; fi
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
run() {
  timeout 150 python bench.py --fade $1 --steps 3 --warmup 3 --no-cpu-baseline $3 --e2e-samples 1e6 > gpurun_out/bench_$2.json 2> gpurun_out/bench_$2.err
  python - <<PY
import json
try:
    b = json.load(open("gpurun_out/bench_$2.json"))
    t = b["tiles"]
    print("%-8s value %.0f ms %.2f dev %.2f slicer_stage %.2f frac %.3f self %s | pipe %d runs %d aborts %d redone %d rep %d fix %d exact %d" % ("$2", b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["frac"], b["selfcheck"] and b["selfcheck"]["identical"], t["pipe_tiles"], t["pipe_runs"], t["pipe_aborts"], t["pipe_redone"], t["repeated_passes"], t["fixpoint_tiles"], t["exact_tiles"]))
except Exception as e:
    print("$2: no line", e)
PY
  tail -2 gpurun_out/bench_$2.err
}
run 0.05 redo1 " "
NFC_PIPE_REDO=0 run 0.05 redo0 --no-selfcheck
NFC_PIPE_COOL=1 run 0.05 redo1cool1 --no-selfcheck
run 0 calm --no-selfcheck
