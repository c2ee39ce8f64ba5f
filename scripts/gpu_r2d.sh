#!/bin/bash
# parity tests, then the default bench traffic with different cool-down settings of the pipelined mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() {
  timeout 600 python bench.py --fade $1 --steps 3 --warmup 3 --no-cpu-baseline ${3:---no-selfcheck} --e2e-samples 1e6 > gpurun_out/bench_$2.json 2> gpurun_out/bench_$2.err
  python - <<PY
import json
b = json.load(open("gpurun_out/bench_$2.json"))
print("$2 fade=$1 value %.0f ms %.2f dev %.2f slicer_stage %.2f frac %.3f self %s tiles %s" % (b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["frac"], b["selfcheck"] and b["selfcheck"]["identical"], b["tiles"]))
PY
}
run 0.05 default " "
NFC_PIPE_COOL=0 run 0.05 cool0
NFC_PIPE_COOL=1 run 0.05 cool1
NFC_PIPE_COOL=4 run 0.05 cool4
NFC_PIPE_COOL=8 NFC_PIPE_MIN=12 run 0.05 cool8
run 0 calm
