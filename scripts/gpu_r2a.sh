#!/bin/bash
# round 2, first contact of the pipelined streaming slicer with a GPU: parity tests, smoke, bench A/B (pipelined / synchronous)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
for v in 1 0; do
  NFC_SLICER_PIPE=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipe$v.json 2> gpurun_out/bench_pipe$v.err
  echo "bench pipe=$v exit $?"
  python - <<PY
import json
try:
    b = json.load(open("gpurun_out/bench_pipe$v.json"))
    print("pipe=$v value %.0f ms %.2f dev %.2f slicer_stage %.2f kern %.3f frac %.3f self %s tiles %s" % (b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["selfcheck"] and b["selfcheck"]["identical"], b["tiles"]))
except Exception as e:
    print("no line:", e)
PY
  tail -3 gpurun_out/bench_pipe$v.err
done
NFC_SLICER_STAGES=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-selfcheck > gpurun_out/bench_pipe_s2.json 2> gpurun_out/bench_pipe_s2.err
python -c "import json; b=json.load(open('gpurun_out/bench_pipe_s2.json')); print('stages=2 value %.0f kern %.3f frac %.3f' % (b['value'], b['roofline']['avg_launch_ms'], b['roofline']['frac']))"
