#!/bin/bash
# parity tests, then bench with the default traffic and with calm traffic (--fade 0), pipelined kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for f in 0.05 0; do
  timeout 600 python bench.py --fade $f --steps 3 --warmup 3 --no-cpu-baseline ${SELF:---no-selfcheck} --e2e-samples 1e6 > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err
  python - <<PY
import json
b = json.load(open("gpurun_out/bench_f$f.json"))
print("fade=$f value %.0f ms %.2f dev %.2f slicer_stage %.2f kern %.3f frac %.3f tiles %s" % (b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["tiles"]))
PY
done
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_pipe_nofade \
   python bench.py --fade 0 --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof_nofade.log 2>&1
echo "ncu exit $?"
fi
