#!/bin/bash
# calm traffic (no fade: no tile needs the precise passes): the pipelined mode's own rate against the synchronous loop's
mkdir -p gpurun_out
for v in 1 0; do
  NFC_SLICER_PIPE=$v timeout 600 python bench.py --fade 0 --steps 3 --warmup 3 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/bench_nofade_pipe$v.json 2> gpurun_out/bench_nofade_pipe$v.err
  python - <<PY
import json
b = json.load(open("gpurun_out/bench_nofade_pipe$v.json"))
print("nofade pipe=$v value %.0f ms %.2f dev %.2f slicer_stage %.2f kern %.3f frac %.3f tiles %s" % (b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["tiles"]))
PY
done
NFC_SLICER_STAGES=2 timeout 600 python bench.py --fade 0 --steps 3 --warmup 3 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/bench_nofade_s2.json 2>/dev/null
python -c "import json; b=json.load(open('gpurun_out/bench_nofade_s2.json')); print('nofade stages=2 value %.0f kern %.3f frac %.3f' % (b['value'], b['roofline']['avg_launch_ms'], b['roofline']['frac']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_pipe_nofade \
   python bench.py --fade 0 --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof_nofade.log 2>&1
echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_pipe_fade \
   python bench.py --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof_fade.log 2>&1
echo "ncu exit $?"
