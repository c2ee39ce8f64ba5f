#!/bin/bash
# full GPU suite, then the bench line (no CPU baseline, no sub-configs) and the C4 batch sub-run
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -8 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu.log | head -30; fi
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_async.json 2> gpurun_out/bench_async.err; echo "bench exit $?"
tail -3 gpurun_out/bench_async.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_async.json"))
print("value %.0f ms %.2f dev %.2f frac %.3f step_frac %.3f e2e %.0f (%.2f ms) self %s windows %s"%(b["value"],b["ms_per_step"],b["device_ms_per_step"],b["roofline"]["frac"],b["roofline"]["step_frac"],b["e2e"]["value"], b["e2e"]["ms_per_step"], b["selfcheck"]["identical"], b["selfcheck"].get("parity_windows",{}).get("identical")))
PY
