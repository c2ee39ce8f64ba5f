#!/bin/bash
# queued slab chains: full GPU suite, then the bench line with per-slab timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -15 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu.log | head -30; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_async.json 2> gpurun_out/bench_async.err; echo "bench exit $?"
tail -3 gpurun_out/bench_async.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_async.json"))
print("value %.0f ms %.2f dev %.2f frac %.3f step_frac %.3f e2e %.0f self %s windows %s retries?"%(b["value"],b["ms_per_step"],b["device_ms_per_step"],b["roofline"]["frac"],b["roofline"]["step_frac"],b["e2e"]["value"], b["selfcheck"]["identical"], b["selfcheck"].get("parity_windows",{}).get("identical")))
PY
NFC_TIMING=1 timeout 150 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-selfcheck --no-configs --e2e-samples 1e6 > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep slab gpurun_out/bench_timing.err | tail -24
NFC_POST_SYNC=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-selfcheck --e2e-samples 1e6 > gpurun_out/bench_sync.json 2> gpurun_out/bench_sync.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_sync.json"))
print("NFC_POST_SYNC=1: value %.0f ms %.2f dev %.2f"%(b["value"],b["ms_per_step"],b["device_ms_per_step"]))
PY
