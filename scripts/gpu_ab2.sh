#!/bin/bash
# fail-fast parity, then A/B of the current build against an older one on calm and default traffic
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
    print("%-10s value %.0f ms %.2f dev %.2f slicer_stage %.2f kern %.3f frac %.3f self %s | pipe %d runs %d aborts %d redone %d rep %d fix %d" % ("$2", b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["selfcheck"] and b["selfcheck"]["identical"], t["pipe_tiles"], t["pipe_runs"], t["pipe_aborts"], t.get("pipe_redone", 0), t["repeated_passes"], t["fixpoint_tiles"]))
except Exception as e:
    print("$2: no line", e)
PY
  tail -2 gpurun_out/bench_$2.err
}
run 0 calm_new --no-selfcheck
USRP_NFC_B200_LIB=$PWD/build_variants/lib_8e63db8.so run 0 calm_old --no-selfcheck
run 0.05 fade_new " "
USRP_NFC_B200_LIB=$PWD/build_variants/lib_8e63db8.so run 0.05 fade_old --no-selfcheck
NFC_PIPE_REDO=0 run 0.05 fade_new_noredo --no-selfcheck
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_calm_new \
   python bench.py --fade 0 --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof_calm_new.log 2>&1
echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_fade_new \
   python bench.py --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof_fade_new.log 2>&1
echo "ncu exit $?"
fi
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
