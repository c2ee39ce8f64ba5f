#!/bin/bash
# A/B of two builds of the library on calm and default traffic; ncu --set full of the current build on calm traffic
mkdir -p gpurun_out
run() {
  timeout 150 python bench.py --fade $1 --steps 3 --warmup 3 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/bench_$2.json 2> gpurun_out/bench_$2.err
  python - <<PY
import json
try:
    b = json.load(open("gpurun_out/bench_$2.json"))
    t = b["tiles"]
    print("%-10s value %.0f ms %.2f dev %.2f slicer_stage %.2f kern %.3f frac %.3f | pipe %d runs %d aborts %d redone %d rep %d" % ("$2", b["value"], b["ms_per_step"], b["device_ms_per_step"], b["slicer_ms_per_step"], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], t["pipe_tiles"], t["pipe_runs"], t["pipe_aborts"], t.get("pipe_redone", 0), t["repeated_passes"]))
except Exception as e:
    print("$2: no line", e)
PY
}
run 0 calm_new
USRP_NFC_B200_LIB=$PWD/build_variants/lib_8e63db8.so run 0 calm_old
run 0 calm_new2
USRP_NFC_B200_LIB=$PWD/build_variants/lib_8e63db8.so run 0 calm_old2
run 0.05 fade_new
USRP_NFC_B200_LIB=$PWD/build_variants/lib_8e63db8.so run 0.05 fade_old
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slicer_fast -s 1 -c 1 -f -o gpurun_out/prof_calm_new \
   python bench.py --fade 0 --samples 5.3e9 --steps 1 --warmup 1 --no-cpu-baseline --no-selfcheck --e2e-samples 1e6 > gpurun_out/prof_calm_new.log 2>&1
echo "ncu exit $?"
