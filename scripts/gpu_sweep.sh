#!/bin/bash
# A/B of the streaming kernel's tunables on the bench traffic (kernel alone and step): pass `VAR=value ...` sets as arguments
mkdir -p gpurun_out
run() {
  env $1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-selfcheck --no-configs --e2e-samples 1e6 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$1" <<'PY'
import json,sys
try:
    b=json.load(open("gpurun_out/sw.json")); t=b["tiles"]
    print("%-44s kern %.3f ms frac %.3f step %.2f ms | pipe_tiles %d runs %d aborts %d rep %d fix %d resum %d"%(sys.argv[1], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["ms_per_step"], t["pipe_tiles"], t["pipe_runs"], t["pipe_aborts"], t["repeated_passes"], t["fixpoint_tiles"], t["ring_resums"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
if [ $# -eq 0 ]; then set -- "NFC_X=0" "NFC_PIPE_COOL=0" "NFC_PIPE_COOL=1" "NFC_PIPE_COOL=4" "NFC_PIPE_MIN=3" "NFC_PIPE_MIN=12" "NFC_SLICER_STAGES=2" "NFC_MEAS_MAX=3" "NFC_MEAS_MAX=4"; fi
for v in "$@"; do run "$v"; done
