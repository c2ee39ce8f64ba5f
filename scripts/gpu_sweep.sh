#!/bin/bash
# A/B of the pipelined mode's tunables on the bench traffic (kernel alone and step)
mkdir -p gpurun_out
run() {
  env $1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-selfcheck --no-configs --e2e-samples 1e6 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$1" <<'PY'
import json,sys
try:
    b=json.load(open("gpurun_out/sw.json")); t=b["tiles"]
    print("%-44s kern %.3f ms frac %.3f step %.2f ms | pipe_tiles %d runs %d aborts %d rep %d fix %d"%(sys.argv[1], b["roofline"]["avg_launch_ms"], b["roofline"]["frac"], b["ms_per_step"], t["pipe_tiles"], t["pipe_runs"], t["pipe_aborts"], t["repeated_passes"], t["fixpoint_tiles"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run "NFC_X=0"
run "NFC_PIPE_COOL=0"
run "NFC_PIPE_COOL=1"
run "NFC_PIPE_COOL=4"
run "NFC_PIPE_MIN=3"
run "NFC_PIPE_MIN=12"
run "NFC_PIPE_COOL=1 NFC_PIPE_MIN=3"
run "NFC_SLICER_STAGES=2"
run "NFC_SUPER_SLAB=8"
run "NFC_SUPER_SLAB=3"
