#!/bin/bash
# the build as it stands: smoke, the complete bench line (sub-configs, CPU baseline), the reference arm, then the ncu launch list
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench.json"))
print("value %.0f ms %.2f dev %.2f frac %.3f step_frac %.3f e2e %.0f (%.2f ms) self %s windows %s"%(b["value"],b["ms_per_step"],b["device_ms_per_step"],b["roofline"]["frac"],b["roofline"]["step_frac"],b["e2e"]["value"], b["e2e"]["ms_per_step"], b["selfcheck"]["identical"], b["selfcheck"].get("parity_windows",{}).get("identical")))
for k,v in b.get("configs",{}).items(): print(k, v.get("ms"), v.get("value"), v.get("frac"))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-selfcheck --no-configs --e2e-samples 1e6 > gpurun_out/launches.log 2>&1
echo "ncu launch list exit $?"
