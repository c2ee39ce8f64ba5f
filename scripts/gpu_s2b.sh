#!/bin/bash
# bench line with the configs block (2 / 20 MS/s, batch, C5) on one GPU
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; echo "bench exit $?"
tail -5 gpurun_out/bench_cfg.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_cfg.json"))
print("value %.0f ms %.2f frac %.3f step_frac %.3f e2e %.0f"%(b["value"],b["ms_per_step"],b["roofline"]["frac"],b["roofline"]["step_frac"],b["e2e"]["value"]))
print("h2d", b["e2e"].get("h2d_ceiling"))
print("parity_windows", b["selfcheck"].get("parity_windows"))
for k,v in (b.get("configs") or {}).items():
    print(k, {kk:vv for kk,vv in v.items() if kk not in ("workload","clocks")})
    print("   clocks", v.get("clocks"))
PY
