#!/bin/bash
# N-GPU bench line as the driver launches it (torchrun), incl. the configs legs
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
grep -v "^$" gpurun_out/bench_n$N.err | tail -8
python - <<PY
import json
b=json.load(open("gpurun_out/bench_n$N.json"))
print("N=%d value %.0f ms %.2f dev %.2f frac %.3f step_frac %.3f e2e %.0f"%(b["n_gpus"],b["value"],b["ms_per_step"],b["device_ms_per_step"],b["roofline"]["frac"],b["roofline"]["step_frac"],b["e2e"]["value"]))
print("sharding", b["sharding"])
print("per_rank", b["per_rank"])
print("c5 phases", (b.get("configs") or {}).get("c5_one_capture_20MS_strong", {}).get("ms"))
print("h2d", b["e2e"].get("h2d_ceiling"))
for k,v in (b.get("configs") or {}).items():
    print(k, {kk:vv for kk,vv in v.items() if kk not in ("workload","clocks")})
PY
