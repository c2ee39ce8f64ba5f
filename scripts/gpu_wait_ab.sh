#!/bin/bash
# N ranks: spinning against blocking host waits (short lines: no configs, no selfcheck)
N=${1:-8}
mkdir -p gpurun_out
for w in spin blocking; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --wait $w --no-configs --no-cpu-baseline --e2e-samples 1e6 > gpurun_out/bench_wait_$w.json 2> gpurun_out/bench_wait_$w.err
  python - $w <<'PY'
import json,sys
b=json.load(open("gpurun_out/bench_wait_%s.json"%sys.argv[1]))
print(sys.argv[1], "value %.0f ms %.2f dev %.2f"%(b["value"],b["ms_per_step"],b["device_ms_per_step"]), b["sharding"]["rank0_phases_ms_last_step"], b["sharding"]["gathered_via"])
PY
done
