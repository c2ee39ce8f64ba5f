#!/bin/bash
# N-GPU run: torchrun bench at N ranks (weak scaling over time shards)
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 2 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
