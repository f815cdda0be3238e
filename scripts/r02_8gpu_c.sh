#!/bin/bash
set -u
tag=${1:-r02zb}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 > gpurun_out/${tag}_bench_1024_8gpu.json 2> gpurun_out/${tag}_bench_1024_8gpu.err; echo "b8 rc=$?"; grep "^{" gpurun_out/${tag}_bench_1024_8gpu.json | cut -c1-1800; tail -3 gpurun_out/${tag}_bench_1024_8gpu.err
