#!/bin/bash
set -u
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --oracle-planes-max 3 > gpurun_out/r02zh_bench_1024_2gpu.json 2> gpurun_out/r02zh_bench_1024_2gpu.err; echo "bench2 rc=$?"; grep "^{" gpurun_out/r02zh_bench_1024_2gpu.json | cut -c1-600
