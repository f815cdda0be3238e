#!/bin/bash
set -u
tag=${1:-r02e}
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/${tag}_pytest_mgpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${tag}_pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/${tag}_bench_1024_2gpu.json 2> gpurun_out/${tag}_bench_1024_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/${tag}_bench_1024_2gpu.json; tail -5 gpurun_out/${tag}_bench_1024_2gpu.err
timeout 600 python bench.py --config 512 --no-cpu-baseline > gpurun_out/${tag}_bench_512.json 2> gpurun_out/${tag}_bench_512.err; echo "b512 rc=$?"; cat gpurun_out/${tag}_bench_512.json; tail -3 gpurun_out/${tag}_bench_512.err
