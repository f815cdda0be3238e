#!/bin/bash
# round-2 final state on two GPUs: the multi-GPU tests and the 2-GPU bench line; then the 1-GPU line again (per-K modes)
set -u
tag=${1:-r02za}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_cli.py -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_multi.log; cat gpurun_out/${tag}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/${tag}_bench_1024_2gpu.json 2> gpurun_out/${tag}_bench_1024_2gpu.err; echo "bench2 rc=$?"; grep "^{" gpurun_out/${tag}_bench_1024_2gpu.json | cut -c1-1200
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_1024.json 2> gpurun_out/${tag}_bench_1024.err; cut -c1-200 gpurun_out/${tag}_bench_1024.json
