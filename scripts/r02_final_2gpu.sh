#!/bin/bash
set -u
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; cat gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_1024.json 2> gpurun_out/${tag}_bench_1024.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench_1024.json; tail -3 gpurun_out/${tag}_bench_1024.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/${tag}_bench_1024_2gpu.json 2> gpurun_out/${tag}_bench_1024_2gpu.err; echo "bench2 rc=$?"; grep "^{" gpurun_out/${tag}_bench_1024_2gpu.json | cut -c1-1500
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2>/dev/null; cat gpurun_out/${tag}_bench_ref.json
