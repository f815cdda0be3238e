#!/bin/bash
set -u
tag=${1:-r02h}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 > gpurun_out/${tag}_bench_1024_8gpu.json 2> gpurun_out/${tag}_bench_1024_8gpu.err; echo "b1024 rc=$?"; cat gpurun_out/${tag}_bench_1024_8gpu.json; tail -3 gpurun_out/${tag}_bench_1024_8gpu.err
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --config 2048 --steps 2 --warmup 2 --oracle-planes-max 8 > gpurun_out/${tag}_bench_2048_8gpu.json 2> gpurun_out/${tag}_bench_2048_8gpu.err; echo "b2048 rc=$?"; cat gpurun_out/${tag}_bench_2048_8gpu.json; tail -3 gpurun_out/${tag}_bench_2048_8gpu.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --config frag300 > gpurun_out/${tag}_bench_frag300_8gpu.json 2> gpurun_out/${tag}_bench_frag300_8gpu.err; echo "bfrag rc=$?"; cat gpurun_out/${tag}_bench_frag300_8gpu.json; tail -3 gpurun_out/${tag}_bench_frag300_8gpu.err
timeout 600 python scripts/demo_gpus_check.py 1024 8 48 > gpurun_out/${tag}_demo_gpus8.json 2> gpurun_out/${tag}_demo_gpus8.err; echo "demo rc=$?"; cat gpurun_out/${tag}_demo_gpus8.json; tail -3 gpurun_out/${tag}_demo_gpus8.err
timeout 300 $TR --master-port 29524 bench.py --gpus 8 --impl reference > gpurun_out/${tag}_bench_ref_8.json 2> /dev/null; cat gpurun_out/${tag}_bench_ref_8.json
