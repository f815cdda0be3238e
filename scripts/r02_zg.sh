#!/bin/bash
# configs 4 and 5 in full on ONE GPU with the final kernels (the 8-GPU lines of these configs predate them)
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --config 2048 --steps 2 --warmup 1 --no-cpu-baseline --oracle-planes-max 2 > gpurun_out/r02zg_bench_2048_1gpu.json 2> gpurun_out/r02zg_bench_2048_1gpu.err; echo "rc=$?"; cut -c1-400 gpurun_out/r02zg_bench_2048_1gpu.json; tail -2 gpurun_out/r02zg_bench_2048_1gpu.err
timeout 600 python bench.py --config frag300 --no-cpu-baseline > gpurun_out/r02zg_bench_frag300_1gpu.json 2> gpurun_out/r02zg_bench_frag300_1gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/r02zg_bench_frag300_1gpu.json; tail -2 gpurun_out/r02zg_bench_frag300_1gpu.err
