#!/bin/bash
# round-2 final state on one GPU: the whole -m gpu suite, smoke(), the default bench line, the reference arm, config 2
set -u
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_1024.json 2> gpurun_out/${tag}_bench_1024.err; cut -c1-300 gpurun_out/${tag}_bench_1024.json
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py --config 512 --no-cpu-baseline > gpurun_out/${tag}_bench_512.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_512.json
ENGINE='regex:stream_|integrate_|tile_mask|fill_kernel|sanitize_depth|build_depth_mips|depth_from_u16|extract_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ENGINE" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_ncu_launch.log 2>&1
python profiles/summarize.py --launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt
cat gpurun_out/${tag}_launch_share.txt
