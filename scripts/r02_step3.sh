#!/bin/bash
set -u
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_stream_engine.py -x -q > gpurun_out/${tag}_pytest_stream.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_pytest_stream.log
timeout 600 python scripts/slab_overhead.py 256 8 32 > gpurun_out/${tag}_slab.log 2>&1; cat gpurun_out/${tag}_slab.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_1024.json 2> gpurun_out/${tag}_bench_1024.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench_1024.json; tail -5 gpurun_out/${tag}_bench_1024.err
timeout 600 python bench.py --config frag300 --fragments 32 --no-cpu-baseline > gpurun_out/${tag}_bench_frag32.json 2> gpurun_out/${tag}_bench_frag32.err; echo "frag rc=$?"; cat gpurun_out/${tag}_bench_frag32.json; tail -5 gpurun_out/${tag}_bench_frag32.err
