#!/bin/bash
# round-2 final state on one GPU: the whole -m gpu suite, smoke(), the default bench line, config 2, config 5 (32 fragments)
set -u
tag=${1:-r02zd}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_1024.json 2> gpurun_out/${tag}_bench_1024.err; cut -c1-300 gpurun_out/${tag}_bench_1024.json
timeout 600 python bench.py --config 512 --no-cpu-baseline > gpurun_out/${tag}_bench_512.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_512.json
timeout 600 python bench.py --config frag300 --fragments 32 --no-cpu-baseline > gpurun_out/${tag}_bench_frag32.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_frag32.json
