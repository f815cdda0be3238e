#!/bin/bash
# first GPU contact of the stream engine: smoke, sanitizer, parity tests, A/B against the legacy launch sequence
set -u
mkdir -p gpurun_out
timeout 300 python scripts/sanitizer_smoke.py > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02_smoke.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitizer_smoke.py > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_memcheck.log
tail -5 gpurun_out/r02_memcheck.log
timeout 900 python -m pytest tests/test_stream_engine.py -x -q > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest_stream.log
timeout 300 python bench.py --quick --frames 256 --steps 2 --warmup 2 > gpurun_out/r02_quick_stream.json 2> gpurun_out/r02_quick_stream.err; cat gpurun_out/r02_quick_stream.json
TSDF_ENGINE=legacy timeout 300 python bench.py --quick --frames 256 --steps 2 --warmup 2 --batch 8 > gpurun_out/r02_quick_legacy.json 2> gpurun_out/r02_quick_legacy.err; cat gpurun_out/r02_quick_legacy.json
