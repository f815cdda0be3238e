#!/bin/bash
set -u
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python scripts/realdata_e2e.py 3 > gpurun_out/${tag}_realdata_e2e.json 2> gpurun_out/${tag}_realdata_e2e.err; cat gpurun_out/${tag}_realdata_e2e.json; tail -3 gpurun_out/${tag}_realdata_e2e.err
