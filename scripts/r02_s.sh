#!/bin/bash
# margin tightening: self-test, parity suites, timing; prefetch A/B
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "margin or rounding or packed" 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_stream_engine.py -q -m gpu -x 2>&1 | tail -3
for k in 32 8; do echo "K=$k"; timeout 300 python scripts/tune_variants.py 256 $k 0 2>&1 | tail -1; done
echo "validation lib, production variant / prefetch variant"
TSDF_USE_VALIDATION=1 timeout 300 python scripts/tune_variants.py 256 32 0 2>&1 | tail -1
TSDF_USE_VALIDATION=1 TSDF_STREAM_DIAG=3 timeout 300 python scripts/tune_variants.py 256 32 0 2>&1 | tail -1
