#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stream_engine.py -q -m gpu -x -k "oracle or exact_only" 2>&1 | tail -2
for k in 32 8 1; do echo "K=$k"; timeout 300 python scripts/tune_variants.py 256 $k 0 2>&1 | tail -1; done
