#!/bin/bash
# multi-block Hi-Z lookup + epoch 16: parity suites, timing (1024^3 x 256 / 1000 frames, fragments, 512^3)
set -u
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_stream_engine.py -q -m gpu -x 2>&1 | tail -3
for k in 32 8; do echo "K=$k"; timeout 300 python scripts/tune_variants.py 256 $k 0 2>&1 | tail -1; done
timeout 300 python scripts/tune_variants.py 1000 32 0 2>&1 | tail -1
timeout 300 python scripts/frag_time.py 8 2>&1 | tail -1
timeout 300 python scripts/frag_time.py 16 2>&1 | tail -1
timeout 600 python bench.py --config 512 --no-cpu-baseline --no-parity 2>&1 | tail -1 | cut -c1-400
