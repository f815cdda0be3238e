#!/bin/bash
set -u
timeout 300 python scripts/e2e_breakdown.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_stream_engine.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
