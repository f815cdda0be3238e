#!/bin/bash
set -u
timeout 400 python scripts/e2e_breakdown.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_stream_engine.py tests/test_gpu_parity.py tests/test_cli.py -q -m gpu -x 2>&1 | tail -2
