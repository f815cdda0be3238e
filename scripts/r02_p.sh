#!/bin/bash
set -u
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_stream_engine.py -q -m gpu -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for ep in 8 6 10 12; do echo "epoch $ep"; TSDF_STREAM_EPOCH=$ep timeout 300 python scripts/tune_variants.py 256 32 0 2>&1 | tail -1; done
timeout 300 python bench.py --config frag300 --fragments 32 --frag-batch 8 --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('frag B=8 value %.3e ms %.2f' % (j['value'], j['ms_per_step']))"
