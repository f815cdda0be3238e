#!/bin/bash
set -u
tag=${1:-r02l}
mkdir -p gpurun_out
for g in 2 4; do
  TSDF_STREAM_GROUPS=$g timeout 900 python -m pytest tests/test_stream_engine.py -x -q -m gpu -k "oracle or exact_only or groups_and" > gpurun_out/${tag}_pytest_g$g.log 2>&1; echo "pytest G=$g rc=$?"; tail -2 gpurun_out/${tag}_pytest_g$g.log
done
L=gpurun_out/${tag}_groups.log
for g in 1 2 4 8; do
  for ep in 8 4; do
    echo "== G=$g epoch=$ep" >> $L; TSDF_STREAM_GROUPS=$g TSDF_STREAM_EPOCH=$ep timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
  done
done
echo "== G=1 epoch=8 again" >> $L; timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
cat $L
