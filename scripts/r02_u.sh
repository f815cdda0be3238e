#!/bin/bash
# epoch length A/B after the margin change (TSDF_STREAM_EPOCH), 1024^3 x 256 frames and 8 fragments
set -u
for e in 8 11 16 32; do
  echo "epoch=$e"
  TSDF_STREAM_EPOCH=$e timeout 300 python scripts/tune_variants.py 256 32 0 2>&1 | tail -1
  TSDF_STREAM_EPOCH=$e timeout 300 python scripts/frag_time.py 8 2>&1 | tail -1
done
