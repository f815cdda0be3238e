#!/bin/bash
# mask pre-pass with the rolled warp-tile loop: parity, timing; occupancy variants (lib_variants/*.so swapped in)
set -u
timeout 900 python -m pytest tests/test_stream_engine.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python scripts/mask_time.py 1000 2>&1 | tail -1
timeout 300 python scripts/tune_variants.py 1000 32 0 2>&1 | tail -1
timeout 300 python scripts/frag_time.py 16 2>&1 | tail -1
cp 3dmatch-toolbox_b200/lib/libtsdf_fusion.so /tmp/orig.so
for v in rolled_minb2 rolled_minb5 rolled_minb6; do
  cp 3dmatch-toolbox_b200/lib_variants/$v.so 3dmatch-toolbox_b200/lib/libtsdf_fusion.so
  echo "== $v"
  timeout 300 python scripts/mask_time.py 1000 2>&1 | tail -1
  timeout 300 python scripts/tune_variants.py 1000 32 0 2>&1 | tail -1
done
cp /tmp/orig.so 3dmatch-toolbox_b200/lib/libtsdf_fusion.so
