#!/bin/bash
set -u
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stream_engine.py -x -q > gpurun_out/${tag}_pytest_stream.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_pytest_stream.log
for minb in 5 6; do
  for ep in 4 8 16 32; do
    echo "== MINB=$minb K=32 EPOCH=$ep" >> gpurun_out/${tag}_sweep.log
    TSDF_STREAM_MINB=$minb TSDF_STREAM_EPOCH=$ep timeout 300 python scripts/tune_variants.py 256 32 0 >> gpurun_out/${tag}_sweep.log 2>&1
  done
done
echo "== MINB=5 K=16 EPOCH=8" >> gpurun_out/${tag}_sweep.log
TSDF_STREAM_EPOCH=8 timeout 300 python scripts/tune_variants.py 256 16 0 >> gpurun_out/${tag}_sweep.log 2>&1
cat gpurun_out/${tag}_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_integrate --launch-skip 1 --launch-count 1 \
    -f -o /tmp/${tag}_stream python scripts/tune_variants.py 64 32 0 > gpurun_out/${tag}_ncu_stream.log 2>&1
python profiles/summarize.py /tmp/${tag}_stream.ncu-rep > gpurun_out/${tag}_stream_k32.txt
ncu -i /tmp/${tag}_stream.ncu-rep --page source --csv > gpurun_out/${tag}_stream_src_sass.csv 2>/dev/null
cat gpurun_out/${tag}_stream_k32.txt
