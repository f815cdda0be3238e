#!/bin/bash
set -u
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tdf.py tests/test_stream_engine.py -x -q -m gpu > gpurun_out/${tag}_pytest_tdf.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest_tdf.log
timeout 600 python scripts/tdf_bench.py > gpurun_out/${tag}_tdf.json 2> gpurun_out/${tag}_tdf.err; cat gpurun_out/${tag}_tdf.json; tail -3 gpurun_out/${tag}_tdf.err
timeout 600 ncu --set full --clock-control none -k regex:tdf_pass -c 3 -f -o /tmp/${tag}_tdf python scripts/tdf_bench.py > /dev/null 2>&1
python profiles/summarize.py /tmp/${tag}_tdf.ncu-rep > gpurun_out/${tag}_tdf_kernels.txt; grep -E "^## |duration|dram__bytes|dram_throughput" gpurun_out/${tag}_tdf_kernels.txt
timeout 600 python scripts/realdata_e2e.py 3 > gpurun_out/${tag}_realdata_e2e.json 2>/dev/null; cat gpurun_out/${tag}_realdata_e2e.json
# per-warp phase B A/B (same box): parity tests under the switch, then timings
TSDF_STREAM_WARPB=1 timeout 900 python -m pytest tests/test_stream_engine.py -x -q -m gpu > gpurun_out/${tag}_pytest_warpb.log 2>&1; echo "pytest warpb rc=$?"; tail -3 gpurun_out/${tag}_pytest_warpb.log
for ep in 8 4; do
  echo "== block phase B, epoch $ep" >> gpurun_out/${tag}_warpb.log; TSDF_STREAM_EPOCH=$ep timeout 300 python scripts/tune_variants.py 256 32 0 >> gpurun_out/${tag}_warpb.log 2>&1
  echo "== warp phase B, epoch $ep" >> gpurun_out/${tag}_warpb.log; TSDF_STREAM_WARPB=1 TSDF_STREAM_EPOCH=$ep timeout 300 python scripts/tune_variants.py 256 32 0 >> gpurun_out/${tag}_warpb.log 2>&1
done
echo "== warp phase B, epoch 2" >> gpurun_out/${tag}_warpb.log; TSDF_STREAM_WARPB=1 TSDF_STREAM_EPOCH=2 timeout 300 python scripts/tune_variants.py 256 32 0 >> gpurun_out/${tag}_warpb.log 2>&1
echo "== warp phase B, epoch 16" >> gpurun_out/${tag}_warpb.log; TSDF_STREAM_WARPB=1 TSDF_STREAM_EPOCH=16 timeout 300 python scripts/tune_variants.py 256 32 0 >> gpurun_out/${tag}_warpb.log 2>&1
cat gpurun_out/${tag}_warpb.log
timeout 600 python scripts/plan_calibrate.py 256 > gpurun_out/${tag}_plan_calibrate.json 2> gpurun_out/${tag}_plan_calibrate.err; cat gpurun_out/${tag}_plan_calibrate.json; tail -3 gpurun_out/${tag}_plan_calibrate.err
