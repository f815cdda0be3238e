#!/bin/bash
# ncu evidence + tuning sweeps of the stream engine (run on the GPU box; text summaries only come back)
set -u
tag=${1:-r02}
mkdir -p gpurun_out
# 1. sweeps (same box, same inputs): chunk size and resident blocks per SM
for minb in 6 5 4; do
  for k in 32 16 8; do
    echo "== MINB=$minb K=$k" >> gpurun_out/${tag}_sweep.log
    TSDF_STREAM_MINB=$minb timeout 300 python scripts/tune_variants.py 256 $k 0 >> gpurun_out/${tag}_sweep.log 2>&1
  done
done
echo "== legacy K=8" >> gpurun_out/${tag}_sweep.log
TSDF_ENGINE=legacy timeout 300 python scripts/tune_variants.py 256 8 0 >> gpurun_out/${tag}_sweep.log 2>&1
cat gpurun_out/${tag}_sweep.log
# 2. ncu full capture of one persistent launch (64 frames = 2 chunks; second call of the script = warm)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_integrate --launch-skip 1 --launch-count 1 \
    -f -o /tmp/${tag}_stream python scripts/tune_variants.py 64 32 0 > gpurun_out/${tag}_ncu_stream.log 2>&1
python profiles/summarize.py /tmp/${tag}_stream.ncu-rep > gpurun_out/${tag}_stream_k32.txt
ncu -i /tmp/${tag}_stream.ncu-rep --page source --csv --print-source cuda > gpurun_out/${tag}_stream_src_cuda.csv 2>/dev/null
ncu -i /tmp/${tag}_stream.ncu-rep --page source --csv > gpurun_out/${tag}_stream_src_sass.csv 2>/dev/null
cat gpurun_out/${tag}_stream_k32.txt
# 3. launch list of a short bench run: every engine kernel
ENGINE='regex:stream_|integrate_|tile_mask|fill_kernel|sanitize_depth|build_depth_mips|depth_from_u16|extract_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ENGINE" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --frames 256 --steps 2 --warmup 3 --quick \
    > gpurun_out/${tag}_ncu_launch.log 2>&1
python profiles/summarize.py --launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt
cat gpurun_out/${tag}_launch_share.txt
du -sh gpurun_out
