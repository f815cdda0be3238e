#!/bin/bash
set -u
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_mask --launch-skip 1 --launch-count 1 \
    -f -o /tmp/r02y_mask python scripts/tune_variants.py 1000 32 0 > gpurun_out/r02y_ncu.log 2>&1
python profiles/summarize.py /tmp/r02y_mask.ncu-rep > gpurun_out/r02y_mask.txt
ncu -i /tmp/r02y_mask.ncu-rep --page source --csv > /tmp/r02y_src.csv 2>/dev/null
python profiles/sass_blocks.py /tmp/r02y_src.csv 1.0 > gpurun_out/r02y_mask_sass_blocks.txt
cat gpurun_out/r02y_mask.txt; head -40 gpurun_out/r02y_mask_sass_blocks.txt
