#!/bin/bash
# final-kernel evidence: ncu counters on the bench workload, one --set full capture, SASS block shares, launch list, MINB A/B
set -u
tag=r02w
mkdir -p gpurun_out
M=lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 900 ncu --metrics $M --clock-control none -k regex:stream_integrate --launch-skip 1 --launch-count 1 --csv \
   --log-file gpurun_out/${tag}_production_1000.csv python scripts/tune_variants.py 1000 32 0 > gpurun_out/${tag}_ncu_prod1000.log 2>&1
grep -v "^==" gpurun_out/${tag}_production_1000.csv | cut -d, -f13- | tr -d '"'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_integrate --launch-skip 1 --launch-count 1 \
    -f -o /tmp/${tag}_stream python scripts/tune_variants.py 128 32 0 > gpurun_out/${tag}_ncu_stream.log 2>&1
python profiles/summarize.py /tmp/${tag}_stream.ncu-rep > gpurun_out/${tag}_stream_k32.txt
ncu -i /tmp/${tag}_stream.ncu-rep --page source --csv > /tmp/${tag}_src_sass.csv 2>/dev/null
python profiles/sass_blocks.py /tmp/${tag}_src_sass.csv 0.5 > gpurun_out/${tag}_stream_sass_blocks.txt
head -45 gpurun_out/${tag}_stream_k32.txt | tail -40
ENGINE='regex:stream_|integrate_|tile_mask|fill_kernel|sanitize_depth|build_depth_mips|depth_from_u16|extract_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ENGINE" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_ncu_launch.log 2>&1
python profiles/summarize.py --launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt
cat gpurun_out/${tag}_launch_share.txt
for m in 4 6; do echo "MINB=$m"; TSDF_USE_VALIDATION=1 TSDF_STREAM_MINB=$m timeout 300 python scripts/tune_variants.py 256 32 0 2>&1 | tail -1; done
