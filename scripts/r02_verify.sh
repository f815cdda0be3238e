#!/bin/bash
set -u
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_stream_engine.py -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python scripts/tune_variants.py 256 32 0 > gpurun_out/${tag}_time.log 2>&1; cat gpurun_out/${tag}_time.log
M=lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 900 ncu --metrics $M --clock-control none -k regex:stream_integrate --launch-skip 1 --launch-count 1 --csv \
   --log-file gpurun_out/${tag}_production_1000.csv python scripts/tune_variants.py 1000 32 0 > gpurun_out/${tag}_ncu_prod1000.log 2>&1
grep -v "^==" gpurun_out/${tag}_production_1000.csv | cut -d, -f13- | tr -d '"'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_integrate --launch-skip 1 --launch-count 1 \
    -f -o /tmp/${tag}_stream python scripts/tune_variants.py 128 32 0 > gpurun_out/${tag}_ncu_stream.log 2>&1
python profiles/summarize.py /tmp/${tag}_stream.ncu-rep > gpurun_out/${tag}_stream_k32.txt
ncu -i /tmp/${tag}_stream.ncu-rep --page source --csv > gpurun_out/${tag}_stream_src_sass.csv 2>/dev/null
python profiles/sass_blocks.py gpurun_out/${tag}_stream_src_sass.csv 0.5 > gpurun_out/${tag}_stream_sass_blocks.txt
head -45 gpurun_out/${tag}_stream_k32.txt | tail -38
