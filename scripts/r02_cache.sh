#!/bin/bash
# Cache / L2-persistence evidence and the counters of the final kernel (run on the GPU box)
set -u
tag=${1:-r02g}
mkdir -p gpurun_out
L=gpurun_out/${tag}_ab.log
echo "== production" >> $L;                       timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
echo "== production + TSDF_L2_WINDOW=1" >> $L;    TSDF_L2_WINDOW=1 timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
echo "== validation lib, DIAG=0 (control)" >> $L; TSDF_USE_VALIDATION=1 timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
echo "== validation lib, DIAG=2 (L2 evict_last hint on the depth gathers)" >> $L; TSDF_USE_VALIDATION=1 TSDF_STREAM_DIAG=2 timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
echo "== production again" >> $L;                 timeout 300 python scripts/tune_variants.py 256 32 0 >> $L 2>&1
cat $L
M=lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
# gather-only launch: no TSDF / weight traffic, every L2 read is a depth gather (plus masks / records)
TSDF_USE_VALIDATION=1 TSDF_STREAM_DIAG=1 timeout 600 ncu --metrics $M --clock-control none -k regex:stream_integrate --launch-skip 1 --launch-count 1 --csv \
   --log-file gpurun_out/${tag}_gather_only.csv python scripts/tune_variants.py 256 32 0 > gpurun_out/${tag}_ncu_gather.log 2>&1
# the production launch, same metrics, same frames
timeout 600 ncu --metrics $M --clock-control none -k regex:stream_integrate --launch-skip 1 --launch-count 1 --csv \
   --log-file gpurun_out/${tag}_production_256.csv python scripts/tune_variants.py 256 32 0 > gpurun_out/${tag}_ncu_prod.log 2>&1
# counters of the bench workload (1000 frames, one launch)
timeout 900 ncu --metrics $M,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:stream_integrate --launch-skip 1 --launch-count 1 --csv \
   --log-file gpurun_out/${tag}_production_1000.csv python scripts/tune_variants.py 1000 32 0 > gpurun_out/${tag}_ncu_prod1000.log 2>&1
for f in gather_only production_256 production_1000; do echo "--- $f"; grep -v "^==" gpurun_out/${tag}_$f.csv | cut -d, -f5,13- | tr -d '"' ; done
# launch list of a bench run (every engine kernel)
ENGINE='regex:stream_|integrate_|tile_mask|fill_kernel|sanitize_depth|build_depth_mips|depth_from_u16|extract_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ENGINE" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${tag}_ncu_launch.log 2>&1
python profiles/summarize.py --launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt; cat gpurun_out/${tag}_launch_share.txt
# full-set capture of the final kernel for the stall / pipe summary and the SASS page
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_integrate --launch-skip 1 --launch-count 1 \
    -f -o /tmp/${tag}_stream python scripts/tune_variants.py 128 32 0 > gpurun_out/${tag}_ncu_stream.log 2>&1
python profiles/summarize.py /tmp/${tag}_stream.ncu-rep > gpurun_out/${tag}_stream_k32.txt
ncu -i /tmp/${tag}_stream.ncu-rep --page source --csv > gpurun_out/${tag}_stream_src_sass.csv 2>/dev/null
cat gpurun_out/${tag}_stream_k32.txt | head -40
