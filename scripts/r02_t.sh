#!/bin/bash
# fragments (config 5): timing + one ncu --set full capture of the multi-volume launch
set -u
mkdir -p gpurun_out
timeout 300 python scripts/frag_time.py 8 2>&1 | tail -1
timeout 300 python scripts/frag_time.py 16 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_integrate --launch-skip 1 --launch-count 1 \
    -f -o /tmp/r02t_frag python scripts/frag_time.py 8 > gpurun_out/r02t_ncu.log 2>&1
python profiles/summarize.py /tmp/r02t_frag.ncu-rep > gpurun_out/r02t_frag_k32.txt
ncu -i /tmp/r02t_frag.ncu-rep --page raw --csv > /tmp/r02t_raw.csv
python - <<'P' >> gpurun_out/r02t_frag_k32.txt
import csv
rows = list(csv.reader(open('/tmp/r02t_raw.csv')))
hdr, units, r = rows[0], rows[1], rows[2]
print("\n## L1 / LSU detail")
for i, h in enumerate(hdr):
    if any(k in h for k in ("l1tex__data_pipe_lsu_wavefronts", "l1tex__t_requests_pipe_lsu_mem_global_op_ld", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld",
                            "l1tex__throughput", "l1tex__lsu_writeback", "l1tex__data_bank", "l1tex__f_wavefronts", "l1tex__t_output_wavefronts",
                            "l1tex__average_t_sectors_per_request", "smsp__inst_executed_op_global_ld", "l1tex__lsuin_requests", "sm__inst_executed_pipe_lsu",
                            "l1tex__data_pipe_lsu_wavefronts_mem_shared", "smsp__warp_issue_stalled_lg_throttle", "smsp__average_warp")):
        print(f"{h:90s} {r[i]:>20s} {units[i]}")
P
ncu -i /tmp/r02t_frag.ncu-rep --page source --csv > gpurun_out/r02t_frag_src_sass.csv 2>/dev/null
python profiles/sass_blocks.py gpurun_out/r02t_frag_src_sass.csv 0.5 > gpurun_out/r02t_frag_sass_blocks.txt
rm -f gpurun_out/r02t_frag_src_sass.csv
cat gpurun_out/r02t_frag_k32.txt
