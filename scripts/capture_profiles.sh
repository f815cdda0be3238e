#!/bin/bash
# Run on the GPU box (gpurun -- 'bash scripts/capture_profiles.sh rNN'): the ncu evidence profiles/README.md quotes,
# reduced to text ON THE BOX (the .ncu-rep files stay in /tmp: gpurun merges at most 64 MiB back).
#   gpurun_out/<tag>_launches.csv      per-launch durations of the engine's kernels over a short bench run
#   gpurun_out/<tag>_integrate_k8.txt  ncu --set full summary of the three sub-slab launches of one K = 8 pass
#   gpurun_out/<tag>_integrate_k1.txt  the same for K = 1
#   gpurun_out/<tag>_k8_src.csv        SASS-level source page of the K = 8 capture (instruction shares)
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
ENGINE='regex:integrate_|tile_mask|fill_kernel|sanitize_depth|build_depth_mips|depth_from_u16|extract_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$ENGINE" -c 800 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --frames 64 --steps 2 --warmup 3 --quick \
    > gpurun_out/${tag}_ncu_launch.log 2>&1
python profiles/summarize.py --launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt
# tune_variants.py F K tuning: the first integrate_batch issues (F / K) passes x 3 sub-slab launches; skip them
ncu --set full --clock-control none --import-source on -k regex:integrate_pending --launch-skip 12 --launch-count 3 \
    -f -o /tmp/${tag}_k8 python scripts/tune_variants.py 32 8 0 > gpurun_out/${tag}_ncu_k8.log 2>&1
python profiles/summarize.py /tmp/${tag}_k8.ncu-rep > gpurun_out/${tag}_integrate_k8.txt
ncu -i /tmp/${tag}_k8.ncu-rep --page source --csv > gpurun_out/${tag}_k8_src.csv 2>/dev/null
ncu --set full --clock-control none -k regex:integrate_pending --launch-skip 12 --launch-count 3 \
    -f -o /tmp/${tag}_k1 python scripts/tune_variants.py 4 1 0 > gpurun_out/${tag}_ncu_k1.log 2>&1
python profiles/summarize.py /tmp/${tag}_k1.ncu-rep > gpurun_out/${tag}_integrate_k1.txt
du -sh gpurun_out
