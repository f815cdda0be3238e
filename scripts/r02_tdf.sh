#!/bin/bash
set -u
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tdf.py tests/test_stream_engine.py -x -q -m gpu > gpurun_out/${tag}_pytest_tdf.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest_tdf.log
timeout 600 python scripts/tdf_bench.py > gpurun_out/${tag}_tdf.json 2> gpurun_out/${tag}_tdf.err; cat gpurun_out/${tag}_tdf.json; tail -3 gpurun_out/${tag}_tdf.err
timeout 600 ncu --set full --clock-control none -k regex:tdf_pass -c 3 -f -o /tmp/${tag}_tdf python scripts/tdf_bench.py > /dev/null 2>&1
python profiles/summarize.py /tmp/${tag}_tdf.ncu-rep > gpurun_out/${tag}_tdf_kernels.txt; grep -E "^## |duration|dram__bytes|dram_throughput" gpurun_out/${tag}_tdf_kernels.txt
timeout 600 python scripts/realdata_e2e.py 3 > gpurun_out/${tag}_realdata_e2e.json 2>/dev/null; cat gpurun_out/${tag}_realdata_e2e.json
