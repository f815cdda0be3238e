#!/bin/bash
# compute-sanitizer on the final kernels (the mask pre-pass now passes its ballots through shared memory)
set -u
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool racecheck python scripts/sanitizer_smoke.py > gpurun_out/r02zi_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02zi_racecheck.log
timeout 100 compute-sanitizer --tool memcheck python scripts/sanitizer_smoke.py > gpurun_out/r02zi_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02zi_memcheck.log
