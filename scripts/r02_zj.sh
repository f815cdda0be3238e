#!/bin/bash
set -u
timeout 100 compute-sanitizer --tool initcheck python scripts/sanitizer_smoke.py > gpurun_out/r02zj_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -2 gpurun_out/r02zj_initcheck.log
timeout 100 compute-sanitizer --tool synccheck python scripts/sanitizer_smoke.py > gpurun_out/r02zj_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/r02zj_synccheck.log
