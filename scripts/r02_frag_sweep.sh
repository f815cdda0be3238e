#!/bin/bash
set -u
tag=${1:-r02n}
mkdir -p gpurun_out
for b in 4 8 16 32; do
  timeout 300 python bench.py --config frag300 --fragments 32 --frag-batch $b --no-parity --no-cpu-baseline > gpurun_out/${tag}_frag_b$b.json 2> gpurun_out/${tag}_frag_b$b.err
  python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/${tag}_frag_b$b.json") if l.startswith("{")][-1])
print("B=$b value %.3e ms %.2f e2e %.3e seq %.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["sequential_handles"]["value"]))
PY
done
for k in 25 13; do
  timeout 300 python bench.py --config frag300 --fragments 32 --frag-batch 8 --batch $k --no-parity --no-cpu-baseline > gpurun_out/${tag}_frag_k$k.json 2> gpurun_out/${tag}_frag_k$k.err
  python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/${tag}_frag_k$k.json") if l.startswith("{")][-1])
print("K=$k value %.3e ms %.2f" % (j["value"], j["ms_per_step"]))
PY
done
