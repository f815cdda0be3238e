"""Diagnostic: wall time of the frame-mask pre-pass on the 1024^3 x F-frame workload, through tsdf_plan_z_cost (the same
tests in cost-only form + sanitize + mips), and the resulting total cost (kept pairs).  python scripts/mask_time.py [F]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
key = sys.argv[2] if len(sys.argv) > 2 else "1024"
cfg = dict(synth.CONFIGS[key], frames=F)
grid = cfg["grid"]
depth, poses = bench.build_inputs(pkg, cfg, torch.device("cuda", 0), 0, 1, None)
ts = []
for _ in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cost = pkg.plan_z_cost(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, depth, poses, frame_step=1, cull_level=2)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"plan_z_cost {key} x {F} frames: {min(ts):.2f} ms (runs: {', '.join(f'{t:.1f}' for t in ts)}); total cost {int(cost.sum())}", flush=True)
