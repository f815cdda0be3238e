"""Calibrates the planner's cost model on the 1024^3 workload: time = a * kept (tile, frame) pairs + b * pairs that reach
a surface band + c per group of 8 planes, fitted over 16 slabs of 64 planes timed one after the other on one GPU.
    python scripts/plan_calibrate.py [frames]"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = dict(synth.CONFIGS["1024"], frames=F)
grid = cfg["grid"]
depth, poses = bench.build_inputs(pkg, cfg, torch.device("cuda", 0), 0, 1, None)


def plan(beta):
    os.environ["TSDF_PLAN_BETA"] = str(beta)
    return pkg.plan_z_cost(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, depth, poses).astype(np.float64)


pairs = plan(0)
band = plan(1) - pairs
del os.environ["TSDF_PLAN_BETA"]
rows, times = [], []
step = 64
for z0 in range(0, grid.dims[2], step):
    with pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, z_range=(z0, z0 + step)) as vol:
        vol.integrate_batch(depth, poses)
        ts = []
        for _ in range(3):
            vol.reset()
            vol.synchronize()
            vol.timer_start()
            vol.integrate_batch(depth, poses)
            ts.append(vol.timer_stop())
    g0, g1 = z0 // 8, (z0 + step) // 8
    rows.append([pairs[g0:g1].sum(), band[g0:g1].sum(), g1 - g0])
    times.append(min(ts))
A, y = np.array(rows), np.array(times)
sol, *_ = np.linalg.lstsq(A, y, rcond=None)
pred = A @ sol
only_pairs, *_ = np.linalg.lstsq(A[:, [0, 2]], y, rcond=None)
print(json.dumps({"frames": F, "slab_planes": step, "ms": [round(t, 2) for t in times],
                  "pairs_M": [round(r[0] / 1e6, 2) for r in rows], "band_pairs_M": [round(r[1] / 1e6, 2) for r in rows],
                  "fit_ms_per_Mpair": sol[0] * 1e6, "fit_ms_per_Mband": sol[1] * 1e6, "fit_ms_per_group": sol[2],
                  "beta": sol[1] / sol[0], "base_cost_rel": sol[2] / (sol[0] * pairs.mean()),
                  "max_rel_err_with_band": float(np.abs(pred / y - 1).max()),
                  "max_rel_err_pairs_only": float(np.abs((A[:, [0, 2]] @ only_pairs) / y - 1).max())}))
