"""Diagnostic: time each of the N equal-work z-slabs of the 1024^3 workload on ONE GPU, one after the other, and
compare the sum with the monolithic volume (isolates sharding overhead from multi-GPU effects).
    python scripts/slab_overhead.py [frames] [n_slabs] [K ...]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = dict(synth.CONFIGS["1024"], frames=F)
grid = cfg["grid"]
device = torch.device("cuda", 0)
depth, poses = bench.build_inputs(pkg, cfg, device, 0, 1, None)
cost = pkg.plan_z_cost(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, depth, poses,
                       frame_step=max(1, F // 128)).astype(np.float64)


def time_slab(z0, z1, batch):
    with pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, z_range=(z0, z1),
                    max_batch=batch) as vol:
        vol.integrate_batch(depth, poses)
        ts = []
        for _ in range(2):
            vol.reset()
            vol.synchronize()
            vol.timer_start()
            vol.integrate_batch(depth, poses)
            ts.append(vol.timer_stop())
        return min(ts)


for batch in [int(a) for a in sys.argv[3:]] or [32]:
    whole = time_slab(0, grid.dims[2], batch)
    for base in (0.05, 0.15):
        ranges = pkg.sharding.split_groups(cost, grid.dims[2], N, base)
        parts = [time_slab(a, b, batch) for a, b in ranges]
        print(f"K={batch} base_cost={base}: whole {whole:.1f} ms, {N} slabs sum {sum(parts):.1f} ms (+{100 * (sum(parts) / whole - 1):.1f} %), "
              f"max {max(parts):.1f} (mean {sum(parts) / N:.1f}) -> speed-up {whole / max(parts):.2f}x, "
              f"{whole / (sum(parts) / N):.2f}x at perfect balance; cuts {[a for a, _ in ranges]} parts {[round(p, 1) for p in parts]}", flush=True)
