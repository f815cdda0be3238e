"""Diagnostic: time each of the N cost-balanced z-slabs of the 1024^3 workload on ONE GPU, one after the
other, and compare the sum with the monolithic volume (isolates sharding overhead from multi-GPU effects)."""
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
cost = pkg.sharding.estimate_plane_cost(grid.dims, grid.origin, grid.voxel_size, synth.CAM_K, poses)


def time_slab(z0, z1, batch, tuning=0):
    with pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, z_range=(z0, z1),
                    max_batch=batch, tuning=tuning) as vol:
        vol.integrate_batch(depth, poses)
        vol.reset()
        vol.synchronize()
        vol.timer_start()
        vol.integrate_batch(depth, poses)
        return vol.timer_stop()


CASES = [tuple(int(v, 0) for v in a.split(':')) for a in sys.argv[3:]] or [(16, 1 << 10), (16, 0), (16, 3 << 10), (32, 0)]
for batch, tuning in CASES:
    whole = time_slab(0, grid.dims[2], batch, tuning)
    parts = [time_slab(a, b, batch, tuning) for a, b in pkg.sharding.weighted_slab_ranges(cost, N)]
    print(f"K={batch} tuning={tuning}: whole {whole:.1f} ms, {N} slabs sum {sum(parts):.1f} ms (+{100 * (sum(parts) / whole - 1):.1f} %), "
          f"max {max(parts):.1f} -> speed-up {whole / max(parts):.2f}x; parts {[round(p, 1) for p in parts]}")


