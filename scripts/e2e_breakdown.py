"""Diagnostic: where the end-to-end step (host frames -> points on the host) spends its time beyond the resident-input
step: device input in one group, device input in the host path's group schedule, pinned host input, extraction, copy."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
cfg = dict(synth.CONFIGS["1024"])
grid, F = cfg["grid"], cfg["frames"]
depth, poses = bench.build_inputs(pkg, cfg, torch.device("cuda", 0), 0, 1, None)
host = torch.empty(depth.shape, dtype=torch.float32, pin_memory=True)
host.copy_(depth)
torch.cuda.synchronize()
vol = pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, max_batch=32)
groups = [(0, 32), (32, 128), (128, 384), (384, 640), (640, 896), (896, 1000)]


def wall(fn, n=3):
    ts = []
    for _ in range(n):
        vol.reset()
        vol.synchronize()
        t0 = time.perf_counter()
        fn()
        vol.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def dev_groups():
    for a, b in groups:
        vol.integrate_batch(depth[a:b], poses[a:b])


for name, fn in (("device input, one group", lambda: vol.integrate_batch(depth, poses)),
                 ("device input, host-path group schedule", dev_groups),
                 ("pinned host input", lambda: vol.integrate_batch(host, poses))):
    fn()
    vol.synchronize()
    print(f"{name:45s} {wall(fn):8.2f} ms", flush=True)
t0 = time.perf_counter(); n = vol.extract_surface_count(); vol.synchronize(); t1 = time.perf_counter()
pts = np.empty((n + 16, 3), np.float32)
pkg.binding._check(vol._lib.tsdf_copy_surface(vol._h, pts.ctypes.data, pts.shape[0]), "copy_surface"); t2 = time.perf_counter()
t3 = time.perf_counter(); n = vol.extract_surface_count(); vol.synchronize(); t4 = time.perf_counter()
print(f"extract_surface (count + emit) {1e3 * (t1 - t0):.2f} ms, again {1e3 * (t4 - t3):.2f} ms; copy_surface of {n} points {1e3 * (t2 - t1):.2f} ms")
t0 = time.perf_counter(); vol.reset(); vol.synchronize(); print(f"reset {1e3 * (time.perf_counter() - t0):.2f} ms")
