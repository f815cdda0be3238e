"""Diagnostic: time the 1024^3 K-frame fusion for a list of kernel tuning words in ONE process (same box, same
inputs), e.g.  python scripts/tune_variants.py 128 16 0 18 24 4194304"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
tunings = [int(a, 0) for a in sys.argv[3:]] or [0]
cfg = dict(synth.CONFIGS["1024"], frames=F)
grid = cfg["grid"]
depth, poses = bench.build_inputs(pkg, cfg, torch.device("cuda", 0), 0, 1, None)
ref = None
import os
VALIDATION = True if os.environ.get("TSDF_USE_VALIDATION") == "1" else None   # None: validation library only for tuning != 0
for tuning in tunings:
    with pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, max_batch=K,
                    tuning=tuning, validation=VALIDATION) as vol:
        vol.integrate_batch(depth, poses)
        times = []
        for _ in range(3):
            vol.reset()
            vol.synchronize()
            vol.timer_start()
            vol.integrate_batch(depth, poses)
            times.append(vol.timer_stop())
        pts = vol.extract_surface()
    if ref is None:
        ref = pts
    same = pts.shape == ref.shape and bool((pts.view("uint32") == ref.view("uint32")).all())
    ms = sorted(times)[1]
    print(f"tuning {tuning:#x}: {ms:8.2f} ms / {F} frames -> {grid.n_voxels * F / ms / 1e6:8.1f} G voxel-updates/s, "
          f"{len(pts)} surface points, identical to first: {same}", flush=True)
