"""Diagnostic: time the fusion of B 300^3 fragments x 50 frames through ONE tsdf_integrate_multi launch (BASELINE
config 5's unit of work), e.g.  python scripts/frag_time.py 8 [K] [dim] [voxel]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = synth.CONFIGS["frag300"]
grid, F = cfg["grid"], cfg["frames"]
if len(sys.argv) > 4:
    grid = synth.GridSpec(int(sys.argv[3]), float(sys.argv[4]))
dev = torch.device("cuda", 0)
mm, poses = [], []
for i in range(B):
    frames = list(range(F * i, F * i + F))
    mm.append(synth.mm_to_u16_tensor(synth.render_depth_mm(frames, base_frame=F * i, device=dev)))
    c2w, b2w = synth.make_poses(frames, base_frame=F * i)
    poses.append(np.stack([pkg.cam2base_from_world(b2w, m).ravel() for m in c2w]).astype(np.float32))
torch.cuda.synchronize()
vols = [pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, max_batch=K) for _ in range(B)]
times = []
for it in range(4):
    for v in vols:
        v.reset()
    vols[0].synchronize()
    for v in vols:
        v.synchronize()
    vols[0].timer_start()
    pkg.integrate_multi(vols, mm, poses)
    times.append(vols[0].timer_stop())
    for v in vols:
        v.synchronize()
n_pts = sum(len(v.extract_surface()) for v in vols)
ms = sorted(times[1:])[1]
print(f"{B} fragments {grid.dims[0]}^3 ({grid.voxel_size * 1e3:.0f} mm) x {F} frames, K={K}: {ms:8.2f} ms -> "
      f"{B * grid.n_voxels * F / ms / 1e6:8.1f} G voxel-updates/s, {n_pts} surface points", flush=True)
for v in vols:
    v.close()
