"""Small fusion + extraction through the C ABI, meant to run under compute-sanitizer
(memcheck / racecheck / initcheck):  compute-sanitizer --tool memcheck python scripts/sanitizer_smoke.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402

pkg, oracle = graft.load_package(), graft.load_oracle()
synth = pkg.synth
TUNING = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0
frames = [0, 9, 21, 140, 333, 500]
mm = synth.mm_to_u16_numpy(synth.render_depth_mm(frames))
depth = pkg.depth_from_u16(mm)
c2w, b2w = synth.make_poses(frames)
poses = np.stack([pkg.cam2base_from_world(b2w, m).ravel() for m in c2w])
for dims, voxel in (((100, 72, 90), 0.03), ((50, 62, 38), 0.06)):   # float4 path (2 sub-slab streams) and scalar path
    trunc = float(np.float32(voxel) * np.float32(5))
    with pkg.Volume(dims, (-1.5, -1.5, 0.5), voxel, trunc, synth.CAM_K, max_batch=4, tuning=TUNING) as vol:
        vol.integrate_batch(depth, poses)
        vol.integrate_batch(mm, poses)
        t, w = vol.download()
        pts = vol.extract_surface()
    t_ref, w_ref = oracle.fuse(synth.CAM_K, np.concatenate([poses, poses]), np.concatenate([depth, depth]), dims,
                               (-1.5, -1.5, 0.5), voxel, trunc)
    assert np.array_equal(t.view(np.uint32), t_ref.view(np.uint32)) and np.array_equal(w, w_ref)
    print(dims, "ok", len(pts), "points")
