"""bin/demo --gpus N vs --gpus 1 on a synthetic sequence written as PNG / pose files: tsdf.ply must be byte-identical
(BASELINE north_star (4)).  python scripts/demo_gpus_check.py [dim] [gpus] [frames]"""
import hashlib
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 48
pkg = graft.load_package()
synth = pkg.synth
voxel = 3.0 / dim
with tempfile.TemporaryDirectory() as tmp:
    tmp = Path(tmp)
    data = tmp / "seq-01"
    data.mkdir()
    frames = [int(v) for v in np.linspace(0, 999, n)]
    mm = synth.mm_to_u16_numpy(synth.render_depth_mm(frames, device="cuda"))
    c2w, _ = synth.make_poses(frames)
    for i in range(n):
        cv2.imwrite(str(data / f"frame-{i:06d}.depth.png"), mm[i])
        (data / f"frame-{i:06d}.pose.txt").write_text("".join("\t".join(f" {v: .8e}" for v in row) + "\t\n" for row in c2w[i]))
    (tmp / "camera-intrinsics.txt").write_text("".join("\t".join(f" {v: .8e}" for v in row) + "\t\n" for row in synth.CAM_K.reshape(3, 3)))
    out = {"dim": dim, "frames": n, "voxel_size": voxel}
    for g in (1, gpus):
        d = tmp / f"g{g}"
        d.mkdir()
        argv = [str(ROOT / "3dmatch-toolbox_b200" / "bin" / "demo"), str(tmp / "camera-intrinsics.txt"), str(data), "0", "0", str(n),
                "-1.5", "-1.5", "0.5", "%f" % voxel, "%f" % (5 * voxel), "--dim", str(dim), "--gpus", str(g), "--timing"]
        t0 = time.perf_counter()
        r = subprocess.run(argv, cwd=d, capture_output=True, text=True, timeout=900)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            out[f"gpus{g}"] = {"error": r.stderr[-600:]}
            continue
        ply = (d / "tsdf.ply").read_bytes()
        timing = [l for l in r.stderr.splitlines() if l.startswith("{")]
        out[f"gpus{g}"] = {"seconds": round(dt, 2), "ply_bytes": len(ply), "sha256": hashlib.sha256(ply).hexdigest(),
                           "timing": json.loads(timing[-1]) if timing else None}
    a, b = out.get("gpus1", {}), out.get(f"gpus{gpus}", {})
    out["byte_identical"] = bool(a.get("sha256") and a.get("sha256") == b.get("sha256"))
    print(json.dumps(out))
