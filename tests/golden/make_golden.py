"""Generates tests/golden/*.npz with the REFERENCE's own code (oracle/_ref/libref_host.so: the
Integrate body of depth-fusion/demo.cu:18-68 and the utils.hpp helpers compiled for the host from
/root/reference where it lies).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The vectors travel to the GPU box, where /root/reference does not exist; tests/test_golden.py checks
the oracle against them and tests/test_gpu_golden.py checks the CUDA engine against them.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

OUT = Path(__file__).resolve().parent
SAMPLE = Path("/root/reference/data/sample/depth-fusion-demo")


def main():
    graft.build()
    pkg, oracle = graft.load_package(), graft.load_oracle()
    assert oracle.have_ref_host(), "needs oracle/_ref/libref_host.so"
    synth = pkg.synth

    # 1. pose algebra: reference multiply_matrix / invert_matrix on general and rigid matrices
    rng = np.random.default_rng(20161128)
    mats = rng.normal(size=(24, 16)).astype(np.float32)
    mats[:8] *= np.float32(100.0)
    for i in range(8):  # rigid poses as the sample's pose files hold them
        mats[16 + i] = synth.cam2world_f32(37 * i).ravel()
    inv = np.stack([oracle.ref_host_invert_matrix(m)[1].ravel() for m in mats])
    mul = np.stack([oracle.ref_host_multiply_matrix(mats[i], mats[(i + 5) % 24]).ravel() for i in range(24)])
    np.savez_compressed(OUT / "pose_algebra.npz", mats=mats, inv=inv, mul=mul)

    # 2. synthetic fusion, small enough to commit: 64x48 depth (intrinsics scaled by 0.1), 40x36x32 grid
    frames = [0, 9, 21, 140, 333]
    im_size = (64, 48)
    K = synth.CAM_K.copy()
    K[[0, 2, 4, 5]] *= np.float32(0.1)
    mm = synth.mm_to_u16_numpy(synth.render_depth_mm(frames, im_size=im_size, cam_K=K))
    depth = oracle.depth_from_u16(mm)
    c2w, b2w = synth.make_poses(frames)
    inv_b = oracle.ref_host_invert_matrix(b2w)[1]
    poses = np.stack([oracle.ref_host_multiply_matrix(inv_b, m).ravel() for m in c2w])
    dims, voxel, origin, trunc = (40, 36, 32), np.float32(0.08), (-1.5, -1.5, 0.5), np.float32(0.4)
    t, w = oracle.ref_host_fuse(K, poses, depth, dims, origin, float(voxel), float(trunc))
    ply = OUT / "_tmp.ply"
    oracle.ref_host_save_surface(ply, dims, float(voxel), origin, t, w)
    np.savez_compressed(OUT / "synthetic_fusion.npz", depth_mm=mm, cam_K=K, cam2world=c2w, base2world=b2w,
                        cam2base=poses, dims=np.array(dims), voxel_size=voxel, origin=np.float32(origin),
                        trunc_margin=trunc, tsdf=t, weight=w, ply=np.frombuffer(ply.read_bytes(), np.uint8))
    ply.unlink()

    # 3. the reference's own fixture: three frames of data/sample/depth-fusion-demo, depth decimated 4x
    #    (nearest) so the vectors stay small; intrinsics scaled accordingly
    if SAMPLE.exists():
        import cv2
        K = oracle.ref_host_load_matrix(SAMPLE / "camera-intrinsics.txt", 3, 3)
        K[[0, 2, 4, 5]] *= np.float32(0.25)
        base = oracle.ref_host_load_matrix(SAMPLE / "rgbd-frames" / "frame-000008.pose.txt", 4, 4)
        inv_b = oracle.ref_host_invert_matrix(base)[1]
        mm, poses, c2ws = [], [], []
        for idx in (8, 30, 57):
            full = cv2.imread(str(SAMPLE / "rgbd-frames" / f"frame-{idx:06d}.depth.png"), cv2.IMREAD_UNCHANGED)
            mm.append(full[::4, ::4].copy())
            c2w = oracle.ref_host_load_matrix(SAMPLE / "rgbd-frames" / f"frame-{idx:06d}.pose.txt", 4, 4)
            c2ws.append(c2w)
            poses.append(oracle.ref_host_multiply_matrix(inv_b, c2w).ravel())
        mm, poses = np.stack(mm), np.stack(poses)
        depth = oracle.depth_from_u16(mm)
        dims, voxel, origin, trunc = (64, 64, 64), np.float32(0.047), (-1.5, -1.5, 0.5), np.float32(0.235)
        t, w = oracle.ref_host_fuse(K, poses, depth, dims, origin, float(voxel), float(trunc))
        np.savez_compressed(OUT / "sample_fusion.npz", depth_mm=mm, cam_K=K, cam2world=np.stack(c2ws),
                            base2world=base, cam2base=poses, dims=np.array(dims), voxel_size=voxel,
                            origin=np.float32(origin), trunc_margin=trunc, tsdf=t, weight=w)
        # one full-resolution depth frame as the PNG-decoder fixture
        src = SAMPLE / "rgbd-frames" / "frame-000008.depth.png"
        (OUT / "frame-000008.depth.png").write_bytes(src.read_bytes())
        (OUT / "frame-000008.pose.txt").write_bytes((SAMPLE / "rgbd-frames" / "frame-000008.pose.txt").read_bytes())
    for f in sorted(OUT.glob("*")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
