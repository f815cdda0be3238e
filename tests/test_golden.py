"""Oracle and host library against the committed golden vectors (tests/golden/*.npz), which were
produced by the reference's own code compiled for the host (tests/golden/make_golden.py).  Runs
anywhere -- this is the parity pin that travels to the GPU box."""
import numpy as np
import pytest

from conftest import GOLDEN, assert_bit_equal


def load(name):
    path = GOLDEN / name
    if not path.exists():
        pytest.skip(f"{name} not generated")
    return np.load(path)


def test_pose_algebra_golden(oracle, pkg):
    g = load("pose_algebra.npz")
    mats = g["mats"]
    for i, m in enumerate(mats):
        assert_bit_equal(oracle.invert_matrix(m)[1].ravel(), g["inv"][i], f"oracle invert #{i}")
        assert_bit_equal(pkg.invert_matrix(m)[1].ravel(), g["inv"][i], f"library invert #{i}")
        other = mats[(i + 5) % len(mats)]
        assert_bit_equal(oracle.multiply_matrix(m, other).ravel(), g["mul"][i], f"oracle multiply #{i}")
        assert_bit_equal(pkg.multiply_matrix(m, other).ravel(), g["mul"][i], f"library multiply #{i}")


@pytest.mark.parametrize("name", ["synthetic_fusion.npz", "sample_fusion.npz"])
def test_fusion_golden(oracle, pkg, name, tmp_path):
    g = load(name)
    depth = oracle.depth_from_u16(g["depth_mm"])
    assert_bit_equal(pkg.depth_from_u16(g["depth_mm"]), depth, "library depth conversion")
    # cam2base = inv(base2world) * cam2world with the reference's rounding (demo.cu:127-128,171)
    for c2w, want in zip(g["cam2world"], g["cam2base"]):
        assert_bit_equal(oracle.cam2base_from_world(g["base2world"], c2w).ravel(), want, "oracle cam2base")
        assert_bit_equal(pkg.cam2base_from_world(g["base2world"], c2w).ravel(), want, "library cam2base")
    dims = tuple(int(d) for d in g["dims"])
    t, w = oracle.fuse(g["cam_K"], g["cam2base"], depth, dims, g["origin"], float(g["voxel_size"]),
                       float(g["trunc_margin"]))
    assert g["weight"].max() >= 2
    assert_bit_equal(w, g["weight"], "weights")
    assert_bit_equal(t, g["tsdf"], "tsdf")
    if "ply" in g:
        xyz = oracle.extract_surface(t, w, dims, g["origin"], float(g["voxel_size"]))
        for writer, tag in ((oracle.write_ply, "oracle"), (pkg.write_ply, "library")):
            path = tmp_path / f"{tag}.ply"
            writer(path, xyz)
            assert path.read_bytes() == g["ply"].tobytes(), f"{tag} tsdf.ply differs from the reference writer's"
