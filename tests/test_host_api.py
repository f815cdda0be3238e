"""Host side of the C ABI without a GPU: the library loads, exports every symbol the header
declares, the pose algebra / file formats match the oracle bit for bit, the PNG reader matches cv2,
and compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, SAMPLE, assert_bit_equal


def test_header_symbols_are_exported(pkg):
    header = (ROOT / "include" / "tsdf_fusion.h").read_text()
    declared = sorted(set(re.findall(r"\b(tsdf_[A-Za-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    lib = pkg.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tsdf_fusion.h but not exported"
    assert sorted(pkg.binding.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.tsdf_version()


def test_desc_struct_layout_matches_header(pkg):
    """ctypes mirror of tsdf_volume_desc: field order / sizes as in the header (checked via sizeof)."""
    assert ctypes.sizeof(pkg.binding.VolumeDesc) == 8 * 5 + 4 * 5 + 4 * 2 + 4 * 9 + 4 * 4 + 8 * 3
    assert ctypes.sizeof(pkg.binding.Stats) == 32


def test_pose_algebra_matches_oracle(pkg, oracle):
    rng = np.random.default_rng(7)
    for i in range(300):
        a = (rng.normal(size=16) * 10 ** rng.uniform(-2, 2)).astype(np.float32)
        b = rng.normal(size=16).astype(np.float32)
        assert_bit_equal(pkg.multiply_matrix(a, b), oracle.multiply_matrix(a, b), f"multiply #{i}")
        ok_p, inv_p = pkg.invert_matrix(a)
        ok_o, inv_o = oracle.invert_matrix(a)
        assert ok_p == ok_o
        assert_bit_equal(inv_p, inv_o, f"invert #{i}")
    assert pkg.invert_matrix(np.zeros(16, np.float32))[0] is False  # utils.hpp:229


def test_depth_conversion_all_u16_values(pkg, oracle):
    mm = np.arange(65536, dtype=np.uint16)
    d = pkg.depth_from_u16(mm)
    assert_bit_equal(d, oracle.depth_from_u16(mm), "depth_from_u16")
    assert d[6000] == np.float32(6.0) and d[6001] == 0.0 and d[65535] == 0.0  # utils.hpp:83-84


def test_ply_and_bin_writers_match_oracle(pkg, oracle, tmp_path):
    rng = np.random.default_rng(8)
    xyz = rng.normal(size=(1234, 3)).astype(np.float32)
    for pts in (xyz, xyz[:0]):
        pkg.write_ply(tmp_path / "a.ply", pts)
        oracle.write_ply(tmp_path / "b.ply", pts)
        data = (tmp_path / "a.ply").read_bytes()
        assert data == (tmp_path / "b.ply").read_bytes()
        head = data.split(b"end_header\n")[0].split(b"\n")
        assert len(head) == 7 and head[6] == b"" and head[2] == b"element vertex %d" % len(pts)  # 7 header lines (core/demo.cu:76-86)
    dims, origin = (7, 5, 3), (-1.5, -1.5, 0.5)
    t = rng.normal(size=7 * 5 * 3).astype(np.float32)
    pkg.write_bin(tmp_path / "a.bin", dims, origin, 0.006, 0.03, t)
    oracle.write_tsdf_bin(tmp_path / "b.bin", dims, origin, 0.006, 0.03, t)
    raw = (tmp_path / "a.bin").read_bytes()
    assert raw == (tmp_path / "b.bin").read_bytes()
    header = np.frombuffer(raw[:32], np.float32)
    assert list(header[:3]) == [7, 5, 3] and len(raw) == 32 + 4 * t.size  # demo.cu:196-213


def test_matrix_loader(pkg, oracle, tmp_path):
    p = tmp_path / "pose.txt"
    p.write_text("  9.0709007e-001\t  2.7616704e-001\t -3.1753206e-001\t -3.4287909e-001\t\n"
                 " -2.7506533e-001\t  9.6011728e-001\t  4.9265359e-002\t  1.2549059e-002\t\n"
                 "  3.1848383e-001\t  4.2655181e-002\t  9.4690639e-001\t  2.9991144e-001\t\n"
                 "  0.0000000e+000\t  0.0000000e+000\t  0.0000000e+000\t  1.0000000e+000\t\n")
    m = pkg.LoadMatrixFromFile(p, 4, 4)
    got, ref = oracle.load_matrix(p, 16)
    assert got == 16
    assert_bit_equal(m.ravel(), ref, "LoadMatrixFromFile")
    assert m[0, 0] == np.float32(0.90709007) and m[3, 3] == 1.0
    with pytest.raises(pkg.TsdfError, match="cannot open"):
        pkg.LoadMatrixFromFile(tmp_path / "missing.txt", 4, 4)
    p.write_text("1 2 3")
    with pytest.raises(pkg.TsdfError, match="too few"):
        pkg.LoadMatrixFromFile(p, 3, 3)


def test_png_reader_matches_cv2(pkg, oracle, tmp_path):
    import cv2
    rng = np.random.default_rng(9)
    cases = {"noise": rng.integers(0, 65536, size=(480, 640)).astype(np.uint16),
             "smooth": (np.add.outer(np.arange(480), np.arange(640)) * 7 % 9000).astype(np.uint16),
             "small": rng.integers(0, 7000, size=(5, 3)).astype(np.uint16)}
    for name, mm in cases.items():
        path = tmp_path / f"{name}.png"
        assert cv2.imwrite(str(path), mm)
        h, w = mm.shape
        assert np.array_equal(pkg.binding.read_depth_png_u16(path, h, w), mm), name
        assert_bit_equal(pkg.ReadDepth(path, h, w), oracle.depth_from_u16(mm), f"ReadDepth {name}")
    fixture = GOLDEN / "frame-000008.depth.png"
    if fixture.exists():  # a frame of the reference's sample sequence (all five PNG filter types occur)
        mm = cv2.imread(str(fixture), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(pkg.binding.read_depth_png_u16(fixture, 480, 640), mm)
    if SAMPLE.exists():
        for idx in (9, 33, 57):
            f = SAMPLE / "rgbd-frames" / f"frame-{idx:06d}.depth.png"
            assert np.array_equal(pkg.binding.read_depth_png_u16(f, 480, 640),
                                  cv2.imread(str(f), cv2.IMREAD_UNCHANGED))
    # error behaviour: wrong size, 8-bit image, garbage, missing file -> status codes, never a crash
    with pytest.raises(pkg.TsdfError, match="expected size"):
        pkg.ReadDepth(tmp_path / "small.png", 480, 640)
    cv2.imwrite(str(tmp_path / "gray8.png"), np.zeros((5, 3), np.uint8))
    with pytest.raises(pkg.TsdfError, match="16-bit"):
        pkg.ReadDepth(tmp_path / "gray8.png", 5, 3)
    (tmp_path / "junk.png").write_bytes(b"not a png at all")
    with pytest.raises(pkg.TsdfError, match="not a PNG"):
        pkg.ReadDepth(tmp_path / "junk.png", 5, 3)
    with pytest.raises(pkg.TsdfError, match="cannot open"):
        pkg.ReadDepth(tmp_path / "missing.png", 5, 3)


def test_argument_validation_without_gpu(pkg):
    lib = pkg.load_library()
    assert lib.tsdf_create(None, None) == 1  # TSDF_ERR_ARG
    with pytest.raises(pkg.TsdfError, match="dimensions"):
        pkg.Volume((0, 4, 4), (0, 0, 0), 0.1, 0.5, np.eye(3).ravel())
    with pytest.raises(pkg.TsdfError, match="slab"):
        pkg.Volume((4, 4, 4), (0, 0, 0), 0.1, 0.5, np.eye(3).ravel(), z_range=(3, 9))
    with pytest.raises(pkg.TsdfError, match="max_batch"):
        pkg.Volume((4, 4, 4), (0, 0, 0), 0.1, 0.5, np.eye(3).ravel(), max_batch=99)


def test_compute_fails_loudly_without_cuda(pkg):
    """No CPU fallback: on a box without a GPU every compute entry point reports a CUDA error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.TsdfError, match="(?i)cuda"):
        pkg.Volume((8, 8, 8), (0, 0, 0), 0.1, 0.5, np.eye(3).ravel())
    with pytest.raises(pkg.TsdfError, match="(?i)cuda"):
        pkg.SaveVoxelGrid2SurfacePointCloud("/tmp/never.ply", 4, 4, 4, 0.1, 0, 0, 0, np.zeros(64), np.zeros(64), 0.2, 1.0)
    with pytest.raises(pkg.TsdfError):
        pkg.binding.selftest_rounding(0)
