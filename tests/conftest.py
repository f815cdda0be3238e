"""Shared fixtures.  `-m "not gpu"`: oracle vs golden vectors / the reference's own host build, host
logic, C-ABI symbol checks.  `-m gpu`: the parity tests proper, through the C ABI on a B200."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import __graft_entry__ as graft  # noqa: E402

REFERENCE = Path("/root/reference")
SAMPLE = REFERENCE / "data" / "sample" / "depth-fusion-demo"
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    if not (graft.PKG_DIR / "lib" / "libtsdf_fusion.so").exists():
        graft.build()
    return graft.load_package()


@pytest.fixture(scope="session")
def oracle():
    mod = graft.load_oracle()
    if not mod.ORACLE_LIB.exists():
        graft.build()
    return mod


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


_seq_cache = {}


def make_sequence(pkg, oracle, frames, base_frame=0, im_size=(640, 480)):
    """(depth float32 [n,H,W], depth_mm uint16 [n,H,W], cam2base float32 [n,16]) of synthetic frames."""
    key = (tuple(frames), base_frame, tuple(im_size))
    if key not in _seq_cache:
        synth = pkg.synth
        K = synth.CAM_K.copy()
        if tuple(im_size) != (640, 480):
            s = im_size[0] / 640.0
            K[[0, 2, 4, 5]] = K[[0, 2, 4, 5]] * np.float32(s)
        mm = synth.mm_to_u16_numpy(synth.render_depth_mm(frames, base_frame=base_frame, im_size=im_size, cam_K=K))
        depth = oracle.depth_from_u16(mm)
        c2w, b2w = synth.make_poses(frames, base_frame)
        poses = np.stack([oracle.cam2base_from_world(b2w, m).ravel() for m in c2w]).astype(np.float32)
        _seq_cache[key] = (depth, mm, poses, K)
    return _seq_cache[key]


@pytest.fixture(scope="session")
def sequence(pkg, oracle):
    return lambda frames, base_frame=0, im_size=(640, 480): make_sequence(pkg, oracle, frames, base_frame, im_size)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(bits(a), bits(b)):
        bad = np.flatnonzero(bits(a).ravel() != bits(b).ravel())
        i = bad[0]
        raise AssertionError(f"{what}: {bad.size} of {a.size} values differ bitwise; first at {i}: "
                             f"{a.ravel()[i]!r} vs {b.ravel()[i]!r}")
