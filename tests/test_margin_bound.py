"""CPU check of pixel_round_margin() (3dmatch-toolbox_b200/csrc/tsdf_kernels.h): the bound that lets the fast
classification take roundf(u) from u' = fma(cam, fl(fx * rcp.approx(z)), c0) instead of the reference's
u = fl(fl(fx * fl(cam / z)) + c0) (depth-fusion/demo.cu:46-47).

The header's own function is compiled for the host and called through ctypes; the two pixel coordinates are formed in
numpy float32 arithmetic (one IEEE rounding per operation) with the reciprocal perturbed by the LARGEST relative error
the PTX ISA allows for rcp.approx (2^-22), in both directions -- the hardware measures 1.65 * 2^-24
(tsdf_selftest_pixel_margin, -m gpu).  Every difference must stay below the margin."""
from __future__ import annotations

import ctypes
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "3dmatch-toolbox_b200" / "csrc"


@pytest.fixture(scope="module")
def margin_fn(tmp_path_factory):
    cuda_inc = Path("/usr/local/cuda/include")
    if shutil.which("g++") is None or not (cuda_inc / "cuda_runtime.h").exists():
        pytest.skip("g++ / cuda_runtime.h not available")
    d = tmp_path_factory.mktemp("margin")
    (d / "m.cpp").write_text('#include "tsdf_kernels.h"\n'
                             'extern "C" float margin(float size, float c0) { return tsdf::pixel_round_margin(size, c0); }\n')
    so = d / "libmargin.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-I", str(CSRC), "-I", str(cuda_inc), "-o", str(so), str(d / "m.cpp")],
                   check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    lib.margin.restype = ctypes.c_float
    lib.margin.argtypes = [ctypes.c_float, ctypes.c_float]
    return lib.margin


def test_margin_values_for_the_sample_intrinsics(margin_fn):
    # 640 x 480, principal point (320, 240): a quarter of a thousandth of a pixel
    mx, my = margin_fn(640.0, 320.0), margin_fn(480.0, 240.0)
    assert 2.0e-4 < mx < 3.0e-4 and 1.5e-4 < my < 2.5e-4
    # grows with the distance between the principal point and the far image border
    assert margin_fn(640.0, 0.0) > mx and margin_fn(640.0, 900.0) > margin_fn(640.0, 0.0)


@pytest.mark.parametrize("fx,c0,size", [(585.0, 320.0, 640.0), (585.0, 240.0, 480.0), (1050.0, 100.5, 1280.0),
                                         (320.0, 700.0, 640.0), (2000.0, -50.0, 1920.0)])
def test_fast_pixel_coordinate_stays_inside_the_margin(margin_fn, fx, c0, size):
    f32 = np.float32
    rng = np.random.default_rng(20161128)
    n = 1_500_000
    margin = f32(margin_fn(size, c0))
    u_t = rng.uniform(-1.5, size + 1.5, n)
    z = (rng.uniform(1.0, 2.0, n) * np.exp2(rng.integers(-6, 8, n))).astype(f32)
    cam = ((u_t - c0) * z.astype(np.float64) / fx).astype(f32)
    fx32, c032 = f32(fx), f32(c0)
    u_ref = (fx32 * (cam / z)).astype(f32) + c032                       # demo.cu:46, one rounding per operation
    assert u_ref.dtype == np.float32
    worst = 0.0
    for rel in (2.0 ** -22 * 0.999, -(2.0 ** -22) * 0.999, None):
        e = rng.uniform(-2.0 ** -22, 2.0 ** -22, n) if rel is None else rel
        r = ((1.0 / z.astype(np.float64)) * (1.0 + e)).astype(f32)     # rcp.approx at the error the ISA allows
        k = (r * fx32).astype(f32)                                      # mul.rn
        u_fast = (cam.astype(np.float64) * k.astype(np.float64) + float(c0)).astype(f32)   # fma.rn (exact product)
        inside = (u_fast >= -1.5) & (u_fast <= size + 1.5)
        worst = max(worst, float(np.abs(u_fast[inside].astype(np.float64) - u_ref[inside].astype(np.float64)).max() / margin))
    assert 0.1 < worst < 1.0, worst
