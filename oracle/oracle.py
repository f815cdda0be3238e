"""ctypes access to the CPU oracle (oracle/_build/liboracle.so) and to the reference's own code
compiled into oracle/_ref/ (libref_host.so: reference kernel body + helpers on host threads;
libref_cuda.so: the reference demo.cu built unchanged for sm_100a).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "_build" / "liboracle.so"
REF_HOST_LIB = HERE / "_ref" / "libref_host.so"
REF_CUDA_LIB = HERE / "_ref" / "libref_cuda.so"
REF_CUDA_NOFMAD_LIB = HERE / "_ref" / "libref_cuda_nofmad.so"

_cache = {}


def _load(path: Path):
    if path not in _cache:
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference for _ref/)")
        _cache[path] = C.CDLL(str(path))
    return _cache[path]


def have_ref_host() -> bool:
    return REF_HOST_LIB.exists()


def have_ref_cuda() -> bool:
    return REF_CUDA_LIB.exists()


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert n is None or a.size == n, (a.size, n)
    return a


def _p(a):
    return C.c_void_p(a.ctypes.data)


# ---- our restatement (oracle/tsdf_oracle.c) --------------------------------------------------------

def lib():
    L = _load(ORACLE_LIB)
    L.oracle_extract_surface.restype = C.c_int64
    L.oracle_surface_normals.restype = C.c_int64
    return L


def num_threads() -> int:
    return lib().oracle_num_threads()


def init_volume(n_voxels):
    return np.ones(n_voxels, np.float32), np.zeros(n_voxels, np.float32)


def fuse(cam_K, cam2base, depth_frames, dims, origin, voxel_size, trunc_margin, z_range=None, tsdf=None,
         weight=None):
    """n reference launches in order on the slab z_range; returns (tsdf, weight) shaped [nz, dy, dx]."""
    dx, dy, dz = (int(v) for v in dims)
    z0, z1 = (0, dz) if z_range is None else (int(z_range[0]), int(z_range[1]))
    poses = _f32(cam2base).reshape(-1, 16)
    depth = _f32(depth_frames)
    n = poses.shape[0]
    h_w = depth.size // n
    im_h, im_w = depth.shape[-2], depth.shape[-1]
    assert im_h * im_w == h_w
    nvox = (z1 - z0) * dy * dx
    if tsdf is None:
        tsdf, weight = init_volume(nvox)
    K = _f32(cam_K, 9)
    lib().oracle_fuse(_p(K), _p(poses), _p(depth), C.c_int(n), C.c_int(im_h), C.c_int(im_w), C.c_int64(dx),
                      C.c_int64(dy), C.c_int64(dz), C.c_int64(z0), C.c_int64(z1), C.c_float(origin[0]),
                      C.c_float(origin[1]), C.c_float(origin[2]), C.c_float(voxel_size), C.c_float(trunc_margin),
                      _p(tsdf), _p(weight))
    return tsdf.reshape(z1 - z0, dy, dx), weight.reshape(z1 - z0, dy, dx)


def multiply_matrix(a, b):
    a, b = _f32(a, 16), _f32(b, 16)
    out = np.empty(16, np.float32)
    lib().oracle_multiply_matrix(_p(a), _p(b), _p(out))
    return out.reshape(4, 4)


def invert_matrix(m):
    m = _f32(m, 16)
    out = np.zeros(16, np.float32)
    ok = lib().oracle_invert_matrix(_p(m), _p(out))
    return bool(ok), out.reshape(4, 4)


def cam2base_from_world(base2world, cam2world):
    _, inv = invert_matrix(base2world)
    return multiply_matrix(inv, cam2world)


def depth_from_u16(mm):
    mm = np.ascontiguousarray(mm, dtype=np.uint16)
    out = np.empty(mm.shape, np.float32)
    lib().oracle_depth_from_u16(_p(mm), C.c_int64(mm.size), _p(out))
    return out


def extract_surface(tsdf, weight, dims, origin, voxel_size, tsdf_thresh=0.2, weight_thresh=1.0, z_range=None):
    dx, dy, dz = (int(v) for v in dims)
    z0, z1 = (0, dz) if z_range is None else (int(z_range[0]), int(z_range[1]))
    t, w = _f32(tsdf, (z1 - z0) * dy * dx), _f32(weight, (z1 - z0) * dy * dx)
    args = [_p(t), _p(w), C.c_int64(dx), C.c_int64(dy), C.c_int64(dz), C.c_int64(z0), C.c_int64(z1),
            C.c_float(origin[0]), C.c_float(origin[1]), C.c_float(origin[2]), C.c_float(voxel_size),
            C.c_float(tsdf_thresh), C.c_float(weight_thresh)]
    n = lib().oracle_extract_surface(*args, None)
    xyz = np.empty((n, 3), np.float32)
    if n:
        lib().oracle_extract_surface(*args, _p(xyz))
    return xyz


def surface_normals(tsdf, weight, tsdf_thresh=0.2, weight_thresh=1.0):
    """One unit normal per surface voxel of the [nz, dy, dx] slab, in extract_surface order (parity unpinned)."""
    t, w = _f32(tsdf), _f32(weight)
    nz, dy, dx = t.shape
    n = int(((np.abs(t) < np.float32(tsdf_thresh)) & (w > np.float32(weight_thresh))).sum())
    out = np.zeros((n, 3), np.float32)
    got = lib().oracle_surface_normals(_p(t), _p(w), C.c_int64(dx), C.c_int64(dy), C.c_int64(nz), C.c_float(tsdf_thresh),
                                       C.c_float(weight_thresh), _p(out))
    assert got == n
    return out


def write_ply(path, xyz):
    xyz = _f32(xyz)
    assert lib().oracle_write_ply(os.fsencode(path), _p(xyz), C.c_int64(xyz.size // 3)) == 0


def write_tsdf_bin(path, dims, origin, voxel_size, trunc_margin, tsdf):
    t = _f32(tsdf)
    assert lib().oracle_write_tsdf_bin(os.fsencode(path), C.c_int64(dims[0]), C.c_int64(dims[1]), C.c_int64(dims[2]),
                                       C.c_float(origin[0]), C.c_float(origin[1]), C.c_float(origin[2]),
                                       C.c_float(voxel_size), C.c_float(trunc_margin), _p(t)) == 0


def load_matrix(path, count):
    out = np.zeros(count, np.float32)
    got = lib().oracle_load_matrix(os.fsencode(path), C.c_int(count), _p(out))
    return got, out


def compute_tdf(occ, dims, voxel_size=0.01, trunc_margin=0.05, tdf=None):
    """core/demo.cu:22-57 brute force; occ shaped [dz, dy, dx]; returns tdf of the same shape."""
    dx, dy, dz = (int(v) for v in dims)
    occ = _f32(occ, dx * dy * dz)
    tdf = np.zeros(dx * dy * dz, np.float32) if tdf is None else _f32(tdf, dx * dy * dz).copy()
    lib().oracle_compute_tdf(_p(occ), _p(tdf), C.c_int(dx), C.c_int(dy), C.c_int(dz), C.c_float(voxel_size),
                             C.c_float(trunc_margin))
    return tdf.reshape(dz, dy, dx)


def tdf_grid_from_points(pts, voxel_size=0.01, padding=15, want_occ=True):
    pts = _f32(pts)
    dims, origin = np.zeros(3, np.int32), np.zeros(3, np.float32)
    lib().oracle_tdf_grid_from_points(_p(pts), C.c_int64(pts.size // 3), C.c_float(voxel_size), C.c_int(padding), _p(dims),
                                      _p(origin), None)
    occ = None
    if want_occ:
        occ = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.float32)
        lib().oracle_tdf_grid_from_points(_p(pts), C.c_int64(pts.size // 3), C.c_float(voxel_size), C.c_int(padding),
                                          _p(dims), _p(origin), _p(occ))
        occ = occ.reshape(int(dims[2]), int(dims[1]), int(dims[0]))
    return dims, origin, occ


REF_TDF_HOST_LIB = HERE / "_ref" / "libref_tdf_host.so"
REF_TDF_CUDA_LIB = HERE / "_ref" / "libref_tdf_cuda.so"


def ref_host_compute_tdf(occ, dims, voxel_size=0.01, trunc_margin=0.05):
    """The reference's ComputeTDF kernel body (core/demo.cu:22-57) on host threads."""
    dx, dy, dz = (int(v) for v in dims)
    occ = _f32(occ, dx * dy * dz).copy()
    tdf = np.zeros(dx * dy * dz, np.float32)
    _load(REF_TDF_HOST_LIB).ref_host_compute_tdf(_p(occ), _p(tdf), C.c_int(dx), C.c_int(dy), C.c_int(dz), C.c_float(0),
                                                 C.c_float(0), C.c_float(0), C.c_float(voxel_size), C.c_float(trunc_margin))
    return tdf.reshape(dz, dy, dx)


def ref_cuda_compute_tdf(occ, dims, voxel_size=0.01, trunc_margin=0.05):
    """The reference's ComputeTDF kernel compiled for sm_100a, launched <<<2880, 512>>> per loop. (tdf, ms)."""
    dx, dy, dz = (int(v) for v in dims)
    occ = _f32(occ, dx * dy * dz)
    tdf = np.zeros(dx * dy * dz, np.float32)
    ms = C.c_float(0)
    rc = _load(REF_TDF_CUDA_LIB).ref_cuda_compute_tdf(_p(occ), _p(tdf), C.c_int(dx), C.c_int(dy), C.c_int(dz), C.c_float(0),
                                                      C.c_float(0), C.c_float(0), C.c_float(voxel_size),
                                                      C.c_float(trunc_margin), C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference ComputeTDF failed: cudaError {rc}")
    return tdf.reshape(dz, dy, dx), ms.value


# ---- the reference's own code on host threads (oracle/_ref/libref_host.so) --------------------------

def ref_host():
    return _load(REF_HOST_LIB)


def ref_host_fuse(cam_K, cam2base, depth_frames, dims, origin, voxel_size, trunc_margin, tsdf=None, weight=None):
    dx, dy, dz = (int(v) for v in dims)
    poses = _f32(cam2base).reshape(-1, 16).copy()
    depth = _f32(depth_frames)
    n = poses.shape[0]
    im_h, im_w = depth.shape[-2], depth.shape[-1]
    depth = depth.reshape(n, im_h * im_w)
    if tsdf is None:
        tsdf, weight = init_volume(dx * dy * dz)
    K = _f32(cam_K, 9).copy()
    L = ref_host()
    for f in range(n):
        L.ref_host_integrate(_p(K), _p(poses[f]), _p(depth[f]), C.c_int(im_h), C.c_int(im_w), C.c_int(dx), C.c_int(dy),
                             C.c_int(dz), C.c_float(origin[0]), C.c_float(origin[1]), C.c_float(origin[2]),
                             C.c_float(voxel_size), C.c_float(trunc_margin), _p(tsdf), _p(weight))
    return tsdf.reshape(dz, dy, dx), weight.reshape(dz, dy, dx)


def ref_host_multiply_matrix(a, b):
    a, b = _f32(a, 16), _f32(b, 16)
    out = np.empty(16, np.float32)
    ref_host().ref_host_multiply_matrix(_p(a), _p(b), _p(out))
    return out.reshape(4, 4)


def ref_host_invert_matrix(m):
    m = _f32(m, 16)
    out = np.zeros(16, np.float32)
    ok = ref_host().ref_host_invert_matrix(_p(m), _p(out))
    return bool(ok), out.reshape(4, 4)


def ref_host_save_surface(path, dims, voxel_size, origin, tsdf, weight, tsdf_thresh=0.2, weight_thresh=1.0):
    t, w = _f32(tsdf).copy(), _f32(weight).copy()
    ref_host().ref_host_save_surface(os.fsencode(path), C.c_int(dims[0]), C.c_int(dims[1]), C.c_int(dims[2]),
                                     C.c_float(voxel_size), C.c_float(origin[0]), C.c_float(origin[1]),
                                     C.c_float(origin[2]), _p(t), _p(w), C.c_float(tsdf_thresh), C.c_float(weight_thresh))


def ref_host_load_matrix(path, rows, cols):
    out = np.zeros(rows * cols, np.float32)
    ref_host().ref_host_load_matrix(os.fsencode(path), C.c_int(rows), C.c_int(cols), _p(out))
    return out


def ref_host_num_threads():
    return ref_host().ref_host_num_threads()


# ---- the reference kernel on the GPU (oracle/_ref/libref_cuda.so) -----------------------------------

def ref_cuda_fuse(cam_K, cam2base, depth_frames, dims, origin, voxel_size, trunc_margin, download=True, fmad=True):
    """Returns (tsdf, weight, kernel_ms) from the reference's Integrate<<<dim_z, dim_y>>> on the GPU
    (fmad=True: built like depth-fusion/compile.sh:18; fmad=False: the same file with --fmad=false)."""
    dx, dy, dz = (int(v) for v in dims)
    poses = _f32(cam2base).reshape(-1, 16)
    depth = _f32(depth_frames)
    n = poses.shape[0]
    im_h, im_w = depth.shape[-2], depth.shape[-1]
    K = _f32(cam_K, 9)
    t = np.empty(dx * dy * dz, np.float32) if download else None
    w = np.empty(dx * dy * dz, np.float32) if download else None
    ms = C.c_float(0)
    rc = _load(REF_CUDA_LIB if fmad else REF_CUDA_NOFMAD_LIB).ref_cuda_fuse(
        _p(K), _p(poses), _p(depth), C.c_int(n), C.c_int(im_h), C.c_int(im_w), C.c_int(dx), C.c_int(dy), C.c_int(dz),
        C.c_float(origin[0]), C.c_float(origin[1]), C.c_float(origin[2]), C.c_float(voxel_size),
        C.c_float(trunc_margin), _p(t) if download else None, _p(w) if download else None, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference CUDA kernel failed: cudaError {rc}")
    if download:
        return t.reshape(dz, dy, dx), w.reshape(dz, dy, dx), ms.value
    return None, None, ms.value
