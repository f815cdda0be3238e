"""ctypes binding of lib/libtsdf_fusion.so (include/tsdf_fusion.h).

Host-side mirror of the reference's depth-fusion driver (depth-fusion/demo.cu:72-216) for the
test-suite and bench.py: the names follow the reference (`Integrate`, `invert_matrix`,
`multiply_matrix`, `SaveVoxelGrid2SurfacePointCloud`, `ReadDepth`, `LoadMatrixFromFile`).
There is no Python or CPU implementation behind any compute call: if the shared library is
missing, or no CUDA device is usable, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libtsdf_fusion.so"

TSDF_MAX_BATCH = 32
LEGACY_ENGINE = 1 << 26   # tuning bit without a kernel-variant meaning: any non-zero tuning selects the legacy engine

# every extern "C" symbol include/tsdf_fusion.h declares
EXPORTED_SYMBOLS = [
    "tsdf_create", "tsdf_destroy", "tsdf_reset", "tsdf_synchronize", "tsdf_get_stats", "tsdf_reset_stats",
    "tsdf_timer_start", "tsdf_timer_stop", "tsdf_integrate", "tsdf_integrate_batch",
    "tsdf_integrate_batch_u16", "tsdf_download", "tsdf_upload", "tsdf_device_ptrs", "tsdf_extract_surface",
    "tsdf_copy_surface", "tsdf_surface_device_ptr", "tsdf_write_ply", "tsdf_write_bin", "tsdf_load_matrix",
    "tsdf_read_depth_png", "tsdf_read_depth_png_u16", "tsdf_depth_from_u16", "tsdf_multiply_matrix",
    "tsdf_invert_matrix", "tsdf_Integrate", "tsdf_SaveVoxelGrid2SurfacePointCloud", "tsdf_last_error",
    "tsdf_version", "tsdf_device_count", "tsdf_selftest_rounding", "tsdf_ComputeTDF", "tsdf_tdf_grid_from_points",
    "tsdf_point_cloud_to_tdf", "tsdf_integrate_multi", "tsdf_plan_z_cost", "tsdf_get_z_cost", "tsdf_split_slabs",
    "tsdf_sharded_create", "tsdf_sharded_destroy", "tsdf_sharded_reset", "tsdf_sharded_synchronize",
    "tsdf_sharded_integrate_batch", "tsdf_sharded_extract_surface", "tsdf_sharded_copy_surface", "tsdf_sharded_write_bin",
    "tsdf_sharded_slab", "tsdf_copy_surface_normals", "tsdf_selftest_packed", "tsdf_selftest_pixel_margin",
]


class TsdfError(RuntimeError):
    pass


class VolumeDesc(C.Structure):
    _fields_ = [
        ("dim_x", C.c_int64), ("dim_y", C.c_int64), ("dim_z", C.c_int64),
        ("z_begin", C.c_int64), ("z_end", C.c_int64),
        ("origin_x", C.c_float), ("origin_y", C.c_float), ("origin_z", C.c_float),
        ("voxel_size", C.c_float), ("trunc_margin", C.c_float),
        ("im_width", C.c_int32), ("im_height", C.c_int32),
        ("cam_K", C.c_float * 9),
        ("device", C.c_int32), ("max_batch", C.c_int32), ("cull_level", C.c_int32), ("reserved0", C.c_int32),
        ("stream", C.c_void_p), ("ext_tsdf", C.c_void_p), ("ext_weight", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("frames_integrated", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


VALIDATION_LIB_PATH = PKG_DIR / "lib" / "libtsdf_fusion_validation.so"
_libs = {}


def load_library(validation: bool = False) -> C.CDLL:
    """dlopen the engine; fails loudly when it has not been built (see __graft_entry__.build).

    validation=True loads lib/libtsdf_fusion_validation.so: the same library compiled with -DTSDF_VALIDATION, which
    additionally carries every earlier kernel generation / tuning variant (selected through
    tsdf_volume_desc.reserved0) for the bit-identity tests.  The shipped library rejects reserved0 != 0."""
    if validation in _libs:
        return _libs[validation]
    path = VALIDATION_LIB_PATH if validation else LIB_PATH
    if not path.exists():
        raise TsdfError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(str(path))
    fp, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int64)
    vp = C.c_void_p
    sig = {
        "tsdf_create": [C.POINTER(VolumeDesc), C.POINTER(vp)],
        "tsdf_reset": [vp], "tsdf_synchronize": [vp],
        "tsdf_get_stats": [vp, C.POINTER(Stats)], "tsdf_reset_stats": [vp],
        "tsdf_timer_start": [vp], "tsdf_timer_stop": [vp, fp],
        "tsdf_integrate": [vp, vp, vp],
        "tsdf_integrate_batch": [vp, C.c_int32, vp, vp],
        "tsdf_integrate_batch_u16": [vp, C.c_int32, vp, vp],
        "tsdf_download": [vp, vp, vp], "tsdf_upload": [vp, vp, vp],
        "tsdf_device_ptrs": [vp, C.POINTER(vp), C.POINTER(vp)],
        "tsdf_extract_surface": [vp, C.c_float, C.c_float, i64p],
        "tsdf_copy_surface": [vp, vp, C.c_int64],
        "tsdf_copy_surface_normals": [vp, vp, C.c_int64],
        "tsdf_surface_device_ptr": [vp, C.POINTER(vp), i64p],
        "tsdf_write_ply": [C.c_char_p, vp, C.c_int64],
        "tsdf_write_bin": [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float,
                           C.c_float, C.c_float, vp],
        "tsdf_load_matrix": [C.c_char_p, C.c_int32, C.c_int32, vp],
        "tsdf_read_depth_png": [C.c_char_p, C.c_int32, C.c_int32, vp],
        "tsdf_read_depth_png_u16": [C.c_char_p, C.c_int32, C.c_int32, vp],
        "tsdf_invert_matrix": [vp, vp],
        "tsdf_Integrate": [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                           C.c_float, C.c_float, C.c_float, vp, vp, vp],
        "tsdf_SaveVoxelGrid2SurfacePointCloud": [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                                 C.c_float, C.c_float, vp, vp, C.c_float, C.c_float],
        "tsdf_ComputeTDF": [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp],
        "tsdf_tdf_grid_from_points": [vp, C.c_int64, C.c_float, C.c_int32, vp, vp],
        "tsdf_point_cloud_to_tdf": [vp, C.c_int64, vp, vp, C.c_float, C.c_float, vp],
        "tsdf_integrate_multi": [vp, C.c_int32, C.c_int32, vp, C.c_int32, vp],
        "tsdf_plan_z_cost": [C.POINTER(VolumeDesc), C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, C.c_int32],
        "tsdf_get_z_cost": [vp, vp, C.c_int32],
        "tsdf_split_slabs": [vp, C.c_int32, C.c_int64, C.c_int32, C.c_double, vp],
        "tsdf_sharded_create": [C.POINTER(VolumeDesc), C.c_int32, vp, vp, C.POINTER(vp)],
        "tsdf_sharded_reset": [vp], "tsdf_sharded_synchronize": [vp],
        "tsdf_sharded_integrate_batch": [vp, C.c_int32, vp, C.c_int32, vp],
        "tsdf_sharded_extract_surface": [vp, C.c_float, C.c_float, i64p],
        "tsdf_sharded_copy_surface": [vp, vp, C.c_int64],
        "tsdf_sharded_write_bin": [vp, C.c_char_p],
        "tsdf_sharded_slab": [vp, C.c_int32, C.POINTER(vp), i64p, i64p, C.POINTER(C.c_int32)],
        "tsdf_device_count": [C.POINTER(C.c_int32)],
        "tsdf_selftest_rounding": [C.c_int32, i64p],
        "tsdf_selftest_packed": [C.c_int32, i64p],
        "tsdf_selftest_pixel_margin": [C.c_int32, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.tsdf_destroy.argtypes = [vp]
    lib.tsdf_destroy.restype = None
    lib.tsdf_sharded_destroy.argtypes = [vp]
    lib.tsdf_sharded_destroy.restype = None
    lib.tsdf_multiply_matrix.argtypes = [vp, vp, vp]
    lib.tsdf_multiply_matrix.restype = None
    lib.tsdf_depth_from_u16.argtypes = [vp, C.c_int64, vp]
    lib.tsdf_depth_from_u16.restype = None
    lib.tsdf_last_error.restype = C.c_char_p
    lib.tsdf_version.restype = C.c_char_p
    _libs[validation] = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        raise TsdfError(f"{what} failed (status {rc}): {load_library().tsdf_last_error().decode()}")


def _f32(a, n=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


def _ptr(a) -> int:
    """Address of a numpy array, a torch tensor (host or CUDA) or a raw int."""
    if isinstance(a, (int, np.integer)):
        return int(a)
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return int(a.ctypes.data)


# ---- pose algebra / file helpers (host; reference names) ------------------------------------------

def multiply_matrix(m1, m2) -> np.ndarray:
    """depth-fusion/utils.hpp:89-109."""
    a, b = _f32(m1, 16), _f32(m2, 16)
    out = np.empty(16, np.float32)
    load_library().tsdf_multiply_matrix(a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out.reshape(4, 4)


def invert_matrix(m):
    """depth-fusion/utils.hpp:112-238.  Returns (ok, inverse)."""
    a = _f32(m, 16)
    out = np.zeros(16, np.float32)
    ok = load_library().tsdf_invert_matrix(a.ctypes.data, out.ctypes.data)
    return bool(ok), out.reshape(4, 4)


def cam2base_from_world(base2world, cam2world) -> np.ndarray:
    """depth-fusion/demo.cu:127-128,171: cam2base = inv(base2world) * cam2world."""
    _, inv = invert_matrix(base2world)
    return multiply_matrix(inv, cam2world)


def LoadMatrixFromFile(path, rows, cols) -> np.ndarray:
    out = np.empty(rows * cols, np.float32)
    _check(load_library().tsdf_load_matrix(os.fsencode(path), rows, cols, out.ctypes.data), "tsdf_load_matrix")
    return out.reshape(rows, cols)


def ReadDepth(path, im_height, im_width) -> np.ndarray:
    out = np.empty((im_height, im_width), np.float32)
    _check(load_library().tsdf_read_depth_png(os.fsencode(path), im_height, im_width, out.ctypes.data),
           "tsdf_read_depth_png")
    return out


def read_depth_png_u16(path, im_height, im_width) -> np.ndarray:
    out = np.empty((im_height, im_width), np.uint16)
    _check(load_library().tsdf_read_depth_png_u16(os.fsencode(path), im_height, im_width, out.ctypes.data),
           "tsdf_read_depth_png_u16")
    return out


def depth_from_u16(mm) -> np.ndarray:
    mm = np.ascontiguousarray(mm, dtype=np.uint16)
    out = np.empty(mm.shape, np.float32)
    load_library().tsdf_depth_from_u16(mm.ctypes.data, mm.size, out.ctypes.data)
    return out


def write_ply(path, xyz):
    xyz = _f32(xyz)
    _check(load_library().tsdf_write_ply(os.fsencode(path), xyz.ctypes.data, xyz.size // 3), "tsdf_write_ply")


def write_bin(path, dims, origin, voxel_size, trunc_margin, tsdf):
    t = _f32(tsdf, int(dims[0]) * int(dims[1]) * int(dims[2]))
    _check(load_library().tsdf_write_bin(os.fsencode(path), dims[0], dims[1], dims[2], origin[0], origin[1],
                                         origin[2], voxel_size, trunc_margin, t.ctypes.data), "tsdf_write_bin")


def SaveVoxelGrid2SurfacePointCloud(file_name, dim_x, dim_y, dim_z, voxel_size, ox, oy, oz, tsdf, weight,
                                    tsdf_thresh, weight_thresh):
    """depth-fusion/utils.hpp:14-17 argument list; host arrays in, tsdf.ply out (scan on the GPU)."""
    n = dim_x * dim_y * dim_z
    t, w = _f32(tsdf, n), _f32(weight, n)
    _check(load_library().tsdf_SaveVoxelGrid2SurfacePointCloud(
        os.fsencode(file_name), dim_x, dim_y, dim_z, voxel_size, ox, oy, oz, t.ctypes.data, w.ctypes.data,
        tsdf_thresh, weight_thresh), "tsdf_SaveVoxelGrid2SurfacePointCloud")


def Integrate(cam_K, cam2base, depth_im, im_height, im_width, dim_x, dim_y, dim_z, ox, oy, oz, voxel_size,
              trunc_margin, tsdf, weight, stream=0):
    """depth-fusion/demo.cu:18-22 argument list; every pointer is a DEVICE pointer (int / torch CUDA tensor)."""
    _check(load_library().tsdf_Integrate(_ptr(cam_K), _ptr(cam2base), _ptr(depth_im), im_height, im_width, dim_x,
                                         dim_y, dim_z, ox, oy, oz, voxel_size, trunc_margin, _ptr(tsdf),
                                         _ptr(weight), stream), "tsdf_Integrate")


def ComputeTDF(occ, tdf, dim_x, dim_y, dim_z, ox, oy, oz, voxel_size, trunc_margin, stream=0):
    """core/demo.cu:22-27 argument list (without CUDA_LOOP_IDX); DEVICE pointers / CUDA tensors."""
    _check(load_library().tsdf_ComputeTDF(_ptr(occ), _ptr(tdf), dim_x, dim_y, dim_z, ox, oy, oz, voxel_size,
                                          trunc_margin, stream), "tsdf_ComputeTDF")


def tdf_grid_from_points(pts, voxel_size=0.01, padding=15):
    """core/demo.cu:100-126 -> (dims int32[3], origin float32[3])."""
    pts = _f32(pts)
    dims, origin = np.zeros(3, np.int32), np.zeros(3, np.float32)
    _check(load_library().tsdf_tdf_grid_from_points(pts.ctypes.data, pts.size // 3, voxel_size, padding, dims.ctypes.data,
                                                    origin.ctypes.data), "tsdf_tdf_grid_from_points")
    return dims, origin


def point_cloud_to_tdf(pts, dims, origin, voxel_size=0.01, trunc_margin=0.05) -> np.ndarray:
    """core/demo.cu:128-163: TDF grid [dz, dy, dx] of a host point cloud (computed on the GPU)."""
    pts = _f32(pts)
    dims = np.ascontiguousarray(dims, np.int32)
    origin = _f32(origin, 3)
    out = np.empty((int(dims[2]), int(dims[1]), int(dims[0])), np.float32)
    _check(load_library().tsdf_point_cloud_to_tdf(pts.ctypes.data, pts.size // 3, dims.ctypes.data, origin.ctypes.data,
                                                  voxel_size, trunc_margin, out.ctypes.data), "tsdf_point_cloud_to_tdf")
    return out


def _frames_arg(depth_frames):
    """(object to keep alive, address, is_u16, element count) of a numpy array / torch tensor of depth frames."""
    if isinstance(depth_frames, np.ndarray):
        u16 = depth_frames.dtype == np.uint16
        a = np.ascontiguousarray(depth_frames, dtype=np.uint16 if u16 else np.float32)
        return a, int(a.ctypes.data), u16, a.size
    import torch
    u16 = depth_frames.dtype in (torch.uint16, torch.int16)
    if not depth_frames.is_contiguous():
        raise ValueError("depth tensor must be contiguous")
    return depth_frames, int(depth_frames.data_ptr()), u16, depth_frames.numel()


def make_desc(dims, origin, voxel_size, trunc_margin, cam_K, im_size=(640, 480), z_range=None, device=0, max_batch=0,
              cull_level=-1, tuning=0) -> VolumeDesc:
    d = VolumeDesc()
    d.dim_x, d.dim_y, d.dim_z = (int(v) for v in dims)
    d.z_begin, d.z_end = (0, d.dim_z) if z_range is None else (int(z_range[0]), int(z_range[1]))
    d.origin_x, d.origin_y, d.origin_z = (float(v) for v in origin)
    d.voxel_size, d.trunc_margin = float(voxel_size), float(trunc_margin)
    d.im_width, d.im_height = int(im_size[0]), int(im_size[1])
    d.cam_K = (C.c_float * 9)(*[float(v) for v in np.asarray(cam_K, np.float32).ravel()])
    d.device, d.max_batch, d.cull_level = int(device), int(max_batch), int(cull_level)
    d.reserved0 = int(tuning)
    return d


def plan_z_cost(dims, origin, voxel_size, trunc_margin, cam_K, depth_frames, cam2base, im_size=(640, 480), device=0,
                frame_step=1, cull_level=-1) -> np.ndarray:
    """Work per group of 8 z-planes of the whole grid for a frame sequence (tsdf_plan_z_cost): uint64 [ceil(dz / 8)]."""
    d = make_desc(dims, origin, voxel_size, trunc_margin, cam_K, im_size, device=device, cull_level=cull_level)
    keep, addr, u16, size = _frames_arg(depth_frames)
    poses = _f32(cam2base)
    n = poses.size // 16
    if size < n * d.im_width * d.im_height:
        raise ValueError("depth_frames smaller than the pose list")
    cost = np.zeros((int(dims[2]) + 7) // 8, np.uint64)
    _check(load_library().tsdf_plan_z_cost(C.byref(d), n, addr, int(u16), poses.ctypes.data, int(frame_step),
                                           cost.ctypes.data, cost.size), "tsdf_plan_z_cost")
    return cost


def split_slabs(cost, dim_z, n_slabs, base_cost=0.05):
    """Equal-cost contiguous z-slabs cut at multiples of 8 planes (tsdf_split_slabs): list of (z_begin, z_end)."""
    cost = np.ascontiguousarray(cost, dtype=np.uint64)
    cuts = np.zeros(n_slabs + 1, np.int64)
    _check(load_library().tsdf_split_slabs(cost.ctypes.data, cost.size, int(dim_z), int(n_slabs), float(base_cost),
                                           cuts.ctypes.data), "tsdf_split_slabs")
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(n_slabs)]


def integrate_multi(volumes, depth_list, pose_list, n_frames=None):
    """tsdf_integrate_multi: independent volumes of identical shape on one device, one launch."""
    n_vols = len(volumes)
    keeps, addrs, u16s, poses = [], [], [], []
    for dep, pos in zip(depth_list, pose_list):
        keep, addr, u16, _ = _frames_arg(dep)
        keeps.append(keep); addrs.append(addr); u16s.append(u16)
        poses.append(_f32(pos))
    if len(set(u16s)) != 1:
        raise ValueError("all volumes must get the same depth type")
    n = poses[0].size // 16 if n_frames is None else int(n_frames)
    vol_arr = (C.c_void_p * n_vols)(*[v._h for v in volumes])
    frame_arr = (C.c_void_p * n_vols)(*addrs)
    pose_arr = (C.c_void_p * n_vols)(*[p.ctypes.data for p in poses])
    for v, k, p in zip(volumes, keeps, poses):
        v._keepalive.append((k, p))
    _check(volumes[0]._lib.tsdf_integrate_multi(vol_arr, n_vols, n, frame_arr, int(u16s[0]), pose_arr),
           "tsdf_integrate_multi")


def device_count() -> int:
    n = C.c_int32(0)
    _check(load_library().tsdf_device_count(C.byref(n)), "tsdf_device_count")
    return n.value


def selftest_packed(device=0) -> int:
    n = C.c_int64(-1)
    _check(load_library().tsdf_selftest_packed(device, C.byref(n)), "tsdf_selftest_packed")
    return n.value


def selftest_pixel_margin(fx, c0, size, device=0):
    """(max relative error of rcp.approx in units of 2^-24, max |u' - u| as a fraction of pixel_round_margin)."""
    a, b = C.c_float(-1), C.c_float(-1)
    _check(load_library().tsdf_selftest_pixel_margin(device, fx, c0, size, C.byref(a), C.byref(b)), "tsdf_selftest_pixel_margin")
    return a.value, b.value


def selftest_rounding(device=0) -> int:
    n = C.c_int64(-1)
    _check(load_library().tsdf_selftest_rounding(device, C.byref(n)), "tsdf_selftest_rounding")
    return n.value


# ---- the engine handle -------------------------------------------------------------------------------

class Volume:
    """A TSDF volume (or one z-slab of it) resident in the HBM of one GPU."""

    def __init__(self, dims, origin, voxel_size, trunc_margin, cam_K, im_size=(640, 480), z_range=None,
                 device=0, max_batch=0, cull_level=-1, stream=0, ext_tsdf=None, ext_weight=None, tuning=0,
                 validation=None):
        # tuning != 0 selects a kernel variant of the validation flavour of the library (LEGACY_ENGINE = the round-1
        # per-pass launch sequence with its default kernel)
        self._lib = load_library(bool(tuning) if validation is None else bool(validation))
        d = VolumeDesc()
        d.dim_x, d.dim_y, d.dim_z = (int(v) for v in dims)
        d.z_begin, d.z_end = (0, d.dim_z) if z_range is None else (int(z_range[0]), int(z_range[1]))
        d.origin_x, d.origin_y, d.origin_z = (float(v) for v in origin)
        d.voxel_size, d.trunc_margin = float(voxel_size), float(trunc_margin)
        d.im_width, d.im_height = int(im_size[0]), int(im_size[1])
        d.cam_K = (C.c_float * 9)(*[float(v) for v in np.asarray(cam_K, np.float32).ravel()])
        d.device, d.max_batch, d.cull_level = int(device), int(max_batch), int(cull_level)
        d.reserved0 = int(tuning)  # bits 0-7: kernel tuning variant, bit 8: exact path only (validation)
        d.stream = stream or None
        d.ext_tsdf = _ptr(ext_tsdf) if ext_tsdf is not None else None
        d.ext_weight = _ptr(ext_weight) if ext_weight is not None else None
        self.desc = d
        self._h = C.c_void_p()
        _check(self._lib.tsdf_create(C.byref(d), C.byref(self._h)), "tsdf_create")
        self.dims = (d.dim_x, d.dim_y, d.dim_z)
        self.z_range = (d.z_begin, d.z_end)
        self.n_voxels = (d.z_end - d.z_begin) * d.dim_y * d.dim_x
        self.frame_px = d.im_width * d.im_height
        self._keepalive = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tsdf_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _check(self._lib.tsdf_reset(self._h), "tsdf_reset")

    def synchronize(self):
        _check(self._lib.tsdf_synchronize(self._h), "tsdf_synchronize")
        self._keepalive.clear()

    def integrate(self, depth, cam2base):
        """One reference launch (demo.cu:173-182). depth: numpy (host) or torch tensor (host / CUDA)."""
        self.integrate_batch(depth, np.asarray(cam2base, np.float32).reshape(1, 16), n_frames=1)

    def integrate_batch(self, depth_frames, cam2base, n_frames=None):
        poses = _f32(cam2base)
        n = poses.size // 16 if n_frames is None else int(n_frames)
        u16 = False
        if isinstance(depth_frames, np.ndarray):
            u16 = depth_frames.dtype == np.uint16
            depth_frames = np.ascontiguousarray(depth_frames, dtype=np.uint16 if u16 else np.float32)
            size = depth_frames.size
        else:  # torch tensor
            import torch
            u16 = depth_frames.dtype in (torch.uint16, torch.int16)
            if not depth_frames.is_contiguous():
                raise ValueError("depth tensor must be contiguous")
            size = depth_frames.numel()
        if size < n * self.frame_px or poses.size < 16 * n:
            raise ValueError("depth_frames / cam2base smaller than n_frames")
        fn = self._lib.tsdf_integrate_batch_u16 if u16 else self._lib.tsdf_integrate_batch
        if len(self._keepalive) >= 64:   # a long loop without synchronize(): wait once instead of growing forever
            self.synchronize()
        self._keepalive.append((depth_frames, poses))  # async H2D / the staging kernel may still read them
        _check(fn(self._h, n, _ptr(depth_frames), poses.ctypes.data), "tsdf_integrate_batch")

    def download(self, want_tsdf=True, want_weight=True):
        nz = self.z_range[1] - self.z_range[0]
        shape = (nz, self.dims[1], self.dims[0])
        t = np.empty(shape, np.float32) if want_tsdf else None
        w = np.empty(shape, np.float32) if want_weight else None
        _check(self._lib.tsdf_download(self._h, t.ctypes.data if want_tsdf else None,
                                       w.ctypes.data if want_weight else None), "tsdf_download")
        self._keepalive.clear()
        return t, w

    def upload(self, tsdf=None, weight=None):
        t = _f32(tsdf, self.n_voxels) if tsdf is not None else None
        w = _f32(weight, self.n_voxels) if weight is not None else None
        _check(self._lib.tsdf_upload(self._h, t.ctypes.data if t is not None else None,
                                     w.ctypes.data if w is not None else None), "tsdf_upload")

    def device_ptrs(self):
        t, w = C.c_void_p(), C.c_void_p()
        _check(self._lib.tsdf_device_ptrs(self._h, C.byref(t), C.byref(w)), "tsdf_device_ptrs")
        return t.value, w.value

    def extract_surface_count(self, tsdf_thresh=0.2, weight_thresh=1.0) -> int:
        n = C.c_int64(0)
        _check(self._lib.tsdf_extract_surface(self._h, tsdf_thresh, weight_thresh, C.byref(n)),
               "tsdf_extract_surface")
        return n.value

    def extract_surface(self, tsdf_thresh=0.2, weight_thresh=1.0, out=None) -> np.ndarray:
        """utils.hpp:20-53 thresholds default to demo.cu:194's 0.2f / 1.0f. Returns (N, 3) float32."""
        n = self.extract_surface_count(tsdf_thresh, weight_thresh)
        if out is None:
            out = np.empty((n, 3), np.float32)
        _check(self._lib.tsdf_copy_surface(self._h, _ptr(out), n), "tsdf_copy_surface")
        return out[:n] if isinstance(out, np.ndarray) else out

    def surface_normals(self) -> np.ndarray:
        """Unit normals of the points of the last extract_surface(), same order (N, 3)."""
        _, n = self.surface_device_ptr()
        out = np.zeros((n, 3), np.float32)
        _check(self._lib.tsdf_copy_surface_normals(self._h, out.ctypes.data, n), "tsdf_copy_surface_normals")
        return out

    def surface_device_ptr(self):
        p, n = C.c_void_p(), C.c_int64(0)
        _check(self._lib.tsdf_surface_device_ptr(self._h, C.byref(p), C.byref(n)), "tsdf_surface_device_ptr")
        return p.value, n.value

    def z_cost(self) -> np.ndarray:
        """Kept (tile, frame) pairs per group of 8 planes of this slab, accumulated by the fusion calls so far."""
        nz = self.z_range[1] - self.z_range[0]
        cost = np.zeros((nz + 7) // 8, np.uint64)
        _check(self._lib.tsdf_get_z_cost(self._h, cost.ctypes.data, cost.size), "tsdf_get_z_cost")
        return cost

    def stats(self) -> dict:
        s = Stats()
        _check(self._lib.tsdf_get_stats(self._h, C.byref(s)), "tsdf_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def reset_stats(self):
        _check(self._lib.tsdf_reset_stats(self._h), "tsdf_reset_stats")

    def timer_start(self):
        _check(self._lib.tsdf_timer_start(self._h), "tsdf_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        _check(self._lib.tsdf_timer_stop(self._h, C.byref(ms)), "tsdf_timer_stop")
        return ms.value


class ShardedVolume:
    """One volume cut into z-slabs over several GPUs, driven from one process (tsdf_sharded_*): frames are uploaded
    once and broadcast with NCCL, slabs fuse independently, points are gathered in z order to the first device."""

    def __init__(self, dims, origin, voxel_size, trunc_margin, cam_K, n_slabs, devices=None, z_cuts=None,
                 im_size=(640, 480), max_batch=0, cull_level=-1):
        self._lib = load_library()
        d = make_desc(dims, origin, voxel_size, trunc_margin, cam_K, im_size, max_batch=max_batch, cull_level=cull_level)
        dev = np.ascontiguousarray(devices, dtype=np.int32) if devices is not None else None
        cuts = np.ascontiguousarray(z_cuts, dtype=np.int64) if z_cuts is not None else None
        if dev is not None and dev.size != n_slabs:
            raise ValueError("one device per slab")
        if cuts is not None and cuts.size != n_slabs + 1:
            raise ValueError("n_slabs + 1 cuts")
        self._h = C.c_void_p()
        _check(self._lib.tsdf_sharded_create(C.byref(d), n_slabs, dev.ctypes.data if dev is not None else None,
                                             cuts.ctypes.data if cuts is not None else None, C.byref(self._h)),
               "tsdf_sharded_create")
        self.dims, self.n_slabs = tuple(int(v) for v in dims), n_slabs
        self._keepalive = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tsdf_sharded_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def slabs(self):
        out = []
        for i in range(self.n_slabs):
            a, b, dv = C.c_int64(0), C.c_int64(0), C.c_int32(0)
            _check(self._lib.tsdf_sharded_slab(self._h, i, None, C.byref(a), C.byref(b), C.byref(dv)), "tsdf_sharded_slab")
            out.append((a.value, b.value, dv.value))
        return out

    def reset(self):
        _check(self._lib.tsdf_sharded_reset(self._h), "tsdf_sharded_reset")

    def synchronize(self):
        _check(self._lib.tsdf_sharded_synchronize(self._h), "tsdf_sharded_synchronize")
        self._keepalive.clear()

    def integrate_batch(self, depth_frames, cam2base, n_frames=None):
        keep, addr, u16, _ = _frames_arg(depth_frames)
        poses = _f32(cam2base)
        n = poses.size // 16 if n_frames is None else int(n_frames)
        self._keepalive.append((keep, poses))
        _check(self._lib.tsdf_sharded_integrate_batch(self._h, n, addr, int(u16), poses.ctypes.data),
               "tsdf_sharded_integrate_batch")

    def extract_surface(self, tsdf_thresh=0.2, weight_thresh=1.0) -> np.ndarray:
        n = C.c_int64(0)
        _check(self._lib.tsdf_sharded_extract_surface(self._h, tsdf_thresh, weight_thresh, C.byref(n)),
               "tsdf_sharded_extract_surface")
        out = np.empty((n.value, 3), np.float32)
        _check(self._lib.tsdf_sharded_copy_surface(self._h, out.ctypes.data, n.value), "tsdf_sharded_copy_surface")
        self._keepalive.clear()
        return out

    def write_bin(self, path):
        _check(self._lib.tsdf_sharded_write_bin(self._h, os.fsencode(path)), "tsdf_sharded_write_bin")
