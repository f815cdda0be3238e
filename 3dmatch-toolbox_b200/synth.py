"""Deterministic synthetic RGB-D sequences for tests and bench.py.

An analytic room (axis-aligned box around the grid, far wall inside the volume) with five spheres
is ray-cast along a smooth camera arc that mimics the reference's sample sequence
(data/sample/depth-fusion-demo: ~0.3 m and ~5 degrees per 50 frames, intrinsics 585/585/320/240).
Depth is produced the way the reference consumes it: z-depth -> uint16 millimetres (the PNG
payload) -> (float)mm / 1000.0f with > 6 m zeroed (depth-fusion/utils.hpp:82-84).  10 % of the
pixels are invalidated (0) and 0.1 % set to the 65535 sentinel by a counter-based hash of
(seed, frame, pixel); seed = 20161128.

Ray casting runs in float64 torch ops on any device (CPU for the tests, the GPU for bench.py);
this is input synthesis, not part of the fusion path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SEED = 20161128
CAM_K = np.array([585.0, 0.0, 320.0, 0.0, 585.0, 240.0, 0.0, 0.0, 1.0], dtype=np.float32)

# room interior (base-frame camera coordinates, metres) and spheres (cx, cy, cz, r)
ROOM_LO = (-1.40, -1.15, -0.80)
ROOM_HI = (1.40, 1.15, 3.20)
SPHERES = ((-0.50, 0.30, 2.00, 0.35), (0.60, -0.20, 1.60, 0.25), (0.10, 0.60, 2.60, 0.40),
           (-0.80, -0.50, 2.80, 0.30), (0.90, 0.50, 2.40, 0.30))


def _rot(ax: float, ay: float, az: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ ry @ rx


def true_pose(frame: int) -> np.ndarray:
    """float64 camera -> scene transform of `frame` (identity at frame 0)."""
    s = float(frame)
    t = np.array([0.45 * math.sin(2 * math.pi * s / 410.0),
                  0.18 * (math.sin(2 * math.pi * s / 290.0 + 0.7) - math.sin(0.7)),
                  0.25 * (1.0 - math.cos(2 * math.pi * s / 530.0))])
    r = _rot(0.06 * math.sin(2 * math.pi * s / 230.0), 0.12 * math.sin(2 * math.pi * s / 370.0),
             0.03 * math.sin(2 * math.pi * s / 610.0))
    m = np.eye(4)
    m[:3, :3] = r
    m[:3, 3] = t
    return m


# scene -> "world" rigid transform, so that base2world is a general pose like the sample's
_WORLD = np.eye(4)
_WORLD[:3, :3] = _rot(0.21, -0.52, 0.13)
_WORLD[:3, 3] = (1.2, -0.7, 0.4)


def cam2world_f32(frame: int) -> np.ndarray:
    """What a frame-%06d.pose.txt would hold: camera -> world, rounded to float32."""
    return (_WORLD @ true_pose(frame)).astype(np.float32)


def _hash32(x: torch.Tensor) -> torch.Tensor:
    """32-bit integer mix on int64 tensors; every product stays below 2^63."""
    m = 0xFFFFFFFF
    x = x & m
    x = (((x >> 16) ^ x) * 0x45D9F3B) & m
    x = (((x >> 16) ^ x) * 0x45D9F3B) & m
    return (x >> 16) ^ x


@torch.no_grad()
def render_depth_mm(frames, base_frame: int = 0, im_size=(640, 480), cam_K=CAM_K, device="cpu",
                    seed: int = SEED) -> torch.Tensor:
    """uint16-range millimetre depth of the listed frame indices: int32 tensor [n, H, W] on `device`.

    The scene is fixed in the coordinates of camera `base_frame`.
    """
    w, h = im_size
    fx, fy, cx, cy = float(cam_K[0]), float(cam_K[4]), float(cam_K[2]), float(cam_K[5])
    dev = torch.device(device)
    f64 = torch.float64
    u = (torch.arange(w, device=dev, dtype=f64) - cx) / fx
    v = (torch.arange(h, device=dev, dtype=f64) - cy) / fy
    dcx = u.view(1, w).expand(h, w)
    dcy = v.view(h, 1).expand(h, w)
    pix = (torch.arange(h, device=dev, dtype=torch.int64).view(h, 1) * w
           + torch.arange(w, device=dev, dtype=torch.int64).view(1, w))
    base_inv = np.linalg.inv(true_pose(base_frame))
    out = torch.empty((len(frames), h, w), dtype=torch.int32, device=dev)
    inf = float("inf")
    for i, frame in enumerate(frames):
        m = base_inv @ true_pose(int(frame))
        r, o = m[:3, :3], m[:3, 3]
        # ray direction in scene coordinates; camera-frame z component is 1, so t is z-depth
        d = [dcx * r[a, 0] + dcy * r[a, 1] + r[a, 2] for a in range(3)]
        t_hit = torch.full((h, w), inf, dtype=f64, device=dev)
        for a in range(3):  # room walls, seen from inside
            wall = torch.where(d[a] > 0, torch.full_like(d[a], ROOM_HI[a]), torch.full_like(d[a], ROOM_LO[a]))
            t = (wall - float(o[a])) / d[a]
            t = torch.where((d[a] != 0) & (t > 0), t, torch.full_like(t, inf))
            t_hit = torch.minimum(t_hit, t)
        dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        for (sx, sy, sz, sr) in SPHERES:
            oc = (float(o[0]) - sx, float(o[1]) - sy, float(o[2]) - sz)
            b = d[0] * oc[0] + d[1] * oc[1] + d[2] * oc[2]
            c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sr * sr
            disc = b * b - dd * c
            t = (-b - torch.sqrt(torch.clamp(disc, min=0.0))) / dd
            t = torch.where((disc > 0) & (t > 0), t, torch.full_like(t, inf))
            t_hit = torch.minimum(t_hit, t)
        mm = torch.round(torch.clamp(t_hit, max=65.0) * 1000.0).to(torch.int64)
        mm = torch.clamp(mm, 0, 65535)
        hsh = _hash32(pix * 2654435761 + int(frame) * 40503 + seed)
        mm = torch.where(hsh % 1000 < 100, torch.zeros_like(mm), mm)          # 10 % invalid
        mm = torch.where(hsh % 1000 == 999, torch.full_like(mm, 65535), mm)   # 0.1 % sentinel
        out[i] = mm.to(torch.int32)
    return out


def mm_to_u16_numpy(mm: torch.Tensor) -> np.ndarray:
    return mm.cpu().numpy().astype(np.uint16)


def mm_to_u16_tensor(mm: torch.Tensor) -> torch.Tensor:
    """Same bits as uint16, carried in an int16 tensor (works on every device)."""
    return ((mm + 32768) % 65536 - 32768).to(torch.int16)


def make_poses(frames, base_frame: int = 0):
    """(cam2world_f32 [n,4,4], base2world_f32 [4,4]) as the pose files of a sequence would hold them."""
    c2w = np.stack([cam2world_f32(int(f)) for f in frames])
    return c2w, cam2world_f32(base_frame)


class GridSpec:
    """Volume placement (depth-fusion/demo.cu:92-99)."""

    def __init__(self, dim, voxel_size, origin=(-1.5, -1.5, 0.5), trunc_margin=None):
        self.dims = (dim, dim, dim) if np.isscalar(dim) else tuple(int(d) for d in dim)
        self.voxel_size = float(np.float32(voxel_size))
        self.origin = tuple(float(np.float32(o)) for o in origin)
        # demo.cu:96: trunc_margin = voxel_size * 5 (float arithmetic)
        self.trunc_margin = float(np.float32(voxel_size) * np.float32(5)) if trunc_margin is None \
            else float(np.float32(trunc_margin))

    @property
    def n_voxels(self):
        return self.dims[0] * self.dims[1] * self.dims[2]


# the BASELINE.json configs
CONFIGS = {
    "demo500": dict(grid=GridSpec(500, 0.006), frames=50),
    "512": dict(grid=GridSpec(512, 0.006), frames=300),
    "1024": dict(grid=GridSpec(1024, 0.003), frames=1000),
    "2048": dict(grid=GridSpec(2048, 0.0015), frames=2000),
    "frag300": dict(grid=GridSpec(300, 0.01), frames=50, fragments=256),
}
