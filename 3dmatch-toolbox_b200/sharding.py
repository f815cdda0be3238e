"""z-slab sharding of one TSDF volume across the GPUs of a box.

Voxels never interact and frames are read-only, so the volume is cut into contiguous z-slabs
(z is the slowest index, depth-fusion/demo.cu:61, hence a slab is one contiguous byte range of
both arrays), every rank fuses ALL frames into its slab, and nothing is exchanged on the data
path.  Collectives (torch.distributed: NCCL on GPUs, gloo in the CPU tests) appear only at the
edges: the broadcast of the input frames from the rank that loaded / rendered them, and the
ordered gather of the per-slab surface points (rank order == ascending z == ascending linear
index, so the concatenation is the reference's tsdf.ply payload, depth-fusion/utils.hpp:36-53).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def slab_range(dim_z: int, rank: int, world: int) -> tuple[int, int]:
    """Planes [z_begin, z_end) of `rank`; slabs tile [0, dim_z) and differ by at most one plane."""
    if not 0 <= rank < world:
        raise ValueError("rank outside world")
    return (dim_z * rank) // world, (dim_z * (rank + 1)) // world


def weighted_slab_ranges(plane_cost, world: int) -> list[tuple[int, int]]:
    """Contiguous slabs with (nearly) equal summed cost, for frustum-dependent load balancing."""
    cost = np.asarray(plane_cost, dtype=np.float64)
    total = cost.sum()
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        z = int(np.searchsorted(cum, total * r / world))
        cuts.append(min(max(z, cuts[-1] + 1), len(cost) - (world - r)))
    cuts.append(len(cost))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def estimate_plane_cost(dims, origin, voxel_size, cam_K, cam2base, im_size=(640, 480), max_frames=48,
                        xy_samples=40, culled_cost=0.12) -> np.ndarray:
    """Relative fusion cost of every z-plane: the fraction of (voxel, frame) pairs that project into
    the image (those run the per-voxel classification) plus a small constant for culled voxels.

    Near planes see a small part of the frustum, far planes almost all of it, so equal-height slabs are
    badly unbalanced (measured 1.31x on 2 GPUs); slabs of equal summed cost are not.  Evaluated on a
    coarse sample (xy_samples^2 voxels per plane, every plane, at most max_frames frames) in float64 --
    it only steers the partition, never the arithmetic.
    """
    dx, dy, dz = (int(d) for d in dims)
    poses = np.asarray(cam2base, dtype=np.float64).reshape(-1, 4, 4)
    if len(poses) > max_frames:
        poses = poses[np.linspace(0, len(poses) - 1, max_frames).astype(int)]
    fx, fy, cx, cy = float(cam_K[0]), float(cam_K[4]), float(cam_K[2]), float(cam_K[5])
    xs = origin[0] + (np.linspace(0, dx - 1, xy_samples)) * voxel_size
    ys = origin[1] + (np.linspace(0, dy - 1, xy_samples)) * voxel_size
    zs = origin[2] + np.arange(dz) * voxel_size
    gx, gy = np.meshgrid(xs, ys, indexing="xy")
    cost = np.zeros(dz)
    for m in poses:
        r, t = m[:3, :3], m[:3, 3]
        # camera coordinates = R^T (p - t), separable in z
        px, py = gx - t[0], gy - t[1]
        bx = r[0, 0] * px + r[1, 0] * py
        by = r[0, 1] * px + r[1, 1] * py
        bz = r[0, 2] * px + r[1, 2] * py
        pz = zs - t[2]
        camx = bx[None] + r[2, 0] * pz[:, None, None]
        camy = by[None] + r[2, 1] * pz[:, None, None]
        camz = bz[None] + r[2, 2] * pz[:, None, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * camx / camz + cx
            v = fy * camy / camz + cy
        inside = (camz > 0) & (u >= -0.5) & (u < im_size[0] - 0.5) & (v >= -0.5) & (v < im_size[1] - 0.5)
        cost += inside.reshape(dz, -1).mean(axis=1)
    return cost / len(poses) + culled_cost


def measure_plane_cost(volume_factory, depth, poses, dim_z, chunk=64, sample_batches=4, batch=16) -> np.ndarray:
    """Per-plane fusion cost MEASURED with the engine itself: the volume is cut into `chunk`-plane pieces,
    every rank times its share of them (CUDA events, `sample_batches` evenly spaced K-frame passes of the
    real sequence) and the results are all-gathered, so every rank derives the same partition.
    volume_factory(z0, z1) must return a binding.Volume for that slab.  A few hundred ms, once per
    trajectory; captures what the analytic estimate cannot (surface bands, invalid pixels).
    """
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n_frames = poses.shape[0]
    starts = sorted({(max(n_frames - batch, 0) * b) // max(sample_batches - 1, 1) for b in range(sample_batches)})
    edges = list(range(0, dim_z, chunk)) + [dim_z]
    n_chunks = len(edges) - 1
    ms = np.zeros(n_chunks)
    for c in range(rank, n_chunks, world):
        with volume_factory(edges[c], edges[c + 1]) as vol:
            def run():
                for s0 in starts:
                    n = min(batch, n_frames - s0)
                    vol.integrate_batch(depth[s0:s0 + n], poses[s0:s0 + n], n_frames=n)
            run()
            vol.reset()
            vol.synchronize()
            vol.timer_start()
            run()
            ms[c] = vol.timer_stop()
    if world > 1:
        t = torch.from_numpy(ms).to(depth.device if hasattr(depth, "device") else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = t.cpu().numpy()
    density = ms / np.diff(edges)                       # cost per plane inside each chunk
    centers = 0.5 * (np.array(edges[:-1]) + np.array(edges[1:]) - 1)
    return np.maximum(np.interp(np.arange(dim_z), centers, density), 1e-9)


def refine_plane_cost(plane_cost, ranges, rank_ms, damping: float = 1.0) -> np.ndarray:
    """One step of feedback balancing: scale the per-plane cost model by measured / modelled time of
    every slab (interpolated between slab centres), so that the next equal-cost partition evens out
    the measured rank times.  `rank_ms[r]` is the time rank r needed for slab `ranges[r]`."""
    cost = np.asarray(plane_cost, dtype=np.float64)
    model = np.array([cost[a:b].sum() for a, b in ranges])
    ratio = np.asarray(rank_ms, dtype=np.float64) / np.maximum(model, 1e-12)
    ratio = (ratio / ratio.mean()) ** damping   # damping < 1: smaller steps, no overshoot on noisy rank times
    centers = np.array([0.5 * (a + b - 1) for a, b in ranges])
    return cost * np.interp(np.arange(len(cost)), centers, ratio)


def balanced_slab_range(dim_z, rank, world, plane_cost=None):
    """slab_range() when no cost is given, else the rank's slab of the equal-cost partition."""
    if plane_cost is None or world == 1:
        return slab_range(dim_z, rank, world)
    return weighted_slab_ranges(plane_cost, world)[rank]


def broadcast_frames(frames: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Depth frames (and anything else tensor-shaped) from `src` to every rank, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(frames, src=src)
    return frames


def gather_surface(points: torch.Tensor, dst: int = 0):
    """Per-slab surface points [n_r, 3] -> on `dst` the [sum n_r, 3] cloud in rank (= z) order.

    Counts travel by all_gather; the payload by a padded all_gather (slabs produce similar counts;
    the padding is a few MB at most).  Returns None on the other ranks.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return points
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([points.shape[0]], dtype=torch.int64, device=points.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, 3), dtype=points.dtype, device=points.device)
    padded[: points.shape[0]] = points
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: counts[r]] for r in range(world)], dim=0)


def gather_slabs(slab: torch.Tensor, dim_z: int, dst: int = 0, ranges=None):
    """Slab arrays [nz_r, dy, dx] -> the full [dim_z, dy, dx] volume on `dst` (for tsdf.bin).

    `ranges`: the (z_begin, z_end) of every rank (equal-cost slabs have different heights); None = the equal-height
    partition of slab_range().  The heights are cross-checked against what the ranks actually hold."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return slab
    world, rank = dist.get_world_size(), dist.get_rank()
    if ranges is None:
        ranges = [slab_range(dim_z, r, world) for r in range(world)]
    ranges = [(int(a), int(b)) for a, b in ranges]
    if len(ranges) != world or ranges[0][0] != 0 or ranges[-1][1] != dim_z or any(a[1] != b[0] for a, b in zip(ranges, ranges[1:])):
        raise ValueError(f"slab ranges {ranges} do not tile [0, {dim_z})")
    n = torch.tensor([slab.shape[0]], dtype=torch.int64, device=slab.device)
    heights = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(heights, n)
    heights = [int(h.item()) for h in heights]
    if heights != [b - a for a, b in ranges]:
        raise ValueError(f"ranks hold slabs of {heights} planes, the ranges say {[b - a for a, b in ranges]}")
    nz_max = max(heights)
    padded = torch.zeros((nz_max,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
    padded[: slab.shape[0]] = slab
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: heights[r]] for r in range(world)], dim=0)


# ---- equal-work slabs from the engine's own cull statistics (tsdf_plan_z_cost / tsdf_split_slabs) ----------------

def split_groups(group_cost, dim_z: int, world: int, base_cost: float = 0.05):
    """Contiguous z-slabs of (nearly) equal summed cost, cut at multiples of 8 planes.  `group_cost`: one value per
    group of 8 planes (any non-negative floats; scaled to integers for the C routine)."""
    from . import binding
    c = np.maximum(np.asarray(group_cost, dtype=np.float64), 0.0)
    scale = (2.0 ** 40) / max(float(c.max()), 1e-300)
    return binding.split_slabs(np.round(c * scale).astype(np.uint64), dim_z, world, base_cost)


def refine_group_cost(group_cost, ranges, rank_ms, base_cost: float = 0.05, damping: float = 0.8) -> np.ndarray:
    """Feedback balancing at group (8-plane) granularity: scale the cost model by measured / modelled time of every
    slab, interpolated between slab centres.  The model includes the same base cost per group as split_groups()."""
    cost = np.asarray(group_cost, dtype=np.float64)
    eff = cost + base_cost * cost.mean()
    g_ranges = [(a // 8, -(-b // 8)) for a, b in ranges]
    model = np.array([eff[a:b].sum() for a, b in g_ranges])
    ratio = np.asarray(rank_ms, dtype=np.float64) / np.maximum(model, 1e-300)
    ratio = (ratio / ratio.mean()) ** damping
    centers = np.array([0.5 * (a + b - 1) for a, b in g_ranges])
    scale = np.interp(np.arange(len(cost)), centers, ratio)
    # scale the effective cost, then take the base part out again so that split_groups() adds it back unchanged
    return np.maximum(eff * scale - base_cost * cost.mean(), 0.0)


# ---- plane-granular variant (bench.py): a group of 8 planes next to the far wall is ~6 % of a rank's work at 8 GPUs,
#      so the feedback loop needs finer cuts than tsdf_split_slabs' multiples of 8 (slab tiles are laid out relative to
#      the slab's first plane, so any cut is legal)

def plane_cost_from_groups(group_cost, dim_z: int, base_cost: float = 0.05) -> np.ndarray:
    """Per-plane cost model from the per-group cull statistics: a group's count spread over its planes, plus the
    base cost of a culled plane (relative to the mean)."""
    g = np.maximum(np.asarray(group_cost, dtype=np.float64), 0.0)
    per_plane = np.repeat(g / 8.0, 8)[:dim_z]
    return per_plane + base_cost * max(per_plane.mean(), 1e-300)


def split_planes(plane_cost, world: int):
    """Contiguous slabs of (nearly) equal summed cost, cut at single planes; every slab keeps at least one plane."""
    cost = np.maximum(np.asarray(plane_cost, dtype=np.float64), 0.0) + 1e-12
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    dim_z = len(cost)
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        z = int(np.searchsorted(cum, target))
        if z > 0 and target - cum[z - 1] < cum[z] - target:
            z -= 1
        cuts.append(min(max(z, cuts[-1] + 1), dim_z - (world - r)))
    cuts.append(dim_z)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def refine_plane_cost_piecewise(plane_cost, ranges, rank_ms, damping: float = 1.0) -> np.ndarray:
    """Feedback balancing: every slab's planes are scaled by (measured time / modelled cost) of that slab, so the model
    reproduces the measured rank times of the current partition exactly (damping = 1); what remains is the distribution
    INSIDE each slab, which the next measurement corrects."""
    cost = np.asarray(plane_cost, dtype=np.float64).copy()
    model = np.array([cost[a:b].sum() for a, b in ranges])
    ratio = np.asarray(rank_ms, dtype=np.float64) / np.maximum(model, 1e-300)
    ratio = (ratio / ratio.mean()) ** damping
    for (a, b), r in zip(ranges, ratio):
        cost[a:b] *= r
    return cost

