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

