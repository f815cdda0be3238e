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


def gather_slabs(slab: torch.Tensor, dim_z: int, dst: int = 0):
    """Slab arrays [nz_r, dy, dx] -> the full [dim_z, dy, dx] volume on `dst` (for tsdf.bin)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return slab
    world, rank = dist.get_world_size(), dist.get_rank()
    nz_max = max(slab_range(dim_z, r, world)[1] - slab_range(dim_z, r, world)[0] for r in range(world))
    padded = torch.zeros((nz_max,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
    padded[: slab.shape[0]] = slab
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        z0, z1 = slab_range(dim_z, r, world)
        parts.append(bufs[r][: z1 - z0])
    return torch.cat(parts, dim=0)
