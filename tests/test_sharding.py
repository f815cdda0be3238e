"""Multi-GPU host logic on CPU: world_size-2/3 gloo process groups.  Each rank fuses its z-slab
(with the oracle standing in for the GPU engine -- there is no GPU here), frames are broadcast
from rank 0, surface points are gathered in rank order; rank 0 checks the result against the
monolithic volume.  The same sharding module drives bench.py / the engine on real GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, weighted):
    import sys
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as graft
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg, oracle = graft.load_package(), graft.load_oracle()
        synth, sharding = pkg.synth, pkg.sharding
        frames = [0, 11, 23, 41]
        dims, voxel, origin, trunc = (48, 40, 37), 0.0625, (-1.5, -1.25, 0.5), 0.3125
        # rank 0 renders, everybody receives (NCCL broadcast on the GPU box, gloo here)
        mm = torch.zeros((len(frames), 48, 64), dtype=torch.int32)
        K = synth.CAM_K.copy()
        K[[0, 2, 4, 5]] *= np.float32(0.1)
        if rank == 0:
            mm = synth.render_depth_mm(frames, im_size=(64, 48), cam_K=K)
        sharding.broadcast_frames(mm, src=0)
        depth = oracle.depth_from_u16(mm.numpy().astype(np.uint16))
        c2w, b2w = synth.make_poses(frames)
        poses = np.stack([oracle.cam2base_from_world(b2w, m).ravel() for m in c2w])
        # equal-cost slabs have different heights: the gather must follow the actual ranges
        ranges = sharding.weighted_slab_ranges(np.linspace(1.0, 6.0, dims[2]), world) if weighted else None
        z0, z1 = ranges[rank] if weighted else sharding.slab_range(dims[2], rank, world)
        t, w = oracle.fuse(K, poses, depth, dims, origin, voxel, trunc, z_range=(z0, z1))
        pts = oracle.extract_surface(t, w, dims, origin, voxel, z_range=(z0, z1))
        cloud = sharding.gather_surface(torch.from_numpy(pts), dst=0)
        full_t = sharding.gather_slabs(torch.from_numpy(t), dims[2], dst=0, ranges=ranges)
        if rank == 0:
            t_ref, w_ref = oracle.fuse(K, poses, depth, dims, origin, voxel, trunc)
            pts_ref = oracle.extract_surface(t_ref, w_ref, dims, origin, voxel)
            ok = (np.array_equal(cloud.numpy().view(np.uint32), pts_ref.view(np.uint32))
                  and np.array_equal(full_t.numpy().view(np.uint32), t_ref.view(np.uint32)) and len(pts_ref) > 50)
            with open(os.path.join(out_dir, "result.txt"), "w") as f:
                f.write("ok" if ok else "mismatch")
        else:
            assert cloud is None and full_t is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,weighted", [(2, False), (3, False), (3, True)])
def test_sharded_fusion_equals_monolithic_gloo(world, weighted, tmp_path, oracle):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), weighted), nprocs=world, join=True)
    assert (tmp_path / "result.txt").read_text() == "ok"


def test_split_slabs_equal_cost_at_group_granularity(pkg):
    """tsdf_split_slabs (host routine of the C ABI): contiguous, cut at multiples of 8 planes, every slab non-empty,
    costs balanced as well as the granularity allows; refine_group_cost moves the cuts towards the slow rank."""
    sh = pkg.sharding
    rng = np.random.default_rng(3)
    for dim_z, world in [(1024, 8), (1024, 3), (500, 4), (2048, 8), (64, 8), (300, 2), (8, 1)]:
        n_groups = (dim_z + 7) // 8
        cost = np.concatenate([np.linspace(1, 40, n_groups - n_groups // 4), np.full(n_groups // 4, 0.2)]) * 1e6
        cost *= 1 + 0.05 * rng.standard_normal(n_groups)
        ranges = sh.split_groups(cost, dim_z, world, base_cost=0.05)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == dim_z
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])) and all(b > a for a, b in ranges)
        assert all(a % 8 == 0 for a, _ in ranges)
        eff = cost + 0.05 * cost.mean()
        sums = np.array([eff[a // 8:-(-b // 8)].sum() for a, b in ranges])
        if n_groups >= 16 * world:
            assert sums.max() / sums.mean() < 1.0 + 2.5 * eff.max() / sums.mean(), (dim_z, world, sums)
    ranges = sh.split_groups(np.ones(128), 1024, 4)
    assert ranges == [(0, 256), (256, 512), (512, 768), (768, 1024)]
    slow_last = sh.refine_plane_cost_piecewise(sh.plane_cost_from_groups(np.ones(128), 1024), ranges, [10.0, 10.0, 10.0, 16.0])
    new = sh.split_planes(slow_last, 4)
    assert new[-1][0] > ranges[-1][0]          # the slow last rank gets a shorter slab
    with pytest.raises(pkg.TsdfError):
        pkg.binding.split_slabs(np.ones(4, np.uint64), 32, 5)   # more slabs than groups


def test_slab_ranges_tile_the_volume(pkg):
    sharding = pkg.sharding
    for dim_z in (1, 7, 500, 1024, 2048):
        for world in (1, 2, 3, 4, 8):
            if world > dim_z:
                continue
            cuts = [sharding.slab_range(dim_z, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == dim_z
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    with pytest.raises(ValueError):
        sharding.slab_range(8, 8, 8)
    ranges = sharding.weighted_slab_ranges(np.linspace(1.0, 3.0, 64), 4)
    assert ranges[0][0] == 0 and ranges[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    costs = [np.linspace(1.0, 3.0, 64)[a:b].sum() for a, b in ranges]
    assert max(costs) / min(costs) < 1.15
