"""GPU parity tests of the stream engine (tsdf_stream_kernels.cu: persistent work-list kernel, 32-frame chunks,
sticky band voxels) through the C ABI: against the CPU oracle on small grids, against the exact-only validation
kernel and the round-1 launch sequence on grids the oracle would take minutes for."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

ORIGIN = (-1.5, -1.5, 0.5)


def fuse_gpu(pkg, K, poses, depth, dims, origin, voxel, trunc, **kw):
    with pkg.Volume(dims, origin, voxel, trunc, K, im_size=(depth.shape[-1], depth.shape[-2]), **kw) as vol:
        vol.integrate_batch(depth, poses)
        t, w = vol.download()
        pts = vol.extract_surface()
    return t, w, pts


@pytest.mark.parametrize("dims,voxel,chunk,cull,n_frames", [
    ((96, 80, 64), 0.03125, 32, 2, 40),    # two chunks (32 + 8)
    ((96, 80, 64), 0.03125, 5, 2, 23),     # ragged last chunk
    ((96, 80, 64), 0.03125, 1, 1, 7),      # one frame per chunk, frustum cull only
    ((100, 100, 100), 0.03, 32, 0, 33),    # no culling: every frame reaches every tile; partial x tiles
    ((64, 36, 20), 0.05, 16, 2, 50),       # ragged y tiles and z tail
    ((32, 32, 8), 0.09, 3, 2, 60),         # ONE block tile, 20 chunks: every item waits for its predecessor
])
def test_stream_bit_exact_vs_oracle(pkg, oracle, sequence, dims, voxel, chunk, cull, n_frames):
    frames = [(7 * i) % 997 for i in range(n_frames)]
    depth, _, poses, K = sequence(frames)
    trunc = float(np.float32(voxel) * np.float32(5))
    t_ref, w_ref = oracle.fuse(K, poses, depth, dims, ORIGIN, voxel, trunc)
    t, w, pts = fuse_gpu(pkg, K, poses, depth, dims, ORIGIN, voxel, trunc, max_batch=chunk, cull_level=cull)
    assert w_ref.max() >= 3
    assert_bit_equal(w, w_ref, "weights")
    assert_bit_equal(t, t_ref, "tsdf")
    assert_bit_equal(pts, oracle.extract_surface(t_ref, w_ref, dims, ORIGIN, voxel), "surface points")


def test_stream_equals_exact_only_and_legacy(pkg, sequence):
    """~2e9 voxel-frames: stream engine (several chunk sizes / cull levels) == every voxel through the plain
    restatement (validation kernel, tuning bit 8) == the round-1 launch sequence."""
    frames = list(range(0, 192, 3)) + [500, 640, 777, 901]
    depth, _, poses, K = sequence(frames)
    for dims, voxel in [((320, 320, 320), 0.0096), ((512, 256, 128), 0.006)]:
        trunc = float(np.float32(voxel) * np.float32(5))
        exact = fuse_gpu(pkg, K, poses, depth, dims, ORIGIN, voxel, trunc, max_batch=8, cull_level=0, tuning=256)
        legacy = fuse_gpu(pkg, K, poses, depth, dims, ORIGIN, voxel, trunc, max_batch=8, cull_level=2,
                          tuning=pkg.binding.LEGACY_ENGINE)
        for a, b, name in zip(legacy, exact, ("tsdf", "weight", "points")):
            assert_bit_equal(a, b, f"legacy {name} dims {dims}")
        for chunk, cull in [(32, 2), (32, 0), (8, 2), (13, 1), (1, 2)]:
            got = fuse_gpu(pkg, K, poses, depth, dims, ORIGIN, voxel, trunc, max_batch=chunk, cull_level=cull)
            for a, b, name in zip(got, exact, ("tsdf", "weight", "points")):
                assert_bit_equal(a, b, f"stream {name} chunk={chunk} cull={cull} dims {dims}")


def test_stream_groups_and_input_kinds(pkg, oracle, sequence):
    """300 frames: host input is fused in groups of 128 frames through two alternating slots (H2D overlapped with
    the fusion of the previous group); pageable, pinned, device float and uint16 inputs must agree bit for bit, also
    when the frames arrive in several calls."""
    import torch
    frames = [(11 * i) % 1000 for i in range(300)]
    depth, mm, poses, K = sequence(frames)
    dims, voxel, trunc = (96, 96, 96), 0.03125, 0.15625
    t_ref, w_ref = oracle.fuse(K, poses, depth, dims, ORIGIN, voxel, trunc)
    variants = {
        "pageable": depth,
        "pinned": torch.from_numpy(depth).pin_memory(),
        "device": torch.from_numpy(depth).cuda(),
        "u16 pageable": mm,
        "u16 pinned": torch.from_numpy(mm.view(np.int16)).pin_memory(),
        "u16 device": torch.from_numpy(mm.view(np.int16)).cuda(),
    }
    for name, src in variants.items():
        with pkg.Volume(dims, ORIGIN, voxel, trunc, K) as vol:
            vol.integrate_batch(src, poses)
            t, w = vol.download()
        assert_bit_equal(w, w_ref, f"weights [{name}]")
        assert_bit_equal(t, t_ref, f"tsdf [{name}]")
    with pkg.Volume(dims, ORIGIN, voxel, trunc, K, max_batch=7) as vol:   # several calls, ragged sizes
        for a, b in [(0, 1), (1, 130), (130, 131), (131, 300)]:
            vol.integrate_batch(depth[a:b], poses[a:b])
        t, w = vol.download()
    assert_bit_equal(w, w_ref, "weights [4 calls]")
    assert_bit_equal(t, t_ref, "tsdf [4 calls]")


def test_nan_depth_pixels_follow_the_reference(pkg, oracle, sequence):
    """demo.cu:53-62 does not reject a NaN depth pixel: both comparisons are false and fmin(1.0f, NaN) = 1.0f, so the
    voxel receives a dist = 1 observation.  Every engine path must do the same (float input; uint16 cannot hold NaN)."""
    import torch
    frames = [0, 4, 9, 15, 300]
    depth, _, poses, K = sequence(frames)
    depth = depth.copy()
    rng = np.random.default_rng(11)
    depth[rng.random(depth.shape) < 0.05] = np.nan
    depth[2, 100:140, 200:300] = np.nan
    depth[3, 50:60, 50:60] = np.inf      # +inf IS rejected by the reference (d > 6)
    depth[3, 70:80, 50:60] = -np.inf
    for dims, voxel in [((96, 80, 64), 0.03125), ((50, 62, 38), 0.06)]:   # VEC = 4 stream engine / VEC = 1 kernel
        trunc = float(np.float32(voxel) * np.float32(5))
        t_ref, w_ref = oracle.fuse(K, poses, depth, dims, ORIGIN, voxel, trunc)
        if oracle.have_ref_host():
            t_r, w_r = oracle.ref_host_fuse(K, poses, depth, dims, ORIGIN, voxel, trunc)
            assert_bit_equal(w_ref, w_r.reshape(w_ref.shape), "oracle vs reference host build (weights)")
            assert_bit_equal(t_ref, t_r.reshape(t_ref.shape), "oracle vs reference host build (tsdf)")
        for kw in [dict(), dict(max_batch=2, cull_level=0), dict(tuning=pkg.binding.LEGACY_ENGINE), dict(tuning=256),
                   dict(tuning=(1 << 22)), dict(tuning=(1 << 20))]:
            t, w, _ = fuse_gpu(pkg, K, poses, depth, dims, ORIGIN, voxel, trunc, **kw)
            assert_bit_equal(w, w_ref, f"weights {dims} {kw}")
            assert_bit_equal(t, t_ref, f"tsdf {dims} {kw}")
        # the reference-shaped entry point reads the caller's frame in place (not sanitized)
        n = dims[0] * dims[1] * dims[2]
        t_dev = torch.ones(n, dtype=torch.float32, device="cuda")
        w_dev = torch.zeros(n, dtype=torch.float32, device="cuda")
        K_dev = torch.from_numpy(K).cuda()
        for f in range(len(frames)):
            pkg.Integrate(K_dev, torch.from_numpy(poses[f]).cuda(), torch.from_numpy(depth[f]).cuda(), 480, 640,
                          dims[0], dims[1], dims[2], *ORIGIN, voxel, trunc, t_dev, w_dev,
                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert_bit_equal(w_dev.cpu().numpy().reshape(w_ref.shape), w_ref, f"tsdf_Integrate weights {dims}")
        assert_bit_equal(t_dev.cpu().numpy().reshape(t_ref.shape), t_ref, f"tsdf_Integrate tsdf {dims}")


def test_misaligned_external_slab_is_rejected(pkg, synth):
    import torch
    buf = torch.zeros(64 * 64 * 64 + 4, dtype=torch.float32, device="cuda")
    ok = torch.zeros(64 * 64 * 64, dtype=torch.float32, device="cuda")
    with pytest.raises(pkg.TsdfError, match="16-byte aligned"):
        pkg.Volume((64, 64, 64), ORIGIN, 0.05, 0.25, synth.CAM_K, ext_tsdf=buf[1:], ext_weight=ok)


def test_shipped_library_rejects_reserved0(pkg, synth):
    with pytest.raises(pkg.TsdfError, match="reserved0"):
        pkg.Volume((64, 64, 64), ORIGIN, 0.05, 0.25, synth.CAM_K, tuning=5, validation=False)


def test_integrate_multi_equals_separate_calls(pkg, oracle, sequence):
    """tsdf_integrate_multi (BASELINE config 5: independent fragment volumes, one launch) == one tsdf_integrate_batch
    per volume == the oracle; device, pinned and pageable uint16 inputs; also the fallback for differing volumes."""
    import torch
    n_frag, n_frames = 5, 40
    dims, voxel, trunc = (120, 100, 76), 0.025, 0.125
    data = []
    for i in range(n_frag):
        frames = list(range(50 * i, 50 * i + n_frames))
        depth, mm, poses, K = sequence(frames, base_frame=50 * i)
        data.append((depth, mm, poses, K))
    K = data[0][3]
    refs = [oracle.fuse(K, d[2], d[0], dims, ORIGIN, voxel, trunc) for d in data[:2]]
    base = []
    for depth, mm, poses, _ in data:
        with pkg.Volume(dims, ORIGIN, voxel, trunc, K) as vol:
            vol.integrate_batch(mm, poses)
            t, w = vol.download()
            base.append((t, w, vol.extract_surface()))
    for (t, w, _), (t_ref, w_ref) in zip(base, refs):
        assert_bit_equal(w, w_ref, "separate weights vs oracle")
        assert_bit_equal(t, t_ref, "separate tsdf vs oracle")
    sources = {
        "u16 device": [torch.from_numpy(d[1].view(np.int16)).cuda() for d in data],
        "u16 pinned": [torch.from_numpy(d[1].view(np.int16)).pin_memory() for d in data],
        "u16 pageable": [d[1] for d in data],
        "float device": [torch.from_numpy(d[0]).cuda() for d in data],
    }
    for name, src in sources.items():
        vols = [pkg.Volume(dims, ORIGIN, voxel, trunc, K) for _ in range(n_frag)]
        for rep in range(2):   # twice: the second launch reuses every buffer
            for v in vols:
                v.reset()
            pkg.integrate_multi(vols, src, [d[2] for d in data])
        for v, (t_b, w_b, p_b) in zip(vols, base):
            t, w = v.download()
            assert_bit_equal(w, w_b, f"multi weights [{name}]")
            assert_bit_equal(t, t_b, f"multi tsdf [{name}]")
            assert_bit_equal(v.extract_surface(), p_b, f"multi points [{name}]")
            v.close()
    # volumes of different shape: the call falls back to one launch per volume
    v_a = pkg.Volume(dims, ORIGIN, voxel, trunc, K)
    v_b = pkg.Volume((dims[0], dims[1], dims[2] - 12), ORIGIN, voxel, trunc, K)
    pkg.integrate_multi([v_a, v_b], [data[0][1], data[1][1]], [data[0][2], data[1][2]])
    t, w = v_a.download()
    assert_bit_equal(t, base[0][0], "fallback tsdf")
    t, w = v_b.download()
    assert_bit_equal(t, base[1][0][: dims[2] - 12], "fallback tsdf (shorter volume)")
    v_a.close(); v_b.close()


def test_plan_z_cost_matches_engine_statistics_and_balances(pkg, sequence):
    """tsdf_plan_z_cost (whole grid, no volume memory) == the counts the fusion kernel's mask pre-pass accumulates
    (tsdf_get_z_cost), is deterministic, and its equal-cost cuts are far better balanced than equal heights."""
    import torch
    frames = list(range(0, 128, 2))
    depth, mm, poses, K = sequence(frames)
    dims, voxel = (256, 256, 256), 0.012
    trunc = float(np.float32(voxel) * np.float32(5))
    cost = pkg.plan_z_cost(dims, ORIGIN, voxel, trunc, K, depth, poses)
    cost_dev = pkg.plan_z_cost(dims, ORIGIN, voxel, trunc, K, torch.from_numpy(mm.view(np.int16)).cuda(), poses)
    assert cost.shape == (32,) and cost.sum() > 0 and np.array_equal(cost, cost_dev)
    # (device input: the engine fuses the 64 frames as ONE group, so its cull margins -- derived from the group's poses --
    #  are the planner's; host input is cut into several groups with slightly different conservative margins)
    depth_d = torch.from_numpy(depth).cuda()
    with pkg.Volume(dims, ORIGIN, voxel, trunc, K) as vol:
        vol.integrate_batch(depth_d, poses)
        assert np.array_equal(vol.z_cost(), cost)
    ranges = pkg.sharding.split_groups(cost, dims[2], 4)
    part = []
    for a, b in ranges:
        with pkg.Volume(dims, ORIGIN, voxel, trunc, K, z_range=(a, b)) as vol:
            vol.integrate_batch(depth_d, poses)
            part.append(int(vol.z_cost().sum()))
    assert sum(part) == int(cost.sum())           # slabs cut at multiples of 8 planes see exactly the same tiles
    eq = [int(cost[a // 8:b // 8].sum()) for a, b in [(0, 64), (64, 128), (128, 192), (192, 256)]]
    assert max(part) / (sum(part) / 4) < max(eq) / (sum(eq) / 4)
    sub = pkg.plan_z_cost(dims, ORIGIN, voxel, trunc, K, depth, poses, frame_step=4)
    assert 0.15 * cost.sum() < sub.sum() < 0.4 * cost.sum()


def test_surface_normals_vs_oracle(pkg, oracle, sequence):
    """SURVEY 8f row 4 (parity unpinned: no reference fixture; the algorithm is PCL-benchmark/pclWrapperFPFH.h:308-330
    + mLib's getNormalized): normals of the extracted surface voxels against the oracle restatement, bit for bit."""
    depth, _, poses, K = sequence(list(range(0, 60, 4)))
    dims, voxel, trunc = (128, 120, 96), 0.0234, 0.117
    with pkg.Volume(dims, ORIGIN, voxel, trunc, K) as vol:
        vol.integrate_batch(depth, poses)
        t, w = vol.download()
        pts = vol.extract_surface()
        nrm = vol.surface_normals()
    ref = oracle.surface_normals(t, w)
    assert nrm.shape == ref.shape == pts.shape and len(pts) > 5000
    assert_bit_equal(nrm, ref, "surface normals")
    length = np.linalg.norm(nrm.astype(np.float64), axis=1)
    assert ((np.abs(length - 1) < 1e-6) | (length == 0)).all() and (length > 0).mean() > 0.9


def test_config2_full_size_all_frames(pkg, oracle, sequence):
    """BASELINE config 2 at full size (512^3, 6 mm voxels, 300 frames): per-frame (K = 1) == K = 32 through whole-volume
    checksums, and oracle planes over ALL 300 frames (voxels are independent, so planes are exact samples)."""
    import torch
    frames = list(range(300))
    depth, _, poses, K = sequence(frames)
    dims, voxel, trunc = (512, 512, 512), 0.006, 0.03
    n = dims[0] * dims[1] * dims[2]
    depth_d = torch.from_numpy(depth).cuda()
    sums = []
    for batch in (32, 1):
        t_dev = torch.empty(n, dtype=torch.float32, device="cuda")
        w_dev = torch.empty(n, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        with pkg.Volume(dims, ORIGIN, voxel, trunc, K, max_batch=batch, ext_tsdf=t_dev, ext_weight=w_dev) as vol:
            vol.reset()
            vol.integrate_batch(depth_d, poses)
            vol.synchronize()
            n_pts = vol.extract_surface_count()
        sums.append((int(t_dev.view(torch.int32).to(torch.int64).sum().item()), float(w_dev.double().sum().item()), n_pts))
        if batch == 32:
            tv, wv = t_dev.view(512, 512, 512), w_dev.view(512, 512, 512)
            for z in (0, 7, 8, 100, 255, 256, 300, 400, 450, 480, 500, 511):
                t_ref, w_ref = oracle.fuse(K, poses, depth, dims, ORIGIN, voxel, trunc, z_range=(z, z + 1))
                assert_bit_equal(wv[z].cpu().numpy(), w_ref[0], f"weights plane {z}")
                assert_bit_equal(tv[z].cpu().numpy(), t_ref[0], f"tsdf plane {z}")
            assert float(w_dev.max().item()) > 250
        del t_dev, w_dev
    assert sums[0] == sums[1], f"K=32 vs K=1 checksums differ: {sums}"
    assert sums[0][2] > 100000


def test_stream_every_voxel_in_the_band(pkg, oracle, sequence):
    """A truncation margin far larger than the scene: every voxel with a valid pixel is inside the band at every frame,
    so every voxel of every tile becomes a sticky record in every epoch (the record lists run at full capacity, 8192 per
    block) and the whole fusion goes through phase B."""
    frames = [(13 * i) % 700 for i in range(45)]
    depth, _, poses, K = sequence(frames)
    dims, voxel, trunc = (64, 64, 24), 0.045, 12.0
    t_ref, w_ref = oracle.fuse(K, poses, depth, dims, ORIGIN, voxel, trunc)
    assert (np.abs(t_ref[w_ref > 0]) < 1).mean() > 0.95
    for chunk in (32, 7):
        t, w, _ = fuse_gpu(pkg, K, poses, depth, dims, ORIGIN, voxel, trunc, max_batch=chunk)
        assert_bit_equal(w, w_ref, f"weights chunk={chunk}")
        assert_bit_equal(t, t_ref, f"tsdf chunk={chunk}")


def test_stream_more_frames_than_one_group(pkg, oracle, sequence):
    """2200 frames in one call: device input is fused in groups of 1024 frames (three persistent launches), host input
    in groups of 32, 96, 256, ... ; with K = 1 a group holds 1024 one-frame chunks (the segment table at its limit)."""
    import torch
    base = [(9 * i) % 500 for i in range(50)]
    depth, _, poses, K = sequence(base)
    reps = 44
    depth_all = np.tile(depth, (reps, 1, 1))
    poses_all = np.tile(poses, (reps, 1))
    dims, voxel, trunc = (32, 32, 16), 0.09, 0.45
    t_ref, w_ref = oracle.fuse(K, poses_all, depth_all, dims, ORIGIN, voxel, trunc)
    assert w_ref.max() > 1000
    for src, chunk in ((torch.from_numpy(depth_all).cuda(), 0), (depth_all, 0), (torch.from_numpy(depth_all).cuda(), 1)):
        t, w, _ = fuse_gpu(pkg, K, poses_all, src, dims, ORIGIN, voxel, trunc, max_batch=chunk)
        assert_bit_equal(w, w_ref, f"weights chunk={chunk}")
        assert_bit_equal(t, t_ref, f"tsdf chunk={chunk}")
