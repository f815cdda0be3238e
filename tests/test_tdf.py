"""Truncated distance function: the reference's ComputeTDF kernel (core/demo.cu:22-57).
CPU part: the oracle's brute-force restatement is pinned bit for bit against the reference kernel body compiled
for the host (oracle/_ref/libref_tdf_host.so).  GPU part: the separable-transform kernels behind
tsdf_ComputeTDF against the oracle and against the reference kernel recompiled for sm_100a."""
import numpy as np
import pytest

from conftest import assert_bit_equal


def random_occupancy(dims, density, seed):
    rng = np.random.default_rng(seed)
    dx, dy, dz = dims
    return (rng.random((dz, dy, dx)) < density).astype(np.float32)


def shell_occupancy(dims):
    """A sphere shell + a plane: surface-like occupancy as a fused fragment would give."""
    dx, dy, dz = dims
    z, y, x = np.meshgrid(np.arange(dz), np.arange(dy), np.arange(dx), indexing="ij")
    r = np.sqrt((x - dx / 2) ** 2 + (y - dy / 2) ** 2 + (z - dz / 2) ** 2)
    occ = (np.abs(r - min(dims) / 3) < 0.6) | (y == dy // 5)
    return occ.astype(np.float32)


@pytest.mark.parametrize("dims,density,seed", [((37, 29, 23), 0.01, 1), ((40, 40, 40), 0.0005, 2), ((16, 9, 50), 0.2, 3)])
def test_oracle_tdf_matches_reference_kernel_body(oracle, dims, density, seed):
    if not oracle.REF_TDF_HOST_LIB.exists():
        pytest.skip("oracle/_ref/libref_tdf_host.so not built (needs /root/reference)")
    occ = random_occupancy(dims, density, seed)
    for voxel, trunc in ((0.01, 0.05), (0.02, 0.06), (0.01, 0.01)):
        assert_bit_equal(oracle.compute_tdf(occ, dims, voxel, trunc), oracle.ref_host_compute_tdf(occ, dims, voxel, trunc),
                         f"tdf R={trunc / voxel}")
    occ = shell_occupancy(dims)
    t = oracle.compute_tdf(occ, dims)
    assert_bit_equal(t, oracle.ref_host_compute_tdf(occ, dims), "shell tdf")
    assert t.max() == 1.0 and 0.0 < t[t > 0].min() < 0.3 and (t == 0).any()


def test_oracle_tdf_grid_from_points(oracle, pkg):
    """core/demo.cu:100-136 (not compilable on its own): library and oracle restatements agree; known answers."""
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1.0, 1.5, size=(2000, 3)).astype(np.float32)
    dims_o, origin_o, occ = oracle.tdf_grid_from_points(pts, 0.01, 15)
    dims_p, origin_p = pkg.tdf_grid_from_points(pts, 0.01, 15)
    assert list(dims_o) == list(dims_p)
    assert_bit_equal(origin_o, origin_p, "origin")
    assert occ.sum() <= len(pts) and occ.sum() > 1900
    span = pts.max(0) - pts.min(0)
    assert list(dims_o) == [int(np.round(np.float32(s) / np.float32(0.01)) + 31) for s in span]


@pytest.mark.gpu
@pytest.mark.parametrize("dims,kind", [((37, 29, 23), "random"), ((64, 64, 64), "shell"), ((101, 7, 55), "random"),
                                       ((20, 20, 20), "empty"), ((20, 20, 20), "full"), ((33, 1, 1), "random")])
def test_ComputeTDF_bit_exact_vs_oracle(pkg, oracle, dims, kind):
    import torch
    occ = {"random": lambda: random_occupancy(dims, 0.01, 7), "shell": lambda: shell_occupancy(dims),
           "empty": lambda: np.zeros(dims[::-1], np.float32), "full": lambda: np.ones(dims[::-1], np.float32)}[kind]()
    for voxel, trunc in ((0.01, 0.05), (0.01, 0.03), (0.005, 0.05)):
        ref = oracle.compute_tdf(occ, dims, voxel, trunc)
        occ_d = torch.from_numpy(occ).cuda()
        tdf_d = torch.zeros_like(occ_d)
        torch.cuda.synchronize()
        pkg.ComputeTDF(occ_d, tdf_d, *dims, 0.0, 0.0, 0.0, voxel, trunc)
        torch.cuda.synchronize()
        assert_bit_equal(tdf_d.cpu().numpy(), ref, f"tdf {kind} R={trunc / voxel:.0f}")
    # read-modify-write like the kernel: a pre-filled TDF is only ever raised (core/demo.cu:54-55)
    pre = np.full(dims[::-1], 0.5, np.float32)
    tdf_d = torch.from_numpy(pre).cuda()
    pkg.ComputeTDF(torch.from_numpy(occ).cuda(), tdf_d, *dims, 0.0, 0.0, 0.0, 0.01, 0.05)
    torch.cuda.synchronize()
    assert_bit_equal(tdf_d.cpu().numpy(), oracle.compute_tdf(occ, dims, 0.01, 0.05, tdf=pre), "tdf with prefilled grid")


@pytest.mark.gpu
def test_ComputeTDF_vs_reference_cuda_kernel_beyond_2pow24(pkg, oracle):
    """A fragment-sized grid (330 x 300 x 260 = 25.7 M voxels > 2^24): the reference decomposes the voxel index in
    float (core/demo.cu:33-35) and mis-addresses neighbourhoods there; the replacement reproduces that, so it
    stays bit-identical to the reference kernel recompiled for sm_100a.  Also prints both kernel times."""
    import torch
    if not oracle.REF_TDF_CUDA_LIB.exists():
        pytest.skip("oracle/_ref/libref_tdf_cuda.so not built")
    dims = (330, 300, 260)
    occ = shell_occupancy(dims)
    ref, ref_ms = oracle.ref_cuda_compute_tdf(occ, dims)
    occ_d = torch.from_numpy(occ).cuda()
    tdf_d = torch.zeros_like(occ_d)
    pkg.ComputeTDF(occ_d, tdf_d, *dims, 0.0, 0.0, 0.0, 0.01, 0.05)   # warm-up
    tdf_d.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()
    e0.record(stream)
    pkg.ComputeTDF(occ_d, tdf_d, *dims, 0.0, 0.0, 0.0, 0.01, 0.05, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1)
    assert_bit_equal(tdf_d.cpu().numpy(), ref, "tdf vs reference CUDA kernel")
    n = dims[0] * dims[1] * dims[2]
    print(f"ComputeTDF {dims}: reference kernel {ref_ms:.2f} ms, tsdf_ComputeTDF {ours_ms:.3f} ms "
          f"({ref_ms / ours_ms:.0f}x), {n * 8 / ours_ms / 1e6:.0f} GB/s of occ+tdf traffic")


@pytest.mark.gpu
def test_fused_surface_to_tdf_pipeline(pkg, oracle, sequence):
    """fusion -> surface points -> TDF grid (core/demo.cu:100-163) entirely through the C ABI, vs the oracle."""
    depth, _, poses, K = sequence([0, 7, 13, 21])
    dims, voxel, trunc = (96, 80, 64), 0.03125, 0.15625
    with pkg.Volume(dims, (-1.5, -1.5, 0.5), voxel, trunc, K) as vol:
        vol.integrate_batch(depth, poses)
        pts = vol.extract_surface()
    assert len(pts) > 500
    g_dims, g_origin = pkg.tdf_grid_from_points(pts, 0.03125, 15)
    tdf = pkg.point_cloud_to_tdf(pts, g_dims, g_origin, 0.03125, 0.15625)
    o_dims, o_origin, occ = oracle.tdf_grid_from_points(pts, 0.03125, 15)
    assert list(o_dims) == list(g_dims)
    assert_bit_equal(g_origin, o_origin, "grid origin")
    assert_bit_equal(tdf, oracle.compute_tdf(occ, o_dims, 0.03125, 0.15625), "tdf of the fused surface")
    assert (tdf == 1.0).sum() == int(occ.sum())
