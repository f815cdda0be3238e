"""Multi-GPU parity on hardware (skipped on boxes with fewer than 2 GPUs): bench.py's own parity block -- every
rank's slab against a monolithic fusion of the same frames, the NCCL-gathered cloud against the monolithic cloud,
oracle planes at the slab cuts -- on a reduced workload, launched exactly like the driver launches the bench."""
import json
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run_bench(n, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "bench.py"), "--gpus", str(n), "--steps", "1", "--warmup", "1",
           "--no-cpu-baseline"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("config,frames", [("512", 96), ("1024", 40)])
def test_sharded_bench_is_bitwise_equal_to_monolithic(config, frames):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    out = _run_bench(n, ["--config", config, "--frames", str(frames), "--rebalance", "2", "--oracle-planes-max", "6"])
    p = out["parity"]
    assert out["n_gpus"] == n and len(out["config"]["slabs"]) == n
    assert p["slabs_equal_monolithic"] is True and p["gathered_cloud_equal_monolithic"] is True
    assert p["oracle_planes_checked"] >= 3 and p["oracle_planes_mismatching"] == 0 and p["all_equal"] is True
    assert p["surface_points"] == p["monolithic_surface_points"] == out["e2e"]["surface_points"] > 1000


def test_fragment_batch_over_ranks():
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    out = _run_bench(2, ["--config", "frag300", "--fragments", "12", "--frag-batch", "4"])
    p = out["parity"]
    assert p["multi_equals_sequential"] is True and p["oracle_fragments_mismatching"] == 0 and p["fragments_hashed"] == 12
    assert out["e2e"]["surface_points"] > 1000


def test_sharded_c_abi_equals_monolithic(tmp_path):
    """tsdf_sharded_* (one process, one handle per slab): ncclBroadcast of the frames, independent fusion, ordered
    ncclSend/Recv gather of the points, slab-ordered tsdf.bin -- byte-identical to the single-volume path.  With one
    GPU the slabs share the device (plain device copies instead of NCCL), with more they are spread over the GPUs."""
    import numpy as np
    import __graft_entry__ as graft
    from conftest import make_sequence
    pkg, oracle = graft.load_package(), graft.load_oracle()
    frames = [(5 * i) % 600 for i in range(150)]          # 150 frames: two broadcast groups (128 + 22)
    depth, mm, poses, K = make_sequence(pkg, oracle, frames)
    dims, origin, voxel, trunc = (160, 128, 144), (-1.5, -1.5, 0.5), 0.02, 0.1
    with pkg.Volume(dims, origin, voxel, trunc, K) as vol:
        vol.integrate_batch(mm, poses)
        t, _ = vol.download()
        pts = vol.extract_surface()
    pkg.write_bin(tmp_path / "mono.bin", dims, origin, voxel, trunc, t)
    n_gpu = _gpus()
    cost = pkg.plan_z_cost(dims, origin, voxel, trunc, K, mm, poses, frame_step=2)
    for n_slabs, devices in [(3, [0, 0, 0]), (2, None), (4, [i % n_gpu for i in range(4)]), (1, [0])]:
        if devices is None and n_gpu < 2:
            continue
        cuts = [a for a, _ in pkg.split_slabs(cost, dims[2], n_slabs)] + [dims[2]]
        for src in (mm, depth):
            with pkg.ShardedVolume(dims, origin, voxel, trunc, K, n_slabs, devices=devices, z_cuts=cuts) as sv:
                assert [s[:2] for s in sv.slabs()] == list(zip(cuts[:-1], cuts[1:]))
                sv.integrate_batch(src[:70], poses[:70])
                sv.integrate_batch(src[70:], poses[70:])
                got = sv.extract_surface()
                sv.write_bin(tmp_path / "sharded.bin")
            assert got.shape == pts.shape and np.array_equal(got.view(np.uint32), pts.view(np.uint32)), (n_slabs, devices)
            assert (tmp_path / "sharded.bin").read_bytes() == (tmp_path / "mono.bin").read_bytes()


def test_cli_gpus_flag_is_byte_identical(tmp_path):
    """bin/demo --gpus N writes the same tsdf.ply / tsdf.bin as --gpus 1 (equal-work slabs, NCCL broadcast, gather)."""
    import __graft_entry__ as graft
    from test_cli import DEMO, write_sequence
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    pkg = graft.load_package()
    frames = list(range(0, 80, 2))
    data, _, _ = write_sequence(tmp_path, pkg.synth, frames, first_idx=0)
    outs = {}
    for g in (1, 2 if n < 4 else 4):
        d = tmp_path / f"run{g}"
        d.mkdir()
        argv = [str(DEMO), str(tmp_path / "camera-intrinsics.txt"), str(data), "0", "0", str(len(frames)), "-1.5", "-1.5", "0.5",
                "0.0117", "0.0585", "--dim", "256", "--bin", "--gpus", str(g)]
        run = subprocess.run(argv, cwd=d, capture_output=True, text=True, timeout=600)
        assert run.returncode == 0, run.stderr
        # (NCCL prints its version banner to stdout when NCCL_DEBUG is set in the environment)
        stdout = "\n".join(l for l in run.stdout.splitlines() if not l.startswith("NCCL version"))
        outs[g] = ((d / "tsdf.ply").read_bytes(), (d / "tsdf.bin").read_bytes(), stdout)
    (p1, b1, s1), (pn, bn, sn) = outs.values()
    assert len(p1) > 10000 and p1 == pn and b1 == bn and s1 == sn
