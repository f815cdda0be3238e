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
