"""Diagnostic: CUPTI timeline (torch.profiler) of one host-input step of the 1024^3 x 1000-frame fusion: when the H2D
copies and the engine's kernels start and end."""
import json
import sys
import tempfile
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
cfg = dict(synth.CONFIGS["1024"])
grid = cfg["grid"]
depth, poses = bench.build_inputs(pkg, cfg, torch.device("cuda", 0), 0, 1, None)
host = torch.empty(depth.shape, dtype=torch.float32, pin_memory=True)
host.copy_(depth)
vol = pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, max_batch=32)
for _ in range(2):
    vol.reset(); vol.integrate_batch(host, poses); vol.synchronize()
vol.reset(); vol.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    vol.integrate_batch(host, poses)
    vol.synchronize()
path = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    big = e["cat"] == "kernel" and e["dur"] > 200 or e["cat"] == "gpu_memcpy" and e["dur"] > 200
    if big:
        print(f"{(e['ts'] - t0) / 1e3:9.2f} ms  +{e['dur'] / 1e3:8.2f} ms  stream {e['args'].get('stream', '?'):>3}  {e['name'][:60]}")
print(f"total span {(max(e['ts'] + e['dur'] for e in ev) - t0) / 1e3:.2f} ms")
