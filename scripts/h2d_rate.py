"""Diagnostic: pinned host -> device copy rate on this box, alone and while the fusion kernel is running."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
synth = pkg.synth
cfg = dict(synth.CONFIGS["1024"], frames=512)
grid = cfg["grid"]
depth, poses = bench.build_inputs(pkg, cfg, torch.device("cuda", 0), 0, 1, None)
host = torch.empty((256, 480, 640), dtype=torch.float32, pin_memory=True)
host.copy_(depth[:256])
dst = torch.empty_like(depth[:256])
side = torch.cuda.Stream()
nbytes = host.numel() * 4


def copy_ms():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record()
        dst.copy_(host, non_blocking=True)
        e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)


copy_ms()
alone = min(copy_ms() for _ in range(3))
vol = pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, max_batch=32)
vol.integrate_batch(depth, poses)
vol.synchronize()
vol.reset(); vol.synchronize()
vol.timer_start(); vol.integrate_batch(depth, poses); quiet = vol.timer_stop()
vol.reset(); vol.synchronize()
vol.timer_start(); vol.integrate_batch(depth, poses)
time.sleep(0.02)
busy = copy_ms()
loaded = vol.timer_stop()
print(f"pinned H2D of {nbytes / 1e6:.0f} MB: alone {alone:.2f} ms = {nbytes / alone / 1e6:.1f} GB/s; while fusing {busy:.2f} ms = "
      f"{nbytes / busy / 1e6:.1f} GB/s; fusion of 512 frames {quiet:.2f} ms alone, {loaded:.2f} ms with the copy in flight", flush=True)
