#!/usr/bin/env python
"""bench.py -- TSDF voxel-updates/s of the fusion hot path, with its rooflines and baselines.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--config 1024|512|demo500|2048|frag300]

Single-volume configs (BASELINE.json configs 1-4): a "step" fuses the configuration's whole synthetic sequence
(F frames of 640x480 depth) into the volume: re-initialisation on the device (demo.cu:131-145) plus the fusion of
all F frames (demo.cu:156-183; one persistent work-list kernel per group of <= 1024 frames).  N > 1 (launched by
torch.distributed.run, one rank per GPU): the volume is cut into contiguous z-slabs of equal work, rank 0 renders
the frames and broadcasts them over NCCL, every rank fuses all frames into its slab; no collective on the data path
(strong scaling).  frag300 (config 5): 256 independent 300^3 fragment volumes x 50 frames, distributed over the
ranks, several fragments per launch (tsdf_integrate_multi) -- the call sequence of bin/fuse_fragments.

`value`        voxel-updates/s (voxels x frames / time) with the frames resident in HBM, CUDA-event timed on the
               engine's stream, max over ranks.
`e2e`          the same metric through the public C-ABI calls with HOST (pinned) depth buffers: H2D of every frame
               (N > 1: one upload on rank 0 + NCCL broadcast), fusion, surface extraction, D2H of the points.
`roofline`     algorithmic bytes (16 B per voxel per pass over the volume) of the fusion kernel / its
               measured duration against the measured HBM copy peak; `roofline_dram` the same with the DRAM bytes
               ncu measured; `roofline_issue` warp-instructions (committed ncu count) / duration against the issue
               slots of the GPU -- the ceiling that actually binds this kernel.
`parity`       checksums computed in this very run: every rank's slab against a monolithic fusion of the same
               frames, the NCCL-gathered cloud against the monolithic cloud, oracle planes at every slab cut.
`cpu_baseline` the reference's Integrate body compiled for the host (oracle/_ref) -- or the oracle port -- on a
               bounded sample of the same workload, all host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

METRIC = "tsdf_voxel_updates_per_sec"
UNIT = "voxel-updates/s"
BYTES_PER_VOXEL_PASS = 16  # TSDF + weight, read + write
N_SMS, SCHEDULERS_PER_SM = 148, 4


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_counters(key):
    """Per-frame counters of the fusion kernel from the committed ncu capture (profiles/counters.json)."""
    p = ROOT / "profiles" / "counters.json"
    try:
        return json.loads(p.read_text()).get(key)
    except Exception:
        return None


def use_all_host_threads():
    """The reference arms run OpenMP code: torch.distributed.run exports OMP_NUM_THREADS=1 to its children, so the
    thread count is set explicitly (environment for libraries not loaded yet, omp_set_num_threads for libgomp)."""
    n = os.cpu_count() or 1
    try:   # NCCL / the launcher may have narrowed this process to the cores next to its GPU: OpenMP workers inherit the mask
        os.sched_setaffinity(0, range(n))
    except (AttributeError, OSError):
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started before the warm-up steps (nvidia-smi takes
    a few hundred ms to produce its first line, longer than the timed region of a multi-GPU run); mark() is called when
    the timed region starts and the summary uses the samples taken after it -- or, if the region was too short to
    catch any, the warm-up samples of the same workload, saying so in "window"."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.mark_at = 0

    def mark(self):
        self.mark_at = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = self.lines[self.mark_at:]
        window = "timed region" if len(timed) >= 2 else "warm-up + timed region"
        for line in (timed if len(timed) >= 2 else self.lines):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


class Dist:
    """torch.distributed plumbing of one bench process (NCCL, one rank per GPU)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def gather_floats(self, x):
        if self.world == 1:
            return [float(x)]
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def gather_int64(self, vec):
        """Variable-length int64 vectors of every rank -> list of numpy arrays (on every rank)."""
        torch = self.torch
        v = torch.as_tensor(np.asarray(vec, dtype=np.int64), device=self.device)
        if self.world == 1:
            return [v.cpu().numpy()]
        n = torch.tensor([v.numel()], dtype=torch.int64, device=self.device)
        ns = [torch.zeros_like(n) for _ in range(self.world)]
        self.dist.all_gather(ns, n)
        ns = [int(x.item()) for x in ns]
        pad = torch.zeros(max(max(ns), 1), dtype=torch.int64, device=self.device)
        pad[: v.numel()] = v
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(bufs, pad)
        return [b[:k].cpu().numpy() for b, k in zip(bufs, ns)]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class _DevArray:
    """torch view of engine-owned device memory (CUDA array interface); the engine was synchronized before."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def plane_hashes(torch, tsdf_ptr, weight_ptr, nz, plane, device, seed=20161128):
    """Position-sensitive 64-bit hash of every z-plane of a slab's TSDF and weight arrays (bit patterns), computed on
    the device: sum over the plane of bits * w[i] with a fixed odd multiplier vector w (wrap-around arithmetic)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = (torch.randint(0, 2 ** 31, (plane,), generator=g, dtype=torch.int64) * 2 + 1).to(device)
    t = torch.as_tensor(_DevArray(tsdf_ptr, (nz, plane), "<i4"), device=device)
    wt = torch.as_tensor(_DevArray(weight_ptr, (nz, plane), "<i4"), device=device)
    out = torch.empty(nz, dtype=torch.int64, device=device)
    for z in range(nz):
        out[z] = (t[z].to(torch.int64) * w).sum() + 31 * (wt[z].to(torch.int64) * w).sum()
    return out.cpu().numpy()


def cloud_hash(pts: np.ndarray) -> int:
    """Order-sensitive hash of a surface point cloud (float bit patterns)."""
    b = np.ascontiguousarray(pts, dtype=np.float32).view(np.uint32).astype(np.uint64).ravel()
    if b.size == 0:
        return 0
    mult = (np.arange(b.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return int((b * mult).sum(dtype=np.uint64))


def build_inputs(pkg, cfg, device, rank, world, dist, want_mm=False):
    """Frames rendered on rank 0's GPU and broadcast over NCCL; poses via the reference's host algebra."""
    import torch
    synth = pkg.synth
    F = cfg["frames"]
    frames = list(range(F))
    mm = torch.empty((F, 480, 640), dtype=torch.int32, device=device)
    if rank == 0:
        for f0 in range(0, F, 50):
            mm[f0:f0 + 50] = synth.render_depth_mm(frames[f0:f0 + 50], device=device)
    pkg.sharding.broadcast_frames(mm, src=0)  # NCCL over NVLink when world > 1
    # utils.hpp:82-84 on the device: (float)mm / 1000.0f (IEEE division), > 6 -> 0
    depth = torch.div(mm.to(torch.float32), torch.full((1,), 1000.0, dtype=torch.float32, device=device))
    depth = torch.where(depth > 6.0, torch.zeros_like(depth), depth).contiguous()
    check = pkg.depth_from_u16(mm[0].cpu().numpy().astype(np.uint16))
    if not np.array_equal(check.view(np.uint32), depth[0].cpu().numpy().view(np.uint32)):
        raise RuntimeError("device depth conversion is not bit-identical to ReadDepth's")
    c2w, b2w = synth.make_poses(frames)
    poses = np.stack([pkg.cam2base_from_world(b2w, m).ravel() for m in c2w]).astype(np.float32)
    del mm
    torch.cuda.synchronize()  # the engine runs on its own stream
    return depth, poses


def time_step(vol, fn, D):
    D.barrier()
    vol.timer_start()
    fn()
    return vol.timer_stop()


# =====================================================================================================================
# single-volume configurations (BASELINE configs 1-4)
# =====================================================================================================================

def run_volume(args, D):
    torch = D.torch
    rank, world, device = D.rank, D.world, D.device
    pkg = graft.load_package()
    synth, sharding = pkg.synth, pkg.sharding
    cfg = dict(synth.CONFIGS[args.config])
    if args.frames:
        cfg["frames"] = args.frames
    grid, F = cfg["grid"], cfg["frames"]
    dim_z = grid.dims[2]
    # every rank watches its own GPU; started first (nvidia-smi needs seconds to produce its first line on an 8-GPU box,
    # longer than the timed region of a multi-GPU run)
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    depth, poses = build_inputs(pkg, cfg, device, rank, world, D.dist)

    def make_volume(z_range, batch=None):
        return pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, z_range=z_range,
                          device=D.local_rank, max_batch=batch or args.batch, cull_level=args.cull, tuning=args.tuning)

    # ---- z-slabs of equal WORK: the engine's own cull statistics (kept (tile, frame) pairs per group of 8 planes,
    #      tsdf_plan_z_cost: deterministic integer counts, so every rank derives the same cuts without communication),
    #      then --rebalance warm-up steps feed the measured rank times back (damped, best measured partition kept)
    ranges = [sharding.slab_range(dim_z, r, world) for r in range(world)]
    group_cost, plan_ms = None, None
    if world > 1 and not args.equal_slabs:
        t0 = time.perf_counter()
        group_cost = pkg.plan_z_cost(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, depth, poses,
                                     device=D.local_rank, frame_step=max(1, F // 128), cull_level=args.cull).astype(np.float64)
        plan_ms = (time.perf_counter() - t0) * 1e3
        plane_cost = sharding.plane_cost_from_groups(group_cost, dim_z, args.base_cost)
        ranges = sharding.split_planes(plane_cost, world)

    def make_warm_volume(z_range):
        """A volume whose lazy allocations (frame store, masks, item lists) have already happened, so that the first
        timed step on it measures the fusion and not cudaMalloc."""
        v = make_volume(z_range)
        v.integrate_batch(depth, poses)
        v.synchronize()
        return v

    vol = make_warm_volume(ranges[rank])
    rebalance_log = []
    if group_cost is not None:
        def switch_to(new_ranges):
            nonlocal vol, ranges
            if new_ranges[rank] != ranges[rank]:
                vol.close()
                vol = make_warm_volume(new_ranges[rank])
            ranges = new_ranges

        best = None   # (slowest rank's ms, partition) over the measured partitions
        for _ in range(args.rebalance):
            vol.reset()
            vol.synchronize()
            rank_t = D.gather_floats(time_step(vol, lambda: vol.integrate_batch(depth, poses), D))
            rebalance_log.append([round(x, 1) for x in rank_t])
            if best is None or max(rank_t) < best[0]:
                best = (max(rank_t), ranges)
            if max(rank_t) / (sum(rank_t) / world) < 1.015:
                break
            # (slightly over-relaxed at first: the plan knows the frusta and the occlusion but not the cost of the
            #  surface band, which sits in one thin layer of the last slab)
            plane_cost = sharding.refine_plane_cost_piecewise(plane_cost, ranges, rank_t, damping=1.4 if len(rebalance_log) <= 3 else 1.0)
            new_ranges = sharding.split_planes(plane_cost, world)
            if new_ranges == ranges:
                break
            switch_to(new_ranges)
        if best is not None and best[1] != ranges:
            switch_to(best[1])   # the last proposal was not measured: keep the best partition that was
    z0, z1 = ranges[rank]
    slab_voxels = vol.n_voxels
    total_voxels = grid.n_voxels

    def step_device():
        vol.reset()
        vol.integrate_batch(depth, poses)

    # ---- value: inputs resident in HBM --------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    D.barrier()
    vol.reset_stats()
    D.barrier()
    sampler.mark()
    vol.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = vol.timer_stop()
    D.barrier()
    rank_ms = [round(x / args.steps, 2) for x in D.gather_floats(ms)]
    clocks = sampler.stop()
    if world > 1:
        clocks["rank_sm_mhz"] = D.gather_floats(clocks["sm_mhz"] or 0.0)
        clocks["ranks_with_slowdown_reasons"] = int(D.sum(float(any(r != "sw_power_cap" for r in clocks["reasons"]))))
    launches = D.sum(vol.stats()["kernel_launches"])
    ms_per_step = D.max(ms) / args.steps
    value = total_voxels * F / (ms_per_step * 1e-3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "ms_per_step": round(ms_per_step, 3),
                              "config_key": args.config, "frames": F, "batch": args.batch, "cull": args.cull,
                              "tuning": args.tuning, "slabs": [list(map(int, r)) for r in ranges], "rank_ms": rank_ms,
                              "sum_rank_ms": round(sum(rank_ms), 1), "rebalance": rebalance_log, "plan_ms": plan_ms,
                              "clocks": clocks, "quick": True}), flush=True)
        vol.close()
        return

    # ---- rooflines of the dominant kernel (stream_integrate_kernel): fusion only, no reset ----------
    vol.reset()
    vol.synchronize()
    ms_int = D.max(time_step(vol, lambda: vol.integrate_batch(depth, poses), D))
    K = args.batch
    n_pass = -(-F // K)                    # passes over the volume = chunks
    n_launch = -(-F // 1024)               # persistent launches (groups of <= 1024 frames)
    peak, peak_src = measured_peak_gbs()
    # N > 1: the whole job's bytes over the slowest rank's time, per GPU (a per-rank figure would credit the rank that
    # holds the cheap near slab with bytes it never had to move)
    alg_bytes = BYTES_PER_VOXEL_PASS * total_voxels * n_pass
    achieved = alg_bytes / (ms_int * 1e-3) / 1e9 / world
    ctr = kernel_counters(f"{args.config}_k{K}") or {}
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": (ctr["dram_bytes_per_frame"] * F / n_launch) if "dram_bytes_per_frame" in ctr and world == 1 else None,
                "peak_source": peak_src, "kernel": "stream_integrate_kernel (persistent, one launch per <= 1024 frames)",
                "frames_per_pass": K, "passes_per_launch": n_pass / n_launch, "launches": n_launch,
                "algorithmic_bytes_per_launch": alg_bytes / n_launch, "avg_launch_ms": round(ms_int / n_launch, 3),
                "updates_per_sec_roofline": peak * 1e9 / BYTES_PER_VOXEL_PASS * K * world, "per_gpu": world > 1,
                "binding": "instruction issue, not HBM: the kernel never loads voxels no frame observes and amortises "
                           "the volume pass over K frames, so the algorithmic-bytes ceiling (peak x K / 16 updates/s) "
                           "rises with K while the per-voxel classification cost does not; see roofline_issue"}
    roofline_dram = roofline_issue = None
    if world == 1 and "dram_bytes_per_frame" in ctr:
        dram = ctr["dram_bytes_per_frame"] * F / (ms_int * 1e-3) / 1e9
        roofline_dram = {"bound": "hbm", "achieved": round(dram, 1), "peak": peak, "unit": "GB/s", "frac": round(dram / peak, 4),
                         "source": ctr.get("source"), "note": "DRAM bytes ncu measured for this kernel (per frame) x frames / measured duration"}
    if world == 1 and "warp_instructions_per_frame" in ctr:
        issued = ctr["warp_instructions_per_frame"] * F / (ms_int * 1e-3)
        slots = N_SMS * SCHEDULERS_PER_SM * sm_hz
        roofline_issue = {"bound": "issue", "achieved": issued, "peak": slots, "unit": "warp-instructions/s",
                          "frac": round(issued / slots, 4), "thread_instructions_per_update":
                          round(ctr["warp_instructions_per_frame"] * ctr.get("threads_per_instruction", 32.0) / total_voxels, 2),
                          "source": ctr.get("source"),
                          "note": "warp-instructions per frame from the committed ncu capture of this kernel on this workload x frames "
                                  "/ duration measured here, against 148 SMs x 4 schedulers x the SM clock sampled in this run"}

    # per-frame (K = 1) and K = 8 modes on a frame subset: BASELINE config 2's "per-frame vs K-frame" comparison and the
    # north-star's K-amortised HBM roofline line
    by_k = {}
    if (world == 1 or args.config == "1024") and slab_voxels * 8 < 40e9:
        for k, nfr in ((1, min(F, 64)), (8, min(F, 256))):
            v1 = make_volume((z0, z1), batch=k)
            v1.integrate_batch(depth, poses, n_frames=nfr)   # the same call once untimed: the lazy allocations (frame
            v1.reset()                                       # store, masks, item lists for nfr frames) happen here
            v1.synchronize()
            ms_k = D.max(time_step(v1, lambda: v1.integrate_batch(depth, poses, n_frames=nfr), D))
            v1.close()
            ach = BYTES_PER_VOXEL_PASS * total_voxels * (-(-nfr // k)) / (ms_k * 1e-3) / 1e9 / world
            by_k[str(k)] = {"frames": nfr, "ms": round(ms_k, 3), "achieved": round(ach, 1), "frac": round(ach / peak, 4),
                            "updates_per_sec": total_voxels * nfr / (ms_k * 1e-3)}
    by_k[str(K)] = {"frames": F, "ms": round(ms_int, 3), "achieved": round(achieved, 1), "frac": round(achieved / peak, 4),
                    "updates_per_sec": total_voxels * F / (ms_int * 1e-3)}

    # the north-star's line: fraction of the K-amortised HBM roofline (16 B per voxel per pass) at K = 8, the batch size it
    # was written for, next to the production K
    target = None
    if "8" in by_k:
        target = {"hbm_roofline_target_frac": 0.70, "k8_mode_frac": by_k["8"]["frac"], "k8_mode_updates_per_sec": by_k["8"]["updates_per_sec"],
                  f"k{K}_frac": by_k[str(K)]["frac"], f"k{K}_updates_per_sec": by_k[str(K)]["updates_per_sec"],
                  "note": "the ceiling peak x K / 16 rises with K, the cost of classifying a voxel does not: K = 8 sits at the HBM "
                          "ceiling (culled voxels are never loaded, so the algorithmic-bytes figure can exceed 1), the production "
                          "K = 32 is faster in absolute terms and bound by instruction issue (roofline_issue)"}

    # ---- e2e: host buffers through the C ABI, H2D + fusion + extraction + D2H of the points ---------
    e2e, gathered_cloud = run_e2e(args, D, pkg, vol, depth, poses, ranges, total_voxels, F)

    # ---- parity: checksums against a monolithic fusion and oracle planes at the slab cuts -----------
    parity = None
    if not args.no_parity:
        parity = run_parity(args, D, pkg, cfg, vol, depth, poses, ranges, gathered_cloud)

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(pkg, cfg, depth[:64].cpu().numpy(), poses[:64], target_s=args.cpu_seconds)
    vol.close()
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_cuda = ref_cuda_baseline(pkg, cfg, depth[:16].cpu().numpy(), poses[:16])
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{grid.dims[0]}^3 grid ({grid.voxel_size * 1e3:.1f} mm voxels), {F} frames 640x480, "
                                   f"K={K} frames per pass, z-slab sharded over {world} GPU(s)",
                       "config_key": args.config, "frames": F, "batch": K, "cull_level": args.cull,
                       "l2": "volume (>= 1 GB per GPU) is larger than L2; no flush needed",
                       "parallelism": f"z-slab x{world}" + ("" if world == 1 else (" equal-height" if args.equal_slabs else " equal-work (cull statistics + feedback)")),
                       "slabs": [list(map(int, r)) for r in ranges], "rank_ms_per_step": rank_ms,
                       "sum_rank_ms": round(sum(rank_ms), 1), "rebalance_warmup_ms": rebalance_log, "plan_ms": plan_ms},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_dram": roofline_dram,
            "roofline_issue": roofline_issue, "roofline_by_k": by_k, "north_star_target": target, "parity": parity,
            "cpu_baseline": cpu, "ref_cuda_kernel_baseline": ref_cuda, "clocks": clocks,
        }
        print(json.dumps(out), flush=True)


def run_e2e(args, D, pkg, vol, depth, poses, ranges, total_voxels, F):
    """tsdf_reset + fusion from HOST frames + tsdf_extract_surface + the output side, timed with the host clock.
    N = 1: tsdf_integrate_batch(pinned host depth), tsdf_copy_surface.  N > 1: rank 0 uploads each group of 128 frames
    once and broadcasts it over NCCL (the upload / broadcast of group g + 1 overlaps the fusion of group g), every rank
    fuses from its device copy; ordered NCCL gather of the per-slab points to rank 0 and D2H of the cloud there."""
    torch = D.torch
    rank, world, device = D.rank, D.world, D.device
    depth_host = None
    if rank == 0 or world == 1:
        depth_host = torch.empty(depth.shape, dtype=torch.float32, pin_memory=True)
        depth_host.copy_(depth)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))
    group = 128
    # N > 1: the broadcast lands in a full-length device buffer (no buffer is reused inside a step, so the host never has
    # to drain the engine's stream between groups)
    stage = torch.empty_like(depth) if world > 1 else None
    state = {"pts": None, "n": 0, "h2d": 0}

    def step():
        vol.reset()
        if world == 1:
            vol.integrate_batch(depth_host, poses)
        else:
            for gi, g0 in enumerate(range(0, F, group)):
                n = min(group, F - g0)
                buf = stage[g0:g0 + n]
                if rank == 0:
                    buf.copy_(depth_host[g0:g0 + n], non_blocking=True)
                    state["h2d"] += n * depth.shape[1] * depth.shape[2] * 4
                D.dist.broadcast(buf, src=0)
                torch.cuda.current_stream().synchronize()   # the engine runs on its own stream
                vol.integrate_batch(buf, poses[g0:g0 + n], n_frames=n)
        n_pts = vol.extract_surface_count()
        if world == 1:
            if state["pts"] is None or state["pts"].shape[0] < n_pts:
                state["pts"] = np.empty((int(n_pts * 1.2) + 16, 3), np.float32)
            pkg.binding._check(vol._lib.tsdf_copy_surface(vol._h, state["pts"].ctypes.data, state["pts"].shape[0]), "copy_surface")
            state["n"] = n_pts
            return
        ptr, n_dev = vol.surface_device_ptr()   # (waits for the extraction kernels)
        local = torch.as_tensor(_DevArray(ptr, (n_dev, 3)), device=device) if n_dev else torch.zeros((0, 3), dtype=torch.float32, device=device)
        cloud = pkg.sharding.gather_surface(local, dst=0)
        if rank == 0:
            state["pts"] = cloud.cpu().numpy()
            state["n"] = state["pts"].shape[0]

    step()
    D.barrier()
    vol.reset_stats()
    state["h2d"] = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step()
    vol.synchronize()
    t_e2e = D.max(time.perf_counter() - t0)
    D.barrier()
    st = vol.stats()
    h2d = st["h2d_bytes"] / e2e_steps + poses.nbytes if world == 1 else D.sum(state["h2d"]) / e2e_steps + poses.nbytes
    d2h = st["d2h_bytes"] / e2e_steps + (int(state["n"]) * 12 if world > 1 and rank == 0 else 0)
    d2h = D.max(d2h)
    n_gathered = int(D.max(state["n"] if rank == 0 else 0))
    e2e = {"value": total_voxels * F / (t_e2e / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": round(t_e2e / e2e_steps * 1e3, 2),
           "surface_points": n_gathered,
           "path": "tsdf_reset + " + ("tsdf_integrate_batch(pinned host depth) + tsdf_extract_surface + tsdf_copy_surface" if world == 1 else
                                      "per 128-frame group: H2D on rank 0, NCCL broadcast, tsdf_integrate_batch(device) on every rank; "
                                      "tsdf_extract_surface + ordered NCCL gather of the slab points to rank 0 + D2H")}
    cloud = state["pts"][: state["n"]] if (rank == 0 and state["pts"] is not None) else None
    return e2e, cloud


def run_parity(args, D, pkg, cfg, vol, depth, poses, ranges, gathered_cloud):
    """Bitwise checks inside the bench run (never timed).  `vol` holds the result of a full fusion of all frames."""
    torch = D.torch
    rank, world, device = D.rank, D.world, D.device
    synth = pkg.synth
    grid, F = cfg["grid"], cfg["frames"]
    dx, dy, dz = grid.dims
    plane = dx * dy
    out = {}
    vol.synchronize()
    tp, wp = vol.device_ptrs()
    z0, z1 = ranges[rank]
    mine = plane_hashes(torch, tp, wp, z1 - z0, plane, device)
    all_hashes = D.gather_int64(mine)
    n_pts_local = vol.extract_surface_count()
    pts_total = int(D.sum(n_pts_local))
    # (1) N > 1: a monolithic fusion of the same frames on rank 0 -> per-plane hashes and the cloud
    if world > 1:
        ok_slabs, ok_cloud, mono_pts = None, None, None
        if rank == 0:
            free, _ = torch.cuda.mem_get_info()
            need = 8 * plane * dz + (2 << 30)
            if free > need:
                with pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, device=D.local_rank,
                                max_batch=args.batch, cull_level=args.cull) as mono:
                    mono.integrate_batch(depth, poses)
                    mono.synchronize()
                    mtp, mwp = mono.device_ptrs()
                    mono_h = plane_hashes(torch, mtp, mwp, dz, plane, device)
                    mono_cloud = mono.extract_surface()
                sharded = np.concatenate(all_hashes)
                ok_slabs = bool(sharded.shape == mono_h.shape and np.array_equal(sharded, mono_h))
                mono_pts = int(mono_cloud.shape[0])
                ok_cloud = bool(gathered_cloud is not None and cloud_hash(gathered_cloud) == cloud_hash(mono_cloud)
                                and gathered_cloud.shape == mono_cloud.shape)
                del mono_cloud
            else:
                out["monolithic"] = f"skipped: {free / 2**30:.0f} GiB free on rank 0, {need / 2**30:.0f} GiB needed"
        out["slabs_equal_monolithic"] = ok_slabs
        out["gathered_cloud_equal_monolithic"] = ok_cloud
        out["monolithic_surface_points"] = mono_pts
    out["surface_points"] = pts_total
    out["planes_hashed"] = int(sum(len(h) for h in all_hashes))
    # (2) oracle planes: at every slab cut (cut - 1, cut, cut + 1) for N > 1, three planes of the volume for N = 1;
    #     all F frames, checked bit for bit on the rank that holds the plane
    if not args.no_oracle_planes:
        if world > 1:
            zs = sorted({z for a, _ in ranges[1:] for z in (a - 1, a, a + 1) if 0 <= z < dz})
        else:
            zs = [dz // 4, dz // 2, (3 * dz) // 4]
        if args.oracle_planes_max and len(zs) > args.oracle_planes_max:
            zs = [zs[i] for i in np.linspace(0, len(zs) - 1, args.oracle_planes_max).astype(int)]
        use_all_host_threads()
        oracle = graft.load_oracle()
        depth_cpu = depth.cpu().numpy() if any(z0 <= z < z1 for z in zs) else None
        bad, checked = 0, 0
        t_arr = torch.as_tensor(_DevArray(tp, (z1 - z0, dy, dx)), device=device)
        w_arr = torch.as_tensor(_DevArray(wp, (z1 - z0, dy, dx)), device=device)
        t0 = time.perf_counter()
        for turn in range(world):       # the ranks take turns: the oracle uses every host core
            if turn == rank:
                for z in zs:
                    if not (z0 <= z < z1):
                        continue
                    t_ref, w_ref = oracle.fuse(synth.CAM_K, poses, depth_cpu, grid.dims, grid.origin, grid.voxel_size,
                                               grid.trunc_margin, z_range=(z, z + 1))
                    t_gpu, w_gpu = t_arr[z - z0].cpu().numpy(), w_arr[z - z0].cpu().numpy()
                    same = np.array_equal(t_gpu.view(np.uint32), t_ref[0].view(np.uint32)) and np.array_equal(w_gpu, w_ref[0])
                    bad += 0 if same else 1
                    checked += 1
            if world > 1:
                D.dist.barrier()
        out["oracle_planes"] = sorted(int(z) for z in zs)
        out["oracle_planes_checked"] = int(D.sum(checked))
        out["oracle_planes_mismatching"] = int(D.sum(bad))
        out["oracle_seconds"] = round(D.max(time.perf_counter() - t0), 1)
    vals = [v for k, v in out.items() if k in ("slabs_equal_monolithic", "gathered_cloud_equal_monolithic") and v is not None]
    out["all_equal"] = bool(all(vals) and out.get("oracle_planes_mismatching", 0) == 0)
    return out


# =====================================================================================================================
# config 5: fuseSceneFragments-style batch of independent fragment volumes
# =====================================================================================================================

def run_fragments(args, D):
    torch = D.torch
    rank, world, device = D.rank, D.world, D.device
    pkg = graft.load_package()
    synth = pkg.synth
    cfg = dict(synth.CONFIGS["frag300"])
    grid, F = cfg["grid"], cfg["frames"]
    n_frag = args.fragments or cfg["fragments"]
    mine = list(range(rank, n_frag, world))              # round-robin, no communication
    B = max(1, min(args.frag_batch, len(mine)))
    vox = grid.n_voxels
    # inputs: fragment i uses trajectory frames 50 i .. 50 i + 49 with base frame 50 i (fuseSceneFragments.m:59-64)
    mm_dev = torch.empty((len(mine), F, 480, 640), dtype=torch.int16, device=device)
    poses = []
    for li, i in enumerate(mine):
        frames = list(range(F * i, F * i + F))
        mm_dev[li] = synth.mm_to_u16_tensor(synth.render_depth_mm(frames, base_frame=F * i, device=device))
        c2w, b2w = synth.make_poses(frames, base_frame=F * i)
        poses.append(np.stack([pkg.cam2base_from_world(b2w, m).ravel() for m in c2w]).astype(np.float32))
    torch.cuda.synchronize()
    vols = [pkg.Volume(grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin, synth.CAM_K, device=D.local_rank,
                       max_batch=args.batch, cull_level=args.cull) for _ in range(B)]
    lead = vols[0]

    def fuse_all(src, multi=True, extract=False, sink=None):
        for b0 in range(0, len(mine), B):
            idx = list(range(b0, min(b0 + B, len(mine))))
            vs = vols[:len(idx)]
            for v in vs:
                v.reset()
            if multi:
                pkg.integrate_multi(vs, [src[li] for li in idx], [poses[li] for li in idx])
            else:
                for v, li in zip(vs, idx):
                    v.integrate_batch(src[li], poses[li])
            if extract:
                for v, li in zip(vs, idx):
                    pts = v.extract_surface()
                    if sink is not None:
                        sink(li, v, pts)
        for v in vols:
            v.synchronize()

    def timed(fn):
        D.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    sampler = ClockSampler(D.local_rank)
    sampler.start()
    for _ in range(args.warmup):
        fuse_all(mm_dev)
    D.barrier()
    for v in vols:
        v.reset_stats()
    sampler.mark()
    t = timed(lambda: [fuse_all(mm_dev) for _ in range(args.steps)])
    clocks = sampler.stop()
    rank_ms = [round(x * 1e3 / args.steps, 2) for x in D.gather_floats(t)]
    ms_per_step = D.max(t) * 1e3 / args.steps
    value = n_frag * vox * F / (ms_per_step * 1e-3)
    launches = D.sum(sum(v.stats()["kernel_launches"] for v in vols))
    # the same fragments through ONE handle, one after the other (the round-1 design of fuse_fragments: one live
    # handle per GPU, re-initialised between fragments)
    def fuse_one_handle():
        for li in range(len(mine)):
            lead.reset()
            lead.integrate_batch(mm_dev[li], poses[li])
        lead.synchronize()
    fuse_one_handle()
    t_seq = D.max(timed(fuse_one_handle))
    # fusion only (no reset) for the roofline
    t_int = D.max(timed(lambda: fuse_all(mm_dev)))   # includes the fills; they are 2 x 108 MB per fragment
    peak, peak_src = measured_peak_gbs()
    n_pass = -(-F // args.batch)
    alg = BYTES_PER_VOXEL_PASS * vox * n_pass * len(mine)
    achieved = alg / t_int / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None, "peak_source": peak_src, "kernel": "stream_integrate_kernel (one launch per batch of fragments)",
                "fragments_per_launch": B, "frames_per_pass": args.batch, "passes_per_fragment": n_pass,
                "binding": "instruction issue (see the 1024^3 line's roofline_issue); small volumes additionally pay the "
                           "launch tail once per batch of fragments"}
    # e2e: pinned host uint16 frames -> H2D inside tsdf_integrate_multi, extraction, D2H of every fragment's points
    mm_host = torch.empty(mm_dev.shape, dtype=torch.int16, pin_memory=True)
    mm_host.copy_(mm_dev)
    torch.cuda.synchronize()
    counts = {}
    fuse_all(mm_host, extract=True)
    for v in vols:
        v.reset_stats()
    t_e2e = D.max(timed(lambda: fuse_all(mm_host, extract=True, sink=lambda li, v, pts: counts.__setitem__(li, len(pts)))))
    h2d = D.sum(sum(v.stats()["h2d_bytes"] for v in vols) + sum(p.nbytes for p in poses))
    d2h = D.sum(sum(v.stats()["d2h_bytes"] for v in vols))
    e2e = {"value": n_frag * vox * F / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": round(t_e2e * 1e3, 2), "surface_points": int(D.sum(sum(counts.values()))),
           "path": "per batch of fragments: tsdf_reset x B + tsdf_integrate_multi(pinned host uint16 frames) + "
                   "tsdf_extract_surface + tsdf_copy_surface per fragment (the call sequence of bin/fuse_fragments)"}
    # parity: multi-volume launch == one handle at a time (all fragments, hashes), and the first fragments of every rank
    # against the oracle (full volume, bit for bit)
    parity = None
    if not args.no_parity:
        plane = grid.dims[0] * grid.dims[1]
        hashes = {}

        def grab(tag):
            def sink(li, v, pts):
                v.synchronize()
                tp, wp = v.device_ptrs()
                h = plane_hashes(torch, tp, wp, grid.dims[2], plane, device)
                hashes[(tag, li)] = (int(h.sum()), cloud_hash(pts))
            return sink
        fuse_all(mm_dev, multi=True, extract=True, sink=grab("multi"))
        fuse_all(mm_dev, multi=False, extract=True, sink=grab("seq"))
        same = all(hashes[("multi", li)] == hashes[("seq", li)] for li in range(len(mine)))
        bad, checked = 0, 0
        if not args.no_oracle_planes:
            use_all_host_threads()
            oracle = graft.load_oracle()
            for li in range(min(args.oracle_fragments, len(mine))):
                lead.reset()
                lead.integrate_batch(mm_dev[li], poses[li])
                t_gpu, w_gpu = lead.download()
                depth_cpu = oracle.depth_from_u16(mm_dev[li].cpu().numpy().view(np.uint16))
                t_ref, w_ref = oracle.fuse(synth.CAM_K, poses[li], depth_cpu, grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin)
                ok = np.array_equal(t_gpu.view(np.uint32), t_ref.view(np.uint32)) and np.array_equal(w_gpu, w_ref)
                hm = hashes[("multi", li)]
                tp, wp = lead.device_ptrs()
                ok = ok and int(plane_hashes(torch, tp, wp, grid.dims[2], plane, device).sum()) == hm[0]
                bad += 0 if ok else 1
                checked += 1
        parity = {"multi_equals_sequential": bool(D.sum(0.0 if same else 1.0) == 0), "fragments_hashed": int(D.sum(len(mine))),
                  "oracle_fragments_checked": int(D.sum(checked)), "oracle_fragments_mismatching": int(D.sum(bad))}
        parity["all_equal"] = bool(parity["multi_equals_sequential"] and parity["oracle_fragments_mismatching"] == 0)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        d0 = pkg.depth_from_u16(mm_dev[0].cpu().numpy().view(np.uint16))
        cpu = cpu_baseline(pkg, cfg, d0, poses[0], target_s=args.cpu_seconds)
    for v in vols:
        v.close()
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"{n_frag} independent {grid.dims[0]}^3 fragment volumes ({grid.voxel_size * 1e3:.0f} mm voxels) x {F} frames "
                                      f"640x480 each (own base pose), round-robin over {world} GPU(s), {B} fragments per launch",
                          "config_key": "frag300", "fragments": n_frag, "frames": F, "batch": args.batch, "fragments_per_launch": B,
                          "l2": "each batch of fragments (>= 1.7 GB of volumes) is larger than L2; no flush needed",
                          "parallelism": f"fragments x{world} (replicas, no collective)", "rank_ms_per_step": rank_ms},
               "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
               "sequential_handles": {"value": n_frag * vox * F / t_seq, "unit": UNIT, "ms_per_step": round(t_seq * 1e3, 2),
                                      "note": "the same fragments through ONE handle, one tsdf_reset + tsdf_integrate_batch per fragment after the other"},
               "parity": parity, "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(out), flush=True)


# =====================================================================================================================
# baselines
# =====================================================================================================================

def sample_slabs(dz, n_slabs, planes):
    """Thin z-slabs spread evenly over the grid, so that the sample averages over cheap (culled / occluded) and
    expensive (surface band) planes the way the whole volume does."""
    n_slabs = max(1, min(n_slabs, dz // max(planes, 1)))
    return [(int(c) - planes // 2, int(c) - planes // 2 + planes)
            for c in (np.arange(n_slabs) + 0.5) * dz / n_slabs]


def cpu_baseline(pkg, cfg, depth, poses, target_s=12.0, force_port=False):
    """Reference Integrate body on host threads (oracle/_ref) or the oracle port, bounded sample."""
    threads = use_all_host_threads()
    oracle = graft.load_oracle()
    grid = cfg["grid"]
    dx, dy, dz = grid.dims
    planes = max(1, min(4, (8 * 1024 * 1024) // (dx * dy)))
    slabs = sample_slabs(dz, 8, planes)
    K = pkg.synth.CAM_K
    use_ref = oracle.have_ref_host() and not force_port

    def run(n_frames):
        t0 = time.perf_counter()
        for a, b in slabs:
            if use_ref:
                # the reference kernel addresses the grid from z = 0: shift the origin by `a` planes
                oz = float(np.float32(grid.origin[2]) + np.float32(a) * np.float32(grid.voxel_size))
                oracle.ref_host_fuse(K, poses[:n_frames], depth[:n_frames], (dx, dy, b - a),
                                     (grid.origin[0], grid.origin[1], oz), grid.voxel_size, grid.trunc_margin)
            else:
                oracle.fuse(K, poses[:n_frames], depth[:n_frames], grid.dims, grid.origin, grid.voxel_size,
                            grid.trunc_margin, z_range=(a, b))
        return time.perf_counter() - t0

    t1 = run(1)
    n = int(max(1, min(len(poses), target_s / max(t1, 1e-3))))
    t = run(n)
    nvox = dx * dy * sum(b - a for a, b in slabs)
    cores = oracle.ref_host_num_threads() if use_ref else oracle.num_threads()
    return {"value": nvox * n / t, "unit": UNIT, "cores": cores, "host_cpus": threads, "kind": "reference" if use_ref else "port",
            "sample": f"{len(slabs)} slabs of {planes} z-planes spread evenly over the {dx}x{dy}x{dz} grid (z = "
                      f"{', '.join(str(a) for a, _ in slabs)}) x {n} frames ({nvox * n / 1e9:.2f} G voxel-updates in {t:.1f} s), "
                      + ("reference Integrate body (demo.cu:18-68) compiled for the host, OpenMP over (z, y)"
                         if use_ref else "oracle port (oracle/tsdf_oracle.c), OpenMP over rows")}


REF_CUDA_SAMPLE_SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as graft
oracle = graft.load_oracle()
d = np.load(sys.argv[2])
dims = tuple(int(v) for v in d["dims"])
origin = tuple(float(v) for v in d["origin"])
args = (d["cam_K"], d["poses"], d["depth"], dims, origin, float(d["voxel"]), float(d["trunc"]))
oracle.ref_cuda_fuse(*args[:1], args[1][:2], args[2][:2], *args[3:], download=False)           # context + first launch
_, _, ms = oracle.ref_cuda_fuse(*args, download=False)
print(json.dumps({"kernel_ms": float(ms)}))
"""


def ref_cuda_baseline(pkg, cfg, depth, poses, n_frames=16):
    """Reported baseline (north-star): the reference's own Integrate kernel, depth-fusion/demo.cu compiled unchanged
    for sm_100a (oracle/_ref/libref_cuda.so), <<<dim_z, dim_y>>> per frame exactly as demo.cu:179-182 launches it, on
    the WHOLE grid of the configuration (same config as the engine's number; fewer frames), CUDA events around the
    launches only.  Runs in its OWN process so that nothing it does can disturb the measured engine; any failure is
    reported, never raised."""
    import tempfile
    try:
        oracle = graft.load_oracle()
        if not oracle.have_ref_cuda():
            return {"unavailable": "oracle/_ref/libref_cuda.so not built"}
        grid = cfg["grid"]
        dx, dy, dz = grid.dims
        if dy > 1024:
            return {"unavailable": "reference launches dim_y threads per block (demo.cu:179): dim_y > 1024 cannot run; "
                                   "its int32 voxel index (demo.cu:61) overflows beyond 2^31 voxels"}
        n = min(n_frames, len(poses))
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "sample.npz")
            np.savez(path, cam_K=np.asarray(pkg.synth.CAM_K, np.float32), poses=np.asarray(poses[:n], np.float32),
                     depth=np.asarray(depth[:n], np.float32), dims=np.array([dx, dy, dz]),
                     origin=np.array(grid.origin, np.float64), voxel=np.float64(grid.voxel_size),
                     trunc=np.float64(grid.trunc_margin))
            r = subprocess.run([sys.executable, "-c", REF_CUDA_SAMPLE_SCRIPT, str(ROOT), path], capture_output=True,
                               text=True, timeout=300)
        if r.returncode != 0:
            return {"unavailable": ("exit %d: " % r.returncode) + r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else "failed"}
        ms = json.loads(r.stdout.strip().splitlines()[-1])["kernel_ms"]
        return {"value": dx * dy * dz * n / (ms * 1e-3), "unit": UNIT, "kind": "reference CUDA kernel rebuilt for sm_100a",
                "kernel_ms": round(ms, 3),
                "sample": f"the whole {dx}x{dy}x{dz} grid (z-planes [0, {dz})) x the first {n} frames, one "
                          f"Integrate<<<{dz}, {dy}>>> launch per frame (demo.cu:179-182), depth and poses resident"}
    except Exception as e:  # a reported baseline must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:240]}


def run_reference(args):
    """--impl reference: the reference's own CPU-executable implementation of the path (its Integrate body
    compiled for the host, oracle/_ref/libref_host.so; else the oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = use_all_host_threads()   # torch.distributed.run exports OMP_NUM_THREADS=1
    pkg = graft.load_package()
    oracle = graft.load_oracle()
    synth = pkg.synth
    cfg = dict(synth.CONFIGS[args.config])
    if args.frames:
        cfg["frames"] = args.frames
    grid, F = cfg["grid"], cfg["frames"]
    n_frames = 8
    frames = [int(v) for v in np.linspace(0, F - 1, n_frames)]   # spread over the trajectory
    mm = synth.mm_to_u16_numpy(synth.render_depth_mm(frames, device="cpu"))
    depth = oracle.depth_from_u16(mm)
    c2w, b2w = synth.make_poses(frames)
    poses = np.stack([oracle.cam2base_from_world(b2w, m).ravel() for m in c2w]).astype(np.float32)
    dx, dy, dz = grid.dims
    planes = max(1, min(4, (4 * 1024 * 1024) // (dx * dy)))
    slabs = sample_slabs(dz, 8, planes)
    use_ref = oracle.have_ref_host()

    def step():
        for a, b in slabs:
            if use_ref:
                oz = float(np.float32(grid.origin[2]) + np.float32(a) * np.float32(grid.voxel_size))
                oracle.ref_host_fuse(synth.CAM_K, poses, depth, (dx, dy, b - a), (grid.origin[0], grid.origin[1], oz),
                                     grid.voxel_size, grid.trunc_margin)
            else:
                oracle.fuse(synth.CAM_K, poses, depth, grid.dims, grid.origin, grid.voxel_size, grid.trunc_margin,
                            z_range=(a, b))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    t = time.perf_counter() - t0
    updates = dx * dy * sum(b - a for a, b in slabs) * n_frames
    value = updates * args.steps / t
    cores = oracle.ref_host_num_threads() if use_ref else oracle.num_threads()
    sample = (f"each step: {len(slabs)} slabs of {planes} z-planes spread evenly over the {dx}x{dy}x{dz} grid (z = "
              f"{', '.join(str(a) for a, _ in slabs)}) x {n_frames} frames spread over the {F}-frame trajectory "
              f"= {updates / 1e6:.0f} M voxel-updates")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{grid.dims[0]}^3 grid ({grid.voxel_size * 1e3:.1f} mm voxels), {F} frames 640x480 "
                               f"(bounded sample per step, see cpu_baseline.sample)", "config_key": args.config},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cpus": threads,
                         "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", default="1024", choices=["1024", "512", "demo500", "2048", "frag300"])
    ap.add_argument("--frames", type=int, default=0, help="override the configuration's frame count")
    ap.add_argument("--batch", type=int, default=32, help="K: frames per chunk = per pass over the volume (stream engine: 32)")
    ap.add_argument("--cull", type=int, default=2)
    ap.add_argument("--tuning", type=int, default=0, help="validation library: kernel variant word (0 = shipped library)")
    ap.add_argument("--quick", action="store_true", help="value only: skip e2e, rooflines, parity and baselines")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-oracle-planes", action="store_true", help="parity block without the CPU oracle planes")
    ap.add_argument("--oracle-planes-max", type=int, default=9, help="cap on the number of oracle planes (0 = all cuts +-1)")
    ap.add_argument("--equal-slabs", action="store_true", help="equal-height z-slabs instead of equal-work ones")
    ap.add_argument("--rebalance", type=int, default=10, help="feedback-balancing warm-up steps for N > 1")
    ap.add_argument("--base-cost", type=float, default=0.05, help="cost of a fully culled z group relative to the mean group")
    ap.add_argument("--fragments", type=int, default=0, help="frag300: number of fragments (default 256)")
    ap.add_argument("--frag-batch", type=int, default=16, help="frag300: fragments per launch")
    ap.add_argument("--oracle-fragments", type=int, default=1, help="frag300: fragments per rank checked against the oracle")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    D = Dist(args)
    try:
        if args.config == "frag300":
            run_fragments(args, D)
        else:
            run_volume(args, D)
    finally:
        D.close()


if __name__ == "__main__":
    main()
