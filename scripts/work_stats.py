#!/usr/bin/env python
"""Diagnostic (CPU, numpy): where do the classified (voxel, frame) pairs of the 1024^3 workload end up?

For a random sample of warp tiles (32 x 4 x 8 voxels) and frames, every voxel of the tile is projected the way
depth-fusion/demo.cu:30-62 does (float64 here: statistics, not parity) and binned:
  out     cam_z <= 0 or pixel outside the image       (demo.cu:43,48)
  invalid depth pixel 0 / > 6                          (demo.cu:53)
  behind  diff <= -trunc                               (demo.cu:57)
  free    e >= trunc (dist = 1)
  band    everything else (needs the exact evaluation)
and the tile is run through (a) the engine's hierarchical-Z test as implemented (square blocks of 8..128 px, the level
whose block covers the larger extent, 2 x 2 blocks) and (b) candidate tests, to see how many kept (tile, frame) pairs
contain no voxel that is updated at all.

  python scripts/work_stats.py [n_tiles] [n_frames] [config]
"""
import importlib.util
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402


def max_pyramid_aniso(depth, sx, sy):
    h, w = depth.shape
    hh, ww = -(-h // sy), -(-w // sx)
    pad = np.full((hh * sy, ww * sx), -np.inf, dtype=depth.dtype)
    pad[:h, :w] = depth
    return pad.reshape(hh, sy, ww, sx).max(axis=(1, 3))


def main():
    n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    key = sys.argv[3] if len(sys.argv) > 3 else "1024"
    pkg = graft.load_package()
    synth = pkg.synth
    cfg = synth.CONFIGS[key]
    grid, F = cfg["grid"], cfg["frames"]
    dx, dy, dz = grid.dims
    vs, tr = grid.voxel_size, grid.trunc_margin
    ox, oy, oz = grid.origin
    fx, fy, cx, cy = 585.0, 585.0, 320.0, 240.0
    W, H = 640, 480
    rng = np.random.default_rng(7)
    frames = sorted(rng.choice(F, n_frames, replace=False).tolist())
    mm = synth.render_depth_mm(frames, device="cpu").numpy()
    depth = mm.astype(np.float64) / 1000.0
    depth[depth > 6.0] = 0.0
    san = np.where(depth > 0, depth, -np.inf)
    levels = [8, 16, 32, 64, 128]
    sq = [[max_pyramid_aniso(s, l, l) for l in levels] for s in san]
    an = [{(a, b): max_pyramid_aniso(s, a, b) for a in levels for b in levels} for s in san]
    c2w, b2w = synth.make_poses(frames)
    poses = [np.linalg.inv(b2w.astype(np.float64)) @ m.astype(np.float64) for m in c2w]

    TX, TY, TZ = (int(v) for v in os.environ.get('TILE', '32,4,8').split(','))
    ntx, nty, ntz = dx // TX, dy // TY, dz // TZ
    tiles = np.stack([rng.integers(0, ntx, n_tiles), rng.integers(0, nty, n_tiles), rng.integers(0, ntz, n_tiles)], 1)
    lx = np.arange(TX)[None, None, :]
    ly = np.arange(TY)[None, :, None]
    lz = np.arange(TZ)[:, None, None]
    tot = dict(out=0, invalid=0, behind=0, free=0, band=0)
    kept_now = kept_aniso = kept_ideal = kept_frustum = 0
    kept_multi = useful_multi = 0
    ZGS = ()
    sub_tot = {z: 0 for z in ZGS}
    sub_kept = {z: 0 for z in ZGS}
    MAXB = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    vox_now = vox_aniso = 0
    useful_now = useful_aniso = 0
    n_pairs = 0
    for fi, P in enumerate(poses):
        R, t = P[:3, :3], P[:3, 3]
        for (tx, ty, tz) in tiles:
            X = ox + (tx * TX + lx) * vs
            Y = oy + (ty * TY + ly) * vs
            Z = oz + (tz * TZ + lz) * vs
            X, Y, Z = np.broadcast_arrays(X, Y, Z)
            p = np.stack([X - t[0], Y - t[1], Z - t[2]], 0).reshape(3, -1)
            cam = R.T @ p
            czv = cam[2]
            pos = czv > 0
            safe = np.where(pos, czv, 1.0)
            u = fx * cam[0] / safe + cx
            v = fy * cam[1] / safe + cy
            px = np.floor(u + 0.5).astype(np.int64)
            py = np.floor(v + 0.5).astype(np.int64)
            inb = pos & (px >= 0) & (px < W) & (py >= 0) & (py < H)
            d = np.zeros_like(czv)
            d[inb] = depth[fi][py[inb], px[inb]]
            valid = inb & (d > 0)
            e = d - czv
            ray = np.sqrt(1 + (cam[0] / safe) ** 2 + (cam[1] / safe) ** 2)
            diff = e * ray
            behind = valid & (diff <= -tr)
            free = valid & (e >= tr)
            band = valid & ~behind & ~free
            n_out, n_inv = int((~inb).sum()), int((inb & ~valid).sum())
            n_b, n_f, n_band = int(behind.sum()), int(free.sum()), int(band.sum())
            n_pairs += 1
            useful = (n_f + n_band) > 0
            kept_ideal += useful
            if inb.any():
                kept_frustum += 1
            # the engine's test (corners; conservative, here on the exact voxel set which is tighter than corners+margin)
            if not inb.any():
                continue
            if not pos.all():
                now = aniso = True
            else:
                x0, x1 = max(0, int(np.floor(u.min() - 1.5))), min(W - 1, int(np.ceil(u.max() + 1.5)))
                y0, y1 = max(0, int(np.floor(v.min() - 1.5))), min(H - 1, int(np.ceil(v.max() + 1.5)))
                czmin = czv.min()
                ext = max(x1 - x0, y1 - y0) + 1
                if ext > 128:
                    now = True
                else:
                    l = 0
                    while levels[l] < ext:
                        l += 1
                    s = levels[l]
                    m = sq[fi][l]
                    dmax = max(m[y0 // s, x0 // s], m[y0 // s, x1 // s], m[y1 // s, x0 // s], m[y1 // s, x1 // s])
                    now = not (dmax + tr * 1.0001 <= czmin)
                ex, ey = x1 - x0 + 1, y1 - y0 + 1
                if ex > 128 or ey > 128:
                    aniso = True
                else:
                    a = next(l for l in levels if l >= ex)
                    b = next(l for l in levels if l >= ey)
                    m = an[fi][(a, b)]
                    dmax = max(m[y0 // b, x0 // a], m[y0 // b, x1 // a], m[y1 // b, x0 // a], m[y1 // b, x1 // a])
                    aniso = not (dmax + tr * 1.0001 <= czmin)
                # (c) finest square level whose block grid over the bounding box has at most MAXB blocks
                multi = True
                for l, sz in enumerate(levels):
                    bx0, bx1, by0, by1 = x0 // sz, x1 // sz, y0 // sz, y1 // sz
                    if (bx1 - bx0 + 1) * (by1 - by0 + 1) <= MAXB:
                        dmax = sq[fi][l][by0:by1 + 1, bx0:bx1 + 1].max()
                        multi = not (dmax + tr * 1.0001 <= czmin)
                        break
            if not pos.all():
                multi = True
            kept_multi += multi
            useful_multi += (multi and useful)
            if multi:
                # z-granular refinement: groups of ZG planes of the kept tile, same test on the group's voxels
                U, V, CZ, INB = (a.reshape(TZ, TY * TX) for a in (u, v, czv, inb))
                upd = (free | band).reshape(TZ, TY * TX)
                for zg in ZGS:
                    for g0 in range(0, TZ, zg):
                        sl = slice(g0, g0 + zg)
                        sub_tot[zg] += 1
                        if not INB[sl].any():
                            continue
                        keep = True
                        if pos.all():
                            gx0, gx1 = max(0, int(np.floor(U[sl].min() - 1.5))), min(W - 1, int(np.ceil(U[sl].max() + 1.5)))
                            gy0, gy1 = max(0, int(np.floor(V[sl].min() - 1.5))), min(H - 1, int(np.ceil(V[sl].max() + 1.5)))
                            if gx1 < gx0 or gy1 < gy0:
                                keep = False
                            else:
                                for l, sz in enumerate(levels):
                                    b0, b1, c0, c1 = gx0 // sz, gx1 // sz, gy0 // sz, gy1 // sz
                                    if (b1 - b0 + 1) * (c1 - c0 + 1) <= MAXB:
                                        keep = not (sq[fi][l][c0:c1 + 1, b0:b1 + 1].max() + tr * 1.0001 <= CZ[sl].min())
                                        break
                        sub_kept[zg] += keep
                        assert keep or not upd[sl].any()
            assert (not useful) or multi
            for k, n in zip(("out", "invalid", "behind", "free", "band"), (n_out, n_inv, n_b, n_f, n_band)):
                if now:
                    tot[k] += n
            kept_now += now
            kept_aniso += aniso
            useful_now += (now and useful)
            useful_aniso += (aniso and useful)
            assert (not useful) or (now and aniso), "a test culled a tile that holds updated voxels"
    nv = TX * TY * TZ
    print(f"config {key}: {n_tiles} warp tiles x {n_frames} frames = {n_pairs} (tile, frame) pairs")
    print(f"  in frustum (any voxel in the image): {kept_frustum / n_pairs:.3f}")
    print(f"  kept by the engine's test:           {kept_now / n_pairs:.3f}   of which useful {useful_now / max(kept_now, 1):.3f}")
    print(f"  kept by the anisotropic pyramid:     {kept_aniso / n_pairs:.3f}   of which useful {useful_aniso / max(kept_aniso, 1):.3f}")
    print(f"  kept by <= {MAXB} blocks of the finest level: {kept_multi / n_pairs:.3f}   of which useful {useful_multi / max(kept_multi, 1):.3f}")
    for z in ZGS:
        print(f"    of the kept tiles' groups of {z} plane(s), kept by the same test: {sub_kept[z] / max(sub_tot[z], 1):.3f}")
    print(f"  ideal (tile holds an updated voxel): {kept_ideal / n_pairs:.3f}")
    s = sum(tot.values())
    print("  voxels of the kept pairs: " + ", ".join(f"{k} {v / s:.3f}" for k, v in tot.items()))
    print(f"  classified voxel-frames per voxel-update of the metric: {s / (n_pairs * nv):.3f}")


if __name__ == "__main__":
    main()
