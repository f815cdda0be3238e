// tsdf_kernels.cu -- hand-written sm_100a kernels of the TSDF fusion engine.
//
// Replaces (reference paths relative to the reference repository root):
//   depth-fusion/demo.cu:18-68    Integrate                       -> integrate_kernel
//   depth-fusion/demo.cu:131-145  host-side init + upload         -> fill_kernel
//   depth-fusion/utils.hpp:82-84  ReadDepth's mm -> m conversion  -> depth_from_u16_kernel
//   depth-fusion/utils.hpp:20-53  surface scan (count, emit)      -> extract_count/scan/emit kernels
//
// Numerical contract: every arithmetic step of demo.cu:30-66 is issued with an explicit
// round-to-nearest intrinsic (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn), so nvcc can
// never contract a multiply-add and the result is the "--fmad=false" evaluation of the reference
// regardless of compiler flags.  The only re-arrangements are bit-exact ones: hoisting of
// products/sums that do not depend on the inner index, powf(q,2) == q*q, skipping divisions
// whose result is known exactly, and culling voxels that provably fail the reference's own
// early-outs (demo.cu:43-49).
#include "tsdf_device.cuh"

namespace tsdf {

namespace {


// ------------------------------------------------------------------------------------------------
// integrate_kernel: K frames fused per pass, each voxel's TSDF/weight held in registers.
//
// Thread -> voxels: one thread owns a patch of VEC consecutive x  x  ZT consecutive z voxels at
// one y (VEC = 4: one float4 per plane).  A block of 256 threads is 2^lx_shift threads along x
// by 256 >> lx_shift rows of y; blockIdx = (x tile, y tile, z group).  Lanes of a warp are
// consecutive along x, so every TSDF/weight access is a fully coalesced 128-bit access, and the
// ZT planes of a patch project to (nearly) the same pixels, so the depth gather hits L1.
//
// For a frame, the camera coordinates of voxel (j, k) of the patch are
//   cam = ((M0*t0[j] + M4*t1) + M8*t2[k])    (demo.cu:39-41, left-to-right)
// so the x/y part is formed once per (j) and the z product once per (k): 3 adds per voxel.
//
// TSDF/weight are loaded lazily -- only when some voxel of the warp is actually observed by one
// of the K frames -- and stored only for planes that changed, so untouched parts of the volume
// cost no HBM traffic at all.
// ------------------------------------------------------------------------------------------------
template <int VEC, int ZT, bool DEVPOSE, int MINB, bool FAST>
__global__ void __launch_bounds__(kBlockThreads, MINB)
integrate_kernel(const __grid_constant__ IntegrateParams P, const float* __restrict__ d_cam_K,
                 const float* __restrict__ d_cam2base) {
  const int lx_mask = (1 << P.lx_shift) - 1;
  const int lx = threadIdx.x & lx_mask;
  const int ly = threadIdx.x >> P.lx_shift;
  const int rows_per_block = kBlockThreads >> P.lx_shift;
  const int xu = (blockIdx.x << P.lx_shift) + lx;
  const int y = blockIdx.y * rows_per_block + ly;
  // z groups are dispatched heaviest first (far planes see more of the frustum) to shorten the tail
  const int zl0 = (P.z_reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z) * ZT;
  const bool active = (xu < P.nx_units) && (y < P.dim_y);
  const int zcount = min(ZT, P.nz - zl0);

  Intrinsics K;
  FramePose dev_pose;
  if (DEVPOSE) {
    K.fx = __ldg(d_cam_K + 0); K.cx = __ldg(d_cam_K + 2); K.fy = __ldg(d_cam_K + 4); K.cy = __ldg(d_cam_K + 5);
    K.lim_x = 0.5f - pixel_round_margin((float)P.im_w, K.cx);
    K.lim_y = 0.5f - pixel_round_margin((float)P.im_h, K.cy);
    dev_pose.m0 = __ldg(d_cam2base + 0); dev_pose.m1 = __ldg(d_cam2base + 1); dev_pose.m2 = __ldg(d_cam2base + 2);
    dev_pose.m4 = __ldg(d_cam2base + 4); dev_pose.m5 = __ldg(d_cam2base + 5); dev_pose.m6 = __ldg(d_cam2base + 6);
    dev_pose.m8 = __ldg(d_cam2base + 8); dev_pose.m9 = __ldg(d_cam2base + 9); dev_pose.m10 = __ldg(d_cam2base + 10);
    dev_pose.tx = __ldg(d_cam2base + 3); dev_pose.ty = __ldg(d_cam2base + 7); dev_pose.tz = __ldg(d_cam2base + 11);
  } else {
    K.fx = P.fx; K.fy = P.fy; K.cx = P.cx; K.cy = P.cy;
    K.lim_x = P.lim_x; K.lim_y = P.lim_y;
  }

  // ---- per-warp frame mask: lane f tests frame f against the bounding box of the warp's voxels
  //      (2^lx_shift*VEC x 32>>lx_shift x ZT); frames whose frustum cannot touch it are never looked at
  unsigned frame_mask = (P.n_frames >= 32) ? 0xffffffffu : ((1u << P.n_frames) - 1u);
  if (!DEVPOSE && P.cull_level >= 1) {
    bool keep = false;
    const int f = threadIdx.x & 31;
    if (f < P.n_frames) {
      const FramePose& ps = P.pose[f];
      const int rows_per_warp = 32 >> P.lx_shift;
      const int x_lo = (blockIdx.x << P.lx_shift) * VEC;
      const int x_hi = min(P.dim_x - 1, x_lo + ((VEC << P.lx_shift) - 1));
      const int y_lo = blockIdx.y * rows_per_block + (threadIdx.x >> 5) * rows_per_warp;
      const int y_hi = min(P.dim_y - 1, y_lo + rows_per_warp - 1);
      const int z_lo = P.z_begin + zl0;
      const int z_hi = z_lo + ZT - 1;
      CullAccum acc;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float bx = fmaf((float)((c & 1) ? x_hi : x_lo), P.voxel_size, P.ox) - ps.tx;
        const float by = fmaf((float)((c & 2) ? y_hi : y_lo), P.voxel_size, P.oy) - ps.ty;
        const float bz = fmaf((float)((c & 4) ? z_hi : z_lo), P.voxel_size, P.oz) - ps.tz;
        acc.add(fmaf(ps.m0, bx, fmaf(ps.m4, by, ps.m8 * bz)), fmaf(ps.m1, bx, fmaf(ps.m5, by, ps.m9 * bz)),
                fmaf(ps.m2, bx, fmaf(ps.m6, by, ps.m10 * bz)), P, K);
      }
      keep = !acc.culled() && (y_lo < P.dim_y);
    }
    frame_mask = __ballot_sync(0xffffffffu, keep);
  }
  if (frame_mask == 0u) return;  // warp-uniform

  // ---- voxel-centre coordinates (demo.cu:30-32), always from the GLOBAL index ----
  float base_x[VEC], base_z[ZT];
#pragma unroll
  for (int j = 0; j < VEC; ++j) base_x[j] = __fadd_rn(P.ox, __fmul_rn((float)(xu * VEC + j), P.voxel_size));
  const float base_y = __fadd_rn(P.oy, __fmul_rn((float)y, P.voxel_size));
#pragma unroll
  for (int k = 0; k < ZT; ++k)
    base_z[k] = __fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size));

  // voxels of the patch that exist (VEC divides dim_x, so there is no x tail inside a patch)
  unsigned exist_mask = 0;
  if (active) {
#pragma unroll
    for (int k = 0; k < ZT; ++k)
      if (k < zcount) exist_mask |= ((1u << VEC) - 1u) << (k * VEC);
  }

  const int64_t plane = (int64_t)P.dim_y * P.dim_x;
  const int64_t idx0 = (int64_t)zl0 * plane + (int64_t)y * P.dim_x + (int64_t)xu * VEC;

  float T[ZT][VEC], W[ZT][VEC];
  bool loaded = false;     // warp-uniform
  unsigned dirty = 0;      // bit k*VEC+j: voxel changed
  constexpr bool use_fast = FAST;  // the launcher falls back to FAST = false when trunc_margin <= 0

  for (unsigned fm = frame_mask; fm != 0u; fm &= fm - 1u) {
    const int f = __ffs(fm) - 1;
    const FramePose& ps = DEVPOSE ? dev_pose : P.pose[f];
    const float* __restrict__ depth = P.depth + (int64_t)f * P.frame_stride;

    // (M0*t0 + M4*t1) per x, M8*t2 per z  -- demo.cu:35-41
    float sx[VEC], sy[VEC], sz[VEC], zx[ZT], zy[ZT], zz[ZT];
    {
      const float t1 = __fsub_rn(base_y, ps.ty);
      const float yx = __fmul_rn(ps.m4, t1), yy = __fmul_rn(ps.m5, t1), yz = __fmul_rn(ps.m6, t1);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float t0 = __fsub_rn(base_x[j], ps.tx);
        sx[j] = __fadd_rn(__fmul_rn(ps.m0, t0), yx);
        sy[j] = __fadd_rn(__fmul_rn(ps.m1, t0), yy);
        sz[j] = __fadd_rn(__fmul_rn(ps.m2, t0), yz);
      }
#pragma unroll
      for (int k = 0; k < ZT; ++k) {
        const float t2 = __fsub_rn(base_z[k], ps.tz);
        zx[k] = __fmul_rn(ps.m8, t2);
        zy[k] = __fmul_rn(ps.m9, t2);
        zz[k] = __fmul_rn(ps.m10, t2);
      }
    }

#pragma unroll
    for (int k = 0; k < ZT; ++k) {
      float dist[VEC];
      unsigned hit = 0;
      if (active) {
        unsigned exact = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          dist[j] = 1.0f;
          bool free_space = false, need_exact = true;
          if (use_fast) {
            if (P.sanitized)
              voxel_classify<true>(__fadd_rn(sx[j], zx[k]), __fadd_rn(sy[j], zy[k]), __fadd_rn(sz[j], zz[k]), K,
                                   depth, P.im_w, P.im_h, P.trunc_margin, free_space, need_exact);
            else
              voxel_classify<false>(__fadd_rn(sx[j], zx[k]), __fadd_rn(sy[j], zy[k]), __fadd_rn(sz[j], zz[k]), K,
                                    depth, P.im_w, P.im_h, P.trunc_margin, free_space, need_exact);
          }
          hit |= (free_space ? 1u : 0u) << j;
          exact |= (need_exact ? 1u : 0u) << j;
        }
        if (exact != 0u) {
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (exact & (1u << j)) {
              const float cam_x = __fadd_rn(sx[j], zx[k]), cam_y = __fadd_rn(sy[j], zy[k]),
                          cam_z = __fadd_rn(sz[j], zz[k]);
              // (the frame may be sanitized: -inf = rejected pixel, +inf = NaN pixel)
              const float dj = P.sanitized
                  ? voxel_observation_exact<use_fast, true>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy, depth, P.im_w,
                                                            P.im_h, P.trunc_margin)
                  : voxel_observation_exact<use_fast, false>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy, depth,
                                                             P.im_w, P.im_h, P.trunc_margin);
              if (dj != kNoObservation) {
                dist[j] = dj;
                hit |= 1u << j;
              }
            }
        }
        hit &= exist_mask >> (k * VEC);
      }
      if (__any_sync(0xffffffffu, hit != 0u)) {
        if (!loaded) {
          loaded = true;
#pragma unroll
          for (int kk = 0; kk < ZT; ++kk)
            if (active && kk < zcount) {
              Row<VEC>::load(P.tsdf + idx0 + kk * plane, T[kk]);
              Row<VEC>::load(P.weight + idx0 + kk * plane, W[kk]);
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (hit & (1u << j)) voxel_update(T[k][j], W[k][j], dist[j]);
        dirty |= hit << (k * VEC);
      }
    }
  }

  // ---- store the planes that changed (warp-uniform decision -> full coalesced lines) ----
#pragma unroll
  for (int k = 0; k < ZT; ++k) {
    const unsigned plane_bits = ((1u << VEC) - 1u) << (k * VEC);
    if (__any_sync(0xffffffffu, (dirty & plane_bits) != 0u)) {
      if (active && k < zcount) {
        Row<VEC>::store(P.tsdf + idx0 + k * plane, T[k]);
        Row<VEC>::store(P.weight + idx0 + k * plane, W[k]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// integrate_pending_kernel: the production K-frame kernel.
//
// Same thread -> voxel mapping, frame masks and classification as integrate_kernel above, but the
// per-voxel state carried across the K frames is not TSDF/weight themselves: it is the COUNT of
// pending free-space observations (dist == 1), 8 bits per voxel, ZT registers per thread.  This is
// exact because
//   * a free-space observation of a voxel with t == 1 leaves t == 1 and adds 1.0f to w
//     (demo.cu:63-66: (1*w + 1)/(w + 1) == 1), and c such additions equal w + (float)c exactly
//     (w is integer-valued and < 2^24);
//   * observations inside the +-trunc band (dist < 1, a few percent) are applied immediately, in frame
//     order, by the owning thread: pending count flushed, then (t*w + dist)/(w + 1), read-modify-write
//     straight in global memory (L2-resident across the K frames);
//   * at the end of the pass the remaining counts are applied with coalesced float4 accesses (for
//     t != 1 by replaying the c updates one by one).
// Every voxel therefore sees its updates in ascending frame order with the reference's arithmetic,
// while the inner loop carries 4 state registers instead of 32, needs no loads at all for voxels that
// are only ever seen as free space until the final write-back, and the rare exact evaluations are
// compacted per warp (each lane pops its own pending exact voxels) instead of being issued slot by slot.
// ------------------------------------------------------------------------------------------------

// Pre-pass of a K-frame launch: the frame masks of every warp tile of the slab, one thread per (tile, frame).
// Inside the fusion kernel this test is a long dependent chain executed by half-empty warps while their block's
// other warps wait; here it runs at full occupancy over all tiles at once (~0.3 % of the fusion kernel's
// instructions) and the fusion warps start with one 8-byte load.  out[tile] = {frame mask, interior mask},
// tile = (z group * n_yt + y tile) * n_xt + x tile, the numbering integrate_pending_kernel derives from its block
// and warp indices.
template <int VEC, int ZT>
__global__ void __launch_bounds__(256)
tile_mask_kernel(const __grid_constant__ IntegrateParams P, int n_xt, int n_yt, int64_t n_tiles, int nf_shift) {
  Intrinsics K;
  K.fx = P.fx; K.fy = P.fy; K.cx = P.cx; K.cy = P.cy; K.lim_x = P.lim_x; K.lim_y = P.lim_y;
  const unsigned lane = threadIdx.x & 31u;
  const int f = (int)(lane & ((1u << nf_shift) - 1u));
  const unsigned sub = lane >> nf_shift;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t tile = warp * (32 >> nf_shift) + sub;
  bool keep = false, inside = false;
  if (tile < n_tiles && f < P.n_frames) {
    const int xt = (int)(tile % n_xt);
    const int64_t r = tile / n_xt;
    const int yt = (int)(r % n_yt), zg = (int)(r / n_yt);
    const int rows_per_warp = 32 >> P.lx_shift;
    tile_frame_test<VEC, ZT>(P, K, f, (xt << P.lx_shift) * VEC, yt * rows_per_warp, rows_per_warp, P.z_begin + zg * ZT, keep,
                             inside);
  }
  const unsigned bk = __ballot_sync(0xffffffffu, keep), bi = __ballot_sync(0xffffffffu, inside);
  if (f == 0 && tile < n_tiles) {
    const unsigned sel = (nf_shift == 5) ? 0xffffffffu : ((1u << (1u << nf_shift)) - 1u);
    P.tile_masks[tile] = make_uint2((bk >> (sub << nf_shift)) & sel, (bi >> (sub << nf_shift)) & sel);
  }
}

constexpr int kExactQueueCap = 256;  // exact-evaluation queue entries per warp
constexpr int kDeferredQueueCap = 128;  // deferred (multi-frame) queue entries per warp
constexpr int kBlockQueueCap = 2048;    // block-shared queue entries (QUEUE == 3)

template <int VEC, int ZT, bool DEVPOSE, int MINB, bool SANITIZED, int PTX_LEVEL, int ROLL, int QUEUE, bool PREMASK = false>
__global__ void __launch_bounds__(kBlockThreads, MINB)
integrate_pending_kernel(const __grid_constant__ IntegrateParams P, const float* __restrict__ d_cam_K,
                         const float* __restrict__ d_cam2base) {
  static_assert(VEC == 4 || VEC == 1, "VEC");
  static_assert(VEC * ZT <= 32, "one bit of the exact mask per voxel of the patch");
  // PTX_LEVEL: 0 = C++ classification, 1 = scalar PTX block, 2 = packed (f32x2) PTX block for voxel pairs
  constexpr bool USE_PTX = PTX_LEVEL >= 1;
  constexpr bool kPacked = SANITIZED && PTX_LEVEL >= 2 && VEC == 4 && ROLL > 0;
  const int lx_mask = (1 << P.lx_shift) - 1;
  const int lx = threadIdx.x & lx_mask;
  const int ly = threadIdx.x >> P.lx_shift;
  const int rows_per_block = kBlockThreads >> P.lx_shift;
  const int xu = (blockIdx.x << P.lx_shift) + lx;
  const int y = blockIdx.y * rows_per_block + ly;
  // z groups are dispatched heaviest first (far planes see more of the frustum) to shorten the tail
  const int zl0 = (P.z_reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z) * ZT;
  const bool active = (xu < P.nx_units) && (y < P.dim_y);
  const int zcount = min(ZT, P.nz - zl0);

  Intrinsics K;
  FramePose dev_pose;
  if (DEVPOSE) {
    K.fx = __ldg(d_cam_K + 0); K.cx = __ldg(d_cam_K + 2); K.fy = __ldg(d_cam_K + 4); K.cy = __ldg(d_cam_K + 5);
    K.lim_x = 0.5f - pixel_round_margin((float)P.im_w, K.cx);
    K.lim_y = 0.5f - pixel_round_margin((float)P.im_h, K.cy);
    dev_pose.m0 = __ldg(d_cam2base + 0); dev_pose.m1 = __ldg(d_cam2base + 1); dev_pose.m2 = __ldg(d_cam2base + 2);
    dev_pose.m4 = __ldg(d_cam2base + 4); dev_pose.m5 = __ldg(d_cam2base + 5); dev_pose.m6 = __ldg(d_cam2base + 6);
    dev_pose.m8 = __ldg(d_cam2base + 8); dev_pose.m9 = __ldg(d_cam2base + 9); dev_pose.m10 = __ldg(d_cam2base + 10);
    dev_pose.tx = __ldg(d_cam2base + 3); dev_pose.ty = __ldg(d_cam2base + 7); dev_pose.tz = __ldg(d_cam2base + 11);
  } else {
    K.fx = P.fx; K.fy = P.fy; K.cx = P.cx; K.cy = P.cy;
    K.lim_x = P.lim_x; K.lim_y = P.lim_y;
  }

  // ---- packed variant: the plane products M8*t2 (demo.cu:32,37,39-41) of every (frame, plane) of this block,
  //      formed once per block with the reference's operations and broadcast from shared memory in the loop
  __shared__ float4 s_ztab[kPacked ? kMaxBatch * ZT : 1];
  constexpr bool kBlockQ = QUEUE == 3 && ROLL > 0;
  auto fill_ztab = [&](unsigned frames) {
    if constexpr (kPacked) {
      const int nf = DEVPOSE ? 1 : P.n_frames;
      for (int e = threadIdx.x; e < nf * ZT; e += kBlockThreads) {
        const int f = e / ZT, k = e - f * ZT;
        if (!((frames >> f) & 1u)) continue;
        const FramePose& ps = DEVPOSE ? dev_pose : P.pose[f];
        const float t2 = __fsub_rn(__fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size)), ps.tz);
        s_ztab[e] = make_float4(__fmul_rn(ps.m8, t2), __fmul_rn(ps.m9, t2), __fmul_rn(ps.m10, t2), 0.0f);
      }
      __syncthreads();
    }
  };
  if constexpr (!kBlockQ) fill_ztab(0xffffffffu);  // (before any warp of the block may leave)

  // ---- per-warp frame mask (see integrate_kernel): read from the pre-pass (tile_mask_kernel) when the launcher
  //      ran it, else computed here with lane f testing frame f ----
  unsigned frame_mask = (P.n_frames >= 32) ? 0xffffffffu : ((1u << P.n_frames) - 1u);
  unsigned interior_mask = 0u;  // frames for which the whole tile projects well inside the image
  if (!DEVPOSE && P.cull_level >= 1) {
    const int rows_per_warp = 32 >> P.lx_shift;
    if constexpr (PREMASK) {
      const int64_t tile = ((int64_t)(zl0 / ZT) * (gridDim.y * (kBlockThreads / 32)) +
                            (blockIdx.y * (kBlockThreads / 32) + (threadIdx.x >> 5))) * gridDim.x + blockIdx.x;
      const uint2 m = __ldg(P.tile_masks + tile);
      frame_mask = m.x;
      interior_mask = m.y;
    } else {
      bool keep = false, inside = false;
      const int f = threadIdx.x & 31;
      if (f < P.n_frames) {
        const int x_lo = (blockIdx.x << P.lx_shift) * VEC;
        const int y_lo = blockIdx.y * rows_per_block + (threadIdx.x >> 5) * rows_per_warp;
        tile_frame_test<VEC, ZT>(P, K, f, x_lo, y_lo, rows_per_warp, P.z_begin + zl0, keep, inside);
      }
      frame_mask = __ballot_sync(0xffffffffu, keep);
      interior_mask = __ballot_sync(0xffffffffu, inside);
    }
  }
  // QUEUE == 3: the exact-evaluation queue is shared by the block, so its 8 warps walk the UNION of their frame
  // masks in lockstep (two barriers per frame); a warp whose own mask lacks the frame only helps with the queue
  __shared__ unsigned s_wmask[kBlockThreads / 32];
  __shared__ unsigned s_bq_count[2], s_bq_valid[2];
  __shared__ unsigned short s_bq[kBlockQ ? kBlockQueueCap : 1];
  unsigned loop_mask = frame_mask;
  if constexpr (kBlockQ) {
    if ((threadIdx.x & 31) == 0) s_wmask[threadIdx.x >> 5] = frame_mask;
    if (threadIdx.x == 0) { s_bq_count[0] = s_bq_count[1] = 0u; s_bq_valid[0] = s_bq_valid[1] = 0u; }
    __syncthreads();
    loop_mask = 0u;
#pragma unroll
    for (int w = 0; w < kBlockThreads / 32; ++w) loop_mask |= s_wmask[w];
    if (loop_mask == 0u) return;  // block-uniform: no frame reaches this block
    fill_ztab(loop_mask);
  } else {
    if (frame_mask == 0u) return;  // warp-uniform
  }

  const float base_y = __fadd_rn(P.oy, __fmul_rn((float)y, P.voxel_size));  // demo.cu:31
  const int64_t plane = (int64_t)P.dim_y * P.dim_x;
  const int64_t idx0 = (int64_t)zl0 * plane + (int64_t)y * P.dim_x + (int64_t)xu * VEC;
  // byte j of pend(k): pending free-space updates of voxel (j, k).  ROLL == 0: ZT registers and fully
  // unrolled plane loops (32 inlined voxel bodies, ~20 KB of code per frame iteration).  ROLL > 0: the
  // counts live in shared memory, which can be indexed dynamically, so the plane loop is unrolled only ROLL
  // times and the hot loop fits the instruction cache.
  constexpr bool kRolled = ROLL > 0;
  __shared__ unsigned s_pend[kRolled ? ZT : 1][kRolled ? kBlockThreads : 1];
  __shared__ unsigned short s_queue[(QUEUE == 1 && kRolled) ? kBlockThreads / 32 : 1][(QUEUE == 1 && kRolled) ? kExactQueueCap : 1];
  // QUEUE == 2: deferred queue.  An entry is (owner lane, voxel slot, frame, snapshot of the voxel's pending count);
  // entries of several frames are evaluated together, 32 at a time, so the lanes are full (see drain_deferred)
  constexpr bool kDeferred = QUEUE == 2 && kRolled;
  __shared__ unsigned s_dq[kDeferred ? kBlockThreads / 32 : 1][kDeferred ? kDeferredQueueCap : 1];
  unsigned dq_count = 0u;  // warp-uniform
  // Evaluates and applies every queued entry.  Evaluation (voxel_observation_exact) does not depend on the voxel's
  // state, so the 32 entries of a round are computed in parallel whatever their frames; the APPLICATION is ordered:
  // entries were appended in frame order, entries of one voxel that share a round are applied in lane (= queue)
  // order (match.any groups, one sub-step per rank), rounds follow each other.  Every entry first flushes the
  // pending free-space count snapshot taken when it was queued -- the observations older than the entry -- so each
  // voxel still sees all its updates in ascending frame order (demo.cu:63-66).
  auto drain_deferred = [&]() {
    if constexpr (kDeferred) {
      const unsigned lane = threadIdx.x & 31u;
      const unsigned* q = s_dq[threadIdx.x >> 5];
      __syncwarp();
      for (unsigned base = 0; base < dq_count; base += 32u) {
        const unsigned i = base + lane;
        const bool valid = i < dq_count;
        float dj = kNoObservation;
        unsigned c = 0u, key = 0x80000000u | lane;  // invalid lanes: unique keys
        int64_t idx = 0;
        if (valid) {
          const unsigned item = q[i];
          key = item & 0x3ffu;
          c = (item >> 16) & 0xffu;
          const int fr = DEVPOSE ? 0 : (int)((item >> 10) & 31u);
          const int owner_t = (int)((threadIdx.x & ~31u) | (item & 31u));
          const int s = (int)((item >> 5) & 31u), j = s & (VEC - 1), k = s / VEC;
          const int o_xu = (blockIdx.x << P.lx_shift) + (owner_t & lx_mask);
          const int o_y = blockIdx.y * rows_per_block + (owner_t >> P.lx_shift);
          FramePose pq;
          if (DEVPOSE) pq = dev_pose; else pq = P.pose[fr];
          // the owner's arithmetic, operation for operation (demo.cu:30-41)
          const float q0 = __fsub_rn(__fadd_rn(P.ox, __fmul_rn((float)(o_xu * VEC + j), P.voxel_size)), pq.tx);
          const float q1 = __fsub_rn(__fadd_rn(P.oy, __fmul_rn((float)o_y, P.voxel_size)), pq.ty);
          const float q2 = __fsub_rn(__fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size)), pq.tz);
          const float cam_x = __fadd_rn(__fadd_rn(__fmul_rn(pq.m0, q0), __fmul_rn(pq.m4, q1)), __fmul_rn(pq.m8, q2));
          const float cam_y = __fadd_rn(__fadd_rn(__fmul_rn(pq.m1, q0), __fmul_rn(pq.m5, q1)), __fmul_rn(pq.m9, q2));
          const float cam_z = __fadd_rn(__fadd_rn(__fmul_rn(pq.m2, q0), __fmul_rn(pq.m6, q1)), __fmul_rn(pq.m10, q2));
          dj = voxel_observation_exact<true, SANITIZED>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy,
                                             P.depth + (int64_t)fr * P.frame_stride, P.im_w, P.im_h, P.trunc_margin);
          idx = ((int64_t)(zl0 + k) * P.dim_y + o_y) * (int64_t)P.dim_x + (int64_t)o_xu * VEC + j;
        }
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        const unsigned rank = __popc(grp & ((1u << lane) - 1u));
        const bool work = valid && (dj != kNoObservation || c != 0u);
        for (unsigned r = 0; __any_sync(0xffffffffu, work && rank >= r); ++r) {
          if (work && rank == r) {
            float t = P.tsdf[idx], w = P.weight[idx];
            apply_pending(t, w, c);
            if (dj != kNoObservation) voxel_update(t, w, dj);
            P.tsdf[idx] = t;
            P.weight[idx] = w;
          }
          __syncwarp();
        }
      }
      dq_count = 0u;
      __syncwarp();
    }
  };
  unsigned pend_reg[kRolled ? 1 : ZT];
  if constexpr (kRolled) {
#pragma unroll
    for (int k = 0; k < ZT; ++k) s_pend[k][threadIdx.x] = 0u;
  } else {
#pragma unroll
    for (int k = 0; k < ZT; ++k) pend_reg[k] = 0u;
  }

  unsigned bq_parity = 0u;
  for (unsigned fm = loop_mask; fm != 0u; fm &= fm - 1u) {
    const int f = __ffs(fm) - 1;
    const FramePose& ps = DEVPOSE ? dev_pose : P.pose[f];
    const float* __restrict__ depth = P.depth + (int64_t)f * P.frame_stride;
    const bool warp_sees_frame = !kBlockQ || ((frame_mask >> f) & 1u);  // warp-uniform

    // demo.cu:30-41: (M0*t0 + M4*t1) per x, M8*t2 per z
    const float t1 = __fsub_rn(base_y, ps.ty);
    const float yx = __fmul_rn(ps.m4, t1), yy = __fmul_rn(ps.m5, t1), yz = __fmul_rn(ps.m6, t1);
    float sx[VEC], sy[VEC], sz[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float t0 = __fsub_rn(__fadd_rn(P.ox, __fmul_rn((float)(xu * VEC + j), P.voxel_size)), ps.tx);
      sx[j] = __fadd_rn(__fmul_rn(ps.m0, t0), yx);
      sy[j] = __fadd_rn(__fmul_rn(ps.m1, t0), yy);
      sz[j] = __fadd_rn(__fmul_rn(ps.m2, t0), yz);
    }

    unsigned exact = 0u;  // bit k*VEC + j: voxel needs voxel_observation_exact
    if (active && warp_sees_frame) {
      const bool tile_interior =
          SANITIZED && USE_PTX && kRolled && !P.no_interior && ((interior_mask >> f) & 1u);  // warp-uniform
      // packed variant: the voxel pairs (0,1) and (2,3) as f32x2 operands; gather base biased by the bit
      // pattern of the rint magic (see voxel_classify_pair)
      f32x2_t sx2[2], sy2[2], sz2[2];
      const char* depth_biased = reinterpret_cast<const char*>(depth) - 4ll * (long long)kRintMagicBits;
      const float im_w_f = (float)P.im_w, half_w = 0.5f * (float)(P.im_w - 1), half_h = 0.5f * (float)(P.im_h - 1);
      if constexpr (kPacked) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          sx2[h] = pack2(sx[2 * h], sx[2 * h + 1]);
          sy2[h] = pack2(sy[2 * h], sy[2 * h + 1]);
          sz2[h] = pack2(sz[2 * h], sz[2 * h + 1]);
        }
      }
      auto plane_body = [&](int k, unsigned& pk, auto interior_tag) {
        constexpr bool kInterior = decltype(interior_tag)::value;
        if constexpr (kPacked) {
          const float4 zt = s_ztab[(DEVPOSE ? 0 : f) * ZT + k];
          unsigned ex = 0u;
#pragma unroll
          for (int h = 0; h < 2; ++h)
            voxel_classify_pair<kInterior>(sx2[h], sy2[h], sz2[h], zt.x, zt.y, zt.z, K, im_w_f, half_w, half_h,
                                           depth_biased, P.trunc_margin, 1u << (16 * h), 1u << (16 * h + 8),
                                           1u << (2 * h), 1u << (2 * h + 1), pk, ex);
          exact |= ex << (k * VEC);
        } else {
        const float t2 = __fsub_rn(__fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size)), ps.tz);
        const float zx = __fmul_rn(ps.m8, t2), zy = __fmul_rn(ps.m9, t2), zz = __fmul_rn(ps.m10, t2);
        unsigned ex = 0u;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if constexpr (SANITIZED && USE_PTX && kInterior) {
            voxel_classify_ptx_interior(__fadd_rn(sx[j], zx), __fadd_rn(sy[j], zy), __fadd_rn(sz[j], zz), K, depth,
                                        P.im_w, P.trunc_margin, 1u << (8 * j), 1u << j, pk, ex);
          } else if constexpr (SANITIZED && USE_PTX) {
            voxel_classify_ptx(__fadd_rn(sx[j], zx), __fadd_rn(sy[j], zy), __fadd_rn(sz[j], zz), K, depth, P.im_w,
                               P.im_h, P.trunc_margin, 1u << (8 * j), 1u << j, pk, ex);
          } else {
            bool free_space, need_exact;
            voxel_classify<SANITIZED>(__fadd_rn(sx[j], zx), __fadd_rn(sy[j], zy), __fadd_rn(sz[j], zz), K, depth,
                                      P.im_w, P.im_h, P.trunc_margin, free_space, need_exact);
            if (free_space) pk += 1u << (8 * j);
            if (need_exact) ex |= 1u << j;
          }
        }
        exact |= ex << (k * VEC);
        }
      };
      // block-uniform bound: only the topmost z group of a slab is ragged
      if constexpr (kRolled) {
        if (tile_interior) {
#pragma unroll ROLL
          for (int k = 0; k < zcount; ++k) {
            unsigned pk = s_pend[k][threadIdx.x];
            plane_body(k, pk, std::true_type{});
            s_pend[k][threadIdx.x] = pk;
          }
        } else {
#pragma unroll ROLL
          for (int k = 0; k < zcount; ++k) {
            unsigned pk = s_pend[k][threadIdx.x];
            plane_body(k, pk, std::false_type{});
            s_pend[k][threadIdx.x] = pk;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < ZT; ++k)
          if (k < zcount) plane_body(k, pend_reg[k], std::false_type{});
      }
    }

    // ---- exact evaluations, redistributed over the warp: the voxels that need the exact path are few
    //      and unevenly spread (a surface band crossing the tile hits 2-3 lanes with up to 32 voxels each),
    //      so they are queued in shared memory and handed out one per lane; results go back to the owner's
    //      pending counts through shared-memory atomics, in-band ones are applied by the lane that computed
    //      them.  Frame order per voxel is kept: the queue is drained before the next frame starts.
    if constexpr (kDeferred) {
      const unsigned lane = threadIdx.x & 31u;
      const unsigned n_mine = __popc(exact);
      const unsigned total = __reduce_add_sync(0xffffffffu, n_mine);
      if (total != 0u) {  // warp-uniform
        if (dq_count + total > (unsigned)kDeferredQueueCap) drain_deferred();
        if (total <= (unsigned)kDeferredQueueCap) {
          unsigned incl = n_mine;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += v;
          }
          unsigned* q = s_dq[threadIdx.x >> 5];
          unsigned pos = dq_count + incl - n_mine;
          while (exact != 0u) {
            const unsigned s = (unsigned)(__ffs(exact) - 1);
            exact &= exact - 1u;
            const unsigned sh = 8u * (s & (unsigned)(VEC - 1));
            // snapshot the voxel's pending count into the entry and restart it: free-space observations counted
            // from here on are younger than this entry
            unsigned* pw = &s_pend[s / VEC][threadIdx.x];
            const unsigned pk = *pw;
            *pw = pk & ~(0xffu << sh);
            q[pos++] = lane | (s << 5) | ((unsigned)f << 10) | (((pk >> sh) & 0xffu) << 16);
          }
          dq_count += total;
        }
        // (a frame with more exact voxels than the queue holds falls through to the per-lane loop below, after
        //  the drain above has applied everything older)
      }
    }
    if constexpr (kBlockQ) {
      // push: a lane reserves room for all its exact voxels with one shared-memory atomic (order inside a frame is
      // irrelevant: the entries of a frame are distinct voxels); lanes that do not fit keep `exact` and fall through
      // to the per-lane loop below
      const unsigned n_mine = __popc(exact);
      if (n_mine != 0u) {
        unsigned pos = atomicAdd(&s_bq_count[bq_parity], n_mine);
        if (pos + n_mine <= (unsigned)kBlockQueueCap) {
          atomicMax(&s_bq_valid[bq_parity], pos + n_mine);
          while (exact != 0u) {
            const unsigned s = (unsigned)(__ffs(exact) - 1);
            exact &= exact - 1u;
            s_bq[pos++] = (unsigned short)(threadIdx.x | (s << 8));
          }
        }
      }
      __syncthreads();
      const unsigned total = s_bq_valid[bq_parity];
      if (threadIdx.x == 0) { s_bq_count[bq_parity ^ 1u] = 0u; s_bq_valid[bq_parity ^ 1u] = 0u; }  // next frame's
      for (unsigned i = threadIdx.x; i < total; i += kBlockThreads) {
        const unsigned item = s_bq[i];
        const int owner_t = (int)(item & 0xffu);
        const int s = (int)(item >> 8), j = s & (VEC - 1), k = s / VEC;
        const int o_xu = (blockIdx.x << P.lx_shift) + (owner_t & lx_mask);
        const int o_y = blockIdx.y * rows_per_block + (owner_t >> P.lx_shift);
        // the owner's arithmetic, operation for operation (demo.cu:30-41)
        const float q0 = __fsub_rn(__fadd_rn(P.ox, __fmul_rn((float)(o_xu * VEC + j), P.voxel_size)), ps.tx);
        const float q1 = __fsub_rn(__fadd_rn(P.oy, __fmul_rn((float)o_y, P.voxel_size)), ps.ty);
        const float q2 = __fsub_rn(__fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size)), ps.tz);
        const float cam_x = __fadd_rn(__fadd_rn(__fmul_rn(ps.m0, q0), __fmul_rn(ps.m4, q1)), __fmul_rn(ps.m8, q2));
        const float cam_y = __fadd_rn(__fadd_rn(__fmul_rn(ps.m1, q0), __fmul_rn(ps.m5, q1)), __fmul_rn(ps.m9, q2));
        const float cam_z = __fadd_rn(__fadd_rn(__fmul_rn(ps.m2, q0), __fmul_rn(ps.m6, q1)), __fmul_rn(ps.m10, q2));
        const float dj = voxel_observation_exact<true, SANITIZED>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy, depth, P.im_w,
                                                       P.im_h, P.trunc_margin);
        if (dj != kNoObservation) {
          unsigned* pw = &s_pend[k][owner_t];
          if (dj == 1.0f) {
            atomicAdd(pw, 1u << (8 * j));
          } else {
            const unsigned old = atomicAnd(pw, ~(0xffu << (8 * j)));
            const int64_t idx = ((int64_t)(zl0 + k) * P.dim_y + o_y) * (int64_t)P.dim_x + (int64_t)o_xu * VEC + j;
            float t = P.tsdf[idx], w = P.weight[idx];
            apply_pending(t, w, (old >> (8 * j)) & 0xffu);
            const float w_new = __fadd_rn(w, 1.0f);
            t = __fdiv_rn(__fadd_rn(__fmul_rn(t, w), dj), w_new);
            P.tsdf[idx] = t;
            P.weight[idx] = w_new;
          }
        }
      }
      __syncthreads();  // every result is in s_pend / global memory before the next frame's classification
      bq_parity ^= 1u;
    }
    if constexpr (QUEUE == 1 && kRolled) {
      const unsigned lane = threadIdx.x & 31u;
      const unsigned n_mine = __popc(exact);
      unsigned incl = n_mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
      }
      const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
      if (total != 0u && total <= (unsigned)kExactQueueCap) {  // warp-uniform
        unsigned short* q = s_queue[threadIdx.x >> 5];
        unsigned pos = incl - n_mine;
        while (exact != 0u) {
          const unsigned s = (unsigned)(__ffs(exact) - 1);
          exact &= exact - 1u;
          q[pos++] = (unsigned short)(lane | (s << 5));
        }
        __syncwarp();
        for (unsigned base = 0; base < total; base += 32u) {
          const unsigned i = base + lane;
          if (i < total) {
            const unsigned item = q[i];
            const int owner_t = (int)((threadIdx.x & ~31u) | (item & 31u));
            const int s = (int)(item >> 5), j = s & (VEC - 1), k = s / VEC;
            const int o_xu = (blockIdx.x << P.lx_shift) + (owner_t & lx_mask);
            const int o_y = blockIdx.y * rows_per_block + (owner_t >> P.lx_shift);
            // the owner's arithmetic, operation for operation (demo.cu:30-41)
            const float q0 = __fsub_rn(__fadd_rn(P.ox, __fmul_rn((float)(o_xu * VEC + j), P.voxel_size)), ps.tx);
            const float q1 = __fsub_rn(__fadd_rn(P.oy, __fmul_rn((float)o_y, P.voxel_size)), ps.ty);
            const float q2 =
                __fsub_rn(__fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size)), ps.tz);
            const float cam_x =
                __fadd_rn(__fadd_rn(__fmul_rn(ps.m0, q0), __fmul_rn(ps.m4, q1)), __fmul_rn(ps.m8, q2));
            const float cam_y =
                __fadd_rn(__fadd_rn(__fmul_rn(ps.m1, q0), __fmul_rn(ps.m5, q1)), __fmul_rn(ps.m9, q2));
            const float cam_z =
                __fadd_rn(__fadd_rn(__fmul_rn(ps.m2, q0), __fmul_rn(ps.m6, q1)), __fmul_rn(ps.m10, q2));
            const float dj = voxel_observation_exact<true, SANITIZED>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy, depth,
                                                           P.im_w, P.im_h, P.trunc_margin);
            if (dj != kNoObservation) {
              unsigned* pw = &s_pend[k][owner_t];
              if (dj == 1.0f) {
                atomicAdd(pw, 1u << (8 * j));
              } else {
                const unsigned old = atomicAnd(pw, ~(0xffu << (8 * j)));
                const int64_t idx =
                    ((int64_t)(zl0 + k) * P.dim_y + o_y) * (int64_t)P.dim_x + (int64_t)o_xu * VEC + j;
                float t = P.tsdf[idx], w = P.weight[idx];
                apply_pending(t, w, (old >> (8 * j)) & 0xffu);
                const float w_new = __fadd_rn(w, 1.0f);
                t = __fdiv_rn(__fadd_rn(__fmul_rn(t, w), dj), w_new);
                P.tsdf[idx] = t;
                P.weight[idx] = w_new;
              }
            }
          }
        }
        __syncwarp();
      }
    }

    // ---- (remaining) exact evaluations: every lane pops its own pending voxels ----
    while (__any_sync(0xffffffffu, exact != 0u)) {
      if (exact != 0u) {
        const int s = __ffs(exact) - 1;
        exact &= exact - 1u;
        const int j = s & (VEC - 1), k = s / VEC;
        const float t0 = __fsub_rn(__fadd_rn(P.ox, __fmul_rn((float)(xu * VEC + j), P.voxel_size)), ps.tx);
        const float t2 = __fsub_rn(__fadd_rn(P.oz, __fmul_rn((float)(P.z_begin + zl0 + k), P.voxel_size)), ps.tz);
        const float cam_x = __fadd_rn(__fadd_rn(__fmul_rn(ps.m0, t0), yx), __fmul_rn(ps.m8, t2));
        const float cam_y = __fadd_rn(__fadd_rn(__fmul_rn(ps.m1, t0), yy), __fmul_rn(ps.m9, t2));
        const float cam_z = __fadd_rn(__fadd_rn(__fmul_rn(ps.m2, t0), yz), __fmul_rn(ps.m10, t2));
        const float dj = voxel_observation_exact<true, SANITIZED>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy, depth, P.im_w,
                                                       P.im_h, P.trunc_margin);
        if (dj != kNoObservation) {
          unsigned c = 0u;
          if constexpr (kRolled) {
            unsigned pk = s_pend[k][threadIdx.x];
            if (dj == 1.0f) {
              pk += 1u << (8 * j);
            } else {
              c = (pk >> (8 * j)) & 0xffu;
              pk &= ~(0xffu << (8 * j));
            }
            s_pend[k][threadIdx.x] = pk;
          } else {
#pragma unroll
            for (int kk = 0; kk < ZT; ++kk)
              if (kk == k) {
                if (dj == 1.0f) {
                  pend_reg[kk] += 1u << (8 * j);
                } else {
                  c = (pend_reg[kk] >> (8 * j)) & 0xffu;
                  pend_reg[kk] &= ~(0xffu << (8 * j));
                }
              }
          }
          if (dj != 1.0f) {
            // in-band observation: flush the voxel's pending free-space updates, then demo.cu:63-66
            const int64_t idx = idx0 + (int64_t)k * plane + j;
            float t = P.tsdf[idx], w = P.weight[idx];
            apply_pending(t, w, c);
            const float w_new = __fadd_rn(w, 1.0f);
            t = __fdiv_rn(__fadd_rn(__fmul_rn(t, w), dj), w_new);
            P.tsdf[idx] = t;
            P.weight[idx] = w_new;
          }
        }
      }
    }
  }

  drain_deferred();

  // ---- write-back of the pending free-space counts: coalesced, warp-uniform per plane ----
  auto write_back = [&](int k, unsigned pk) {
    if (__any_sync(0xffffffffu, pk != 0u)) {
      if (active && k < zcount) {
        float T[VEC], W[VEC];
        Row<VEC>::load(P.tsdf + idx0 + k * plane, T);
        Row<VEC>::load(P.weight + idx0 + k * plane, W);
#pragma unroll
        for (int j = 0; j < VEC; ++j) apply_pending(T[j], W[j], (pk >> (8 * j)) & 0xffu);
        Row<VEC>::store(P.tsdf + idx0 + k * plane, T);
        Row<VEC>::store(P.weight + idx0 + k * plane, W);
      }
    }
  };
  if constexpr (kRolled) {
#pragma unroll 1
    for (int k = 0; k < ZT; ++k) write_back(k, s_pend[k][threadIdx.x]);
  } else {
#pragma unroll
    for (int k = 0; k < ZT; ++k) write_back(k, pend_reg[k]);
  }
}

// ------------------------------------------------------------------------------------------------
// fill: TSDF <- 1.0f, weight <- 0 (demo.cu:131-135) directly in HBM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ tsdf, float* __restrict__ weight, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    __stcs(reinterpret_cast<float4*>(tsdf) + i, one);
    __stcs(reinterpret_cast<float4*>(weight) + i, zero);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    tsdf[i] = 1.f;
    weight[i] = 0.f;
  }
}

// Rewrites a depth frame for the SANITIZED kernels: pixels the reference rejects (demo.cu:53: d <= 0 or d > 6)
// become -inf; NaN pixels, which the reference does not reject and which end as a dist = 1 observation
// (fmin(1.0f, NaN), demo.cu:62), become +inf (e = +inf - cam_z >= trunc: counted as free space); everything else is
// copied unchanged.
__global__ void __launch_bounds__(256) sanitize_depth_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = in[i];
    out[i] = (d > 0.0f && d <= 6.0f) ? d : ((d != d) ? INFINITY : -INFINITY);
  }
}

// Max-pyramid of sanitized depth frames for the occlusion test (CullAccum::occluded): levels 3..7, i.e. the
// maximum over aligned blocks of 8, 16, 32, 64 and 128 pixels (pixels outside the image and invalid pixels
// count as -inf).  One 256-thread block reduces one 128 x 128 pixel region of one frame.
__global__ void __launch_bounds__(256) build_depth_mips_kernel(const float* __restrict__ frames, float* __restrict__ mips,
                                                               int im_w, int im_h, int64_t frame_stride, MipLayout L) {
  __shared__ float s[256];
  const int frame = blockIdx.z;
  const float* src = frames + (int64_t)frame * frame_stride;
  float* dst = mips + (int64_t)frame * L.stride;
  const int t = threadIdx.x;
  // level 3: thread t owns the 8 x 8 block (bx, by) of the 16 x 16 blocks of this region; bits of t are
  // interleaved (Morton order) so that every further 2 x 2 reduction combines 4 consecutive entries
  int bx = 0, by = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) { bx |= ((t >> (2 * b)) & 1) << b; by |= ((t >> (2 * b + 1)) & 1) << b; }
  const int px0 = blockIdx.x * 128 + bx * 8, py0 = blockIdx.y * 128 + by * 8;
  float m = -INFINITY;
  for (int yy = 0; yy < 8; ++yy) {
    const int py = py0 + yy;
    if (py >= im_h) break;
    for (int xx = 0; xx < 8; ++xx) {
      const int px = px0 + xx;
      if (px < im_w) m = fmaxf(m, src[(int64_t)py * im_w + px]);
    }
  }
  s[t] = m;
  if (px0 < im_w && py0 < im_h) dst[L.off[0] + (py0 >> 3) * L.w[0] + (px0 >> 3)] = m;
  __syncthreads();
#pragma unroll
  for (int level = 1; level < 5; ++level) {
    const int n = 256 >> (2 * level);      // entries of this level inside the region
    float v = -INFINITY;
    if (t < n) v = fmaxf(fmaxf(s[4 * t], s[4 * t + 1]), fmaxf(s[4 * t + 2], s[4 * t + 3]));
    __syncthreads();
    if (t < n) {
      s[t] = v;
      int cx = 0, cy = 0;                  // de-interleave t -> block coordinates at this level
      for (int b = 0; b < 4 - level; ++b) { cx |= ((t >> (2 * b)) & 1) << b; cy |= ((t >> (2 * b + 1)) & 1) << b; }
      const int gx = blockIdx.x * (16 >> level) + cx, gy = blockIdx.y * (16 >> level) + cy;
      if (gx < L.w[level] && gy < L.h[level]) dst[L.off[level] + gy * L.w[level] + gx] = v;
    }
    __syncthreads();
  }
}

// utils.hpp:82-84: (float)mm / 1000.0f, then > 6.0f -> 0.
__global__ void __launch_bounds__(256) depth_from_u16_kernel(const uint16_t* __restrict__ mm,
                                                             float* __restrict__ depth, int64_t n, bool sanitize) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = __fdiv_rn((float)mm[i], 1000.0f);
    if (d > 6.0f) d = 0.0f;
    if (sanitize && !(d > 0.0f)) d = -INFINITY;
    depth[i] = d;
  }
}

// ------------------------------------------------------------------------------------------------
// Surface extraction (utils.hpp:20-53) as an order-preserving stream compaction:
//   count : per-tile number of surface voxels (warp ballot + popc)
//   scan  : exclusive prefix sum of the tile counts (single block, 64-bit)
//   emit  : per-tile re-evaluation, rank = tile offset + warp prefix + popc(ballot & lanes below)
// A tile is kExtractTile consecutive voxels; in round r thread t looks at voxel
// tile*4096 + r*256 + t, so ranks ascend with the linear voxel index exactly like the
// reference's serial scan.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_surface(const ExtractParams& P, int64_t i) {
  // utils.hpp:22 / :39: std::abs(tsdf) < tsdf_thresh && weight > weight_thresh
  return (i < P.n) && (fabsf(__ldcs(P.tsdf + i)) < P.tsdf_thresh) && (__ldcs(P.weight + i) > P.weight_thresh);
}

constexpr int kExtractRounds = kExtractTile / kBlockThreads;  // 16

__global__ void __launch_bounds__(kBlockThreads)
extract_count_kernel(const __grid_constant__ ExtractParams P, unsigned long long* __restrict__ tile_counts,
                     int64_t n_tiles) {
  __shared__ unsigned s_warp[kBlockThreads / 32];
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kExtractTile + threadIdx.x;
    unsigned cnt = 0;
#pragma unroll
    for (int r = 0; r < kExtractRounds; ++r) cnt += is_surface(P, base + r * kBlockThreads) ? 1u : 0u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned total = 0;
#pragma unroll
      for (int w = 0; w < kBlockThreads / 32; ++w) total += s_warp[w];
      tile_counts[tile] = total;
    }
    __syncthreads();
  }
}

// In-place exclusive scan of tile_counts[0..n_tiles) by one block; *total receives the sum
// (the engine points it at tile_counts[n_tiles], giving emit an n_tiles + 1 offset table).
__global__ void __launch_bounds__(1024)
extract_scan_kernel(unsigned long long* __restrict__ tile_counts, int64_t n_tiles, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long s_part[1024];
  const int t = threadIdx.x;
  const int64_t per = (n_tiles + 1023) / 1024;
  const int64_t lo = min(n_tiles, (int64_t)t * per), hi = min(n_tiles, lo + per);
  unsigned long long sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += tile_counts[i];
  s_part[t] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 1024 partial sums
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned long long v = (t >= off) ? s_part[t - off] : 0ull;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  unsigned long long run = (t == 0) ? 0ull : s_part[t - 1];
  for (int64_t i = lo; i < hi; ++i) {
    const unsigned long long c = tile_counts[i];
    tile_counts[i] = run;
    run += c;
  }
  if (t == 1023) *total = s_part[1023];
}

__global__ void __launch_bounds__(kBlockThreads)
extract_emit_kernel(const __grid_constant__ ExtractParams P, const unsigned long long* __restrict__ tile_offsets,
                    int64_t n_tiles, float* __restrict__ xyz, long long* __restrict__ voxel_index) {
  __shared__ unsigned s_cnt[kExtractRounds][kBlockThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t plane = (int64_t)P.dim_x * P.dim_y;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // tiles with no surface voxel (the vast majority) are skipped without touching the volume;
    // tile_offsets has n_tiles + 1 entries, the last one being the total
    const unsigned long long tile_base = tile_offsets[tile];
    if (tile_offsets[tile + 1] == tile_base) continue;  // block-uniform
    const int64_t base = tile * kExtractTile + threadIdx.x;
    unsigned flags = 0;
    unsigned my_rank[kExtractRounds];
#pragma unroll
    for (int r = 0; r < kExtractRounds; ++r) {
      const bool s = is_surface(P, base + r * kBlockThreads);
      const unsigned b = __ballot_sync(0xffffffffu, s);
      if (s) flags |= 1u << r;
      my_rank[r] = __popc(b & ((1u << lane) - 1u));
      if (lane == 0) s_cnt[r][warp] = __popc(b);
    }
    __syncthreads();
    if (flags != 0u) {
      // prefix over (round, warp) pairs that precede this thread's (r, warp)
      unsigned run = 0;
#pragma unroll
      for (int r = 0; r < kExtractRounds; ++r) {
        unsigned before = 0, round_total = 0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) {
          const unsigned c = s_cnt[r][w];
          before += (w < warp) ? c : 0u;
          round_total += c;
        }
        if (flags & (1u << r)) {
          const int64_t i = base + r * kBlockThreads;
          // utils.hpp:42-49: linear index -> (x, y, z) -> origin + (float)idx * voxel_size
          const int64_t z = i / plane;
          const int64_t rem = i - z * plane;
          const int yy = (int)(rem / P.dim_x);
          const int xx = (int)(rem - (int64_t)yy * P.dim_x);
          const unsigned long long rank = tile_base + run + before + my_rank[r];
          if (voxel_index != nullptr) voxel_index[rank] = (long long)i;
          float* out = xyz + 3ull * rank;
          out[0] = __fadd_rn(P.ox, __fmul_rn((float)xx, P.voxel_size));
          out[1] = __fadd_rn(P.oy, __fmul_rn((float)yy, P.voxel_size));
          out[2] = __fadd_rn(P.oz, __fmul_rn((float)(z + P.z_begin), P.voxel_size));
        }
        run += round_total;
      }
    }
    __syncthreads();
  }
}

// Surface normals: the normalised central difference of the TSDF at a surface voxel, as
// PCL-benchmark/pclWrapperFPFH.h:308-330 (computeSurfaceNormal) computes it; mLib's getNormalized() is
// val = 1 / sqrt(x*x + y*y + z*z), (x*val, y*val, z*val).  (0, 0, 0) where a neighbour lies outside the slab or the
// gradient vanishes.
__global__ void __launch_bounds__(256) surface_normals_kernel(const __grid_constant__ ExtractParams P, int64_t nz,
                                                              const long long* __restrict__ voxel_index, int64_t n,
                                                              float* __restrict__ normals) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int64_t i = voxel_index[p];
  const int64_t plane = (int64_t)P.dim_x * P.dim_y;
  const int64_t z = i / plane, rem = i - z * plane;
  const int y = (int)(rem / P.dim_x), x = (int)(rem - (int64_t)y * P.dim_x);
  float nx = 0.0f, ny = 0.0f, nzv = 0.0f;
  if (x > 0 && x < P.dim_x - 1 && y > 0 && y < P.dim_y - 1 && z > 0 && z < nz - 1) {
    const float gx = __fsub_rn(__ldg(P.tsdf + i + 1), __ldg(P.tsdf + i - 1));
    const float gy = __fsub_rn(__ldg(P.tsdf + i + P.dim_x), __ldg(P.tsdf + i - P.dim_x));
    const float gz = __fsub_rn(__ldg(P.tsdf + i + plane), __ldg(P.tsdf + i - plane));
    if (!(gx == 0.0f && gy == 0.0f && gz == 0.0f)) {
      const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
      const float val = __fdiv_rn(1.0f, __fsqrt_rn(len2));
      nx = __fmul_rn(gx, val); ny = __fmul_rn(gy, val); nzv = __fmul_rn(gz, val);
    }
  }
  normals[3 * p + 0] = nx;
  normals[3 * p + 1] = ny;
  normals[3 * p + 2] = nzv;
}

// All 2^32 float bit patterns: round_pixel() vs roundf() + the reference's bounds test.
__global__ void __launch_bounds__(256) selftest_rounding_kernel(unsigned long long* __restrict__ mismatches) {
  const float sizes[3] = {640.0f, 480.0f, 1.0f};
  unsigned long long bad = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
    const float v = __uint_as_float((unsigned)b);
    const float r = roundf(v);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const bool ref_in = (r >= 0.0f) && (r < sizes[s]);
      int pix;
      const bool in = round_pixel(v, sizes[s], pix);
      if (in != ref_in || (in && pix != (int)r)) ++bad;
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

thread_local int g_integrate_launches = 0;  // kernels the last launch_integrate*() call enqueued

// threads of a warp along x: the power of two in {8, 16, 32} that wastes the fewest lanes.  8 is preferred: a warp
// then covers 32 x 4 x ZT voxels (128-byte row segments), which keeps thin surface bands and frustum boundaries
// warp-coherent (measured: 1.2x over 32 x 1 on the 1024^3 workload).
int choose_lx(int nx_units, int force_lx) {
  int best_lx = 8;
  int64_t best_waste = -1;
  for (int lx = 8; lx <= 32; lx <<= 1) {
    const int64_t covered = (int64_t)((nx_units + lx - 1) / lx) * lx;
    const int64_t waste = covered - nx_units;
    if (best_waste < 0 || waste < best_waste) { best_waste = waste; best_lx = lx; }
  }
  if (force_lx == 4 || force_lx == 8 || force_lx == 16 || force_lx == 32) best_lx = force_lx;  // tuning
  return best_lx;
}

int ilog2(int v) {
  int s = 0;
  while ((1 << (s + 1)) <= v) ++s;
  return s;
}

template <int VEC, int ZT, bool DEVPOSE, int MINB, bool FAST = true, bool PENDING = false, int ROLL = 0,
          int QUEUE = 0>
cudaError_t launch_integrate_t(IntegrateParams p, const float* d_K, const float* d_pose, cudaStream_t s) {
  p.nx_units = p.dim_x / VEC;
  p.lim_x = 0.5f - pixel_round_margin((float)p.im_w, p.cx);
  p.lim_y = 0.5f - pixel_round_margin((float)p.im_h, p.cy);
  const int best_lx = choose_lx(p.nx_units, p.force_lx);
  p.lx_shift = ilog2(best_lx);
  const int rows = kBlockThreads / best_lx;
  dim3 grid((unsigned)((p.nx_units + best_lx - 1) / best_lx), (unsigned)((p.dim_y + rows - 1) / rows),
            (unsigned)((p.nz + ZT - 1) / ZT));
  if (grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidConfiguration;
  g_integrate_launches = 1;
  if (FAST && (p.fast_level == 0 || !(p.trunc_margin > 0.0f)))
    return launch_integrate_t<VEC, ZT, DEVPOSE, MINB, false, false, 0, false>(p, d_K, d_pose, s);
  if constexpr (PENDING) {
    // the packed classification needs |M + iy*W + ix| < 2^24 for its floating-point gather index
    const bool can_pack = (int64_t)p.im_w * p.im_h < (1 << 22) && VEC == 4 && ROLL > 0;
    const bool packed = p.sanitized && !p.no_ptx && !p.no_packed && can_pack;
    // frame-mask pre-pass into the engine's scratch buffer (production kernel only; otherwise the kernel computes
    // its masks itself)
    bool premask = false;
    if constexpr (!DEVPOSE) {
      const int64_t n_xt = grid.x, n_yt = (int64_t)grid.y * (kBlockThreads / 32), n_tiles = n_xt * n_yt * grid.z;
      if (packed && p.tile_masks != nullptr && p.masks_ready && p.cull_level >= 1) {
        premask = true;  // launch_tile_masks() already filled them for the whole slab (engine, main stream)
      } else if (packed && p.tile_masks != nullptr && p.cull_level >= 1 && n_tiles <= p.tile_mask_cap) {
        int nf_shift = 0;
        while ((1 << nf_shift) < p.n_frames) ++nf_shift;
        const int64_t threads = ((n_tiles + (32 >> nf_shift) - 1) / (32 >> nf_shift)) * 32;
        tile_mask_kernel<VEC, ZT><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p, (int)n_xt, (int)n_yt, n_tiles, nf_shift);
        ++g_integrate_launches;
        premask = true;
      }
    }
    if (!premask) p.tile_masks = nullptr;
    if (packed && premask) {
      if constexpr (!DEVPOSE)
        integrate_pending_kernel<VEC, ZT, DEVPOSE, MINB, true, 2, ROLL, QUEUE, true><<<grid, kBlockThreads, 0, s>>>(p, d_K, d_pose);
    } else if (packed)
      integrate_pending_kernel<VEC, ZT, DEVPOSE, MINB, true, 2, ROLL, QUEUE><<<grid, kBlockThreads, 0, s>>>(p, d_K, d_pose);
    else if (p.sanitized && !p.no_ptx)
      integrate_pending_kernel<VEC, ZT, DEVPOSE, MINB, true, 1, ROLL, QUEUE><<<grid, kBlockThreads, 0, s>>>(p, d_K, d_pose);
    else if (p.sanitized)
      integrate_pending_kernel<VEC, ZT, DEVPOSE, MINB, true, 0, ROLL, QUEUE><<<grid, kBlockThreads, 0, s>>>(p, d_K, d_pose);
    else
      integrate_pending_kernel<VEC, ZT, DEVPOSE, MINB, false, 0, ROLL, QUEUE><<<grid, kBlockThreads, 0, s>>>(p, d_K, d_pose);
  }
  else
    integrate_kernel<VEC, ZT, DEVPOSE, MINB, FAST><<<grid, kBlockThreads, 0, s>>>(p, d_K, d_pose);
  return cudaGetLastError();
}

}  // namespace

int last_integrate_launches() { return g_integrate_launches; }

// Frame masks of EVERY warp tile of the slab described by p (production kernel shape: VEC = 4, ZT = 8), in one
// launch on stream s.  The engine runs it on its main stream, where it overlaps the previous pass' fusion kernels,
// and hands each sub-slab launch its z range of the result (tiles are numbered z group-major).  Returns the number
// of tiles per z group, or 0 when the pre-pass does not apply (the fusion kernel then derives its own masks).
int64_t launch_tile_masks(IntegrateParams p, cudaStream_t s) {
  constexpr int VEC = 4, ZT = 8;
  if (p.dim_x % VEC != 0 || p.variant != 0 || p.cull_level < 1 || p.tile_masks == nullptr || p.no_ptx || p.no_packed ||
      !p.sanitized || p.fast_level == 0 || !(p.trunc_margin > 0.0f) || (int64_t)p.im_w * p.im_h >= (1 << 22))
    return 0;
  p.nx_units = p.dim_x / VEC;
  p.lim_x = 0.5f - pixel_round_margin((float)p.im_w, p.cx);
  p.lim_y = 0.5f - pixel_round_margin((float)p.im_h, p.cy);
  const int lx = choose_lx(p.nx_units, p.force_lx);
  p.lx_shift = ilog2(lx);
  const int rows = kBlockThreads / lx;
  const int64_t n_xt = (p.nx_units + lx - 1) / lx;
  const int64_t n_yt = (int64_t)((p.dim_y + rows - 1) / rows) * (kBlockThreads / 32);
  const int64_t n_zg = (p.nz + ZT - 1) / ZT;
  const int64_t n_tiles = n_xt * n_yt * n_zg;
  if (n_tiles > p.tile_mask_cap) return 0;
  int nf_shift = 0;
  while ((1 << nf_shift) < p.n_frames) ++nf_shift;
  const int64_t threads = ((n_tiles + (32 >> nf_shift) - 1) / (32 >> nf_shift)) * 32;
  tile_mask_kernel<VEC, ZT><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p, (int)n_xt, (int)n_yt, n_tiles, nf_shift);
  return cudaGetLastError() == cudaSuccess ? n_xt * n_yt : 0;
}

cudaError_t launch_integrate(const IntegrateParams& p, cudaStream_t s) {
  if (p.n_frames < 1 || p.n_frames > kMaxBatch) return cudaErrorInvalidValue;
  if (p.dim_x % 4 != 0) return launch_integrate_t<1, 8, false, 6, true, true, 2, 3>(p, nullptr, nullptr, s);
#ifdef TSDF_VALIDATION
  // validation flavour of the library only (libtsdf_fusion_validation.so): earlier kernel generations and tuning
  // variants, selected through tsdf_volume_desc.reserved0, all bit-identical (tests/test_gpu_parity.py)
  switch (p.variant) {
    case 1: return launch_integrate_t<4, 4, false, 2>(p, nullptr, nullptr, s);
    case 2: return launch_integrate_t<4, 4, false, 3>(p, nullptr, nullptr, s);
    case 3: return launch_integrate_t<4, 2, false, 4>(p, nullptr, nullptr, s);
    case 4: return launch_integrate_t<4, 2, false, 5>(p, nullptr, nullptr, s);
    case 5: return launch_integrate_t<4, 4, false, 4>(p, nullptr, nullptr, s);
    case 6: return launch_integrate_t<4, 4, false, 4, true, true>(p, nullptr, nullptr, s);
    case 7: return launch_integrate_t<4, 4, false, 5, true, true>(p, nullptr, nullptr, s);
    case 8: return launch_integrate_t<4, 8, false, 4, true, true>(p, nullptr, nullptr, s);
    case 9: return launch_integrate_t<4, 8, false, 3, true, true>(p, nullptr, nullptr, s);
    case 10: return launch_integrate_t<4, 4, false, 6, true, true>(p, nullptr, nullptr, s);
    case 11: return launch_integrate_t<4, 8, false, 4, true, true>(p, nullptr, nullptr, s);
    case 12: return launch_integrate_t<4, 8, false, 5, true, true, 1>(p, nullptr, nullptr, s);
    case 13: return launch_integrate_t<4, 8, false, 5, true, true, 2>(p, nullptr, nullptr, s);
    case 14: return launch_integrate_t<4, 8, false, 6, true, true, 2>(p, nullptr, nullptr, s);
    case 16: return launch_integrate_t<4, 8, false, 5, true, true, 4>(p, nullptr, nullptr, s);
    case 17: return launch_integrate_t<4, 8, false, 5, true, true>(p, nullptr, nullptr, s);
    case 18: return launch_integrate_t<4, 8, false, 5, true, true, 2, true>(p, nullptr, nullptr, s);
    case 19: return launch_integrate_t<4, 8, false, 6, true, true, 2, true>(p, nullptr, nullptr, s);
    case 20: return launch_integrate_t<4, 8, false, 5, true, true, 2>(p, nullptr, nullptr, s);
    case 21: return launch_integrate_t<4, 4, false, 6, true, true, 2, true>(p, nullptr, nullptr, s);
    case 22: return launch_integrate_t<4, 4, false, 6, true, true, 4, true>(p, nullptr, nullptr, s);
    case 23: return launch_integrate_t<4, 2, false, 6, true, true, 2, true>(p, nullptr, nullptr, s);
    case 24: return launch_integrate_t<4, 8, false, 4, true, true, 2, true>(p, nullptr, nullptr, s);
    case 25: return launch_integrate_t<4, 8, false, 5, true, true, 4, true>(p, nullptr, nullptr, s);
    case 26: return launch_integrate_t<4, 8, false, 5, true, true, 1, true>(p, nullptr, nullptr, s);
    case 27: return launch_integrate_t<4, 8, false, 6, true, true, 2, 2>(p, nullptr, nullptr, s);
    case 28: return launch_integrate_t<4, 8, false, 5, true, true, 2, 2>(p, nullptr, nullptr, s);
    case 29: return launch_integrate_t<4, 8, false, 6, true, true, 2, 3>(p, nullptr, nullptr, s);
    case 30: return launch_integrate_t<4, 8, false, 5, true, true, 2, 3>(p, nullptr, nullptr, s);
    case 31: return launch_integrate_t<4, 8, false, 6, true, true, 2, 1>(p, nullptr, nullptr, s);  // per-warp queue
    default: return launch_integrate_t<4, 8, false, 6, true, true, 2, 3>(p, nullptr, nullptr, s);
  }
#else
  return launch_integrate_t<4, 8, false, 6, true, true, 2, 3>(p, nullptr, nullptr, s);
#endif
}

cudaError_t launch_integrate_devpose(IntegrateParams p, const float* d_cam_K, const float* d_cam2base,
                                     cudaStream_t s) {
  p.n_frames = 1;
  p.cull_level = 0;  // pose unknown on the host: no margins can be derived
  p.fast_level = 1;
  p.sanitized = 0;   // the caller's depth image is read in place
  const bool vec4 = (p.dim_x % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.tsdf) & 15u) == 0) &&
                    ((reinterpret_cast<uintptr_t>(p.weight) & 15u) == 0);
  if (vec4) return launch_integrate_t<4, 8, true, 6, true, true, 2, true>(p, d_cam_K, d_cam2base, s);
  return launch_integrate_t<1, 8, true, 6, true, true, 2, true>(p, d_cam_K, d_cam2base, s);
}

int64_t tile_mask_entries(int dim_x, int dim_y, int64_t nz) {
  // production shape: VEC = 4 (or 1), ZT = 8, lx in {8, 16, 32} threads along x, 8 warps of 32 / lx rows per block
  const int64_t zg = (nz + 7) / 8;
  int64_t worst = 0;
  for (int vec = 1; vec <= 4; vec += 3) {
    if (vec == 4 && dim_x % 4 != 0) continue;
    const int64_t nxu = dim_x / vec;
    for (int lx = 4; lx <= 32; lx <<= 1) {
      const int64_t xt = (nxu + lx - 1) / lx, rows = kBlockThreads / lx;
      const int64_t yt = ((dim_y + rows - 1) / rows) * (kBlockThreads / 32);
      if (xt * yt * zg > worst) worst = xt * yt * zg;
    }
  }
  return worst;
}

cudaError_t launch_fill(float* tsdf, float* weight, int64_t n, cudaStream_t s) {
  const int blocks = 148 * 8;
  fill_kernel<<<blocks, 256, 0, s>>>(tsdf, weight, n);
  return cudaGetLastError();
}

cudaError_t launch_depth_from_u16(const uint16_t* mm, float* depth, int64_t n, bool sanitize, cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  depth_from_u16_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(mm, depth, n, sanitize);
  return cudaGetLastError();
}

MipLayout make_mip_layout(int im_w, int im_h) {
  MipLayout L;
  int off = 0;
  for (int l = 0; l < 5; ++l) {
    const int b = 8 << l;
    L.w[l] = (im_w + b - 1) / b;
    L.h[l] = (im_h + b - 1) / b;
    L.off[l] = off;
    off += L.w[l] * L.h[l];
  }
  L.stride = off;
  return L;
}

cudaError_t launch_build_depth_mips(const float* frames, float* mips, int n_frames, int im_w, int im_h,
                                    int64_t frame_stride, cudaStream_t s) {
  const MipLayout L = make_mip_layout(im_w, im_h);
  dim3 grid((unsigned)((im_w + 127) / 128), (unsigned)((im_h + 127) / 128), (unsigned)n_frames);
  build_depth_mips_kernel<<<grid, 256, 0, s>>>(frames, mips, im_w, im_h, frame_stride, L);
  return cudaGetLastError();
}

cudaError_t launch_sanitize_depth(const float* in, float* out, int64_t n, cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  sanitize_depth_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launch_extract_count(const ExtractParams& p, unsigned long long* tile_counts, int64_t n_tiles,
                                 cudaStream_t s) {
  const int blocks = (int)(n_tiles < 148 * 16 ? n_tiles : 148 * 16);
  extract_count_kernel<<<blocks > 0 ? blocks : 1, kBlockThreads, 0, s>>>(p, tile_counts, n_tiles);
  return cudaGetLastError();
}

cudaError_t launch_extract_scan(unsigned long long* tile_counts, int64_t n_tiles, unsigned long long* total,
                                cudaStream_t s) {
  extract_scan_kernel<<<1, 1024, 0, s>>>(tile_counts, n_tiles, total);
  return cudaGetLastError();
}

cudaError_t launch_extract_emit(const ExtractParams& p, const unsigned long long* tile_offsets, int64_t n_tiles,
                                float* xyz, long long* voxel_index, cudaStream_t s) {
  const int blocks = (int)(n_tiles < 148 * 16 ? n_tiles : 148 * 16);
  extract_emit_kernel<<<blocks > 0 ? blocks : 1, kBlockThreads, 0, s>>>(p, tile_offsets, n_tiles, xyz, voxel_index);
  return cudaGetLastError();
}

cudaError_t launch_surface_normals(const ExtractParams& p, int64_t nz, const long long* voxel_index, int64_t n,
                                   float* normals, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  surface_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, nz, voxel_index, n, normals);
  return cudaGetLastError();
}

// Small host -> device transfers of the stream path (poses, job descriptors) and the zeroing of its counters, as a KERNEL
// on the fusion stream instead of cudaMemcpyAsync / cudaMemsetAsync: the copy engine executes what became runnable
// first, and a 48-byte pose copy that becomes runnable together with the next group's 300 MB frame upload waited behind
// it -- the whole upload ended up serialised with the fusion (CUPTI timeline, profiles/r02_e2e_timeline.txt).  `src` is
// pinned host memory (cudaHostAlloc: device-accessible under unified addressing) or nullptr for a zero fill.
__global__ void __launch_bounds__(256) small_copy_kernel(unsigned* __restrict__ dst, const unsigned* src, int64_t n_words) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src ? src[i] : 0u;
}

cudaError_t launch_small_copy(void* dst, const void* pinned_src_or_null, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return cudaSuccess;
  if (bytes % 4 != 0) return cudaErrorInvalidValue;
  const int64_t n = (int64_t)(bytes / 4);
  const int blocks = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  small_copy_kernel<<<blocks, 256, 0, s>>>(static_cast<unsigned*>(dst), static_cast<const unsigned*>(pinned_src_or_null), n);
  return cudaGetLastError();
}

cudaError_t launch_selftest_rounding(unsigned long long* mismatches, cudaStream_t s) {
  selftest_rounding_kernel<<<148 * 16, 256, 0, s>>>(mismatches);
  return cudaGetLastError();
}

}  // namespace tsdf
