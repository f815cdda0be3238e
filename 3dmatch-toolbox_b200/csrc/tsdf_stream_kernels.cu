// tsdf_stream_kernels.cu -- the production fusion path: one persistent work-list kernel per group of frames.
//
// Replaces depth-fusion/demo.cu:18-68 (Integrate) + the frame loop of demo.cu:156-183 for a whole batch of frames.
//
// Work decomposition
//   * block tile  = 32 x 32 x 8 voxels (x, y, z): 256 threads, each owning 4 consecutive x by 8 consecutive z at one y
//                   (one float4 of TSDF and of weight per plane); a warp is 8 lanes along x by 4 rows of y.
//   * chunk       = up to 32 consecutive frames (one 32-bit frame mask per warp tile).
//   * item        = (block tile, chunk) with at least one frame that can touch the tile.
//   stream_mask_kernel tests every (warp tile, frame) of the group up front (frustum + hierarchical-Z occlusion,
//   tsdf_device.cuh) and appends the non-empty items to per-chunk lists.  stream_integrate_kernel is ONE launch of
//   resident blocks that take items from a global ticket, chunk after chunk, so all blocks work on the same <= 32
//   depth frames at a time (39 MB: L2-resident) while the volume streams through once per chunk.  Items of one
//   tile are ordered by a per-tile progress word (release / acquire); items are fetched in ticket order, so the
//   item waited for is always held by a running block and the wait cannot deadlock.
//
// Inside an item
//   Phase A (per warp, no block barrier): for every frame of the warp tile's mask, the packed bit-exact
//     classification (voxel_classify_pair) of the thread's 32 voxels.  Certain free-space observations (dist == 1)
//     only bump an 8-bit pending count per voxel (exact: see apply_pending); a voxel whose observation is NOT certain
//     (inside the +-trunc band, or pixel rounding within the margin) becomes STICKY for the rest of the item: a
//     record {owner thread, voxel slot, first uncertain frame f0, pending count so far} is appended to the block's
//     record list and phase A's further counts for it are ignored.
//   Write-back: pending counts of the non-sticky voxels, coalesced float4 read-modify-write, once per item.
//   Phase B (whole block, full lanes): one thread per record loads the voxel's TSDF/weight, applies the pending
//     count (all older than f0), then evaluates frames f >= f0 of the owner's frame mask with the plain restatement
//     of the reference (voxel_observation_exact) in ascending order, keeping TSDF/weight in registers; one store.
//   Every voxel therefore receives exactly the reference's sequence of updates (demo.cu:63-66) in ascending frame
//   order; there is no barrier per frame and no shared queue traffic per frame.
#include "tsdf_device.cuh"

#include <cstdlib>

namespace tsdf {

namespace {

constexpr int VEC = 4, ZT = 8;
constexpr int kWarps = kBlockThreads / 32;
constexpr int kRecCap = kBlockThreads * VEC * ZT;  // every voxel of a block tile becomes sticky at most once per item

// Packed fp32 add / sub with a scalar (broadcast) second operand: one IEEE rounding per lane.
// NOTE: ptxas (12.9) contracts a packed multiply whose only use is a packed add into FFMA2 -- even with explicit .rn
// modifiers, even with --fmad=false, even when the multiply is written as fma(a, b, -0) -- which would turn the
// reference's two roundings of (M0*t0 + M4*t1) into one.  Products that feed an add are therefore formed with the
// scalar mul.rn.f32 (never contracted) and only the add is packed; tsdf_selftest_packed() pins the result bit for bit
// against the scalar intrinsics.
__device__ __forceinline__ f32x2_t sub2s(f32x2_t a, float s) {
  f32x2_t r;
  asm("{\n\t.reg .b64 b;\n\tmov.b64 b, {%2, %2};\n\tsub.rn.f32x2 %0, %1, b;\n\t}" : "=l"(r) : "l"(a), "f"(s));
  return r;
}
__device__ __forceinline__ f32x2_t add2s(f32x2_t a, float s) {
  f32x2_t r;
  asm("{\n\t.reg .b64 b;\n\tmov.b64 b, {%2, %2};\n\tadd.rn.f32x2 %0, %1, b;\n\t}" : "=l"(r) : "l"(a), "f"(s));
  return r;
}

// demo.cu:35-41 for two x-adjacent voxels: fl(fl(m * t0) + y) with t0 = fl(bx - tx), one rounding per operation.
__device__ __forceinline__ f32x2_t setup_pair(f32x2_t t0, float m, float y) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(t0));
  return add2s(pack2(__fmul_rn(lo, m), __fmul_rn(hi, m)), y);
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Frame masks + item lists.  One warp per (job, block tile); lane = frame inside the chunk.  The block tile is
// tested first (a frame that cannot touch the 32 x 32 x 8 box cannot touch any of its warp tiles), then its 8 warp
// tiles (32 x 4 x 8).
// ------------------------------------------------------------------------------------------------
// kCostOnly: no masks / items are written, only the per-z-group counts of kept (warp tile, frame) pairs (slab planning).
template <bool kCostOnly>
__global__ void __launch_bounds__(256, TSDF_MASK_MINB) stream_mask_kernel(const __grid_constant__ StreamLaunch L) {
  __shared__ uint2 s_ballot[256 / 32][kWarps];   // (keep, interior) ballots of the warp's eight warp tiles
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (int64_t)L.n_jobs * L.n_btiles) return;
  const int job = (int)(gw / L.n_btiles);
  int btile = (int)(gw - (int64_t)job * L.n_btiles);
  if (L.z_reverse) btile = L.n_btiles - 1 - btile;   // far z groups (more of the frusta) enter the lists first
  const StreamJob& J = L.jobs[job];
  Intrinsics K;
  K.fx = L.fx; K.fy = L.fy; K.cx = L.cx; K.cy = L.cy; K.lim_x = L.lim_x; K.lim_y = L.lim_y;
  const int xt = btile % L.n_xt;
  const int r = btile / L.n_xt;
  const int yb = r % L.n_yb, zg = r / L.n_yb;
  const int rows_per_warp = 32 >> L.lx_shift, rows_per_block = kBlockThreads >> L.lx_shift;
  const int x_lo = (xt << L.lx_shift) * VEC, x_hi = min(L.dim_x - 1, x_lo + ((VEC << L.lx_shift) - 1));
  const int yb_lo = yb * rows_per_block, yb_hi = min(L.dim_y - 1, yb_lo + rows_per_block - 1);
  const int z_lo = J.z_begin + zg * ZT, z_hi = z_lo + ZT - 1;
  // Lanes span FRAMES: a lane block covers per_blk = chunk * floor(32 / chunk) consecutive frames, i.e. a whole number
  // of chunks, so the ballots of one round hold the masks of up to 32 chunks (chunk = 1) and lane ci then finishes
  // chunk ci of the block (mask words, item append) in parallel with the other lanes.
  const int chunks_per_blk = 32 / L.chunk, per_blk = chunks_per_blk * L.chunk;
  const unsigned sel = (L.chunk >= 32) ? 0xffffffffu : ((1u << L.chunk) - 1u);
  unsigned carry_prev = 0u;  // (last non-empty chunk of this tile so far) + 1; warp-uniform
  unsigned long long cost = 0ull;
  for (int fb = 0; fb < L.n_frames; fb += per_blk) {
    const int f = fb + lane;
    const bool valid = lane < per_blk && f < L.n_frames;
    unsigned keep_w[kWarps], in_w[kWarps];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) keep_w[w] = in_w[w] = 0u;
    if (L.cull_level <= 0) {
      const unsigned all = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int w = 0; w < kWarps; ++w) keep_w[w] = (yb_lo + w * rows_per_warp < L.dim_y) ? all : 0u;
    } else {
      FramePose ps = {};
      const float* mips_f = nullptr;
      bool bk = false, bi = false;
      if (valid) {
        ps = J.poses[f];
        if (J.mips != nullptr) mips_f = J.mips + (int64_t)f * L.mip_stride;
        box_frame_test(L, K, ps, mips_f, x_lo, x_hi, yb_lo, yb_hi, z_lo, z_hi, bk, bi);
      }
      if (__any_sync(0xffffffffu, bk)) {
        // (kept rolled: eight inlined copies of the test made the kernel 94 KB of SASS and the instruction cache its
        //  bottleneck -- 48 % of the stall samples were no_instructions; the ballots go through shared memory because a
        //  rolled loop cannot index a register array)
#pragma unroll 1
        for (int w = 0; w < kWarps; ++w) {
          const int y_lo = yb_lo + w * rows_per_warp;
          bool k = false, i = false, nr = false;
          if (bk && y_lo < L.dim_y)
            box_frame_test(L, K, ps, mips_f, x_lo, x_hi, y_lo, min(L.dim_y - 1, y_lo + rows_per_warp - 1), z_lo, z_hi, k, i, &nr);
          const unsigned kb = __ballot_sync(0xffffffffu, k), ib = __ballot_sync(0xffffffffu, i);
          if (lane == 0) s_ballot[threadIdx.x >> 5][w] = make_uint2(kb, ib);
          // work statistics: a tile that reaches the band of the farthest surface behind it pays for exact evaluations
          cost += (unsigned long long)L.plan_beta * (unsigned)__popc(__ballot_sync(0xffffffffu, nr));
        }
        __syncwarp();
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          const uint2 b = s_ballot[threadIdx.x >> 5][w];
          keep_w[w] = b.x;
          in_w[w] = b.y;
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int w = 0; w < kWarps; ++w) cost += (unsigned)__popc(keep_w[w]);
    if constexpr (!kCostOnly) {
      // lane ci: chunk c = first chunk of the block + ci
      const int c0 = fb / L.chunk, c = c0 + lane;
      const bool mine = lane < chunks_per_blk && c < L.n_chunks;
      const int sh = (lane < chunks_per_blk) ? lane * L.chunk : 0;
      unsigned uni = 0u;
      if (mine) {
        uint2* out = J.masks + ((int64_t)c * L.n_btiles + btile) * kWarps;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          const unsigned kw = (keep_w[w] >> sh) & sel;
          uni |= kw;
          out[w] = make_uint2(kw, (in_w[w] >> sh) & sel);
        }
      }
      const unsigned nonempty = __ballot_sync(0xffffffffu, mine && uni != 0u);
      if (mine && uni != 0u) {
        const unsigned below = nonempty & ((1u << lane) - 1u);
        const unsigned prev_plus1 = below ? (unsigned)(c0 + (31 - __clz(below)) + 1) : carry_prev;
        const unsigned pos = atomicAdd(&J.item_count[c], 1u);
        J.items[(int64_t)c * L.n_btiles + pos] = (unsigned long long)(unsigned)btile | ((unsigned long long)prev_plus1 << 32);
      }
      if (nonempty != 0u) carry_prev = (unsigned)(c0 + (31 - __clz(nonempty)) + 1);
    }
  }
  if (lane == 0 && J.zcost != nullptr && cost != 0ull) atomicAdd(&J.zcost[zg], cost);
}

// Exclusive prefix of the per-(job, chunk) item counts -> seg_begin[0 .. n_seg], and the ticket reset.
__global__ void __launch_bounds__(kStreamMaxSegments) stream_segment_kernel(const __grid_constant__ StreamLaunch L) {
  __shared__ unsigned s_warp[kStreamMaxSegments / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n_seg = L.n_jobs * L.n_chunks;
  const unsigned cnt = (t < n_seg) ? L.jobs[t / L.n_chunks].item_count[t % L.n_chunks] : 0u;
  unsigned incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned v = s_warp[lane], inc2 = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, inc2, d);
      if (lane >= d) inc2 += u;
    }
    s_warp[lane] = inc2 - v;
  }
  __syncthreads();
  const unsigned excl = s_warp[warp] + incl - cnt;
  if (t < n_seg) L.seg_begin[t] = excl;
  if (t == n_seg - 1) L.seg_begin[n_seg] = excl + cnt;
  if (t == 0) *L.ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// The persistent fusion kernel.
// ------------------------------------------------------------------------------------------------
// DIAG (validation library only, results are NOT valid): 1 = no TSDF / weight traffic at all, so that every L2 read of
// the launch is a depth gather (cache evidence in profiles/); 2 = depth gathers with an L2 evict_last cache hint (A/B).
// (Measured and dropped: every thread prefetching the depth line of its first voxel in the warp's next frame into L1 --
// 84.95 ms against 79.49 ms on the 256-frame run, profiles/r02_margin_and_prefetch_ab.txt.)
// G: phase B is pooled over 256 / G threads: G = 1 the whole block (fullest lanes, one block barrier per epoch), G = 2 / 4
// half / quarter blocks synchronised by named barriers (fewer warps wait for each other), G = 8 per warp (no barrier at
// all, emptiest lanes).
template <int MINB, int DIAG = 0, int G = 1>
__global__ void __launch_bounds__(kBlockThreads, MINB) stream_integrate_kernel(const __grid_constant__ StreamLaunch L) {
  __shared__ float4 s_ztab[kChunkMax * ZT];        // plane products M8*t2 of every (frame, plane) of the item
  __shared__ float4 s_pose[kChunkMax * 3];         // (m0 m1 m2 tx) (m4 m5 m6 ty) (m8 m9 m10 tz) per frame
  __shared__ unsigned s_pend[ZT][kBlockThreads];   // byte j of [k][t]: pending free-space updates of voxel (j, k) of thread t
  __shared__ unsigned s_wmask[kWarps];
  __shared__ unsigned s_rec_count[G][3];           // record counters of every group, rotating over the epochs
  __shared__ unsigned s_hdr[4];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx_mask = (1 << L.lx_shift) - 1;
  const int lx = threadIdx.x & lx_mask, ly = threadIdx.x >> L.lx_shift;
  const int rows_per_block = kBlockThreads >> L.lx_shift;
  unsigned* const rec_base = L.rec_scratch + (size_t)blockIdx.x * (2 * kRecCap);  // two lists, alternating epochs
  Intrinsics K;
  K.fx = L.fx; K.fy = L.fy; K.cx = L.cx; K.cy = L.cy; K.lim_x = L.lim_x; K.lim_y = L.lim_y;
  const float im_w_f = (float)L.im_w, half_w = 0.5f * (float)(L.im_w - 1), half_h = 0.5f * (float)(L.im_h - 1);
  const int64_t plane = (int64_t)L.dim_y * L.dim_x;
  [[maybe_unused]] unsigned long long l2_policy = 0ull;
#ifdef TSDF_VALIDATION
  if constexpr (DIAG == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_policy));
#endif
  const unsigned n_seg = (unsigned)(L.n_jobs * L.n_chunks);

  // thread 0: ticket of the next item (fetched one item ahead, so the atomic's latency is off the critical path)
  unsigned seg = 0u, ticket = 0u, total = 0u;
  if (threadIdx.x == 0) {
    ticket = atomicAdd(L.ticket, 1u);
    total = L.seg_begin[n_seg];
  }

  for (;;) {
    if (threadIdx.x == 0) {
      if (ticket >= total) {
        s_hdr[0] = 0xffffffffu;
      } else {
        while (ticket >= L.seg_begin[seg + 1u]) ++seg;
        const unsigned job = seg / (unsigned)L.n_chunks, chunk = seg - job * (unsigned)L.n_chunks;
        const StreamJob& J0 = L.jobs[job];
        const unsigned long long it = J0.items[(int64_t)chunk * L.n_btiles + (ticket - L.seg_begin[seg])];
        const unsigned bt = (unsigned)it, need = (unsigned)(it >> 32);
        s_hdr[0] = job; s_hdr[1] = chunk; s_hdr[2] = bt;
        // the tile's previous item (an earlier chunk) must be complete: its ticket is smaller than ours, so it is
        // held by a running block and this wait terminates
        if (need != 0u)
          while (ld_acquire(J0.progress + bt) < need) __nanosleep(64);
        ticket = atomicAdd(L.ticket, 1u);
      }
    }
    __syncthreads();
    const unsigned job = s_hdr[0];
    if (job == 0xffffffffu) break;
    const int chunk = (int)s_hdr[1], btile = (int)s_hdr[2];
    const StreamJob& J = L.jobs[job];
    const int xt = btile % L.n_xt;
    const int rr = btile / L.n_xt;
    const int yb = rr % L.n_yb, zg = rr / L.n_yb;
    const int xu = (xt << L.lx_shift) + lx;
    const int y = yb * rows_per_block + ly;
    const int zl0 = zg * ZT;
    const bool active = (xu < L.nx_units) && (y < L.dim_y);
    const int zcount = min(ZT, L.nz - zl0);
    const int f_base = chunk * L.chunk;
    const int nf = min(L.chunk, L.n_frames - f_base);
    const FramePose* __restrict__ poses = J.poses + f_base;

    const uint2 wm2 = __ldg(J.masks + ((int64_t)chunk * L.n_btiles + btile) * kWarps + warp);
    const unsigned frame_mask = wm2.x, interior_mask = wm2.y;
    if (lane == 0) s_wmask[warp] = frame_mask;
    if (threadIdx.x < 3 * G) (&s_rec_count[0][0])[threadIdx.x] = 0u;
    for (int e = threadIdx.x; e < nf * ZT; e += kBlockThreads) {
      const int f = e / ZT, k = e - f * ZT;
      const FramePose ps = poses[f];
      if (k < 3) {
        const float4 row = (k == 0) ? make_float4(ps.m0, ps.m1, ps.m2, ps.tx)
                                    : (k == 1) ? make_float4(ps.m4, ps.m5, ps.m6, ps.ty)
                                               : make_float4(ps.m8, ps.m9, ps.m10, ps.tz);
        s_pose[f * 3 + k] = row;
      }
      // demo.cu:32,37,39-41: the plane's t2 = base_z - tz and its three products, with the reference's operations
      const float t2 = __fsub_rn(__fadd_rn(L.oz, __fmul_rn((float)(J.z_begin + zl0 + k), L.voxel_size)), ps.tz);
      s_ztab[e] = make_float4(__fmul_rn(ps.m8, t2), __fmul_rn(ps.m9, t2), __fmul_rn(ps.m10, t2), 0.0f);
    }
#pragma unroll
    for (int k = 0; k < ZT; ++k) s_pend[k][threadIdx.x] = 0u;
    __syncthreads();

    const int64_t idx0 = (int64_t)zl0 * plane + (int64_t)y * L.dim_x + (int64_t)xu * VEC;
    float* const tsdf = J.tsdf;
    float* const weight = J.weight;
    const float* const depth0 = J.depth + (int64_t)f_base * L.frame_stride;

    // voxel-centre coordinates of the thread's patch (demo.cu:30-31), from the GLOBAL indices; x as f32x2 pairs
    const float base_y = __fadd_rn(L.oy, __fmul_rn((float)y, L.voxel_size));
    f32x2_t bx2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
      bx2[h] = pack2(__fadd_rn(L.ox, __fmul_rn((float)(xu * VEC + 2 * h), L.voxel_size)),
                     __fadd_rn(L.ox, __fmul_rn((float)(xu * VEC + 2 * h + 1), L.voxel_size)));

    constexpr int kWarpsPerGroup = kWarps / G, kGroupThreads = kBlockThreads / G, kGroupRecCap = kRecCap / G;
    const int grp = warp / kWarpsPerGroup;
    auto group_barrier = [&]() {
      if constexpr (G == 1) __syncthreads();
      else if constexpr (G == kWarps) __syncwarp();
      else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kGroupThreads) : "memory");
    };
    unsigned group_mask = 0u;
#pragma unroll
    for (int w8 = 0; w8 < kWarpsPerGroup; ++w8) group_mask |= s_wmask[grp * kWarpsPerGroup + w8];

    // The chunk is walked in EPOCHS of L.epoch frames: phase A (classification, per warp) of the epoch's frames, then
    // phase B (exact evaluation of the voxels that became sticky in this epoch, pooled over the group), then the sticky
    // marks are dropped.  A voxel flagged only because its pixel rounding was within the margin at one frame therefore
    // costs at most one epoch of exact evaluations, while voxels inside the +-trunc band are simply flagged again.
    // One barrier per epoch: the record lists alternate between two buffers and the counters rotate over three words,
    // so a warp that has finished its share of phase B goes straight on to the next epoch's phase A.
    unsigned n_epoch = 0u;
    for (int e0 = 0; e0 < nf; e0 += L.epoch) {
      const unsigned epoch_bits = ((L.epoch >= 32) ? 0xffffffffu : ((1u << L.epoch) - 1u)) << e0;
      if ((group_mask & epoch_bits) == 0u) continue;  // group-uniform
      const unsigned slot = n_epoch % 3u;
      unsigned* const rec = rec_base + grp * (2 * kGroupRecCap) + (n_epoch & 1u) * kGroupRecCap;
      unsigned* const counter = &s_rec_count[grp][slot];
      ++n_epoch;

      // ---------------- phase A: classification, per warp ----------------
      unsigned sticky = 0u;  // bit k*VEC + j: voxel handed to phase B of this epoch
      for (unsigned fm = frame_mask & epoch_bits; fm != 0u; fm &= fm - 1u) {
        const int f = __ffs(fm) - 1;
        const float4 p0 = s_pose[f * 3 + 0], p1 = s_pose[f * 3 + 1];
        // demo.cu:35-41: (M0*t0 + M4*t1) per x with the reference's operations, one rounding each (setup_pair); the M8*t2
        // part comes from s_ztab
        const float t1 = __fsub_rn(base_y, p1.w);
        const float yx = __fmul_rn(p1.x, t1), yy = __fmul_rn(p1.y, t1), yz = __fmul_rn(p1.z, t1);
        unsigned exact = 0u;  // bit k*VEC + j: the observation of this frame is not certain
        if (active) {
          f32x2_t sx2[2], sy2[2], sz2[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const f32x2_t t0 = sub2s(bx2[h], p0.w);
            sx2[h] = setup_pair(t0, p0.x, yx);
            sy2[h] = setup_pair(t0, p0.y, yy);
            sz2[h] = setup_pair(t0, p0.z, yz);
          }
          // gather base biased by the bit pattern of the rint magic (see voxel_classify_pair)
          const char* depth_biased =
              reinterpret_cast<const char*>(depth0 + (int64_t)f * L.frame_stride) - 4ll * (long long)kRintMagicBits;
          auto plane_body = [&](int k, unsigned& pk, auto interior_tag) {
            constexpr bool kInterior = decltype(interior_tag)::value;
            const float4 zt = s_ztab[f * ZT + k];
            unsigned ex = 0u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#ifdef TSDF_VALIDATION
              if constexpr (DIAG == 2) {
                voxel_classify_pair_hint<kInterior>(sx2[h], sy2[h], sz2[h], zt.x, zt.y, zt.z, K, im_w_f, half_w, half_h,
                                                    depth_biased, L.trunc_margin, 1u << (16 * h), 1u << (16 * h + 8),
                                                    1u << (2 * h), 1u << (2 * h + 1), pk, ex, l2_policy);
                continue;
              }
#endif
              voxel_classify_pair<kInterior>(sx2[h], sy2[h], sz2[h], zt.x, zt.y, zt.z, K, im_w_f, half_w, half_h,
                                             depth_biased, L.trunc_margin, 1u << (16 * h), 1u << (16 * h + 8),
                                             1u << (2 * h), 1u << (2 * h + 1), pk, ex);
            }
            exact |= ex << (k * VEC);
          };
          if ((interior_mask >> f) & 1u) {  // warp-uniform
#pragma unroll 2
            for (int k = 0; k < zcount; ++k) {
              unsigned pk = s_pend[k][threadIdx.x];
              plane_body(k, pk, std::true_type{});
              s_pend[k][threadIdx.x] = pk;
            }
          } else {
#pragma unroll 2
            for (int k = 0; k < zcount; ++k) {
              unsigned pk = s_pend[k][threadIdx.x];
              plane_body(k, pk, std::false_type{});
              s_pend[k][threadIdx.x] = pk;
            }
          }
        }
        unsigned fresh = exact & ~sticky;
        if (fresh != 0u) {
          unsigned pos = atomicAdd(counter, (unsigned)__popc(fresh));
          while (fresh != 0u) {
            const unsigned s = (unsigned)(__ffs(fresh) - 1);
            fresh &= fresh - 1u;
            // the free-space observations counted so far are all older than frame f (this frame did not count: the
            // classification counts only certain observations)
            const unsigned c = (s_pend[s / VEC][threadIdx.x] >> (8u * (s & (unsigned)(VEC - 1)))) & 0xffu;
            rec[pos++] = threadIdx.x | (s << 8) | ((unsigned)f << 13) | (c << 18);
          }
        }
        sticky |= exact;
      }
      // the sticky voxels' counts now belong to their records (and whatever phase A counted for them after f0 is
      // void): restart them
      if (sticky != 0u) {
#pragma unroll
        for (int k = 0; k < ZT; ++k) {
          const unsigned st = (sticky >> (k * VEC)) & ((1u << VEC) - 1u);
          if (st != 0u) {
            unsigned keep = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
              if ((st >> j) & 1u) keep &= ~(0xffu << (8 * j));
            s_pend[k][threadIdx.x] &= keep;
          }
        }
      }
      group_barrier();  // the epoch's records are complete

      // ---------------- phase B: sticky voxels, one record per thread, frames in ascending order ----------------
      const unsigned n_rec = *counter;
      // the counter of the epoch after the next one was last read in the previous epoch's phase B, i.e. before this
      // barrier, and is first written after the next epoch's barrier
      if ((threadIdx.x & (kGroupThreads - 1)) == 0) s_rec_count[grp][(slot + 2u) % 3u] = 0u;
      for (unsigned base = (unsigned)(warp - grp * kWarpsPerGroup) * 32u; base < n_rec; base += (unsigned)kGroupThreads) {  // warp-uniform trip count
        const unsigned i = base + (unsigned)lane;
        const bool valid = i < n_rec;
        unsigned wm = 0u, c = 0u;
        int64_t idx = 0;
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, t = 0.f, w = 0.f;
        if (valid) {
          const unsigned item = __ldcg(rec + i);
          const int owner_t = (int)(item & 0xffu);
          const int s = (int)((item >> 8) & 31u), j = s & (VEC - 1), k = s / VEC;
          const unsigned f0 = (item >> 13) & 31u;
          c = (item >> 18) & 0xffu;
          wm = s_wmask[owner_t >> 5] & (0xffffffffu << f0) & epoch_bits;
          const int o_xu = (xt << L.lx_shift) + (owner_t & lx_mask);
          const int o_y = yb * rows_per_block + (owner_t >> L.lx_shift);
          // voxel centre (demo.cu:30-32), from the GLOBAL indices
          q0 = __fadd_rn(L.ox, __fmul_rn((float)(o_xu * VEC + j), L.voxel_size));
          q1 = __fadd_rn(L.oy, __fmul_rn((float)o_y, L.voxel_size));
          q2 = __fadd_rn(L.oz, __fmul_rn((float)(J.z_begin + zl0 + k), L.voxel_size));
          idx = ((int64_t)(zl0 + k) * L.dim_y + o_y) * (int64_t)L.dim_x + (int64_t)o_xu * VEC + j;
          if constexpr (DIAG == 1) {
            t = 1.0f;
            w = (float)(item >> 26);   // (opaque to the compiler; always 0)
          } else {
            t = __ldcg(tsdf + idx);
            w = __ldcg(weight + idx);
          }
          apply_pending(t, w, c);
        }
        bool changed = c != 0u;
        for (unsigned um = __reduce_or_sync(0xffffffffu, wm); um != 0u; um &= um - 1u) {
          const int f = __ffs(um) - 1;
          if ((wm >> f) & 1u) {
            const float4 p0 = s_pose[f * 3 + 0], p1 = s_pose[f * 3 + 1], p2 = s_pose[f * 3 + 2];
            // demo.cu:35-41, operation for operation
            const float t0 = __fsub_rn(q0, p0.w), t1 = __fsub_rn(q1, p1.w), t2 = __fsub_rn(q2, p2.w);
            const float cam_x = __fadd_rn(__fadd_rn(__fmul_rn(p0.x, t0), __fmul_rn(p1.x, t1)), __fmul_rn(p2.x, t2));
            const float cam_y = __fadd_rn(__fadd_rn(__fmul_rn(p0.y, t0), __fmul_rn(p1.y, t1)), __fmul_rn(p2.y, t2));
            const float cam_z = __fadd_rn(__fadd_rn(__fmul_rn(p0.z, t0), __fmul_rn(p1.z, t1)), __fmul_rn(p2.z, t2));
            // (a chunk's frames span < 2^31 elements: 32-bit frame offset)
            const float dj = voxel_observation_exact_inline<true, true>(cam_x, cam_y, cam_z, K.fx, K.fy, K.cx, K.cy, depth0,
                                                                        L.im_w, L.im_h, L.trunc_margin,
                                                                        (unsigned)f * (unsigned)L.frame_stride);
            if (dj != kNoObservation) {
              voxel_update(t, w, dj);  // demo.cu:63-66
              changed = true;
            }
          }
        }
        if (valid && changed && (DIAG != 1 || t == -7.25f)) {   // (DIAG 1: never true, but keeps the evaluations alive)
          __stcg(tsdf + idx, t);
          __stcg(weight + idx, w);
        }
      }
    }
    group_barrier();  // every phase B store of the group precedes the write-back's loads

    // ---------------- write-back of the pending counts (certain free-space observations), coalesced ----------------
    // (after the last phase B: the float4 rows are re-read here, so they carry phase B's stores)
    if (frame_mask != 0u && DIAG != 1) {  // warp-uniform
      // the rows the write-back will touch are prefetched into L2 first: with few frames per chunk the write-back is the
      // HBM-bound part of the item, and the plane loop below (kept rolled: apply_pending is a data-dependent loop) has
      // only one plane's loads in flight
      if (active) {
#pragma unroll
        for (int k = 0; k < ZT; ++k)
          if (k < zcount && s_pend[k][threadIdx.x] != 0u) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tsdf + idx0 + k * plane));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(weight + idx0 + k * plane));
          }
      }
#pragma unroll 1
      for (int k = 0; k < ZT; ++k) {
        const unsigned pk = s_pend[k][threadIdx.x];
        if (__any_sync(0xffffffffu, pk != 0u)) {
          if (active && k < zcount) {
            float4 T = __ldcg(reinterpret_cast<const float4*>(tsdf + idx0 + k * plane));
            float4 W = __ldcg(reinterpret_cast<const float4*>(weight + idx0 + k * plane));
            apply_pending(T.x, W.x, pk & 0xffu);
            apply_pending(T.y, W.y, (pk >> 8) & 0xffu);
            apply_pending(T.z, W.z, (pk >> 16) & 0xffu);
            apply_pending(T.w, W.w, (pk >> 24) & 0xffu);
            __stcs(reinterpret_cast<float4*>(tsdf + idx0 + k * plane), T);
            __stcs(reinterpret_cast<float4*>(weight + idx0 + k * plane), W);
          }
        }
      }
    }

    // ---------------- completion: publish the tile's progress ----------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(J.progress + btile, (unsigned)chunk + 1u);
  }
}

// Self-test of the packed set-up arithmetic against the scalar round-to-nearest intrinsics on pseudo-random operands
// (all magnitudes the path sees, signs mixed): counts lanes whose bits differ.
__global__ void __launch_bounds__(256) selftest_packed_kernel(unsigned long long* __restrict__ mismatches, unsigned n_per_thread) {
  unsigned long long state = 0x9E3779B97F4A7C15ull * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1ull);
  auto next = [&]() {   // splitmix64 -> float in (-4, 4) with a random exponent spread
    state += 0x9E3779B97F4A7C15ull;
    unsigned long long z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);           // [0, 1)
    const float scale = __uint_as_float(((unsigned)((z >> 8) & 7u) + 122u) << 23);  // 2^-5 .. 2^2
    return ((z & 1ull) ? -u : u) * scale;
  };
  unsigned long long bad = 0;
  for (unsigned i = 0; i < n_per_thread; ++i) {
    const float b0 = next(), b1 = next(), tx = next(), m = next(), y = next();
    const f32x2_t got = setup_pair(sub2s(pack2(b0, b1), tx), m, y);
    float g0, g1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(g0), "=f"(g1) : "l"(got));
    const float r0 = __fadd_rn(__fmul_rn(__fsub_rn(b0, tx), m), y), r1 = __fadd_rn(__fmul_rn(__fsub_rn(b1, tx), m), y);
    bad += (__float_as_uint(g0) != __float_as_uint(r0)) + (__float_as_uint(g1) != __float_as_uint(r1));
  }
  if (bad) atomicAdd(mismatches, bad);
}

// Self-test of pixel_round_margin(): (1) the relative error of rcp.approx.ftz.f32 over every mantissa of [1, 2) (the
// exponent is handled exactly), in units of 2^-24; (2) |u' - u| between the fast pixel coordinate and the reference's
// (demo.cu:46-47) on pseudo-random (cam, z) pairs whose pixel coordinate covers [-1.5, size + 1.5] and whose depth spans
// 2^-6 .. 2^7, as a fraction of the margin.  out[0] = max of (1) as float bits, out[1] = max of (2) as float bits.
__global__ void __launch_bounds__(256) selftest_margin_kernel(unsigned* __restrict__ out, float fx, float c0, float size,
                                                              unsigned n_per_thread) {
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x, n_thr = gridDim.x * blockDim.x;
  float worst_rcp = 0.0f, worst_u = 0.0f;
  for (unsigned m = gid; m < (1u << 23); m += n_thr) {
    const float x = __uint_as_float(0x3f800000u | m);
    const double exact = 1.0 / (double)x;
    worst_rcp = fmaxf(worst_rcp, (float)(fabs((double)rcp_approx(x) - exact) / exact * 16777216.0));
  }
  const float margin = pixel_round_margin(size, c0);
  unsigned long long state = 0xD1B54A32D192ED03ull * ((unsigned long long)gid + 1ull);
  for (unsigned i = 0; i < n_per_thread; ++i) {
    state += 0x9E3779B97F4A7C15ull;
    unsigned long long z64 = state;
    z64 = (z64 ^ (z64 >> 30)) * 0xBF58476D1CE4E5B9ull;
    z64 = (z64 ^ (z64 >> 27)) * 0x94D049BB133111EBull;
    z64 ^= z64 >> 31;
    const float u_t = -1.5f + (size + 3.0f) * ((float)(unsigned)(z64 >> 40) * (1.0f / 16777216.0f));
    const float zm = 1.0f + (float)(unsigned)((z64 >> 16) & 0xffffffu) * (1.0f / 16777216.0f);   // [1, 2)
    const float z = zm * __uint_as_float(((unsigned)(z64 & 15u) + 121u) << 23);                 // x 2^-6 .. 2^9
    const float cam = (u_t - c0) * z / fx;
    const float u_ref = __fadd_rn(__fmul_rn(fx, __fdiv_rn(cam, z)), c0);                         // demo.cu:46
    const float u_fast = __fmaf_rn(cam, __fmul_rn(rcp_approx(z), fx), c0);                       // voxel_classify_pair
    if (u_fast >= -1.5f && u_fast <= size + 1.5f) worst_u = fmaxf(worst_u, fabsf(u_fast - u_ref) / margin);
  }
  atomicMax(out + 0, __float_as_uint(worst_rcp));   // (non-negative floats order like their bit patterns)
  atomicMax(out + 1, __float_as_uint(worst_u));
}

int g_resident_blocks[64] = {};

}  // namespace

bool stream_prepare_geometry(StreamLaunch* L) {
  if (L->dim_x % VEC != 0 || !(L->trunc_margin > 0.0f) || (int64_t)L->im_w * L->im_h >= (1 << 22)) return false;
  if (L->chunk < 1 || L->chunk > kChunkMax) return false;
  if (L->epoch < 1) L->epoch = 16;   // (8 / 11 / 16 / 32: 79.4 / 78.9 / 78.5 / 78.8 ms on the 256-frame run)
  L->plan_beta = 3;   // (calibrated on the 1024^3 workload, profiles/README.md)
  if (const char* e = getenv("TSDF_PLAN_BETA")) {
    const int v = atoi(e);
    if (v >= 0 && v <= 64) L->plan_beta = v;
  }
  if (const char* e = getenv("TSDF_STREAM_EPOCH")) {   // tuning / A-B measurements
    const int v = atoi(e);
    if (v >= 1 && v <= kChunkMax) L->epoch = v;
  }
  L->nx_units = L->dim_x / VEC;
  // threads of a warp along x: the power of two in {8, 16, 32} that wastes the fewest lanes (8 preferred, see
  // choose_lx in tsdf_kernels.cu)
  int best_lx = 8;
  int64_t best_waste = -1;
  for (int lx = 8; lx <= 32; lx <<= 1) {
    const int64_t waste = (int64_t)((L->nx_units + lx - 1) / lx) * lx - L->nx_units;
    if (best_waste < 0 || waste < best_waste) { best_waste = waste; best_lx = lx; }
  }
  int shift = 0;
  while ((1 << (shift + 1)) <= best_lx) ++shift;
  L->lx_shift = shift;
  const int rows = kBlockThreads / best_lx;
  L->n_xt = (L->nx_units + best_lx - 1) / best_lx;
  L->n_yb = (L->dim_y + rows - 1) / rows;
  L->n_zg = (L->nz + ZT - 1) / ZT;
  const int64_t nb = (int64_t)L->n_xt * L->n_yb * L->n_zg;
  if (nb > INT32_MAX / 16) return false;
  L->n_btiles = (int)nb;
  L->lim_x = 0.5f - pixel_round_margin((float)L->im_w, L->cx);
  L->lim_y = 0.5f - pixel_round_margin((float)L->im_h, L->cy);
  return true;
}

// The shipped library carries ONE instantiation: 5 resident blocks per SM (48 registers), phase B pooled over the
// block.  The validation library additionally carries the variants measured in profiles/README.md, selected through the
// environment: TSDF_STREAM_MINB = 4 | 6 (64 / 40 registers), TSDF_STREAM_GROUPS = 2 | 4 | 8 (phase B pooled over
// 128 / 64 / 32 threads), TSDF_STREAM_DIAG = 1 | 2 (see stream_integrate_kernel).
struct StreamVariant { int minb, groups, diag; };

StreamVariant stream_variant() {
  static StreamVariant v = {0, 1, 0};
  if (v.minb == 0) {
    v.minb = 5;
#ifdef TSDF_VALIDATION
    if (const char* e = getenv("TSDF_STREAM_MINB")) { const int x = atoi(e); if (x == 4 || x == 6) v.minb = x; }
    if (const char* e = getenv("TSDF_STREAM_GROUPS")) { const int x = atoi(e); if (x == 2 || x == 4 || x == 8) v.groups = x; }
    if (const char* e = getenv("TSDF_STREAM_DIAG")) { const int x = atoi(e); if (x == 1 || x == 2) v.diag = x; }
#endif
  }
  return v;
}

using StreamKernel = void (*)(const StreamLaunch);

StreamKernel stream_kernel() {
#ifdef TSDF_VALIDATION
  const StreamVariant v = stream_variant();
  if (v.diag == 1) return stream_integrate_kernel<5, 1>;
  if (v.diag == 2) return stream_integrate_kernel<5, 2>;
  if (v.groups == 2) return stream_integrate_kernel<5, 0, 2>;
  if (v.groups == 4) return stream_integrate_kernel<5, 0, 4>;
  if (v.groups == 8) return stream_integrate_kernel<5, 0, 8>;
  if (v.minb == 4) return stream_integrate_kernel<4>;
  if (v.minb == 6) return stream_integrate_kernel<6>;
#endif
  return stream_integrate_kernel<5>;
}

int stream_resident_blocks(int device) {
  if (device < 0 || device >= 64) return 0;
  if (g_resident_blocks[device] == 0) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream_kernel(), kBlockThreads, 0) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
      return 0;
    g_resident_blocks[device] = per_sm * sms;
  }
  return g_resident_blocks[device];
}

cudaError_t launch_stream_masks(const StreamLaunch& L, cudaStream_t s) {
  if (L.n_jobs < 1 || L.n_jobs * L.n_chunks > kStreamMaxSegments) return cudaErrorInvalidValue;
  const int64_t warps = (int64_t)L.n_jobs * L.n_btiles;
  const int64_t blocks = (warps * 32 + 255) / 256;
  if (blocks > INT32_MAX) return cudaErrorInvalidConfiguration;
  stream_mask_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(L);
  stream_segment_kernel<<<1, kStreamMaxSegments, 0, s>>>(L);
  return cudaGetLastError();
}

cudaError_t launch_selftest_packed(unsigned long long* mismatches, cudaStream_t s) {
  selftest_packed_kernel<<<148 * 8, 256, 0, s>>>(mismatches, 4096);   // 2.5e9 lanes
  return cudaGetLastError();
}

cudaError_t launch_selftest_margin(unsigned* out2, float fx, float c0, float size, cudaStream_t s) {
  selftest_margin_kernel<<<148 * 8, 256, 0, s>>>(out2, fx, c0, size, 8192);   // 2.5e9 pairs
  return cudaGetLastError();
}

cudaError_t launch_stream_cost(const StreamLaunch& L, cudaStream_t s) {
  if (L.n_jobs < 1) return cudaErrorInvalidValue;
  const int64_t blocks = ((int64_t)L.n_jobs * L.n_btiles * 32 + 255) / 256;
  if (blocks > INT32_MAX) return cudaErrorInvalidConfiguration;
  stream_mask_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(L);
  return cudaGetLastError();
}

cudaError_t launch_stream_integrate(const StreamLaunch& L, int n_blocks, cudaStream_t s) {
  if (n_blocks < 1) return cudaErrorInvalidValue;
  stream_kernel()<<<n_blocks, kBlockThreads, 0, s>>>(L);
  return cudaGetLastError();
}

}  // namespace tsdf
