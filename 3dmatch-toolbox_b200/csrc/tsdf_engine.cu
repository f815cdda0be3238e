// tsdf_engine.cu -- C-ABI implementation (include/tsdf_fusion.h): engine handle, host<->device
// staging, launch sequencing.  All arithmetic lives in tsdf_kernels.cu (device) and
// tsdf_host.cpp (pose algebra / file formats); there is no CPU fallback for the compute path.
#include "tsdf_engine_internal.h"
#include "tsdf_host.h"

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

thread_local char g_err[512] = "";

}  // namespace

int tsdf_errf(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

void tsdf_set_error_text(const char* text) { snprintf(g_err, sizeof g_err, "%s", text); }

namespace {

int use_device(const tsdf_volume* v) {
  CK(cudaSetDevice(v->d.device));
  return TSDF_OK;
}

// Margins for the conservative frustum culls of integrate_kernel.  delta bounds the distance
// between any voxel's fp32 camera coordinates (as the kernel computes them) and the exact affine
// map of its position; the cull half-spaces are pushed out by a multiple of it.
void make_cull_margins(const tsdf_volume* v, const tsdf::FramePose* poses, int n, tsdf::IntegrateParams* p) {
  float rmax = 1.f, tmax = 0.f;
  bool finite = true;
  for (int f = 0; f < n; ++f) {
    const float* m = &poses[f].m0;
    for (int i = 0; i < 9; ++i) { rmax = fmaxf(rmax, fabsf(m[i])); finite = finite && std::isfinite(m[i]); }
    for (int i = 9; i < 12; ++i) { tmax = fmaxf(tmax, fabsf(m[i])); finite = finite && std::isfinite(m[i]); }
  }
  const float coord = v->max_abs_corner + tmax;
  const float delta = 64.f * FLT_EPSILON * 3.f * rmax * (coord + 1.f);
  if (!finite || !std::isfinite(delta)) {
    p->cull_level = 0;
    p->cull_z = p->cull_gx = p->cull_gy = 0.f;
    return;
  }
  p->cull_z = 4.f * delta;
  p->cull_gx = 4.f * delta * (fabsf(p->fx) + fabsf(p->cx) + (float)p->im_w + 4.f);
  p->cull_gy = 4.f * delta * (fabsf(p->fy) + fabsf(p->cy) + (float)p->im_h + 4.f);
}

void fill_params(const tsdf_volume* v, tsdf::IntegrateParams* p) {
  memset(p, 0, sizeof *p);
  p->tsdf = v->tsdf;
  p->weight = v->weight;
  p->frame_stride = v->frame_px;
  p->dim_x = (int)v->d.dim_x;
  p->dim_y = (int)v->d.dim_y;
  p->nz = (int)v->nz;
  p->z_begin = (int)v->d.z_begin;
  p->ox = v->d.origin_x; p->oy = v->d.origin_y; p->oz = v->d.origin_z;
  p->voxel_size = v->d.voxel_size;
  p->trunc_margin = v->d.trunc_margin;
  p->fx = v->d.cam_K[0]; p->cx = v->d.cam_K[2]; p->fy = v->d.cam_K[4]; p->cy = v->d.cam_K[5];
  p->im_w = v->d.im_width; p->im_h = v->d.im_height;
  p->cull_level = v->cull;
  p->fast_level = (v->tuning & 0x100) ? 0 : 1;
  p->variant = v->tuning & 0xff;
  p->force_lx = (v->tuning >> 12) & 0xff;
  p->no_ptx = (v->tuning >> 20) & 1;
  p->no_interior = (v->tuning >> 21) & 1;
  p->no_packed = (v->tuning >> 22) & 1;
}

void pose_from_cam2base(const float* c, tsdf::FramePose* o) {
  o->m0 = c[0]; o->m1 = c[1]; o->m2 = c[2];
  o->m4 = c[4]; o->m5 = c[5]; o->m6 = c[6];
  o->m8 = c[8]; o->m9 = c[9]; o->m10 = c[10];
  o->tx = c[3]; o->ty = c[7]; o->tz = c[11];
}

// Launch one pass over the slab for `n` (<= batch) frames whose (sanitized) depth sits in ring slot
// `slot`.  The producer of the slot ran on v->stream; the sub-slab kernels fork from there onto the
// sub-streams and are joined back only by join_substreams() at the end of an API call.
int launch_chunk(tsdf_volume* v, int slot, const float* cam2base, int n) {
  tsdf::IntegrateParams p;
  fill_params(v, &p);
  p.depth = v->dev_depth[slot];
  if (v->cull >= 2) {
    // max-pyramids of the (sanitized) frames of this slot, for the occlusion test of the frame masks
    CK(tsdf::launch_build_depth_mips(v->dev_depth[slot], v->dev_mips[slot], n, v->d.im_width, v->d.im_height,
                                     v->frame_px, v->stream));
    v->stats.kernel_launches += 1;
    p.mips = v->dev_mips[slot];
    p.mip_stride = v->mip.stride;
    for (int l = 0; l < 5; ++l) { p.mip_w[l] = v->mip.w[l]; p.mip_off[l] = v->mip.off[l]; }
  }
  p.sanitized = 1;  // every depth frame reaches the kernel through sanitize / depth_from_u16(sanitize)
  p.n_frames = n;
  for (int f = 0; f < n; ++f) pose_from_cam2base(cam2base + 16 * f, &p.pose[f]);
  make_cull_margins(v, p.pose, n, &p);
  {
    // planes further from the cameras (along z) intersect more of the frusta and cost more: start there
    float tz = 0.f;
    for (int f = 0; f < n; ++f) tz += p.pose[f].tz;
    const float mid = v->d.origin_z + 0.5f * (float)(v->d.z_begin + v->d.z_end) * v->d.voxel_size;
    p.z_reverse = ((tz / (float)n) < mid) ? 1 : 0;
    if (v->tuning & 0x200) p.z_reverse = !p.z_reverse;
  }
  const int64_t plane = v->d.dim_x * v->d.dim_y;
  // frame masks of the whole slab, on the main stream: they are ready before the sub-slab streams get to this
  // chunk (the main stream runs ahead of the fusion kernels by up to kSlots - 1 chunks)
  int64_t tiles_per_zgroup = 0;
  const bool want_masks = v->cull >= 1 && !(v->tuning & (1 << 23));
  if (want_masks) {
    if (!v->slot_masks[slot]) {
      v->slot_mask_cap = tsdf::tile_mask_entries(p.dim_x, p.dim_y, p.nz);
      CK(cudaMalloc(&v->slot_masks[slot], (size_t)v->slot_mask_cap * sizeof(uint2)));
    }
    tsdf::IntegrateParams m = p;
    m.tile_masks = v->slot_masks[slot];
    m.tile_mask_cap = v->slot_mask_cap;
    tiles_per_zgroup = tsdf::launch_tile_masks(m, v->stream);
    if (tiles_per_zgroup > 0) v->stats.kernel_launches += 1;
  }
  CK(cudaEventRecord(v->ev_ready[slot], v->stream));
  for (int i = 0; i < v->n_sub; ++i) {
    tsdf::IntegrateParams q = p;
    q.tsdf = v->tsdf + v->sub_z[i] * plane;
    q.weight = v->weight + v->sub_z[i] * plane;
    q.nz = (int)(v->sub_z[i + 1] - v->sub_z[i]);
    q.z_begin = (int)(v->d.z_begin + v->sub_z[i]);
    cudaStream_t st = (v->n_sub == 1) ? v->stream : v->sub_stream[i];
    if (v->n_sub > 1) CK(cudaStreamWaitEvent(st, v->ev_ready[slot], 0));
    if (tiles_per_zgroup > 0) {
      q.tile_masks = v->slot_masks[slot] + (v->sub_z[i] / 8) * tiles_per_zgroup;  // sub-slabs start at multiples of 8 planes
      q.tile_mask_cap = v->slot_mask_cap - (v->sub_z[i] / 8) * tiles_per_zgroup;
      q.masks_ready = 1;
    } else if (want_masks) {
      // scratch for the frame-mask pre-pass; passes of one sub-slab are ordered on its stream, so one buffer each
      if (!v->tile_masks[i]) {
        v->tile_mask_cap[i] = tsdf::tile_mask_entries(q.dim_x, q.dim_y, q.nz);
        CK(cudaMalloc(&v->tile_masks[i], (size_t)v->tile_mask_cap[i] * sizeof(uint2)));
      }
      q.tile_masks = v->tile_masks[i];
      q.tile_mask_cap = v->tile_mask_cap[i];
    }
    CK(tsdf::launch_integrate(q, st));
    CK(cudaEventRecord(v->ev_join[slot][i], st));
    v->stats.kernel_launches += tsdf::last_integrate_launches();
  }
  v->slot_busy[slot] = true;
  v->last_slot = slot;
  v->stats.frames_integrated += n;
  v->n_points = -1;
  return TSDF_OK;
}

// v->stream (and through it every later API call) waits for the sub-slab kernels launched so far.
int join_substreams(tsdf_volume* v) {
  if (v->n_sub > 1 && v->last_slot >= 0)
    for (int i = 0; i < v->n_sub; ++i) CK(cudaStreamWaitEvent(v->stream, v->ev_join[v->last_slot][i], 0));
  return TSDF_OK;
}

// Before ring slot `slot` is overwritten: `stream` waits for the kernels that still read it.
int wait_slot_free(tsdf_volume* v, int slot, cudaStream_t stream, bool also_host) {
  if (!v->slot_busy[slot]) return TSDF_OK;
  for (int i = 0; i < v->n_sub; ++i) {
    CK(cudaStreamWaitEvent(stream, v->ev_join[slot][i], 0));
    if (also_host) CK(cudaEventSynchronize(v->ev_join[slot][i]));
  }
  return TSDF_OK;
}

int ensure_ring(tsdf_volume* v, bool u16) {
  const size_t px = (size_t)v->frame_px * v->batch;
  for (int s = 0; s < kSlots; ++s) {
    if (!v->dev_depth[s]) CK(cudaMalloc(&v->dev_depth[s], px * sizeof(float)));
    if (!v->dev_mips[s]) CK(cudaMalloc(&v->dev_mips[s], (size_t)v->mip.stride * v->batch * sizeof(float)));
    if (!v->ev_copied[s]) CK(cudaEventCreateWithFlags(&v->ev_copied[s], cudaEventDisableTiming));
    if (!v->ev_ready[s]) CK(cudaEventCreateWithFlags(&v->ev_ready[s], cudaEventDisableTiming));
    for (int i = 0; i < kMaxSub; ++i)
      if (!v->ev_join[s][i]) CK(cudaEventCreateWithFlags(&v->ev_join[s][i], cudaEventDisableTiming));
    if (u16 && !v->dev_mm[s]) CK(cudaMalloc(&v->dev_mm[s], px * sizeof(uint16_t)));
  }
  if (!v->copy_stream) CK(cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking));
  return TSDF_OK;
}

int ensure_pinned(tsdf_volume* v, bool u16) {
  const size_t px = (size_t)v->frame_px * v->batch;
  for (int s = 0; s < kSlots; ++s) {
    if (!u16 && !v->pin_depth[s]) CK(cudaHostAlloc(&v->pin_depth[s], px * sizeof(float), cudaHostAllocDefault));
    if (u16 && !v->pin_mm[s]) CK(cudaHostAlloc(&v->pin_mm[s], px * sizeof(uint16_t), cudaHostAllocDefault));
  }
  return TSDF_OK;
}

enum class MemKind { Device, PinnedHost, PageableHost };

MemKind classify(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return MemKind::PageableHost;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return MemKind::Device;
  if (a.type == cudaMemoryTypeHost) return MemKind::PinnedHost;
  return MemKind::PageableHost;
}

// Host-resident input: K-frame chunks are copied on copy_stream into a ring of device buffers
// while the previous chunk is being fused on the compute stream.
template <typename Src>
int integrate_from_host(tsdf_volume* v, int n_frames, const Src* src, const float* cam2base, MemKind kind) {
  constexpr bool u16 = sizeof(Src) == 2;
  int rc = ensure_ring(v, u16);
  if (rc) return rc;
  if (kind == MemKind::PageableHost && (rc = ensure_pinned(v, u16))) return rc;
  for (int f0 = 0; f0 < n_frames; f0 += v->batch) {
    const int n = (n_frames - f0 < v->batch) ? n_frames - f0 : v->batch;
    const int s = (int)(v->ring_pos++ % kSlots);
    const size_t elems = (size_t)n * v->frame_px;
    const Src* chunk = src + (size_t)f0 * v->frame_px;
    // the pinned staging buffer and the device slot are reused: wait for their last consumers
    if ((rc = wait_slot_free(v, s, v->copy_stream, kind == MemKind::PageableHost))) return rc;
    const void* from = chunk;
    if (kind == MemKind::PageableHost) {
      void* pin = u16 ? (void*)v->pin_mm[s] : (void*)v->pin_depth[s];
      memcpy(pin, chunk, elems * sizeof(Src));
      from = pin;
    }
    void* to = u16 ? (void*)v->dev_mm[s] : (void*)v->dev_depth[s];
    CK(cudaMemcpyAsync(to, from, elems * sizeof(Src), cudaMemcpyHostToDevice, v->copy_stream));
    v->stats.h2d_bytes += (int64_t)(elems * sizeof(Src));
    CK(cudaEventRecord(v->ev_copied[s], v->copy_stream));
    CK(cudaStreamWaitEvent(v->stream, v->ev_copied[s], 0));
    if (u16)
      CK(tsdf::launch_depth_from_u16(v->dev_mm[s], v->dev_depth[s], (int64_t)elems, true, v->stream));
    else
      CK(tsdf::launch_sanitize_depth(v->dev_depth[s], v->dev_depth[s], (int64_t)elems, v->stream));
    v->stats.kernel_launches += 1;
    rc = launch_chunk(v, s, cam2base + 16 * (size_t)f0, n);
    if (rc) return rc;
  }
  return join_substreams(v);
}


// ================================================================================================================
// Stream engine: groups of frames fused by ONE persistent work-list kernel each (tsdf_stream_kernels.cu).
// ================================================================================================================

template <typename T>
int grow(T** ptr, int64_t* cap, int64_t need, bool pinned = false) {
  if (need <= *cap && *ptr) return TSDF_OK;
  if (*ptr) {
    if (pinned) cudaFreeHost(*ptr); else cudaFree(*ptr);   // (cudaFree waits for the work that may still use it)
    *ptr = nullptr;
    *cap = 0;
  }
  if (pinned) CK(cudaHostAlloc((void**)ptr, (size_t)need * sizeof(T), cudaHostAllocDefault));
  else CK(cudaMalloc((void**)ptr, (size_t)need * sizeof(T)));
  *cap = need;
  return TSDF_OK;
}

// Conservative margins of the frustum / occlusion tests for a set of poses (see make_cull_margins).
void stream_cull_margins(float max_abs_corner, const tsdf::FramePose* poses, int64_t n, tsdf::StreamLaunch* L) {
  float rmax = 1.f, tmax = 0.f;
  bool finite = true;
  for (int64_t f = 0; f < n; ++f) {
    const float* m = &poses[f].m0;
    for (int i = 0; i < 9; ++i) { rmax = fmaxf(rmax, fabsf(m[i])); finite = finite && std::isfinite(m[i]); }
    for (int i = 9; i < 12; ++i) { tmax = fmaxf(tmax, fabsf(m[i])); finite = finite && std::isfinite(m[i]); }
  }
  const float coord = max_abs_corner + tmax;
  const float delta = 64.f * FLT_EPSILON * 3.f * rmax * (coord + 1.f);
  if (!finite || !std::isfinite(delta)) {
    L->cull_level = 0;
    L->cull_z = L->cull_gx = L->cull_gy = 0.f;
    return;
  }
  L->cull_z = fmaxf(L->cull_z, 4.f * delta);
  L->cull_gx = fmaxf(L->cull_gx, 4.f * delta * (fabsf(L->fx) + fabsf(L->cx) + (float)L->im_w + 4.f));
  L->cull_gy = fmaxf(L->cull_gy, 4.f * delta * (fabsf(L->fy) + fabsf(L->cy) + (float)L->im_h + 4.f));
}

int stream_init(tsdf_volume* v) {
  tsdf::StreamLaunch& L = v->sl;
  memset(&L, 0, sizeof L);
  L.frame_stride = v->frame_px;
  L.mip_stride = v->mip.stride;
  for (int l = 0; l < 5; ++l) { L.mip_w[l] = v->mip.w[l]; L.mip_off[l] = v->mip.off[l]; }
  L.dim_x = (int)v->d.dim_x; L.dim_y = (int)v->d.dim_y; L.nz = (int)v->nz;
  L.ox = v->d.origin_x; L.oy = v->d.origin_y; L.oz = v->d.origin_z;
  L.voxel_size = v->d.voxel_size; L.trunc_margin = v->d.trunc_margin;
  L.fx = v->d.cam_K[0]; L.cx = v->d.cam_K[2]; L.fy = v->d.cam_K[4]; L.cy = v->d.cam_K[5];
  L.im_w = v->d.im_width; L.im_h = v->d.im_height;
  L.chunk = v->batch;
  v->stream_ok = tsdf::stream_prepare_geometry(&L);
  // float4 accesses of the slab rows need 16-byte aligned arrays (caller-owned slabs may be offset views)
  if (((reinterpret_cast<uintptr_t>(v->tsdf) | reinterpret_cast<uintptr_t>(v->weight)) & 15u) != 0) v->stream_ok = false;
  return TSDF_OK;
}

int stream_slot_events(StreamSlot& S) {
  if (!S.ev_copied) CK(cudaEventCreateWithFlags(&S.ev_copied, cudaEventDisableTiming));
  if (!S.ev_staged) CK(cudaEventCreateWithFlags(&S.ev_staged, cudaEventDisableTiming));
  if (!S.ev_done) CK(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
  return TSDF_OK;
}

// Everything a group needs besides the sanitized frames themselves (already in S.depth, produced on `st`):
// poses on the device, depth pyramids, zeroed counters / progress words.  Fills *job with the device pointers.
int stream_prepare_job(tsdf_volume* v, StreamSlot& S, int ng, const float* cam2base, tsdf::StreamJob* job,
                       tsdf::StreamLaunch* L, cudaStream_t st) {
  const int n_chunks = (ng + L->chunk - 1) / L->chunk;
  if (int rc = grow(&S.poses_dev, &S.poses_cap, (int64_t)ng)) return rc;
  if (S.poses_pin_cap < ng) {
    if (S.staged) CK(cudaEventSynchronize(S.ev_staged));
    if (int rc = grow(&S.poses_pin, &S.poses_pin_cap, (int64_t)ng, true)) return rc;
  }
  if (v->cull >= 2)
    if (int rc = grow(&S.mips, &S.mips_cap, (int64_t)ng * v->mip.stride)) return rc;
  if (int rc = grow(&S.masks, &S.masks_cap, (int64_t)n_chunks * L->n_btiles * 8)) return rc;
  if (int rc = grow(&S.items, &S.items_cap, (int64_t)n_chunks * L->n_btiles)) return rc;
  if (!S.counters) CK(cudaMalloc(&S.counters, kCounterWords * sizeof(unsigned)));
  if (!S.progress) CK(cudaMalloc(&S.progress, (size_t)L->n_btiles * sizeof(unsigned)));
  if (!v->zcost) {
    CK(cudaMalloc(&v->zcost, (size_t)L->n_zg * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(v->zcost, 0, (size_t)L->n_zg * sizeof(unsigned long long), st));
  }
  if (!v->rec_scratch) {
    v->resident_blocks = tsdf::stream_resident_blocks(v->d.device);
    if (v->resident_blocks <= 0) return set_err(TSDF_ERR_CUDA, "occupancy query of the fusion kernel failed");
    CK(cudaMalloc(&v->rec_scratch, (size_t)v->resident_blocks * 2 * 8192 * sizeof(unsigned)));
  }
  // the pinned pose buffer may still be read by the previous group's copy
  if (S.staged) CK(cudaEventSynchronize(S.ev_staged));
  for (int f = 0; f < ng; ++f) pose_from_cam2base(cam2base + 16 * (size_t)f, &S.poses_pin[f]);
  stream_cull_margins(v->max_abs_corner, S.poses_pin, ng, L);
  // (kernels, not the copy engine: see launch_small_copy)
  CK(tsdf::launch_small_copy(S.poses_dev, S.poses_pin, (size_t)ng * sizeof(tsdf::FramePose), st));
  CK(tsdf::launch_small_copy(S.counters, nullptr, kCounterWords * sizeof(unsigned), st));
  CK(tsdf::launch_small_copy(S.progress, nullptr, (size_t)L->n_btiles * sizeof(unsigned), st));
  v->stats.kernel_launches += 3;
  if (v->cull >= 2) {
    CK(tsdf::launch_build_depth_mips(S.depth, S.mips, ng, v->d.im_width, v->d.im_height, v->frame_px, st));
    v->stats.kernel_launches += 1;
  }
  memset(job, 0, sizeof *job);
  job->tsdf = v->tsdf;
  job->weight = v->weight;
  job->depth = S.depth;
  job->mips = (v->cull >= 2) ? S.mips : nullptr;
  job->poses = S.poses_dev;
  job->masks = S.masks;
  job->items = S.items;
  job->item_count = S.counters;
  job->progress = S.progress;
  job->zcost = v->zcost;
  job->z_begin = (int)v->d.z_begin;
  return TSDF_OK;
}

// Launch parameters of a group of `ng` frames for volumes shaped like `v`.
void stream_group_params(const tsdf_volume* v, int ng, tsdf::StreamLaunch* L) {
  *L = v->sl;
  L->n_frames = ng;
  L->n_chunks = (ng + L->chunk - 1) / L->chunk;
  L->chunk = (ng + L->n_chunks - 1) / L->n_chunks;   // chunks of equal length (50 frames: 25 + 25, not 32 + 18)
  L->n_chunks = (ng + L->chunk - 1) / L->chunk;
  if (L->epoch > 0) {   // epochs of equal length too (a 25-frame chunk: 7 + 6 + 6 + 6, not 8 + 8 + 8 + 1)
    const int n_epochs = (L->chunk + L->epoch - 1) / L->epoch;
    L->epoch = (L->chunk + n_epochs - 1) / n_epochs;
  }
  L->cull_level = v->cull;
}

// Frame masks + item lists + the persistent fusion kernel for the jobs described by L (device array L->jobs).
int stream_launch(tsdf_volume* leader, StreamSlot& S, tsdf::StreamLaunch* L, cudaStream_t st) {
  L->seg_begin = S.counters + tsdf::kStreamMaxSegments;
  L->ticket = S.counters + 2 * tsdf::kStreamMaxSegments + 4;
  L->rec_scratch = leader->rec_scratch;
  CK(tsdf::launch_stream_masks(*L, st));
  static const bool l2_window = getenv("TSDF_L2_WINDOW") && atoi(getenv("TSDF_L2_WINDOW")) == 1;
  if (l2_window && L->n_jobs == 1) {
    // A/B switch for profiles/: an L2-persisting access-policy window over the group's depth frames (BASELINE
    // north_star (2)).  The window cannot follow the chunk the kernel is working on, so with more frames than fit the
    // persisting carve-out only a hitRatio fraction of the lines is pinned.
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, leader->d.device));
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize));
    size_t bytes = (size_t)L->n_frames * L->frame_stride * sizeof(float);
    if (bytes > (size_t)prop.accessPolicyMaxWindowSize) bytes = (size_t)prop.accessPolicyMaxWindowSize;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = (void*)S.depth;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = bytes > 0 ? fminf(1.0f, (float)prop.persistingL2CacheMaxSize / (float)bytes) : 0.f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
  }
  CK(tsdf::launch_stream_integrate(*L, leader->resident_blocks, st));
  leader->stats.kernel_launches += 3;
  return TSDF_OK;
}

int stream_job_array(StreamSlot& S, int n) {
  if (S.jobs_cap >= n) return TSDF_OK;
  if (S.staged) CK(cudaEventSynchronize(S.ev_staged));
  if (S.job_dev) { cudaFree(S.job_dev); S.job_dev = nullptr; }
  if (S.job_pin) { cudaFreeHost(S.job_pin); S.job_pin = nullptr; }
  S.jobs_cap = 0;
  CK(cudaMalloc(&S.job_dev, (size_t)n * sizeof(tsdf::StreamJob)));
  CK(cudaHostAlloc((void**)&S.job_pin, (size_t)n * sizeof(tsdf::StreamJob), cudaHostAllocDefault));
  S.jobs_cap = n;
  return TSDF_OK;
}

// One group, one volume: S.depth holds the sanitized frames (work queued on v->stream).
int stream_fuse_group(tsdf_volume* v, StreamSlot& S, int ng, const float* cam2base) {
  tsdf::StreamLaunch L;
  stream_group_params(v, ng, &L);
  if (int rc = stream_job_array(S, 1)) return rc;
  tsdf::StreamJob job;
  if (int rc = stream_prepare_job(v, S, ng, cam2base, &job, &L, v->stream)) return rc;
  {
    // planes further from the cameras (along z) intersect more of the frusta and cost more: they enter the lists first
    float tz = 0.f;
    for (int f = 0; f < ng; ++f) tz += S.poses_pin[f].tz;
    const float mid = v->d.origin_z + 0.5f * (float)(v->d.z_begin + v->d.z_end) * v->d.voxel_size;
    L.z_reverse = ((tz / (float)ng) < mid) ? 1 : 0;
  }
  *S.job_pin = job;
  CK(tsdf::launch_small_copy(S.job_dev, S.job_pin, sizeof job, v->stream));
  v->stats.kernel_launches += 1;
  CK(cudaEventRecord(S.ev_staged, v->stream));
  S.staged = true;
  L.jobs = S.job_dev;
  L.n_jobs = 1;
  if (int rc = stream_launch(v, S, &L, v->stream)) return rc;
  CK(cudaEventRecord(S.ev_done, v->stream));
  S.busy = true;
  v->stats.frames_integrated += ng;
  v->n_points = -1;
  return TSDF_OK;
}

// Sanitized copy of `ng` frames into S.depth.  Device input: one kernel on v->stream.  Host input: H2D on the copy
// stream (through the slot's pinned buffer when the source is pageable), then the kernel.
template <typename Src>
int stream_stage_frames(tsdf_volume* v, StreamSlot& S, const Src* src, int ng, MemKind kind) {
  constexpr bool u16 = sizeof(Src) == 2;
  const int64_t elems = (int64_t)ng * v->frame_px;
  if (int rc = stream_slot_events(S)) return rc;
  if (int rc = grow(&S.depth, &S.depth_cap, elems)) return rc;
  if (kind == MemKind::Device) {
    if (u16) CK(tsdf::launch_depth_from_u16(reinterpret_cast<const uint16_t*>(src), S.depth, elems, true, v->stream));
    else CK(tsdf::launch_sanitize_depth(reinterpret_cast<const float*>(src), S.depth, elems, v->stream));
    v->stats.kernel_launches += 1;
    return TSDF_OK;
  }
  if (!v->copy_stream) CK(cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking));
  if (u16)
    if (int rc = grow(&S.mm, &S.mm_cap, elems)) return rc;
  // the slot's buffers are overwritten: its previous group must have been fused
  if (S.busy) CK(cudaStreamWaitEvent(v->copy_stream, S.ev_done, 0));
  const void* from = src;
  if (kind == MemKind::PageableHost) {
    const size_t bytes = (size_t)elems * sizeof(Src);
    if (S.pin_cap < bytes) {
      if (S.pin) { CK(cudaStreamSynchronize(v->copy_stream)); cudaFreeHost(S.pin); S.pin = nullptr; S.pin_cap = 0; }
      CK(cudaHostAlloc(&S.pin, bytes, cudaHostAllocDefault));
      S.pin_cap = bytes;
    } else if (S.busy) {
      CK(cudaEventSynchronize(S.ev_copied));  // the previous H2D out of this pinned buffer
    }
    memcpy(S.pin, src, bytes);
    from = S.pin;
  }
  void* to = u16 ? (void*)S.mm : (void*)S.depth;
  CK(cudaMemcpyAsync(to, from, (size_t)elems * sizeof(Src), cudaMemcpyHostToDevice, v->copy_stream));
  v->stats.h2d_bytes += elems * (int64_t)sizeof(Src);
  CK(cudaEventRecord(S.ev_copied, v->copy_stream));
  CK(cudaStreamWaitEvent(v->stream, S.ev_copied, 0));
  if (u16) CK(tsdf::launch_depth_from_u16(S.mm, S.depth, elems, true, v->stream));
  else CK(tsdf::launch_sanitize_depth(S.depth, S.depth, elems, v->stream));
  v->stats.kernel_launches += 1;
  return TSDF_OK;
}

template <typename Src>
int stream_integrate(tsdf_volume* v, int n_frames, const Src* src, const float* cam2base, MemKind kind) {
  const int seg_cap = tsdf::kStreamMaxSegments * v->sl.chunk;
  int group_max = (kind == MemKind::Device) ? kStreamGroupDevice : kStreamGroupHost;
  if (group_max > seg_cap) group_max = seg_cap;
  // host input: the first group is small so that the fusion starts after a short upload, the following ones grow (fewer
  // launch tails); every group's upload overlaps the fusion of the previous one.  (Measured, scripts/e2e_breakdown.py,
  // 1024^3 x 1000 frames: device input 275.1 ms, device input in this group schedule 276.8 ms, pinned host input
  // 298.4 ms; growing by 7/4 from 16 frames instead of 3 from 32 changes nothing: 299.7 ms.)
  int group = (kind == MemKind::Device) ? group_max : (group_max < 32 ? group_max : 32);
  for (int g0 = 0, ng = 0; g0 < n_frames; g0 += ng) {
    ng = (n_frames - g0 < group) ? n_frames - g0 : group;
    group = (group * 3 < group_max) ? group * 3 : group_max;
    // device input is staged and fused in stream order, so one slot suffices; host input alternates two slots so
    // that the copy of group g + 1 overlaps the fusion of group g
    StreamSlot& S = v->ss[(kind == MemKind::Device) ? 0 : (int)(v->stream_group++ % kStreamSlots)];
    if (int rc = stream_stage_frames(v, S, src + (size_t)g0 * v->frame_px, ng, kind)) return rc;
    if (int rc = stream_fuse_group(v, S, ng, cam2base + 16 * (size_t)g0)) return rc;
  }
  return TSDF_OK;
}

}  // namespace

extern "C" {

const char* tsdf_last_error(void) { return g_err; }
const char* tsdf_version(void) { return "tsdf-fusion-b200 0.1 (sm_100a)"; }

int tsdf_device_count(int32_t* count) {
  if (!count) return set_err(TSDF_ERR_ARG, "count is NULL");
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  *count = n;
  return TSDF_OK;
}

int tsdf_create(const tsdf_volume_desc* desc, tsdf_volume** out) {
  if (!desc || !out) return set_err(TSDF_ERR_ARG, "desc/out is NULL");
  *out = nullptr;
  tsdf_volume_desc d = *desc;
  if (d.dim_x <= 0 || d.dim_y <= 0 || d.dim_z <= 0) return set_err(TSDF_ERR_ARG, "grid dimensions must be positive");
  if (d.dim_x > INT32_MAX / 2 || d.dim_y > INT32_MAX / 2 || d.dim_z > INT32_MAX / 2)
    return set_err(TSDF_ERR_ARG, "grid dimension too large");
  if (d.z_begin == 0 && d.z_end == 0) d.z_end = d.dim_z;
  if (d.z_begin < 0 || d.z_end > d.dim_z || d.z_begin >= d.z_end)
    return set_err(TSDF_ERR_ARG, "z slab [%lld, %lld) outside [0, %lld)", (long long)d.z_begin, (long long)d.z_end,
                   (long long)d.dim_z);
  if (d.im_width <= 0 || d.im_height <= 0) return set_err(TSDF_ERR_ARG, "image size must be positive");
  if ((int64_t)d.im_width * d.im_height > INT32_MAX) return set_err(TSDF_ERR_ARG, "image too large");
  if (d.max_batch < 0 || d.max_batch > TSDF_MAX_BATCH) return set_err(TSDF_ERR_ARG, "max_batch out of range");
  if ((d.ext_tsdf == nullptr) != (d.ext_weight == nullptr))
    return set_err(TSDF_ERR_ARG, "ext_tsdf and ext_weight must be given together");
  if (((reinterpret_cast<uintptr_t>(d.ext_tsdf) | reinterpret_cast<uintptr_t>(d.ext_weight)) & 15u) != 0)
    return set_err(TSDF_ERR_ARG, "ext_tsdf / ext_weight must be 16-byte aligned (rows are accessed as float4)");

  tsdf_volume* v = new (std::nothrow) tsdf_volume();
  if (!v) return set_err(TSDF_ERR_ARG, "out of host memory");
  v->d = d;
  v->nz = d.z_end - d.z_begin;
  v->n_vox = v->nz * d.dim_y * d.dim_x;
  v->frame_px = (int64_t)d.im_width * d.im_height;
  v->mip = tsdf::make_mip_layout(d.im_width, d.im_height);
  v->cull = (d.cull_level < 0 || d.cull_level > 2) ? 2 : d.cull_level;
#ifdef TSDF_VALIDATION
  v->tuning = d.reserved0;   // validation flavour of the library: kernel variants / earlier kernel generations
#else
  if (d.reserved0 != 0) { delete v; return set_err(TSDF_ERR_ARG, "reserved0 must be 0"); }
  v->tuning = 0;
#endif
  // TSDF_ENGINE=legacy (environment): the round-1 per-pass launch sequence instead of the persistent work-list kernel
  // (kept for A/B measurements; both are bit-identical)
  const char* eng = getenv("TSDF_ENGINE");
  const bool want_legacy = (eng && strcmp(eng, "legacy") == 0) || (v->tuning != 0);
  // K: frames per chunk of the stream engine (one 32-bit frame mask) / per pass of the legacy launch sequence
  v->batch = d.max_batch ? d.max_batch : (want_legacy ? 8 : 32);
  {
    const float ex = fmaxf(fabsf(d.origin_x), fabsf(d.origin_x + (float)d.dim_x * d.voxel_size));
    const float ey = fmaxf(fabsf(d.origin_y), fabsf(d.origin_y + (float)d.dim_y * d.voxel_size));
    const float ez = fmaxf(fabsf(d.origin_z), fabsf(d.origin_z + (float)d.dim_z * d.voxel_size));
    v->max_abs_corner = fmaxf(ex, fmaxf(ey, ez));
  }
  int rc = TSDF_OK;
  auto fail = [&](int code) { tsdf_destroy(v); return code; };
  cudaError_t e = cudaSetDevice(d.device);
  if (e != cudaSuccess)
    return fail(set_err(TSDF_ERR_CUDA, "cudaSetDevice(%d): %s (no CPU fallback exists)", d.device, cudaGetErrorString(e)));
  if (d.stream) {
    v->stream = (cudaStream_t)d.stream;
  } else {
    // The main stream carries the small per-chunk kernels (sanitize, depth pyramids, frame masks) that the sub-slab
    // fusion launches wait for.  It gets the HIGHEST priority: the block scheduler otherwise serves the thousands of
    // pending fusion blocks of the previous chunk first, the pre-pass of the next chunk only starts in their tail and
    // every pass pays its latency chain (~50-100 us, 20 % of a 1/8-slab pass).  tuning bit 25: equal priorities.
    int prio_lo = 0, prio_hi = 0;
    if (!(v->tuning & (1 << 25))) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    e = cudaStreamCreateWithPriority(&v->stream, cudaStreamNonBlocking, prio_hi);
    if (e != cudaSuccess) return fail(set_err(TSDF_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)));
    v->owns_stream = true;
  }
  if (d.ext_tsdf) {
    v->tsdf = d.ext_tsdf;
    v->weight = d.ext_weight;
  } else {
    e = cudaMalloc(&v->tsdf, (size_t)v->n_vox * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&v->weight, (size_t)v->n_vox * sizeof(float));
    if (e != cudaSuccess)
      return fail(set_err(TSDF_ERR_CUDA, "cudaMalloc of %lld voxels: %s", (long long)v->n_vox, cudaGetErrorString(e)));
    v->owns_volume = true;
  }
  {
    // sub-slabs: cut at multiples of 8 planes (the z extent of a thread's voxel patch)
    int want = (v->tuning >> 10) & 3;
    if (v->tuning & (1 << 24)) want = 4;
    if (want == 0) want = 3;   // measured on 1/8 slabs of the 1024^3 grid: 1 stream +38 %, 2 +25 %, 3 +19 % over the
                               // cost of the same planes inside one big launch
    if (v->nz < 64) want = 1;
    v->n_sub = want;
    for (int i = 0; i <= want; ++i) v->sub_z[i] = (i == want) ? v->nz : ((v->nz * i / want) / 8) * 8;
    for (int i = 0; i < want && want > 1; ++i) {
      e = cudaStreamCreateWithFlags(&v->sub_stream[i], cudaStreamNonBlocking);
      if (e != cudaSuccess) return fail(set_err(TSDF_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)));
    }
  }
  e = cudaEventCreate(&v->t0);
  if (e == cudaSuccess) e = cudaEventCreate(&v->t1);
  if (e != cudaSuccess) return fail(set_err(TSDF_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)));
  if (!want_legacy && (rc = stream_init(v)) != TSDF_OK) return fail(rc);
  if (v->owns_volume && (rc = tsdf_reset(v)) != TSDF_OK) return fail(rc);
  *out = v;
  return TSDF_OK;
}

void tsdf_destroy(tsdf_volume* v) {
  if (!v) return;
  cudaSetDevice(v->d.device);
  if (v->stream) cudaStreamSynchronize(v->stream);
  if (v->copy_stream) { cudaStreamSynchronize(v->copy_stream); cudaStreamDestroy(v->copy_stream); }
  for (int i = 0; i < kMaxSub; ++i)
    if (v->sub_stream[i]) { cudaStreamSynchronize(v->sub_stream[i]); cudaStreamDestroy(v->sub_stream[i]); }
  for (int s = 0; s < kSlots; ++s) {
    cudaFree(v->dev_depth[s]);
    cudaFree(v->dev_mips[s]);
    cudaFree(v->dev_mm[s]);
    if (v->pin_depth[s]) cudaFreeHost(v->pin_depth[s]);
    if (v->pin_mm[s]) cudaFreeHost(v->pin_mm[s]);
    if (v->ev_copied[s]) cudaEventDestroy(v->ev_copied[s]);
    if (v->ev_ready[s]) cudaEventDestroy(v->ev_ready[s]);
    for (int i = 0; i < kMaxSub; ++i)
      if (v->ev_join[s][i]) cudaEventDestroy(v->ev_join[s][i]);
  }
  for (StreamSlot& S : v->ss) {
    cudaFree(S.depth); cudaFree(S.mips); cudaFree(S.mm); cudaFree(S.poses_dev); cudaFree(S.masks); cudaFree(S.items);
    cudaFree(S.counters); cudaFree(S.progress); cudaFree(S.job_dev);
    if (S.pin) cudaFreeHost(S.pin);
    if (S.poses_pin) cudaFreeHost(S.poses_pin);
    if (S.job_pin) cudaFreeHost(S.job_pin);
    if (S.ev_copied) cudaEventDestroy(S.ev_copied);
    if (S.ev_staged) cudaEventDestroy(S.ev_staged);
    if (S.ev_done) cudaEventDestroy(S.ev_done);
  }
  cudaFree(v->rec_scratch);
  cudaFree(v->zcost);
  for (int i = 0; i < kMaxSub; ++i) cudaFree(v->tile_masks[i]);
  for (int s = 0; s < kSlots; ++s) cudaFree(v->slot_masks[s]);
  if (v->owns_volume) { cudaFree(v->tsdf); cudaFree(v->weight); }
  cudaFree(v->tile_offsets);
  cudaFree(v->xyz);
  cudaFree(v->vidx);
  cudaFree(v->normals);
  if (v->t0) cudaEventDestroy(v->t0);
  if (v->t1) cudaEventDestroy(v->t1);
  if (v->owns_stream && v->stream) cudaStreamDestroy(v->stream);
  cudaGetLastError();
  delete v;
}

int tsdf_reset(tsdf_volume* v) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (int rc = use_device(v)) return rc;
  CK(tsdf::launch_fill(v->tsdf, v->weight, v->n_vox, v->stream));
  v->stats.kernel_launches += 1;
  v->n_points = -1;
  return TSDF_OK;
}

int tsdf_synchronize(tsdf_volume* v) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (int rc = use_device(v)) return rc;
  if (v->copy_stream) CK(cudaStreamSynchronize(v->copy_stream));
  CK(cudaStreamSynchronize(v->stream));
  return TSDF_OK;
}

int tsdf_get_stats(tsdf_volume* v, tsdf_stats* out) {
  if (!v || !out) return set_err(TSDF_ERR_ARG, "NULL argument");
  *out = v->stats;
  return TSDF_OK;
}

int tsdf_reset_stats(tsdf_volume* v) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  v->stats = tsdf_stats{};
  return TSDF_OK;
}

int tsdf_timer_start(tsdf_volume* v) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (int rc = use_device(v)) return rc;
  CK(cudaEventRecord(v->t0, v->stream));
  return TSDF_OK;
}

int tsdf_timer_stop(tsdf_volume* v, float* elapsed_ms) {
  if (!v || !elapsed_ms) return set_err(TSDF_ERR_ARG, "NULL argument");
  if (int rc = use_device(v)) return rc;
  CK(cudaEventRecord(v->t1, v->stream));
  CK(cudaEventSynchronize(v->t1));
  CK(cudaEventElapsedTime(elapsed_ms, v->t0, v->t1));
  return TSDF_OK;
}

int tsdf_integrate(tsdf_volume* v, const float* depth, const float* cam2base) {
  return tsdf_integrate_batch(v, 1, depth, cam2base);
}

int tsdf_integrate_batch(tsdf_volume* v, int32_t n_frames, const float* depth_frames, const float* cam2base) {
  if (!v || !depth_frames || !cam2base) return set_err(TSDF_ERR_ARG, "NULL argument");
  if (n_frames < 0) return set_err(TSDF_ERR_ARG, "n_frames < 0");
  if (int rc = use_device(v)) return rc;
  const MemKind kind = classify(depth_frames);
  if (v->stream_ok) return stream_integrate<float>(v, n_frames, depth_frames, cam2base, kind);
  if (kind == MemKind::Device) {
    if (int rc = ensure_ring(v, false)) return rc;
    for (int f0 = 0; f0 < n_frames; f0 += v->batch) {
      const int n = (n_frames - f0 < v->batch) ? n_frames - f0 : v->batch;
      const int s = (int)(v->ring_pos++ % kSlots);
      if (int rc = wait_slot_free(v, s, v->stream, false)) return rc;
      // the caller's frames stay untouched: invalid pixels are marked in the engine's own copy
      CK(tsdf::launch_sanitize_depth(depth_frames + (size_t)f0 * v->frame_px, v->dev_depth[s], (int64_t)n * v->frame_px,
                                     v->stream));
      v->stats.kernel_launches += 1;
      if (int rc = launch_chunk(v, s, cam2base + 16 * (size_t)f0, n)) return rc;
    }
    return join_substreams(v);
  }
  return integrate_from_host<float>(v, n_frames, depth_frames, cam2base, kind);
}

int tsdf_integrate_batch_u16(tsdf_volume* v, int32_t n_frames, const uint16_t* depth_mm, const float* cam2base) {
  if (!v || !depth_mm || !cam2base) return set_err(TSDF_ERR_ARG, "NULL argument");
  if (n_frames < 0) return set_err(TSDF_ERR_ARG, "n_frames < 0");
  if (int rc = use_device(v)) return rc;
  const MemKind kind = classify(depth_mm);
  if (v->stream_ok) return stream_integrate<uint16_t>(v, n_frames, depth_mm, cam2base, kind);
  if (kind == MemKind::Device) {
    if (int rc = ensure_ring(v, true)) return rc;
    for (int f0 = 0; f0 < n_frames; f0 += v->batch) {
      const int n = (n_frames - f0 < v->batch) ? n_frames - f0 : v->batch;
      const int s = (int)(v->ring_pos++ % kSlots);
      if (int rc = wait_slot_free(v, s, v->stream, false)) return rc;
      CK(tsdf::launch_depth_from_u16(depth_mm + (size_t)f0 * v->frame_px, v->dev_depth[s], (int64_t)n * v->frame_px,
                                     true, v->stream));
      v->stats.kernel_launches += 1;
      if (int rc = launch_chunk(v, s, cam2base + 16 * (size_t)f0, n)) return rc;
    }
    return join_substreams(v);
  }
  return integrate_from_host<uint16_t>(v, n_frames, depth_mm, cam2base, kind);
}

int tsdf_download(tsdf_volume* v, float* tsdf_host, float* weight_host) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (int rc = tsdf_synchronize(v)) return rc;
  const size_t bytes = (size_t)v->n_vox * sizeof(float);
  if (tsdf_host) { CK(cudaMemcpy(tsdf_host, v->tsdf, bytes, cudaMemcpyDeviceToHost)); v->stats.d2h_bytes += bytes; }
  if (weight_host) { CK(cudaMemcpy(weight_host, v->weight, bytes, cudaMemcpyDeviceToHost)); v->stats.d2h_bytes += bytes; }
  return TSDF_OK;
}

int tsdf_upload(tsdf_volume* v, const float* tsdf_host, const float* weight_host) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (int rc = tsdf_synchronize(v)) return rc;
  const size_t bytes = (size_t)v->n_vox * sizeof(float);
  if (tsdf_host) { CK(cudaMemcpy(v->tsdf, tsdf_host, bytes, cudaMemcpyHostToDevice)); v->stats.h2d_bytes += bytes; }
  if (weight_host) { CK(cudaMemcpy(v->weight, weight_host, bytes, cudaMemcpyHostToDevice)); v->stats.h2d_bytes += bytes; }
  v->n_points = -1;
  return TSDF_OK;
}

int tsdf_device_ptrs(tsdf_volume* v, float** tsdf_dev, float** weight_dev) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (tsdf_dev) *tsdf_dev = v->tsdf;
  if (weight_dev) *weight_dev = v->weight;
  return TSDF_OK;
}

int tsdf_extract_surface(tsdf_volume* v, float tsdf_thresh, float weight_thresh, int64_t* n_points) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (int rc = use_device(v)) return rc;
  const int64_t n_tiles = (v->n_vox + tsdf::kExtractTile - 1) / tsdf::kExtractTile;
  if (v->tile_cap < n_tiles + 1) {
    cudaFree(v->tile_offsets);
    v->tile_offsets = nullptr;
    v->tile_cap = 0;
    CK(cudaMalloc(&v->tile_offsets, (size_t)(n_tiles + 1) * sizeof(unsigned long long)));
    v->tile_cap = n_tiles + 1;
  }
  tsdf::ExtractParams p;
  p.tsdf = v->tsdf; p.weight = v->weight; p.n = v->n_vox;
  p.dim_x = (int)v->d.dim_x; p.dim_y = (int)v->d.dim_y; p.z_begin = (int)v->d.z_begin;
  p.ox = v->d.origin_x; p.oy = v->d.origin_y; p.oz = v->d.origin_z; p.voxel_size = v->d.voxel_size;
  p.tsdf_thresh = tsdf_thresh; p.weight_thresh = weight_thresh;
  CK(tsdf::launch_extract_count(p, v->tile_offsets, n_tiles, v->stream));
  CK(tsdf::launch_extract_scan(v->tile_offsets, n_tiles, v->tile_offsets + n_tiles, v->stream));
  v->stats.kernel_launches += 2;
  unsigned long long total = 0;
  CK(cudaMemcpyAsync(&total, v->tile_offsets + n_tiles, sizeof total, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  v->stats.d2h_bytes += sizeof total;
  if ((int64_t)total > v->xyz_cap) {
    cudaFree(v->xyz);
    cudaFree(v->vidx);
    v->xyz = nullptr;
    v->vidx = nullptr;
    v->xyz_cap = 0;
    CK(cudaMalloc(&v->xyz, (size_t)total * 3 * sizeof(float)));
    CK(cudaMalloc(&v->vidx, (size_t)total * sizeof(long long)));
    v->xyz_cap = (int64_t)total;
  }
  if (total > 0) {
    CK(tsdf::launch_extract_emit(p, v->tile_offsets, n_tiles, v->xyz, v->vidx, v->stream));
    v->stats.kernel_launches += 1;
  }
  v->n_points = (int64_t)total;
  if (n_points) *n_points = (int64_t)total;
  return TSDF_OK;
}

int tsdf_copy_surface(tsdf_volume* v, float* xyz_host, int64_t capacity_points) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (v->n_points < 0) return set_err(TSDF_ERR_STATE, "tsdf_extract_surface has not been called since the last update");
  if (capacity_points < v->n_points) return set_err(TSDF_ERR_ARG, "capacity %lld < %lld points", (long long)capacity_points, (long long)v->n_points);
  if (v->n_points == 0) return TSDF_OK;
  if (!xyz_host) return set_err(TSDF_ERR_ARG, "xyz_host is NULL");
  if (int rc = use_device(v)) return rc;
  const size_t bytes = (size_t)v->n_points * 3 * sizeof(float);
  CK(cudaMemcpyAsync(xyz_host, v->xyz, bytes, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  v->stats.d2h_bytes += bytes;
  return TSDF_OK;
}

int tsdf_copy_surface_normals(tsdf_volume* v, float* normals_host, int64_t capacity_points) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (v->n_points < 0) return set_err(TSDF_ERR_STATE, "tsdf_extract_surface has not been called since the last update");
  if (capacity_points < v->n_points) return set_err(TSDF_ERR_ARG, "capacity %lld < %lld points", (long long)capacity_points, (long long)v->n_points);
  if (v->n_points == 0) return TSDF_OK;
  if (!normals_host) return set_err(TSDF_ERR_ARG, "normals_host is NULL");
  if (int rc = use_device(v)) return rc;
  if (v->normals_cap < v->n_points) {
    cudaFree(v->normals);
    v->normals = nullptr;
    v->normals_cap = 0;
    CK(cudaMalloc(&v->normals, (size_t)v->n_points * 3 * sizeof(float)));
    v->normals_cap = v->n_points;
  }
  tsdf::ExtractParams p;
  memset(&p, 0, sizeof p);
  p.tsdf = v->tsdf; p.weight = v->weight; p.n = v->n_vox;
  p.dim_x = (int)v->d.dim_x; p.dim_y = (int)v->d.dim_y; p.z_begin = (int)v->d.z_begin;
  CK(tsdf::launch_surface_normals(p, v->nz, v->vidx, v->n_points, v->normals, v->stream));
  v->stats.kernel_launches += 1;
  const size_t bytes = (size_t)v->n_points * 3 * sizeof(float);
  CK(cudaMemcpyAsync(normals_host, v->normals, bytes, cudaMemcpyDeviceToHost, v->stream));
  CK(cudaStreamSynchronize(v->stream));
  v->stats.d2h_bytes += bytes;
  return TSDF_OK;
}

int tsdf_surface_device_ptr(tsdf_volume* v, float** xyz_dev, int64_t* n_points) {
  if (!v) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (v->n_points < 0) return set_err(TSDF_ERR_STATE, "tsdf_extract_surface has not been called since the last update");
  // the emit kernel runs on the engine's own (non-blocking) stream: the caller's stream is not ordered after it, so
  // the buffer is handed out only once it is complete
  if (int rc = use_device(v)) return rc;
  CK(cudaStreamSynchronize(v->stream));
  if (xyz_dev) *xyz_dev = v->xyz;
  if (n_points) *n_points = v->n_points;
  return TSDF_OK;
}

int tsdf_Integrate(const float* cam_K, const float* cam2base, const float* depth_im, int im_height, int im_width,
                   int voxel_grid_dim_x, int voxel_grid_dim_y, int voxel_grid_dim_z, float voxel_grid_origin_x,
                   float voxel_grid_origin_y, float voxel_grid_origin_z, float voxel_size, float trunc_margin,
                   float* voxel_grid_TSDF, float* voxel_grid_weight, void* stream) {
  if (!cam_K || !cam2base || !depth_im || !voxel_grid_TSDF || !voxel_grid_weight)
    return set_err(TSDF_ERR_ARG, "NULL device pointer");
  if (im_height <= 0 || im_width <= 0 || voxel_grid_dim_x <= 0 || voxel_grid_dim_y <= 0 || voxel_grid_dim_z <= 0)
    return set_err(TSDF_ERR_ARG, "non-positive size");
  tsdf::IntegrateParams p;
  memset(&p, 0, sizeof p);
  p.tsdf = voxel_grid_TSDF; p.weight = voxel_grid_weight; p.depth = depth_im;
  p.frame_stride = (int64_t)im_height * im_width;
  p.dim_x = voxel_grid_dim_x; p.dim_y = voxel_grid_dim_y; p.nz = voxel_grid_dim_z; p.z_begin = 0;
  p.ox = voxel_grid_origin_x; p.oy = voxel_grid_origin_y; p.oz = voxel_grid_origin_z;
  p.voxel_size = voxel_size; p.trunc_margin = trunc_margin;
  p.im_w = im_width; p.im_h = im_height;
  CK(tsdf::launch_integrate_devpose(p, cam_K, cam2base, (cudaStream_t)stream));
  return TSDF_OK;
}

int tsdf_SaveVoxelGrid2SurfacePointCloud(const char* file_name, int voxel_grid_dim_x, int voxel_grid_dim_y,
                                         int voxel_grid_dim_z, float voxel_size, float voxel_grid_origin_x,
                                         float voxel_grid_origin_y, float voxel_grid_origin_z,
                                         const float* voxel_grid_TSDF, const float* voxel_grid_weight,
                                         float tsdf_thresh, float weight_thresh) {
  if (!file_name || !voxel_grid_TSDF || !voxel_grid_weight) return set_err(TSDF_ERR_ARG, "NULL argument");
  tsdf_volume_desc d;
  memset(&d, 0, sizeof d);
  d.dim_x = voxel_grid_dim_x; d.dim_y = voxel_grid_dim_y; d.dim_z = voxel_grid_dim_z;
  d.origin_x = voxel_grid_origin_x; d.origin_y = voxel_grid_origin_y; d.origin_z = voxel_grid_origin_z;
  d.voxel_size = voxel_size; d.trunc_margin = voxel_size;
  d.im_width = 1; d.im_height = 1;
  d.cull_level = -1;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  d.device = dev;
  tsdf_volume* v = nullptr;
  int rc = tsdf_create(&d, &v);
  if (rc) return rc;
  int64_t n = 0;
  std::vector<float> pts;
  rc = tsdf_upload(v, voxel_grid_TSDF, voxel_grid_weight);
  if (!rc) rc = tsdf_extract_surface(v, tsdf_thresh, weight_thresh, &n);
  if (!rc) {
    pts.resize((size_t)n * 3);
    rc = tsdf_copy_surface(v, pts.data(), n);
  }
  tsdf_destroy(v);
  if (rc) return rc;
  return tsdf_write_ply(file_name, pts.data(), n);
}

int tsdf_ComputeTDF(const float* voxel_grid_occ, float* voxel_grid_TDF, int voxel_grid_dim_x, int voxel_grid_dim_y,
                    int voxel_grid_dim_z, float voxel_grid_origin_x, float voxel_grid_origin_y, float voxel_grid_origin_z,
                    float voxel_size, float trunc_margin, void* stream) {
  (void)voxel_grid_origin_x; (void)voxel_grid_origin_y; (void)voxel_grid_origin_z;  // unused by the reference kernel too
  if (!voxel_grid_occ || !voxel_grid_TDF) return set_err(TSDF_ERR_ARG, "NULL device pointer");
  if (voxel_grid_dim_x <= 0 || voxel_grid_dim_y <= 0 || voxel_grid_dim_z <= 0) return set_err(TSDF_ERR_ARG, "non-positive size");
  const int64_t n = (int64_t)voxel_grid_dim_x * voxel_grid_dim_y * voxel_grid_dim_z;
  if (n > INT32_MAX) return set_err(TSDF_ERR_ARG, "grid exceeds the reference's int voxel index");
  const int radius = (int)roundf(trunc_margin / voxel_size);  // core/demo.cu:37
  if (radius < 1 || radius > 100) return set_err(TSDF_ERR_ARG, "search radius %d out of range", radius);
  cudaStream_t s = (cudaStream_t)stream;
  // scratch (two intermediate grids) from a per-device pool that only grows: no allocation in the steady state.
  // Calls that share a device are serialised on the pool by stream order: the pool remembers the last stream and
  // makes the next one wait for it.
  struct Pool { void* ptr = nullptr; size_t bytes = 0; cudaEvent_t last = nullptr; };
  static Pool pools[64];
  static std::mutex pool_lock;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return set_err(TSDF_ERR_ARG, "device ordinal out of range");
  const size_t need = tsdf::compute_tdf_scratch_bytes(voxel_grid_dim_x, voxel_grid_dim_y, voxel_grid_dim_z, radius);
  std::lock_guard<std::mutex> lock(pool_lock);
  Pool& P = pools[dev];
  if (!P.last) CK(cudaEventCreateWithFlags(&P.last, cudaEventDisableTiming));
  if (need > P.bytes) {
    if (P.ptr) { CK(cudaEventSynchronize(P.last)); CK(cudaFree(P.ptr)); P.ptr = nullptr; P.bytes = 0; }
    CK(cudaMalloc(&P.ptr, need));
    P.bytes = need;
  } else {
    CK(cudaStreamWaitEvent(s, P.last, 0));
  }
  cudaError_t e = tsdf::launch_compute_tdf(voxel_grid_occ, voxel_grid_TDF, P.ptr, voxel_grid_dim_x, voxel_grid_dim_y,
                                           voxel_grid_dim_z, radius, s);
  CK(e);
  CK(cudaEventRecord(P.last, s));
  return TSDF_OK;
}

int tsdf_point_cloud_to_tdf(const float* pts, int64_t n_pts, const int32_t dims[3], const float origin[3],
                            float voxel_size, float trunc_margin, float* tdf_host) {
  if (!pts || !dims || !origin || !tdf_host || n_pts < 0) return set_err(TSDF_ERR_ARG, "bad argument");
  const int64_t n = (int64_t)dims[0] * dims[1] * dims[2];
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || n > INT32_MAX) return set_err(TSDF_ERR_ARG, "bad grid size");
  float *d_pts = nullptr, *d_occ = nullptr, *d_tdf = nullptr;
  int rc = TSDF_OK;
  cudaError_t e = cudaMalloc(&d_pts, (size_t)(n_pts > 0 ? n_pts : 1) * 3 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_occ, (size_t)n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_tdf, (size_t)n * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(d_pts, pts, (size_t)n_pts * 3 * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_occ, 0, (size_t)n * sizeof(float));   // core/demo.cu:129-130
  if (e == cudaSuccess) e = cudaMemset(d_tdf, 0, (size_t)n * sizeof(float));   // core/demo.cu:140-141
  if (e == cudaSuccess)
    e = tsdf::launch_tdf_scatter_points(d_pts, n_pts, d_occ, dims[0], dims[1], dims[2], origin[0], origin[1], origin[2],
                                        voxel_size, 0);
  if (e != cudaSuccess) rc = set_err(TSDF_ERR_CUDA, "tsdf_point_cloud_to_tdf: %s", cudaGetErrorString(e));
  if (!rc) rc = tsdf_ComputeTDF(d_occ, d_tdf, dims[0], dims[1], dims[2], origin[0], origin[1], origin[2], voxel_size,
                                trunc_margin, nullptr);
  if (!rc) {
    e = cudaMemcpy(tdf_host, d_tdf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = set_err(TSDF_ERR_CUDA, "tsdf_point_cloud_to_tdf: %s", cudaGetErrorString(e));
  }
  cudaFree(d_pts); cudaFree(d_occ); cudaFree(d_tdf);
  return rc;
}

int tsdf_integrate_multi(tsdf_volume* const* vols, int32_t n_vols, int32_t n_frames, const void* const* frames,
                         int32_t frames_are_u16, const float* const* cam2base) {
  if (!vols || !frames || !cam2base || n_vols < 1) return set_err(TSDF_ERR_ARG, "bad argument");
  if (n_frames < 0) return set_err(TSDF_ERR_ARG, "n_frames < 0");
  for (int i = 0; i < n_vols; ++i)
    if (!vols[i] || !frames[i] || !cam2base[i]) return set_err(TSDF_ERR_ARG, "NULL entry %d", i);
  if (n_frames == 0) return TSDF_OK;
  tsdf_volume* lead = vols[0];
  bool same = lead->stream_ok;
  for (int i = 1; i < n_vols && same; ++i) {
    const tsdf_volume* v = vols[i];
    same = v->stream_ok && v->d.device == lead->d.device && v->d.dim_x == lead->d.dim_x && v->d.dim_y == lead->d.dim_y &&
           v->nz == lead->nz && v->d.origin_x == lead->d.origin_x && v->d.origin_y == lead->d.origin_y &&
           v->d.origin_z == lead->d.origin_z && v->d.voxel_size == lead->d.voxel_size &&
           v->d.trunc_margin == lead->d.trunc_margin && v->d.im_width == lead->d.im_width &&
           v->d.im_height == lead->d.im_height && memcmp(v->d.cam_K, lead->d.cam_K, sizeof lead->d.cam_K) == 0 &&
           v->batch == lead->batch && v->cull == lead->cull;
    for (int j = 0; j < i && same; ++j) same = vols[j] != v;
  }
  const int n_chunks = (n_frames + lead->batch - 1) / lead->batch;
  if (!same || n_vols == 1 || (int64_t)n_vols * n_chunks > tsdf::kStreamMaxSegments || n_frames > kStreamGroupDevice) {
    // not one launch's worth of identical volumes: one call per volume (still asynchronous, on each volume's stream)
    for (int i = 0; i < n_vols; ++i) {
      const int rc = frames_are_u16 ? tsdf_integrate_batch_u16(vols[i], n_frames, (const uint16_t*)frames[i], cam2base[i])
                                    : tsdf_integrate_batch(vols[i], n_frames, (const float*)frames[i], cam2base[i]);
      if (rc) return rc;
    }
    return TSDF_OK;
  }
  if (int rc = use_device(lead)) return rc;
  tsdf::StreamLaunch L;
  stream_group_params(lead, n_frames, &L);
  StreamSlot& LS = lead->ss[0];
  if (int rc = stream_slot_events(LS)) return rc;
  if (int rc = stream_job_array(LS, n_vols)) return rc;
  if (LS.staged) CK(cudaEventSynchronize(LS.ev_staged));   // the pinned job array may still be read
  float tz = 0.f;
  for (int i = 0; i < n_vols; ++i) {
    tsdf_volume* v = vols[i];
    StreamSlot& S = v->ss[0];
    const MemKind kind = classify(frames[i]);
    // staging and the per-volume pre-pass run on the volume's own stream (after whatever is queued there, e.g. a
    // reset); the leader's stream waits for it
    int rc = frames_are_u16 ? stream_stage_frames<uint16_t>(v, S, (const uint16_t*)frames[i], n_frames, kind)
                            : stream_stage_frames<float>(v, S, (const float*)frames[i], n_frames, kind);
    if (rc) return rc;
    if ((rc = stream_prepare_job(v, S, n_frames, cam2base[i], &LS.job_pin[i], &L, v->stream))) return rc;
    CK(cudaEventRecord(S.ev_staged, v->stream));
    S.staged = true;
    if (v != lead) CK(cudaStreamWaitEvent(lead->stream, S.ev_staged, 0));
    for (int f = 0; f < n_frames; ++f) tz += S.poses_pin[f].tz;
    v->stats.frames_integrated += n_frames;
    v->n_points = -1;
  }
  if (!lead->rec_scratch) return set_err(TSDF_ERR_STATE, "record scratch missing");
  {
    const float mid = lead->d.origin_z + 0.5f * (float)(lead->d.z_begin + lead->d.z_end) * lead->d.voxel_size;
    L.z_reverse = ((tz / ((float)n_frames * (float)n_vols)) < mid) ? 1 : 0;
  }
  CK(tsdf::launch_small_copy(LS.job_dev, LS.job_pin, (size_t)n_vols * sizeof(tsdf::StreamJob), lead->stream));
  lead->stats.kernel_launches += 1;
  CK(cudaEventRecord(LS.ev_staged, lead->stream));
  LS.staged = true;
  L.jobs = LS.job_dev;
  L.n_jobs = n_vols;
  // (the item counters / segment table / ticket of the launch are the leader's)
  if (int rc = stream_launch(lead, LS, &L, lead->stream)) return rc;
  CK(cudaEventRecord(LS.ev_done, lead->stream));
  LS.busy = true;
  for (int i = 1; i < n_vols; ++i) {
    tsdf_volume* v = vols[i];
    CK(cudaStreamWaitEvent(v->stream, LS.ev_done, 0));
    CK(cudaEventRecord(v->ss[0].ev_done, v->stream));
    v->ss[0].busy = true;
  }
  return TSDF_OK;
}

int tsdf_get_z_cost(tsdf_volume* v, uint64_t* cost, int32_t n_groups) {
  if (!v || !cost) return set_err(TSDF_ERR_ARG, "NULL argument");
  const int64_t n = (v->nz + 7) / 8;
  if (n_groups < n) return set_err(TSDF_ERR_ARG, "cost holds %d groups, the slab has %lld", n_groups, (long long)n);
  if (int rc = tsdf_synchronize(v)) return rc;
  if (!v->zcost) {
    memset(cost, 0, (size_t)n * sizeof *cost);
    return TSDF_OK;
  }
  CK(cudaMemcpy(cost, v->zcost, (size_t)n * sizeof *cost, cudaMemcpyDeviceToHost));
  return TSDF_OK;
}

int tsdf_plan_z_cost(const tsdf_volume_desc* desc, int32_t n_frames, const void* frames, int32_t frames_are_u16,
                     const float* cam2base, int32_t frame_step, uint64_t* cost, int32_t n_groups) {
  if (!desc || !frames || !cam2base || !cost || n_frames < 1) return set_err(TSDF_ERR_ARG, "bad argument");
  if (frame_step < 1) frame_step = 1;
  const int64_t groups = (desc->dim_z + 7) / 8;
  if (n_groups < groups) return set_err(TSDF_ERR_ARG, "cost holds %d groups, the grid has %lld", n_groups, (long long)groups);
  CK(cudaSetDevice(desc->device));
  tsdf::StreamLaunch L;
  memset(&L, 0, sizeof L);
  const tsdf::MipLayout mip = tsdf::make_mip_layout(desc->im_width, desc->im_height);
  const int64_t px = (int64_t)desc->im_width * desc->im_height;
  L.frame_stride = px;
  L.mip_stride = mip.stride;
  for (int l = 0; l < 5; ++l) { L.mip_w[l] = mip.w[l]; L.mip_off[l] = mip.off[l]; }
  L.dim_x = (int)desc->dim_x; L.dim_y = (int)desc->dim_y; L.nz = (int)desc->dim_z;
  L.ox = desc->origin_x; L.oy = desc->origin_y; L.oz = desc->origin_z;
  L.voxel_size = desc->voxel_size; L.trunc_margin = desc->trunc_margin;
  L.fx = desc->cam_K[0]; L.cx = desc->cam_K[2]; L.fy = desc->cam_K[4]; L.cy = desc->cam_K[5];
  L.im_w = desc->im_width; L.im_h = desc->im_height;
  L.chunk = tsdf::kChunkMax;
  L.cull_level = (desc->cull_level < 0 || desc->cull_level > 2) ? 2 : desc->cull_level;
  if (!tsdf::stream_prepare_geometry(&L)) {
    // the stream kernels do not apply to this grid: uniform cost (equal-height slabs)
    for (int64_t g = 0; g < groups; ++g) cost[g] = 1;
    return TSDF_OK;
  }
  const int ns = (n_frames + frame_step - 1) / frame_step;
  L.n_frames = ns;
  L.n_chunks = (ns + L.chunk - 1) / L.chunk;
  L.n_jobs = 1;
  std::vector<tsdf::FramePose> poses((size_t)ns);
  for (int i = 0; i < ns; ++i) pose_from_cam2base(cam2base + 16 * (size_t)i * frame_step, &poses[i]);
  {
    const float ex = fmaxf(fabsf(desc->origin_x), fabsf(desc->origin_x + (float)desc->dim_x * desc->voxel_size));
    const float ey = fmaxf(fabsf(desc->origin_y), fabsf(desc->origin_y + (float)desc->dim_y * desc->voxel_size));
    const float ez = fmaxf(fabsf(desc->origin_z), fabsf(desc->origin_z + (float)desc->dim_z * desc->voxel_size));
    stream_cull_margins(fmaxf(ex, fmaxf(ey, ez)), poses.data(), ns, &L);
  }
  float *d_depth = nullptr, *d_mips = nullptr;
  uint16_t* d_mm = nullptr;
  tsdf::FramePose* d_poses = nullptr;
  tsdf::StreamJob* d_job = nullptr;
  unsigned long long* d_cost = nullptr;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_depth); cudaFree(d_mips); cudaFree(d_mm); cudaFree(d_poses); cudaFree(d_job); cudaFree(d_cost);
    if (st) cudaStreamDestroy(st);
  };
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&d_depth, (size_t)ns * px * sizeof(float));
  if (e == cudaSuccess && L.cull_level >= 2) e = cudaMalloc(&d_mips, (size_t)ns * mip.stride * sizeof(float));
  if (e == cudaSuccess && frames_are_u16) e = cudaMalloc(&d_mm, (size_t)px * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_poses, (size_t)ns * sizeof(tsdf::FramePose));
  if (e == cudaSuccess) e = cudaMalloc(&d_job, sizeof(tsdf::StreamJob));
  if (e == cudaSuccess) e = cudaMalloc(&d_cost, (size_t)groups * sizeof(unsigned long long));
  const MemKind kind = classify(frames);
  for (int i = 0; i < ns && e == cudaSuccess; ++i) {
    const size_t off = (size_t)i * frame_step * px;
    float* dst = d_depth + (size_t)i * px;
    if (frames_are_u16) {
      const uint16_t* src = (const uint16_t*)frames + off;
      if (kind != MemKind::Device) {
        e = cudaMemcpyAsync(d_mm, src, (size_t)px * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
        src = d_mm;
      }
      if (e == cudaSuccess) e = tsdf::launch_depth_from_u16(src, dst, px, true, st);
    } else {
      const float* src = (const float*)frames + off;
      if (kind != MemKind::Device) {
        e = cudaMemcpyAsync(dst, src, (size_t)px * sizeof(float), cudaMemcpyHostToDevice, st);
        src = dst;
      }
      if (e == cudaSuccess) e = tsdf::launch_sanitize_depth(src, dst, px, st);
    }
  }
  if (e == cudaSuccess && L.cull_level >= 2)
    e = tsdf::launch_build_depth_mips(d_depth, d_mips, ns, desc->im_width, desc->im_height, px, st);
  tsdf::StreamJob job;
  memset(&job, 0, sizeof job);
  job.depth = d_depth;
  job.mips = (L.cull_level >= 2) ? d_mips : nullptr;
  job.poses = d_poses;
  job.zcost = d_cost;
  job.z_begin = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_poses, poses.data(), (size_t)ns * sizeof(tsdf::FramePose), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_job, &job, sizeof job, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_cost, 0, (size_t)groups * sizeof(unsigned long long), st);
  L.jobs = d_job;
  if (e == cudaSuccess) e = tsdf::launch_stream_cost(L, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cost, d_cost, (size_t)groups * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) return set_err(TSDF_ERR_CUDA, "tsdf_plan_z_cost: %s", cudaGetErrorString(e));
  return TSDF_OK;
}

int tsdf_split_slabs(const uint64_t* cost, int32_t n_groups, int64_t dim_z, int32_t n_slabs, double base_cost,
                     int64_t* z_cuts) {
  if (!cost || !z_cuts || n_groups < 1 || n_slabs < 1 || dim_z < 1) return set_err(TSDF_ERR_ARG, "bad argument");
  if ((dim_z + 7) / 8 != n_groups) return set_err(TSDF_ERR_ARG, "n_groups must be ceil(dim_z / 8)");
  if (n_slabs > n_groups) return set_err(TSDF_ERR_ARG, "more slabs than groups of 8 planes");
  // cost of a group = its kept (tile, frame) pairs + base_cost x the mean (culled tiles are cheap, not free)
  double mean = 0;
  for (int g = 0; g < n_groups; ++g) mean += (double)cost[g];
  mean /= n_groups;
  std::vector<double> cum((size_t)n_groups + 1, 0.0);
  for (int g = 0; g < n_groups; ++g) cum[g + 1] = cum[g] + (double)cost[g] + base_cost * mean + 1e-9;
  z_cuts[0] = 0;
  int g = 0;
  for (int r = 1; r < n_slabs; ++r) {
    const double target = cum[n_groups] * r / n_slabs;
    while (g < n_groups && cum[g + 1] <= target) ++g;
    // the cut nearer to the target, keeping at least one group per slab on both sides
    int cut = (g < n_groups && target - cum[g] > cum[g + 1] - target) ? g + 1 : g;
    const int lo = (int)(z_cuts[r - 1] / 8) + 1, hi = n_groups - (n_slabs - r);
    if (cut < lo) cut = lo;
    if (cut > hi) cut = hi;
    z_cuts[r] = (int64_t)cut * 8;
  }
  z_cuts[n_slabs] = dim_z;
  return TSDF_OK;
}

int tsdf_selftest_packed(int32_t device, int64_t* mismatches) {
  if (!mismatches) return set_err(TSDF_ERR_ARG, "mismatches is NULL");
  CK(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, sizeof *d));
  CK(cudaMemset(d, 0, sizeof *d));
  cudaError_t e = tsdf::launch_selftest_packed(d, 0);
  unsigned long long h = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CK(e);
  *mismatches = (int64_t)h;
  return TSDF_OK;
}

int tsdf_selftest_pixel_margin(int32_t device, float fx, float c0, float size, float* rcp_error_ulp24, float* worst_fraction) {
  if (!rcp_error_ulp24 || !worst_fraction) return set_err(TSDF_ERR_ARG, "output pointer is NULL");
  if (!(fx > 0.f) || !(size >= 1.f)) return set_err(TSDF_ERR_ARG, "fx and size must be positive");
  CK(cudaSetDevice(device));
  unsigned* d = nullptr;
  CK(cudaMalloc(&d, 2 * sizeof *d));
  CK(cudaMemset(d, 0, 2 * sizeof *d));
  cudaError_t e = tsdf::launch_selftest_margin(d, fx, c0, size, 0);
  float h[2] = {0.f, 0.f};
  if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CK(e);
  *rcp_error_ulp24 = h[0];
  *worst_fraction = h[1];
  return TSDF_OK;
}

int tsdf_selftest_rounding(int32_t device, int64_t* mismatches) {
  if (!mismatches) return set_err(TSDF_ERR_ARG, "mismatches is NULL");
  CK(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, sizeof *d));
  CK(cudaMemset(d, 0, sizeof *d));
  cudaError_t e = tsdf::launch_selftest_rounding(d, 0);
  unsigned long long h = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CK(e);
  *mismatches = (int64_t)h;
  return TSDF_OK;
}

}  // extern "C"
