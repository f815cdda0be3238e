// tsdf_engine_internal.h -- state of an engine handle, shared by the C-ABI implementation files (tsdf_engine.cu,
// tsdf_sharded.cu).  Internal to the library; the public surface is include/tsdf_fusion.h.
#pragma once
#include "tsdf_fusion.h"
#include "tsdf_kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

// Sets the calling thread's tsdf_last_error() text (printf-style) and returns `code`.
int tsdf_errf(int code, const char* fmt, ...);
#define set_err tsdf_errf

#define CK(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return set_err(TSDF_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)


constexpr int kSlots = 3;   // staging ring depth (K-frame depth slots)
constexpr int kMaxSub = 4;  // concurrent sub-slab streams per handle

// ---- stream engine (tsdf_stream_kernels.cu) ----
constexpr int kStreamSlots = 2;
constexpr int kStreamGroupDevice = 1024;  // frames per persistent launch when the input is device-resident
constexpr int kStreamGroupHost = 256;     // ... at most, when it is host memory (double-buffered against the H2D copies)
constexpr int kCounterWords = 2 * tsdf::kStreamMaxSegments + 8;  // item_count[], seg_begin[], ticket

// Buffers of one group of frames in flight: the engine's own sanitized copy of the frames, their depth pyramids and
// poses, and the frame masks / item lists / progress words of the persistent kernel.
struct StreamSlot {
  float* depth = nullptr;   int64_t depth_cap = 0;   // frames
  float* mips = nullptr;    int64_t mips_cap = 0;    // frames
  uint16_t* mm = nullptr;   int64_t mm_cap = 0;      // frames (raw millimetres, uint16 input)
  void* pin = nullptr;      size_t pin_cap = 0;      // bytes: pinned staging for pageable host input
  tsdf::FramePose* poses_dev = nullptr;
  tsdf::FramePose* poses_pin = nullptr;
  int64_t poses_cap = 0, poses_pin_cap = 0;          // frames
  uint2* masks = nullptr;   int64_t masks_cap = 0;   // entries
  unsigned long long* items = nullptr; int64_t items_cap = 0;
  unsigned* counters = nullptr;                      // kCounterWords
  unsigned* progress = nullptr;                      // n_btiles
  tsdf::StreamJob* job_dev = nullptr;                // [jobs_cap]
  tsdf::StreamJob* job_pin = nullptr;
  int jobs_cap = 0;
  cudaEvent_t ev_copied = nullptr;   // H2D of the group's frames finished
  cudaEvent_t ev_staged = nullptr;   // the pinned pose / job / frame staging buffers have been consumed
  cudaEvent_t ev_done = nullptr;     // the group's fusion kernel finished (slot reusable)
  bool staged = false, busy = false;
};


struct tsdf_volume {
  tsdf_volume_desc d;
  int64_t nz = 0, n_vox = 0, frame_px = 0;
  int batch = 0, cull = 2, tuning = 0;
  float* tsdf = nullptr;
  float* weight = nullptr;
  bool owns_volume = false;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  bool owns_stream = false;
  // staging ring (allocated on first host-input integrate)
  float* dev_depth[kSlots] = {};
  float* dev_mips[kSlots] = {};   // max-pyramids of the slot's frames (occlusion culling)
  tsdf::MipLayout mip = {};
  float* pin_depth[kSlots] = {};
  uint16_t* dev_mm[kSlots] = {};
  uint16_t* pin_mm[kSlots] = {};
  cudaEvent_t ev_copied[kSlots] = {}, ev_ready[kSlots] = {};
  cudaEvent_t ev_join[kSlots][kMaxSub] = {};  // sub-stream i finished reading slot s
  bool slot_busy[kSlots] = {};
  // the slab is fused as n_sub z sub-slabs on their own streams: consecutive K-frame passes of one
  // sub-slab are ordered, different sub-slabs overlap, which hides the tail of every launch
  int n_sub = 1;
  cudaStream_t sub_stream[kMaxSub] = {};
  int64_t sub_z[kMaxSub + 1] = {};  // sub-slab i = local planes [sub_z[i], sub_z[i+1])
  uint2* tile_masks[kMaxSub] = {};  // per sub-slab scratch of the frame-mask pre-pass (non-production variants)
  int64_t tile_mask_cap[kMaxSub] = {};
  uint2* slot_masks[kSlots] = {};   // whole-slab frame masks of the ring slot's frames (production kernel)
  int64_t slot_mask_cap = 0;
  int last_slot = -1;
  int64_t ring_pos = 0;
  // stream engine
  bool stream_ok = false;              // the persistent work-list kernel applies to this volume
  tsdf::StreamLaunch sl = {};          // geometry / intrinsics part of the launch parameters
  StreamSlot ss[kStreamSlots];
  int64_t stream_group = 0;
  unsigned* rec_scratch = nullptr;
  int resident_blocks = 0;
  unsigned long long* zcost = nullptr; // per z group: kept (warp tile, frame) pairs since tsdf_reset_stats
  // surface extraction
  unsigned long long* tile_offsets = nullptr;
  int64_t tile_cap = 0;
  float* xyz = nullptr;
  long long* vidx = nullptr;     // slab-local linear voxel index of every extracted point
  float* normals = nullptr;
  int64_t xyz_cap = 0, normals_cap = 0, n_points = -1;
  float last_tsdf_thresh = 0.f, last_weight_thresh = 0.f;
  // timing / stats
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  tsdf_stats stats = {};
  float max_abs_corner = 0.f;
};

