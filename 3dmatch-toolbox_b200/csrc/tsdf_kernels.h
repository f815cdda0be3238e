// tsdf_kernels.h -- launch interface between the engine (tsdf_engine.cu) and the sm_100a kernels
// (tsdf_kernels.cu).  Internal to the library; the public surface is include/tsdf_fusion.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsdf {

constexpr int kMaxBatch = 32;    // == TSDF_MAX_BATCH
constexpr int kBlockThreads = 256;

// The 12 entries of a row-major 4x4 cam2base that demo.cu:35-41 reads, grouped by the camera
// axis they feed: cam_x = (m0*t0 + m4*t1) + m8*t2, cam_y = (m1*t0 + m5*t1) + m9*t2, ...
struct FramePose {
  float m0, m1, m2;    // multiply t0 = base_x - tx
  float m4, m5, m6;    // multiply t1 = base_y - ty
  float m8, m9, m10;   // multiply t2 = base_z - tz
  float tx, ty, tz;    // cam2base[3], [7], [11]
};

// Max-pyramid of a depth frame: levels 3..7 (blocks of 8..128 pixels), level l at off[l], w[l] x h[l] floats.
struct MipLayout {
  int w[5], h[5], off[5];
  int stride;  // floats per frame
};
MipLayout make_mip_layout(int im_w, int im_h);

struct IntegrateParams {
  float* tsdf;           // slab, x fastest
  float* weight;
  const float* depth;    // frame 0 of this batch; frame f at depth + f * frame_stride
  int64_t frame_stride;  // im_height * im_width
  int dim_x, dim_y;      // grid row length / rows per plane
  int nz;                // planes in the slab
  int z_begin;           // global index of the slab's first plane
  int nx_units;          // ceil(dim_x / VEC)
  int lx_shift;          // log2(threads of a block laid along x)
  float ox, oy, oz, voxel_size, trunc_margin;
  float fx, fy, cx, cy;  // cam_K[0], [4], [2], [5]
  int im_w, im_h;
  int n_frames;          // 1..kMaxBatch
  int cull_level;        // 0 none, 1 per-thread frustum cull, 2 + per-tile frame mask
  float cull_gx, cull_gy, cull_z;  // conservative margins (see make_cull_margins)
  const float* mips;     // max-pyramids of the batch's depth frames (nullptr = no occlusion test)
  int mip_stride, mip_w[5], mip_off[5];
  int fast_level;        // 0: every voxel through the exact path, 1: fast classification first
  int variant;           // tuning variant of the kernel instantiation (0 = production)
  int no_interior;       // validation: never use the interior-tile classification
  int no_ptx;            // validation: use the C++ classification instead of the PTX block
  int no_packed;         // validation: scalar PTX block instead of the packed (f32x2) pair block
  int no_premask;        // validation: frame masks computed inside the fusion kernel instead of the pre-pass
  uint2* tile_masks;     // scratch for the frame-mask pre-pass (nullptr = none), tile_mask_cap entries
  int64_t tile_mask_cap;
  int masks_ready;       // tile_masks were filled by launch_tile_masks() for this launch's z range
  int z_reverse;         // dispatch the z groups of the slab in descending order
  int force_lx;          // tuning: threads of a warp laid along x (0 = choose by lane waste)
  int sanitized;         // depth frames went through sanitize_depth_kernel (invalid = -inf)
  float lim_x, lim_y;    // filled by the launcher: 0.5 - pixel_round_margin()
  FramePose pose[kMaxBatch];
};

struct ExtractParams {
  const float* tsdf;
  const float* weight;
  int64_t n;             // voxels in the slab
  int dim_x, dim_y;
  int z_begin;
  float ox, oy, oz, voxel_size;
  float tsdf_thresh, weight_thresh;
};

// ---- stream engine (tsdf_stream_kernels.cu): the production fusion path -----------------------------------------
// A call's frames are cut into CHUNKS of <= kChunkMax consecutive frames (one 32-bit frame mask per warp tile and
// chunk); a work ITEM is (block tile, chunk).  stream_mask_kernel computes the frame masks of every (warp tile,
// chunk) of the call up front and appends the non-empty items to per-chunk lists; stream_integrate_kernel is ONE
// persistent launch that walks those lists chunk after chunk (so the chunk's frames stay L2-resident), with a
// per-tile progress word ordering the items of one tile.
constexpr int kChunkMax = 32;
constexpr int kStreamMaxSegments = 1024;   // jobs x chunks per launch

struct StreamJob {              // device memory; one volume (z-slab) x one group of frames
  float* tsdf;
  float* weight;
  const float* depth;           // sanitized frames of the group, frame f at depth + f * frame_stride
  const float* mips;            // depth max-pyramids, frame f at mips + f * mip_stride (nullptr: no occlusion test)
  const FramePose* poses;       // n_frames entries
  uint2* masks;                 // [n_chunks][n_btiles][8 warps]: {frame mask, interior mask}
  unsigned long long* items;    // [n_chunks][n_btiles]: block tile | (previous non-empty chunk of the tile + 1) << 32
  unsigned* item_count;         // [n_chunks], zeroed before stream_mask_kernel
  unsigned* progress;           // [n_btiles], zeroed per group: (last completed chunk of the tile) + 1
  unsigned long long* zcost;    // [n_zg] kept (warp tile, frame) pairs per z group, or nullptr
  int z_begin;                  // global index of the slab's first plane
  int pad_;
};

struct StreamLaunch {           // kernel parameter block (constant bank): what every job of the launch shares
  const StreamJob* jobs;
  int n_jobs;
  int n_frames, chunk, n_chunks;      // frames of the group, frames per chunk, chunks
  int epoch;                          // frames per epoch (phase A / phase B alternation inside a chunk)
  unsigned* seg_begin;                // [n_jobs * n_chunks + 1] exclusive prefix of the item counts (segment = job, chunk)
  unsigned* ticket;                   // item counter of the persistent kernel
  unsigned* rec_scratch;              // sticky-voxel records, 2 x 8192 per resident block
  int64_t frame_stride;
  int mip_stride, mip_w[5], mip_off[5];
  int dim_x, dim_y, nz, nx_units, lx_shift;
  int n_xt, n_yb, n_zg, n_btiles;     // block tiles along x, y, z and per slab
  float ox, oy, oz, voxel_size, trunc_margin;
  float fx, fy, cx, cy, lim_x, lim_y;
  int im_w, im_h;
  int cull_level;
  float cull_z, cull_gx, cull_gy;
  int z_reverse;
  int plan_beta;                      // work statistics: extra cost of a (tile, frame) pair that reaches a surface band
};

// Fills the geometry fields of L that follow from dim_x / dim_y / nz (nx_units, lx_shift, n_xt, n_yb, n_zg, n_btiles,
// lim_x, lim_y).  Returns false when the stream kernels do not apply (dim_x % 4 != 0, image too large, trunc <= 0).
bool stream_prepare_geometry(StreamLaunch* L);
int stream_resident_blocks(int device);   // blocks of the persistent kernel that fit the device
cudaError_t launch_stream_masks(const StreamLaunch& L, cudaStream_t s);      // + segment prefix + ticket reset (2 kernels)
cudaError_t launch_stream_integrate(const StreamLaunch& L, int n_blocks, cudaStream_t s);
// Only the per-z-group work counts (jobs[].zcost) of the frames in L; no masks, no item lists (slab planning).
cudaError_t launch_stream_cost(const StreamLaunch& L, cudaStream_t s);

// Bound on |u' - u| between the fast pixel coordinate u' = fma(cam, fl(fx * rcp.approx(z)), c0) and the reference's
// u = fl(fl(fx * fl(cam / z)) + c0) (demo.cu:46-47) wherever the two can disagree about the pixel, i.e. for
// u' in [-1.5, size + 1.5] (further out both round to a pixel outside the image, demo.cu:48).  With P = fx cam / z:
//   reference   fl(cam / z) = (cam / z)(1 + d1), fl(fx * .) = (.)(1 + d2), |d| <= 2^-24, final sum + r1, |r1| <= 2^-24 |u|
//   fast        rcp.approx(z) = (1 / z)(1 + e1): the PTX ISA bounds the ABSOLUTE error by 2^-23 on [1, 2), i.e. the
//               relative error by 2^-22 = 4 * 2^-24 (the exponent is handled exactly; tsdf_selftest_pixel_margin measures it
//               over every mantissa); fl(fx * .) = (.)(1 + e2), |e2| <= 2^-24; the fma rounds once: + r2, |r2| <= 2^-24 |u'|
//   |u - u'| <= |P| (|d1| + |d2| + |e1| + |e2|) + |r1| + |r2| <= 7 * 2^-24 * pmax + 2 * 2^-24 * mag   (+ O(2^-46) terms)
// where mag = size + 1.5 bounds |u|, |u'| and pmax = max |u' - c0| over that range bounds |P| (P is within the margin of
// u' - c0).  The returned margin is 1.25 times that bound plus 1e-6; round 1 / 2 used the looser |P| <= mag + |c0| and a
// factor 2, which sent 2.75 times as many voxels to the exact evaluation for nothing.  tsdf_selftest_pixel_margin() checks
// the bound on 2.5e9 pseudo-random (cam, z) pairs per intrinsic set.
__host__ __device__ inline float pixel_round_margin(float size, float c0) {
  const float mag = size + 1.5f;
  const float a = c0 + 1.5f < 0.f ? -(c0 + 1.5f) : c0 + 1.5f, b = mag - c0 < 0.f ? c0 - mag : mag - c0;
  const float pmax = (a > b ? a : b) + 1.0f;
  return 1.25f * (4.1724e-7f * pmax + 1.1921e-7f * mag) + 1e-6f;
}

constexpr int kExtractTile = 4096;  // voxels per extraction tile (256 threads x 16)

// Entries of IntegrateParams::tile_masks that cover a slab of nz planes for every warp shape the launcher may pick.
int64_t tile_mask_entries(int dim_x, int dim_y, int64_t nz);
cudaError_t launch_fill(float* tsdf, float* weight, int64_t n, cudaStream_t s);
cudaError_t launch_integrate(const IntegrateParams& p, cudaStream_t s);
int64_t launch_tile_masks(IntegrateParams p, cudaStream_t s);  // whole-slab mask pre-pass; tiles per z group or 0
int last_integrate_launches();  // kernels the last launch_integrate*() call of this thread enqueued (1 or 2)
// Pose and intrinsics read from DEVICE memory (reference-style Integrate argument list).
cudaError_t launch_integrate_devpose(IntegrateParams p, const float* d_cam_K, const float* d_cam2base,
                                     cudaStream_t s);
cudaError_t launch_depth_from_u16(const uint16_t* mm, float* depth, int64_t n, bool sanitize, cudaStream_t s);
cudaError_t launch_sanitize_depth(const float* in, float* out, int64_t n, cudaStream_t s);
cudaError_t launch_build_depth_mips(const float* frames, float* mips, int n_frames, int im_w, int im_h,
                                    int64_t frame_stride, cudaStream_t s);
cudaError_t launch_extract_count(const ExtractParams& p, unsigned long long* tile_counts, int64_t n_tiles,
                                 cudaStream_t s);
cudaError_t launch_extract_scan(unsigned long long* tile_counts, int64_t n_tiles, unsigned long long* total,
                                cudaStream_t s);
cudaError_t launch_extract_emit(const ExtractParams& p, const unsigned long long* tile_offsets, int64_t n_tiles,
                                float* xyz, long long* voxel_index, cudaStream_t s);
// Unit normals (normalised central TSDF differences) of n surface voxels given by their slab-local linear indices.
cudaError_t launch_surface_normals(const ExtractParams& p, int64_t nz, const long long* voxel_index, int64_t n,
                                   float* normals, cudaStream_t s);
// dst <- pinned host memory (or zeros when the source is nullptr), `bytes` a multiple of 4, by a kernel on `s` (not the copy engine)
cudaError_t launch_small_copy(void* dst, const void* pinned_src_or_null, size_t bytes, cudaStream_t s);
cudaError_t launch_selftest_rounding(unsigned long long* mismatches, cudaStream_t s);
cudaError_t launch_selftest_packed(unsigned long long* mismatches, cudaStream_t s);   // tsdf_stream_kernels.cu
// out2[0] / out2[1] (float bits, zeroed by the caller): max relative error of rcp.approx in units of 2^-24; max |u' - u| of
// the fast pixel coordinate as a fraction of pixel_round_margin(size, c0)
cudaError_t launch_selftest_margin(unsigned* out2, float fx, float c0, float size, cudaStream_t s);
// tdf_kernels.cu: scratch = compute_tdf_scratch_bytes() bytes of device memory (two intermediate grids, 8 or 16 bit)
size_t compute_tdf_scratch_bytes(int dim_x, int dim_y, int dim_z, int radius);
cudaError_t launch_compute_tdf(const float* occ, float* tdf, void* scratch, int dim_x, int dim_y, int dim_z,
                               int radius, cudaStream_t s);
cudaError_t launch_tdf_scatter_points(const float* pts, int64_t n_pts, float* occ, int dim_x, int dim_y, int dim_z,
                                      float ox, float oy, float oz, float voxel_size, cudaStream_t s);

}  // namespace tsdf
