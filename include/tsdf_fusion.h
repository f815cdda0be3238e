/*
 * include/tsdf_fusion.h -- C ABI of the B200-native multi-frame TSDF fusion engine.
 *
 * Drop-in boundary for the hot path of andyzeng/3dmatch-toolbox `depth-fusion/`
 * (reference paths below are relative to the reference repository root):
 *
 *   depth-fusion/demo.cu:18-68     __global__ Integrate            -> tsdf_integrate*, tsdf_Integrate
 *   depth-fusion/demo.cu:131-153   volume init + upload            -> tsdf_create / tsdf_reset
 *   depth-fusion/demo.cu:156-183   per-frame staging + launch loop -> tsdf_integrate_batch
 *   depth-fusion/demo.cu:186-188   volume download                 -> tsdf_download
 *   depth-fusion/utils.hpp:14-56   SaveVoxelGrid2SurfacePointCloud -> tsdf_extract_surface + tsdf_write_ply,
 *                                                                     tsdf_SaveVoxelGrid2SurfacePointCloud
 *   depth-fusion/demo.cu:196-213   tsdf.bin writer (commented out) -> tsdf_write_bin
 *   depth-fusion/utils.hpp:60-70   LoadMatrixFromFile              -> tsdf_load_matrix
 *   depth-fusion/utils.hpp:74-86   ReadDepth                       -> tsdf_read_depth_png, tsdf_depth_from_u16
 *   depth-fusion/utils.hpp:89-109  multiply_matrix                 -> tsdf_multiply_matrix
 *   depth-fusion/utils.hpp:112-238 invert_matrix                   -> tsdf_invert_matrix
 *   depth-fusion/fuseSceneFragments.m:36-81 fragment loop          -> bin/fuse_fragments (built on this ABI),
 *                                                                     tsdf_integrate_multi
 *   (no counterpart: the reference is single-GPU)                  -> tsdf_sharded_* (z-slabs over the GPUs of a box),
 *                                                                     tsdf_plan_z_cost, tsdf_split_slabs
 *   core/demo.cu:22-57,100-163     ComputeTDF + grid/occupancy     -> tsdf_ComputeTDF, tsdf_tdf_grid_from_points,
 *                                                                     tsdf_point_cloud_to_tdf   
 *
 * Conventions: plain pointers and sizes only; every entry point returns an int status
 * (TSDF_OK == 0) unless stated; the library never prints and never calls exit();
 * tsdf_last_error() returns the calling thread's last error text.  Device memory is owned
 * by the handle (unless adopted through ext_tsdf/ext_weight); host memory is owned by the
 * caller.  A handle is not thread-safe; distinct handles are independent.  There is NO CPU
 * fallback: every compute entry point fails with TSDF_ERR_CUDA when no sm_100 device is usable.
 *
 * Numerical contract: fp32, one IEEE round-to-nearest per source-level
 * operation of demo.cu:30-66 (the "--fmad=false" evaluation), powf(q,2) == q*q, frames applied
 * to every voxel in ascending frame order.  Results are independent of batching (K), of the
 * z-slab decomposition and of the culling level.
 */
#ifndef TSDF_FUSION_H
#define TSDF_FUSION_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDF_OK            0
#define TSDF_ERR_ARG       1 /* invalid argument */
#define TSDF_ERR_CUDA      2 /* CUDA runtime error (see tsdf_last_error) */
#define TSDF_ERR_IO        3 /* file could not be opened / read / written */
#define TSDF_ERR_FORMAT    4 /* file content not in the expected format */
#define TSDF_ERR_STATE     5 /* call sequence error (e.g. copy_surface before extract_surface) */

#define TSDF_MAX_BATCH     32 /* frames fused per pass over the volume (K) */

typedef struct tsdf_volume tsdf_volume; /* opaque engine handle */

typedef struct tsdf_volume_desc {
  int64_t dim_x, dim_y, dim_z;          /* GLOBAL grid dimensions, x fastest (demo.cu:97-99, :61) */
  int64_t z_begin, z_end;               /* z-planes held by this handle; 0,0 or 0,dim_z = whole grid */
  float origin_x, origin_y, origin_z;   /* grid origin in base-frame camera coordinates (demo.cu:92-94) */
  float voxel_size;                     /* demo.cu:95 */
  float trunc_margin;                   /* demo.cu:96 */
  int32_t im_width, im_height;          /* depth image size (demo.cu:87-88) */
  float cam_K[9];                       /* row-major 3x3 intrinsics (demo.cu:116-117) */
  int32_t device;                       /* CUDA device ordinal */
  int32_t max_batch;                    /* K: frames per chunk (one pass over the volume), 1..TSDF_MAX_BATCH, 0 = default (32) */
  int32_t cull_level;                   /* 0 none, 1 per-thread frustum cull, 2 + per-tile frame mask (default: -1 = 2) */
  int32_t reserved0;                    /* must be 0 */
  void *stream;                         /* cudaStream_t to run on; NULL = the engine creates its own */
  float *ext_tsdf, *ext_weight;         /* optional caller-owned DEVICE slabs to adopt (NULL = engine-owned);
                                           16-byte aligned */
} tsdf_volume_desc;

typedef struct tsdf_stats {
  int64_t kernel_launches;              /* engine kernels launched since create / last reset_stats */
  int64_t frames_integrated;
  int64_t h2d_bytes, d2h_bytes;
} tsdf_stats;

/* ---- engine lifecycle ------------------------------------------------------------------------ */
int tsdf_create(const tsdf_volume_desc *desc, tsdf_volume **out);
void tsdf_destroy(tsdf_volume *v);
/* TSDF <- 1.0f, weight <- 0 on the device (demo.cu:131-145 without the PCIe round trip). */
int tsdf_reset(tsdf_volume *v);
int tsdf_synchronize(tsdf_volume *v);
int tsdf_get_stats(tsdf_volume *v, tsdf_stats *out);
int tsdf_reset_stats(tsdf_volume *v);
/* CUDA-event stopwatch on the engine's own stream (the stream the kernels are launched on). */
int tsdf_timer_start(tsdf_volume *v);
int tsdf_timer_stop(tsdf_volume *v, float *elapsed_ms); /* records, synchronizes, returns ms */

/* ---- integration ----------------------------------------------------------------------------- */
/* One reference launch (demo.cu:173-182).  depth: im_height*im_width floats (metres, 0 = invalid),
 * host or device memory (detected).  cam2base: 16 host floats, row-major 4x4. */
int tsdf_integrate(tsdf_volume *v, const float *depth, const float *cam2base);
/* n_frames reference launches in ascending order (demo.cu:156-183).  depth_frames is
 * [n_frames][im_height*im_width] contiguous, host (pageable or pinned) or device; cam2base is
 * [n_frames][16] on the host.  Frames are fused K at a time with each voxel's TSDF/weight held in
 * registers across the K frames.  Asynchronous for device input; host input is staged through
 * pinned buffers and overlapped with the kernels; call tsdf_synchronize() to wait. */
int tsdf_integrate_batch(tsdf_volume *v, int32_t n_frames, const float *depth_frames, const float *cam2base);
/* Same, for depth supplied as raw uint16 millimetres (the PNG payload): converted on the device
 * exactly as ReadDepth does (utils.hpp:82-84). */
int tsdf_integrate_batch_u16(tsdf_volume *v, int32_t n_frames, const uint16_t *depth_mm, const float *cam2base);

/* n_vols independent volumes (fragments, fuseSceneFragments.m:36-81) of identical shape on ONE device, each with
 * its own frames and poses, fused by one launch: frames[i] = n_frames depth frames of volume i (float metres or, with
 * frames_are_u16, raw uint16 millimetres; host or device), cam2base[i] = [n_frames][16] on the host.  Small volumes
 * then fill the GPU together instead of one after the other.  Results are identical to n_vols tsdf_integrate_batch
 * calls (to which the call falls back when the volumes differ in shape). */
int tsdf_integrate_multi(tsdf_volume *const *vols, int32_t n_vols, int32_t n_frames, const void *const *frames,
                         int32_t frames_are_u16, const float *const *cam2base);

/* ---- z-slab planning (multi-GPU) ----------------------------------------------- */
/* Work estimate per group of 8 z-planes of the WHOLE grid of `desc` (its z range and device memory are not used)
 * for a frame sequence: the number of (32x4x8-voxel tile, frame) pairs that survive the frustum and occlusion culls,
 * i.e. what the fusion kernel has to classify.  Every frame_step-th frame is sampled.  cost: ceil(dim_z / 8) entries.
 * Deterministic (integer counts), so every rank of a multi-process run derives the same partition. */
int tsdf_plan_z_cost(const tsdf_volume_desc *desc, int32_t n_frames, const void *frames, int32_t frames_are_u16,
                     const float *cam2base, int32_t frame_step, uint64_t *cost, int32_t n_groups);
/* The same counts as accumulated by a handle's own fusion calls since tsdf_create (its slab only). */
int tsdf_get_z_cost(tsdf_volume *v, uint64_t *cost, int32_t n_groups);
/* n_slabs contiguous z-slabs of (nearly) equal summed cost, cut at multiples of 8 planes: z_cuts[0] = 0 ...
 * z_cuts[n_slabs] = dim_z.  base_cost: cost of a fully culled group relative to the mean group (0.05 is typical). */
int tsdf_split_slabs(const uint64_t *cost, int32_t n_groups, int64_t dim_z, int32_t n_slabs, double base_cost,
                     int64_t *z_cuts);

/* ---- one volume over several GPUs (extends demo.cu:72-216 to a multi-GPU box) -- */
/* The grid of `desc` cut into n_slabs contiguous z-slabs, slab i resident on CUDA device devices[i] (NULL: devices
 * 0 .. n_slabs-1; a device may hold several slabs).  z_cuts: n_slabs + 1 ascending plane indices from 0 to dim_z
 * (NULL: equal heights; see tsdf_plan_z_cost / tsdf_split_slabs for equal-work cuts).  desc->device, z_begin/z_end,
 * stream and ext_* are ignored.  One host thread drives all devices; NCCL (libnccl.so.2, loaded at run time) carries
 * the frame broadcast and the surface gather between distinct devices. */
typedef struct tsdf_sharded tsdf_sharded;
int tsdf_sharded_create(const tsdf_volume_desc *desc, int32_t n_slabs, const int32_t *devices, const int64_t *z_cuts,
                        tsdf_sharded **out);
void tsdf_sharded_destroy(tsdf_sharded *s);
int tsdf_sharded_reset(tsdf_sharded *s);
int tsdf_sharded_synchronize(tsdf_sharded *s);
/* n_frames reference launches in order (demo.cu:156-183) on every slab: the frames (float metres or, with
 * frames_are_u16, raw uint16 millimetres; host memory or memory of devices[0]) are uploaded once, 128 at a time, and
 * broadcast with ncclBroadcast while the previous 128 are being fused.  Asynchronous. */
int tsdf_sharded_integrate_batch(tsdf_sharded *s, int32_t n_frames, const void *frames, int32_t frames_are_u16,
                                 const float *cam2base);
/* Per-slab extraction, then the ordered gather (ncclSend / ncclRecv) of the slab points to devices[0]: slab order is
 * ascending z, i.e. ascending linear voxel index, so the gathered list is the reference's tsdf.ply payload. */
int tsdf_sharded_extract_surface(tsdf_sharded *s, float tsdf_thresh, float weight_thresh, int64_t *n_points);
int tsdf_sharded_copy_surface(tsdf_sharded *s, float *xyz_host, int64_t capacity_points);
/* tsdf.bin (demo.cu:196-213) written slab by slab at each slab's byte offset; the volume never has to fit one host
 * buffer. */
int tsdf_sharded_write_bin(tsdf_sharded *s, const char *path);
/* Slab i: its engine handle (owned by s), plane range and device.  Any output pointer may be NULL. */
int tsdf_sharded_slab(tsdf_sharded *s, int32_t i, tsdf_volume **vol, int64_t *z_begin, int64_t *z_end, int32_t *device);

/* ---- results --------------------------------------------------------------------------------- */
/* Device->host copy of the slab (demo.cu:186-187).  Either pointer may be NULL. */
int tsdf_download(tsdf_volume *v, float *tsdf_host, float *weight_host);
int tsdf_upload(tsdf_volume *v, const float *tsdf_host, const float *weight_host);
/* Device pointers of the slab (x fastest, (z - z_begin) slowest). */
int tsdf_device_ptrs(tsdf_volume *v, float **tsdf_dev, float **weight_dev);
/* Surface extraction (utils.hpp:20-53) on the device: voxels with |tsdf| < tsdf_thresh and
 * weight > weight_thresh, emitted as voxel-centre xyz in ascending linear index.  Points stay
 * on the device until tsdf_copy_surface. */
int tsdf_extract_surface(tsdf_volume *v, float tsdf_thresh, float weight_thresh, int64_t *n_points);
int tsdf_copy_surface(tsdf_volume *v, float *xyz_host, int64_t capacity_points);
/* Unit surface normals of the points of the last tsdf_extract_surface, in the same order: the
 * normalised central difference of the TSDF at the voxel, as PCL-benchmark/pclWrapperFPFH.h:308-330
 * (computeSurfaceNormal) computes it; (0, 0, 0) where a neighbour lies outside the slab or the gradient vanishes. */
int tsdf_copy_surface_normals(tsdf_volume *v, float *normals_host, int64_t capacity_points);
/* Device buffer of the extracted points.  The call waits for the extraction kernels (they run on the engine's own
 * stream), so the buffer may be read from any stream afterwards; it stays valid until the next tsdf_extract_surface,
 * integrate, reset or destroy on this handle. */
int tsdf_surface_device_ptr(tsdf_volume *v, float **xyz_dev, int64_t *n_points);

/* ---- file formats (host) --------------------------------------------------------------------- */
/* tsdf.ply: the reference's 7-line header + little-endian float xyz (utils.hpp:26-53). */
int tsdf_write_ply(const char *path, const float *xyz, int64_t n_points);
/* tsdf.bin: 8-float header + dim^3 TSDF floats (demo.cu:196-213). */
int tsdf_write_bin(const char *path, int64_t dim_x, int64_t dim_y, int64_t dim_z, float origin_x,
                   float origin_y, float origin_z, float voxel_size, float trunc_margin,
                   const float *tsdf_host);
/* M*N whitespace-separated floats, row-major (utils.hpp:60-70).  Fails instead of crashing when
 * the file is missing or short. */
int tsdf_load_matrix(const char *path, int32_t rows, int32_t cols, float *out);
/* 16-bit grayscale PNG (millimetres) -> float metres, > 6 m -> 0 (utils.hpp:74-86). */
int tsdf_read_depth_png(const char *path, int32_t im_height, int32_t im_width, float *depth);
int tsdf_read_depth_png_u16(const char *path, int32_t im_height, int32_t im_width, uint16_t *depth_mm);
void tsdf_depth_from_u16(const uint16_t *depth_mm, int64_t n, float *depth);

/* ---- pose algebra (host, reference rounding) ------------------------------------------------- */
void tsdf_multiply_matrix(const float m1[16], const float m2[16], float m_out[16]); /* utils.hpp:89-109 */
int tsdf_invert_matrix(const float m[16], float inv_out[16]); /* utils.hpp:112-238; 1 = ok, 0 = singular */

/* ---- function-level drop-ins (same argument lists as the reference) -------------------------- */
/* Replaces `Integrate<<<dim_z, dim_y>>>(...)` (demo.cu:18-22, :179-182): all pointers are DEVICE
 * pointers owned by the caller; TSDF/weight are updated in place on `stream` (cudaStream_t, NULL =
 * default stream); asynchronous like a kernel launch; the return value replaces cudaGetLastError(). */
int tsdf_Integrate(const float *cam_K, const float *cam2base, const float *depth_im, int im_height,
                   int im_width, int voxel_grid_dim_x, int voxel_grid_dim_y, int voxel_grid_dim_z,
                   float voxel_grid_origin_x, float voxel_grid_origin_y, float voxel_grid_origin_z,
                   float voxel_size, float trunc_margin, float *voxel_grid_TSDF,
                   float *voxel_grid_weight, void *stream);
/* Replaces SaveVoxelGrid2SurfacePointCloud (utils.hpp:14-17): HOST volume pointers; the scan runs
 * on the device and the file is byte-identical to the reference writer's. */
int tsdf_SaveVoxelGrid2SurfacePointCloud(const char *file_name, int voxel_grid_dim_x, int voxel_grid_dim_y,
                                         int voxel_grid_dim_z, float voxel_size, float voxel_grid_origin_x,
                                         float voxel_grid_origin_y, float voxel_grid_origin_z,
                                         const float *voxel_grid_TSDF, const float *voxel_grid_weight,
                                         float tsdf_thresh, float weight_thresh);

/* ---- TSDF consumer: truncated distance function of a point cloud (core/demo.cu) ------------------ */
/* Replaces the launch loop `for CUDA_LOOP_IDX: ComputeTDF<<<2880, 512>>>(CUDA_LOOP_IDX, ...)`
 * (core/demo.cu:22-27, :153-160): one call covers the whole grid.  Device pointers owned by the caller;
 * voxel_grid_TDF is read-modify-write exactly like the kernel (zero it first, core/demo.cu:140-141).
 * Bit-identical to the reference kernel, via a separable exact distance transform instead of the
 * (2R+1)^3 brute-force window. */
int tsdf_ComputeTDF(const float *voxel_grid_occ, float *voxel_grid_TDF, int voxel_grid_dim_x, int voxel_grid_dim_y,
                    int voxel_grid_dim_z, float voxel_grid_origin_x, float voxel_grid_origin_y,
                    float voxel_grid_origin_z, float voxel_size, float trunc_margin, void *stream);
/* core/demo.cu:100-126 (host): grid dimensions and origin from the cloud's bounding box + padding voxels. */
int tsdf_tdf_grid_from_points(const float *pts, int64_t n_pts, float voxel_size, int32_t padding, int32_t dims[3],
                              float origin[3]);
/* core/demo.cu:128-163: occupancy scatter + TDF of a HOST point cloud; tdf_host receives
 * dims[0]*dims[1]*dims[2] floats (x fastest), writable as tdf.bin with tsdf_write_bin. */
int tsdf_point_cloud_to_tdf(const float *pts, int64_t n_pts, const int32_t dims[3], const float origin[3],
                            float voxel_size, float trunc_margin, float *tdf_host);

/* ---- diagnostics ----------------------------------------------------------------------------- */
const char *tsdf_last_error(void);
const char *tsdf_version(void);
int tsdf_device_count(int32_t *count);
/* Device self-test of the pixel rounding helper against roundf() + the reference bounds test
 * (demo.cu:46-49) over all 2^32 float bit patterns; *mismatches receives the count (0 expected). */
int tsdf_selftest_rounding(int32_t device, int64_t *mismatches);
/* Device self-test of the packed (f32x2) set-up arithmetic of the fusion kernel against the scalar round-to-nearest
 * intrinsics (one IEEE rounding per operation of demo.cu:35-41) on 2.5e9 pseudo-random lanes; 0 expected. */
int tsdf_selftest_packed(int32_t device, int64_t *mismatches);
/* Device self-test of the pixel-rounding margin of the fast classification (the bound on the difference between the
 * kernel's pixel coordinate and the reference's, demo.cu:46-47): *rcp_error_ulp24 receives the largest relative error
 * of rcp.approx over every mantissa in units of 2^-24 (the bound assumes <= 4), *worst_fraction the largest observed
 * difference on 2.5e9 pseudo-random (cam, z) pairs for the intrinsics (fx, c0) and image extent `size`, as a fraction
 * of the margin (must stay below 1; the derivation leaves a factor 1.25). */
int tsdf_selftest_pixel_margin(int32_t device, float fx, float c0, float size, float *rcp_error_ulp24,
                               float *worst_fraction);

#ifdef __cplusplus
}
#endif
#endif /* TSDF_FUSION_H */
