/*
 * oracle/tsdf_oracle.h -- CPU restatement of the 3dmatch-toolbox depth-fusion path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (3dmatch-toolbox_b200/, include/)
 * links, imports or executes this code; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do, and only as the checker / the
 * reported CPU baseline.
 *
 * Parity pin: the reference ships no golden vectors or tests for this path (SURVEY.md
 * section 4).  The restatement is therefore pinned against the reference's OWN code,
 * compiled from /root/reference where it lies into oracle/_ref/ (see oracle/Makefile):
 *   - libref_host.so : the reference kernel body (depth-fusion/demo.cu:18-68) and the
 *     utils.hpp helpers compiled for the host by g++ (-ffp-contract=off), run here;
 *   - libref_cuda.so : the same file compiled unchanged by nvcc for sm_100a, run on
 *     the GPU box.
 * tests/test_oracle_pin.py checks every function below against libref_host.so.
 */
#ifndef TSDF_ORACLE_H
#define TSDF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One reference launch of Integrate (depth-fusion/demo.cu:18-68) restricted to the
 * z-planes [z_begin, z_end) of a dim_x*dim_y*dim_z grid.  tsdf/weight hold only that
 * slab, x fastest: slab[(z-z_begin)*dim_y*dim_x + y*dim_x + x].  Voxel coordinates
 * always use the GLOBAL index, so a slab result equals the same planes of a full run. */
void oracle_integrate(const float cam_K[9], const float cam2base[16], const float *depth_im,
                      int im_height, int im_width,
                      int64_t dim_x, int64_t dim_y, int64_t dim_z,
                      int64_t z_begin, int64_t z_end,
                      float origin_x, float origin_y, float origin_z,
                      float voxel_size, float trunc_margin,
                      float *tsdf, float *weight);

/* demo.cu:131-135: TSDF <- 1.0f, weight <- 0. */
void oracle_init_volume(int64_t n_voxels, float *tsdf, float *weight);

/* demo.cu:156-183 without the file I/O: n_frames launches in ascending frame order.
 * depth_frames is [n_frames][im_height*im_width], cam2base is [n_frames][16]. */
void oracle_fuse(const float cam_K[9], const float *cam2base, const float *depth_frames,
                 int n_frames, int im_height, int im_width,
                 int64_t dim_x, int64_t dim_y, int64_t dim_z, int64_t z_begin, int64_t z_end,
                 float origin_x, float origin_y, float origin_z,
                 float voxel_size, float trunc_margin, float *tsdf, float *weight);

/* utils.hpp:89-109 and :112-238. */
void oracle_multiply_matrix(const float m1[16], const float m2[16], float out[16]);
int  oracle_invert_matrix(const float m[16], float inv_out[16]);

/* utils.hpp:82-84: u16 millimetres -> float metres, >6 m -> 0. */
void oracle_depth_from_u16(const uint16_t *mm, int64_t n, float *depth);

/* utils.hpp:20-23 / :36-53 on the slab [z_begin, z_end): count, then emit xyz triples in
 * ascending linear index.  xyz may be NULL (count only).  Returns the number of points. */
int64_t oracle_surface_normals(const float *tsdf, const float *weight, int64_t dim_x, int64_t dim_y, int64_t nz,
                               float tsdf_thresh, float weight_thresh, float *normals);
int64_t oracle_extract_surface(const float *tsdf, const float *weight,
                               int64_t dim_x, int64_t dim_y, int64_t dim_z,
                               int64_t z_begin, int64_t z_end,
                               float origin_x, float origin_y, float origin_z, float voxel_size,
                               float tsdf_thresh, float weight_thresh, float *xyz);

/* utils.hpp:26-53: 7-line header + little-endian float triples.  Returns 0 on success. */
int oracle_write_ply(const char *path, const float *xyz, int64_t n_points);

/* demo.cu:196-213 (commented-out writer): 8-float header + dim^3 TSDF floats. */
int oracle_write_tsdf_bin(const char *path, int64_t dim_x, int64_t dim_y, int64_t dim_z,
                          float origin_x, float origin_y, float origin_z,
                          float voxel_size, float trunc_margin, const float *tsdf);

/* utils.hpp:60-70: M*N whitespace-separated floats, row-major.  Returns values read. */
int oracle_load_matrix(const char *path, int count, float *out);

/* core/demo.cu:22-57 -- ComputeTDF over the whole grid (all CUDA_LOOP_IDX passes), brute force, including the
 * reference's float index decomposition (core/demo.cu:33-35).  tdf is read-modify-write like the kernel
 * (the caller zero-initialises it, core/demo.cu:140-141). */
void oracle_compute_tdf(const float *occ, float *tdf, int dim_x, int dim_y, int dim_z,
                        float voxel_size, float trunc_margin);

/* core/demo.cu:100-136 -- grid placement and occupancy from a point cloud.  dims/origin are outputs;
 * occ (dims[0]*dims[1]*dims[2] floats, zero-initialised by the caller) may be NULL to query the size only. */
void oracle_tdf_grid_from_points(const float *pts, int64_t n_pts, float voxel_size, int padding,
                                 int dims[3], float origin[3], float *occ);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
