/*
 * oracle/tsdf_oracle.c -- CPU restatement of the 3dmatch-toolbox depth-fusion path.
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.h).  Plain C, fp32 throughout, one rounding
 * per source-level operation: build with -ffp-contract=off and without -ffast-math so
 * that no multiply-add is contracted (the "--fmad=false" contract of BASELINE.json).
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  The restatement is pinned against the reference's own code
 * compiled for the host (oracle/_ref/libref_host.so) by tests/test_oracle_pin.py.
 */
#include "tsdf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------
 * One voxel of depth-fusion/demo.cu:29-66.  gx/gy/gz are GLOBAL grid indices.
 * Returns 1 when the voxel was updated.
 *
 * Float -> int pixel conversion (demo.cu:46-47): the reference converts roundf(...) to
 * int and then range-checks (demo.cu:48).  For every value whose conversion is defined
 * this equals range-checking the rounded float; values outside int range / NaN are
 * undefined behaviour on the host (x86 yields INT_MIN -> rejected) and saturate on the
 * device (-> rejected).  The restatement rejects them.
 * ---------------------------------------------------------------------------------- */
static inline int fuse_voxel(const float *K, const float *c2b, const float *depth_im,
                             int im_height, int im_width,
                             int64_t gx, int64_t gy, int64_t gz,
                             float ox, float oy, float oz, float voxel_size, float trunc_margin,
                             float *tsdf_cell, float *weight_cell)
{
    /* demo.cu:30-32 -- voxel centre in base-frame coordinates (int -> float, mul, add) */
    float bx = ox + (float)gx * voxel_size;
    float by = oy + (float)gy * voxel_size;
    float bz = oz + (float)gz * voxel_size;

    /* demo.cu:35-41 -- base frame -> camera frame: R^T (p - t), row-major 4x4 */
    float dx = bx - c2b[3];
    float dy = by - c2b[7];
    float dz = bz - c2b[11];
    float cam_x = c2b[0] * dx + c2b[4] * dy + c2b[8] * dz;
    float cam_y = c2b[1] * dx + c2b[5] * dy + c2b[9] * dz;
    float cam_z = c2b[2] * dx + c2b[6] * dy + c2b[10] * dz;

    /* demo.cu:43-44 */
    if (cam_z <= 0)
        return 0;

    /* demo.cu:46-49 -- pinhole projection, round half away from zero, bounds */
    float u = roundf(K[0] * (cam_x / cam_z) + K[2]);
    float v = roundf(K[4] * (cam_y / cam_z) + K[5]);
    if (!(u >= 0.0f && u < (float)im_width && v >= 0.0f && v < (float)im_height))
        return 0;
    int px = (int)u;
    int py = (int)v;

    /* demo.cu:51-54 -- the depth gather */
    float d = depth_im[(int64_t)py * im_width + px];
    if (d <= 0 || d > 6)
        return 0;

    /* demo.cu:56-58 -- ray-length-corrected signed distance, sum evaluated left to right.
     * powf(q, 2) is written q*q: that is what a host compiler makes of the reference line (GCC
     * folds pow(x, 2) to x*x at every optimisation level -- oracle/_ref/libref_host.so imports
     * no powf, see tests/test_oracle_pin.py), whereas libm's powf differs from q*q by 1 ulp for
     * ~0.07 % of arguments.  The device powf of the reference's nvcc build is a third, few-ulp
     * variant; it is compared with a tolerance only. */
    float qx = cam_x / cam_z, qy = cam_y / cam_z;
    float diff = (d - cam_z) * sqrtf(1 + qx * qx + qy * qy);
    if (diff <= -trunc_margin)
        return 0;

    /* demo.cu:62-66 -- truncated running weighted average */
    float dist = fminf(1.0f, diff / trunc_margin);
    float w_old = *weight_cell;
    float w_new = w_old + 1.0f;
    *weight_cell = w_new;
    *tsdf_cell = (*tsdf_cell * w_old + dist) / w_new;
    return 1;
}

/* demo.cu:18-68 with blockIdx.x -> z, threadIdx.x -> y, serial x (demo.cu:24-27). */
void oracle_integrate(const float cam_K[9], const float cam2base[16], const float *depth_im,
                      int im_height, int im_width,
                      int64_t dim_x, int64_t dim_y, int64_t dim_z,
                      int64_t z_begin, int64_t z_end,
                      float origin_x, float origin_y, float origin_z,
                      float voxel_size, float trunc_margin,
                      float *tsdf, float *weight)
{
    (void)dim_z;
    int64_t rows = (z_end - z_begin) * dim_y;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t row = 0; row < rows; ++row) {
        int64_t z = z_begin + row / dim_y;
        int64_t y = row % dim_y;
        float *t_row = tsdf + row * dim_x;
        float *w_row = weight + row * dim_x;
        for (int64_t x = 0; x < dim_x; ++x)
            fuse_voxel(cam_K, cam2base, depth_im, im_height, im_width, x, y, z,
                       origin_x, origin_y, origin_z, voxel_size, trunc_margin,
                       t_row + x, w_row + x);
    }
}

/* demo.cu:131-135 */
void oracle_init_volume(int64_t n_voxels, float *tsdf, float *weight)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_voxels; ++i) {
        tsdf[i] = 1.0f;
        weight[i] = 0.0f;
    }
}

/* demo.cu:156-183 -- frames applied in ascending order */
void oracle_fuse(const float cam_K[9], const float *cam2base, const float *depth_frames,
                 int n_frames, int im_height, int im_width,
                 int64_t dim_x, int64_t dim_y, int64_t dim_z, int64_t z_begin, int64_t z_end,
                 float origin_x, float origin_y, float origin_z,
                 float voxel_size, float trunc_margin, float *tsdf, float *weight)
{
    int64_t px = (int64_t)im_height * im_width;
    for (int f = 0; f < n_frames; ++f)
        oracle_integrate(cam_K, cam2base + 16 * (int64_t)f, depth_frames + px * f,
                         im_height, im_width, dim_x, dim_y, dim_z, z_begin, z_end,
                         origin_x, origin_y, origin_z, voxel_size, trunc_margin, tsdf, weight);
}

/* utils.hpp:89-109 -- each entry ((a0*b0 + a1*b1) + a2*b2) + a3*b3, left to right */
void oracle_multiply_matrix(const float m1[16], const float m2[16], float out[16])
{
    float tmp[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float acc = m1[4 * r] * m2[c];
            acc = acc + m1[4 * r + 1] * m2[4 + c];
            acc = acc + m1[4 * r + 2] * m2[8 + c];
            acc = acc + m1[4 * r + 3] * m2[12 + c];
            tmp[4 * r + c] = acc;
        }
    memcpy(out, tmp, sizeof tmp);
}

/*
 * utils.hpp:112-238 -- adjugate / determinant inverse.  Each adjugate entry is a sum of
 * six signed triple products (m[a]*m[b])*m[c] accumulated left to right in the
 * reference's term order; the table lists {sign, a, b, c} per term.  A leading unary
 * minus in the reference ((-m[a])*m[b]*m[c]) equals negating the product (exact).
 */
typedef struct { signed char s, a, b, c; } adj_term;
static const adj_term ADJ[16][6] = {
    /* inv[0]  utils.hpp:115-120 */ {{+1,5,10,15},{-1,5,11,14},{-1,9,6,15},{+1,9,7,14},{+1,13,6,11},{-1,13,7,10}},
    /* inv[1]  utils.hpp:143-148 */ {{-1,1,10,15},{+1,1,11,14},{+1,9,2,15},{-1,9,3,14},{-1,13,2,11},{+1,13,3,10}},
    /* inv[2]  utils.hpp:171-176 */ {{+1,1,6,15},{-1,1,7,14},{-1,5,2,15},{+1,5,3,14},{+1,13,2,7},{-1,13,3,6}},
    /* inv[3]  utils.hpp:199-204 */ {{-1,1,6,11},{+1,1,7,10},{+1,5,2,11},{-1,5,3,10},{-1,9,2,7},{+1,9,3,6}},
    /* inv[4]  utils.hpp:122-127 */ {{-1,4,10,15},{+1,4,11,14},{+1,8,6,15},{-1,8,7,14},{-1,12,6,11},{+1,12,7,10}},
    /* inv[5]  utils.hpp:150-155 */ {{+1,0,10,15},{-1,0,11,14},{-1,8,2,15},{+1,8,3,14},{+1,12,2,11},{-1,12,3,10}},
    /* inv[6]  utils.hpp:178-183 */ {{-1,0,6,15},{+1,0,7,14},{+1,4,2,15},{-1,4,3,14},{-1,12,2,7},{+1,12,3,6}},
    /* inv[7]  utils.hpp:206-211 */ {{+1,0,6,11},{-1,0,7,10},{-1,4,2,11},{+1,4,3,10},{+1,8,2,7},{-1,8,3,6}},
    /* inv[8]  utils.hpp:129-134 */ {{+1,4,9,15},{-1,4,11,13},{-1,8,5,15},{+1,8,7,13},{+1,12,5,11},{-1,12,7,9}},
    /* inv[9]  utils.hpp:157-162 */ {{-1,0,9,15},{+1,0,11,13},{+1,8,1,15},{-1,8,3,13},{-1,12,1,11},{+1,12,3,9}},
    /* inv[10] utils.hpp:185-190 */ {{+1,0,5,15},{-1,0,7,13},{-1,4,1,15},{+1,4,3,13},{+1,12,1,7},{-1,12,3,5}},
    /* inv[11] utils.hpp:213-218 */ {{-1,0,5,11},{+1,0,7,9},{+1,4,1,11},{-1,4,3,9},{-1,8,1,7},{+1,8,3,5}},
    /* inv[12] utils.hpp:136-141 */ {{-1,4,9,14},{+1,4,10,13},{+1,8,5,14},{-1,8,6,13},{-1,12,5,10},{+1,12,6,9}},
    /* inv[13] utils.hpp:164-169 */ {{+1,0,9,14},{-1,0,10,13},{-1,8,1,14},{+1,8,2,13},{+1,12,1,10},{-1,12,2,9}},
    /* inv[14] utils.hpp:192-197 */ {{-1,0,5,14},{+1,0,6,13},{+1,4,1,14},{-1,4,2,13},{-1,12,1,6},{+1,12,2,5}},
    /* inv[15] utils.hpp:220-225 */ {{+1,0,5,10},{-1,0,6,9},{-1,4,1,10},{+1,4,2,9},{+1,8,1,6},{-1,8,2,5}},
};

int oracle_invert_matrix(const float m[16], float inv_out[16])
{
    float adj[16];
    for (int i = 0; i < 16; ++i) {
        float acc = 0.0f;
        for (int t = 0; t < 6; ++t) {
            const adj_term *k = &ADJ[i][t];
            float prod = (m[k->a] * m[k->b]) * m[k->c];
            if (t == 0)
                acc = (k->s > 0) ? prod : -prod;
            else
                acc = (k->s > 0) ? acc + prod : acc - prod;
        }
        adj[i] = acc;
    }
    /* utils.hpp:227-229 */
    float det = m[0] * adj[0] + m[1] * adj[4] + m[2] * adj[8] + m[3] * adj[12];
    if (det == 0)
        return 0;
    /* utils.hpp:232 -- double divide, stored to float */
    det = (float)(1.0 / (double)det);
    /* utils.hpp:234-235 */
    for (int i = 0; i < 16; ++i)
        inv_out[i] = adj[i] * det;
    return 1;
}

/* utils.hpp:82-84 */
void oracle_depth_from_u16(const uint16_t *mm, int64_t n, float *depth)
{
    for (int64_t i = 0; i < n; ++i) {
        float d = (float)mm[i] / 1000.0f;
        if (d > 6.0f)
            d = 0;
        depth[i] = d;
    }
}

/* utils.hpp:20-23 (count) and :36-53 (emit), 64-bit indices instead of int */
int64_t oracle_extract_surface(const float *tsdf, const float *weight,
                               int64_t dim_x, int64_t dim_y, int64_t dim_z,
                               int64_t z_begin, int64_t z_end,
                               float origin_x, float origin_y, float origin_z, float voxel_size,
                               float tsdf_thresh, float weight_thresh, float *xyz)
{
    (void)dim_z;
    int64_t n = (z_end - z_begin) * dim_y * dim_x;
    int64_t count = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (fabsf(tsdf[i]) < tsdf_thresh && weight[i] > weight_thresh) {
            if (xyz) {
                /* utils.hpp:42-44 -- linear index -> (x, y, z) */
                int64_t z = i / (dim_x * dim_y);
                int64_t rem = i - z * dim_x * dim_y;
                int64_t y = rem / dim_x;
                int64_t x = rem - y * dim_x;
                /* utils.hpp:47-49 */
                xyz[3 * count + 0] = origin_x + (float)x * voxel_size;
                xyz[3 * count + 1] = origin_y + (float)y * voxel_size;
                xyz[3 * count + 2] = origin_z + (float)(z + z_begin) * voxel_size;
            }
            ++count;
        }
    }
    return count;
}

/* Surface normals, parity UNPINNED: the only reference code is
 * PCL-benchmark/pclWrapperFPFH.h:308-330 (computeSurfaceNormal): central differences of the TSDF at the voxel,
 * normalised with mLib's vec3f::getNormalized() -- mLib (niessner/mLib, not vendored in the reference tree) computes
 * `val = 1.0f / sqrtf(x*x + y*y + z*z); (x*val, y*val, z*val)`; restated here with one rounding per operation.
 * One normal per surface voxel of the slab, in the order of oracle_extract_surface; (0, 0, 0) where a neighbour lies
 * outside the slab (isValidCoordinate fails, :309-315) or the gradient is zero (:320). */
int64_t oracle_surface_normals(const float *tsdf, const float *weight, int64_t dim_x, int64_t dim_y, int64_t nz,
                               float tsdf_thresh, float weight_thresh, float *normals)
{
    int64_t n = nz * dim_y * dim_x, plane = dim_x * dim_y;
    int64_t count = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (fabsf(tsdf[i]) < tsdf_thresh && weight[i] > weight_thresh) {
            int64_t z = i / plane, rem = i - z * plane, y = rem / dim_x, x = rem - y * dim_x;
            float nx = 0.0f, ny = 0.0f, nzv = 0.0f;
            if (x > 0 && x < dim_x - 1 && y > 0 && y < dim_y - 1 && z > 0 && z < nz - 1) {
                float gx = tsdf[i + 1] - tsdf[i - 1];
                float gy = tsdf[i + dim_x] - tsdf[i - dim_x];
                float gz = tsdf[i + plane] - tsdf[i - plane];
                if (!(gx == 0.0f && gy == 0.0f && gz == 0.0f)) {
                    float len2 = (gx * gx + gy * gy) + gz * gz;
                    float val = 1.0f / sqrtf(len2);
                    nx = gx * val; ny = gy * val; nzv = gz * val;
                }
            }
            normals[3 * count + 0] = nx;
            normals[3 * count + 1] = ny;
            normals[3 * count + 2] = nzv;
            ++count;
        }
    }
    return count;
}

/* utils.hpp:26-33 header, :50-52 payload */
int oracle_write_ply(const char *path, const float *xyz, int64_t n_points)
{
    FILE *fp = fopen(path, "w");
    if (!fp)
        return -1;
    fprintf(fp, "ply\n");
    fprintf(fp, "format binary_little_endian 1.0\n");
    fprintf(fp, "element vertex %lld\n", (long long)n_points);
    fprintf(fp, "property float x\n");
    fprintf(fp, "property float y\n");
    fprintf(fp, "property float z\n");
    fprintf(fp, "end_header\n");
    if (n_points > 0 && fwrite(xyz, sizeof(float) * 3, (size_t)n_points, fp) != (size_t)n_points) {
        fclose(fp);
        return -1;
    }
    return fclose(fp);
}

/* demo.cu:196-213 */
int oracle_write_tsdf_bin(const char *path, int64_t dim_x, int64_t dim_y, int64_t dim_z,
                          float origin_x, float origin_y, float origin_z,
                          float voxel_size, float trunc_margin, const float *tsdf)
{
    FILE *fp = fopen(path, "wb");
    if (!fp)
        return -1;
    float header[8] = {(float)dim_x, (float)dim_y, (float)dim_z, origin_x, origin_y, origin_z,
                       voxel_size, trunc_margin};
    size_t n = (size_t)(dim_x * dim_y * dim_z);
    int ok = fwrite(header, sizeof(float), 8, fp) == 8 && fwrite(tsdf, sizeof(float), n, fp) == n;
    return (fclose(fp) == 0 && ok) ? 0 : -1;
}

/* utils.hpp:60-70 */
int oracle_load_matrix(const char *path, int count, float *out)
{
    FILE *fp = fopen(path, "r");
    if (!fp)
        return -1;
    int got = 0;
    for (; got < count; ++got)
        if (fscanf(fp, "%f", &out[got]) != 1)
            break;
    fclose(fp);
    return got;
}

/* core/demo.cu:22-57.  One "thread" per voxel index; the index -> (x, y, z) decomposition is done in float
 * exactly as the kernel does (:33-35), which mis-addresses voxels once the index exceeds 2^24. */
void oracle_compute_tdf(const float *occ, float *tdf, int dim_x, int dim_y, int dim_z,
                        float voxel_size, float trunc_margin)
{
    const int64_t total = (int64_t)dim_x * dim_y * dim_z;
    /* core/demo.cu:37: (int)round(trunc_margin / voxel_size); nvcc resolves round(float) to roundf */
    const int radius = (int)roundf(trunc_margin / voxel_size);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t idx = 0; idx < total; ++idx) {
        int v = (int)idx;
        int pz = (int)floorf((float)v / ((float)dim_x * (float)dim_y));
        int py = (int)floorf(((float)v - ((float)pz * (float)dim_x * (float)dim_y)) / (float)dim_x);
        int px = (int)((float)v - ((float)pz * (float)dim_x * (float)dim_y) - ((float)py * (float)dim_x));
        if (occ[idx] > 0) {                                   /* :39-42 */
            tdf[idx] = 1.0f;
            continue;
        }
        int x0 = px - radius > 0 ? px - radius : 0, x1 = px + radius + 1 < dim_x ? px + radius + 1 : dim_x;
        int y0 = py - radius > 0 ? py - radius : 0, y1 = py + radius + 1 < dim_y ? py + radius + 1 : dim_y;
        int z0 = pz - radius > 0 ? pz - radius : 0, z1 = pz + radius + 1 < dim_z ? pz + radius + 1 : dim_z;
        for (int ix = x0; ix < x1; ++ix)                      /* :45-56 */
            for (int iy = y0; iy < y1; ++iy)
                for (int iz = z0; iz < z1; ++iz) {
                    int64_t n = (int64_t)iz * dim_x * dim_y + (int64_t)iy * dim_x + ix;
                    if (occ[n] > 0) {
                        float xd = (float)(px - ix), yd = (float)(py - iy), zd = (float)(pz - iz);
                        float dist = sqrtf(xd * xd + yd * yd + zd * zd) / (float)radius;
                        if ((1.0f - dist) > tdf[idx])
                            tdf[idx] = 1.0f - dist;
                    }
                }
    }
}

static float minf_(float a, float b) { return a < b ? a : b; }
static float maxf_(float a, float b) { return a > b ? a : b; }

/* core/demo.cu:100-136 */
void oracle_tdf_grid_from_points(const float *pts, int64_t n_pts, float voxel_size, int padding,
                                 int dims[3], float origin[3], float *occ)
{
    float lo[3] = {pts[0], pts[1], pts[2]}, hi[3] = {pts[0], pts[1], pts[2]};
    for (int64_t i = 0; i < n_pts; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = minf_(lo[a], pts[3 * i + a]);
            hi[a] = maxf_(hi[a], pts[3 * i + a]);
        }
    for (int a = 0; a < 3; ++a) {
        /* :119-121 -- round() of a float expression (C++ picks the float overload) + ints */
        dims[a] = (int)(roundf((hi[a] - lo[a]) / voxel_size) + 1 + padding * 2);
        /* :123-125 */
        origin[a] = lo[a] - padding * voxel_size + voxel_size / 2;
    }
    if (!occ)
        return;
    for (int64_t i = 0; i < n_pts; ++i) {                     /* :131-136 */
        int gx = (int)roundf((pts[3 * i + 0] - origin[0]) / voxel_size);
        int gy = (int)roundf((pts[3 * i + 1] - origin[1]) / voxel_size);
        int gz = (int)roundf((pts[3 * i + 2] - origin[2]) / voxel_size);
        occ[(int64_t)gz * dims[1] * dims[0] + (int64_t)gy * dims[0] + gx] = 1.0f;
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
