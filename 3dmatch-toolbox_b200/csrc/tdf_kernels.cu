// tdf_kernels.cu -- truncated distance function of an occupancy grid: replaces the reference's brute-force
// ComputeTDF kernel (core/demo.cu:22-57) and the occupancy scatter of core/demo.cu:128-136.
//
// The reference looks, for every empty voxel, at all (2R+1)^3 neighbours (R = round(trunc_margin / voxel_size),
// 1331 loads for the default R = 5) and keeps max(1 - sqrt(dx^2+dy^2+dz^2) / R).  1 - sqrt(d2)/R is monotone in
// the integer d2, so the result is that expression at the MINIMUM squared distance inside the cube window,
// and that minimum is separable:
//   g1(x,y,z) = min over |dx| <= R, occupied (x+dx,y,z)   of dx^2
//   g2(x,y,z) = min over |dy| <= R                        of dy^2 + g1(x,y+dy,z)
//   d2(x,y,z) = min over |dz| <= R                        of dz^2 + g2(x,y,z+dz)
// (3 x (2R+1) cached loads).  All values are small integers, the final expression uses the reference's own
// float operations (IEEE sqrt, divide, subtract), so the TDF is bit-identical -- including the reference's
// float index decomposition (core/demo.cu:33-35), which addresses the wrong neighbourhood once the voxel
// index exceeds 2^24; it is reproduced because downstream results depend on it.
#include "tsdf_kernels.h"

namespace tsdf {
namespace {

constexpr unsigned short kNone = 0xffffu;  // no occupied voxel in the window

__global__ void __launch_bounds__(256) tdf_pass_x_kernel(const float* __restrict__ occ, unsigned short* __restrict__ g1,
                                                         int dim_x, int64_t n, int radius) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % dim_x);
  const int lo = max(0, x - radius), hi = min(dim_x, x + radius + 1);
  unsigned best = kNone;
  for (int ix = lo; ix < hi; ++ix)
    if (__ldg(occ + i + (ix - x)) > 0.0f) best = min(best, (unsigned)((ix - x) * (ix - x)));
  g1[i] = (unsigned short)best;
}

__global__ void __launch_bounds__(256) tdf_pass_y_kernel(const unsigned short* __restrict__ g1,
                                                         unsigned short* __restrict__ g2, int dim_x, int dim_y, int64_t n,
                                                         int radius) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int y = (int)((i / dim_x) % dim_y);
  const int lo = max(0, y - radius), hi = min(dim_y, y + radius + 1);
  unsigned best = kNone;
  for (int iy = lo; iy < hi; ++iy) {
    const unsigned v = __ldg(g1 + i + (int64_t)(iy - y) * dim_x);
    if (v != kNone) best = min(best, v + (unsigned)((iy - y) * (iy - y)));
  }
  g2[i] = (unsigned short)best;
}

// Final pass = the reference thread of voxel_idx (core/demo.cu:29-56) with the two inner loops replaced by g2.
__global__ void __launch_bounds__(256) tdf_pass_z_kernel(const float* __restrict__ occ, const unsigned short* __restrict__ g2,
                                                         float* __restrict__ tdf, int dim_x, int dim_y, int dim_z, int64_t n,
                                                         int radius) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (__ldg(occ + i) > 0.0f) {  // core/demo.cu:39-42
    tdf[i] = 1.0f;
    return;
  }
  // core/demo.cu:33-35, in float like the reference
  const int v = (int)i;
  const float plane_f = __fmul_rn((float)dim_x, (float)dim_y);
  const int pz = (int)floorf(__fdiv_rn((float)v, plane_f));
  const float rem = __fsub_rn((float)v, __fmul_rn(__fmul_rn((float)pz, (float)dim_x), (float)dim_y));
  const int py = (int)floorf(__fdiv_rn(rem, (float)dim_x));
  const int px = (int)__fsub_rn(rem, __fmul_rn((float)py, (float)dim_x));
  const int z0 = max(0, pz - radius), z1 = min(dim_z, pz + radius + 1);
  unsigned best = kNone;
  if (px >= 0 && px < dim_x && py >= 0 && py < dim_y) {
    const int64_t col = (int64_t)py * dim_x + px;
    const int64_t plane = (int64_t)dim_x * dim_y;
    for (int iz = z0; iz < z1; ++iz) {
      const unsigned g = __ldg(g2 + (int64_t)iz * plane + col);
      if (g != kNone) best = min(best, g + (unsigned)((pz - iz) * (pz - iz)));
    }
  } else {
    // decomposed column outside the grid (only possible beyond 2^24 voxels): the reference's clipped loops
    const int x0 = max(0, px - radius), x1 = min(dim_x, px + radius + 1);
    const int y0 = max(0, py - radius), y1 = min(dim_y, py + radius + 1);
    for (int ix = x0; ix < x1; ++ix)
      for (int iy = y0; iy < y1; ++iy)
        for (int iz = z0; iz < z1; ++iz)
          if (__ldg(occ + ((int64_t)iz * dim_y + iy) * dim_x + ix) > 0.0f)
            best = min(best, (unsigned)((px - ix) * (px - ix) + (py - iy) * (py - iy) + (pz - iz) * (pz - iz)));
  }
  if (best != kNone) {
    // core/demo.cu:52-54 at the nearest occupied voxel (the maximum of a monotone function of d2)
    const float dist = __fdiv_rn(__fsqrt_rn((float)best), (float)radius);
    const float val = __fsub_rn(1.0f, dist);
    if (val > tdf[i]) tdf[i] = val;
  }
}

// core/demo.cu:131-136: occupancy[round((p - origin) / voxel_size)] = 1
__global__ void __launch_bounds__(256) tdf_scatter_points_kernel(const float* __restrict__ pts, int64_t n_pts,
                                                                 float* __restrict__ occ, int dim_x, int dim_y, int dim_z,
                                                                 float ox, float oy, float oz, float voxel_size) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const int gx = (int)roundf(__fdiv_rn(__fsub_rn(pts[3 * i + 0], ox), voxel_size));
  const int gy = (int)roundf(__fdiv_rn(__fsub_rn(pts[3 * i + 1], oy), voxel_size));
  const int gz = (int)roundf(__fdiv_rn(__fsub_rn(pts[3 * i + 2], oz), voxel_size));
  if (gx < 0 || gx >= dim_x || gy < 0 || gy >= dim_y || gz < 0 || gz >= dim_z) return;  // (the reference would write out of bounds)
  occ[((int64_t)gz * dim_y + gy) * dim_x + gx] = 1.0f;
}

}  // namespace

cudaError_t launch_compute_tdf(const float* occ, float* tdf, unsigned short* scratch, int dim_x, int dim_y, int dim_z,
                               int radius, cudaStream_t s) {
  const int64_t n = (int64_t)dim_x * dim_y * dim_z;
  unsigned short* g1 = scratch;
  unsigned short* g2 = scratch + n;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  tdf_pass_x_kernel<<<blocks, 256, 0, s>>>(occ, g1, dim_x, n, radius);
  tdf_pass_y_kernel<<<blocks, 256, 0, s>>>(g1, g2, dim_x, dim_y, n, radius);
  tdf_pass_z_kernel<<<blocks, 256, 0, s>>>(occ, g2, tdf, dim_x, dim_y, dim_z, n, radius);
  return cudaGetLastError();
}

cudaError_t launch_tdf_scatter_points(const float* pts, int64_t n_pts, float* occ, int dim_x, int dim_y, int dim_z,
                                      float ox, float oy, float oz, float voxel_size, cudaStream_t s) {
  if (n_pts <= 0) return cudaSuccess;
  tdf_scatter_points_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, s>>>(pts, n_pts, occ, dim_x, dim_y, dim_z, ox, oy,
                                                                           oz, voxel_size);
  return cudaGetLastError();
}

}  // namespace tsdf
