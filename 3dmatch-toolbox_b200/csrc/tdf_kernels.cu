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
// All values are small integers (d2 <= 3 R^2: 8 bits for R <= 9, else 16), the final expression uses the reference's own
// float operations (IEEE sqrt, divide, subtract), so the TDF is bit-identical -- including the reference's
// float index decomposition (core/demo.cu:33-35), which addresses the wrong neighbourhood once the voxel
// index exceeds 2^24; it is reproduced because downstream results depend on it.
#include "tsdf_kernels.h"

namespace tsdf {
namespace {

// Kernels (HBM-bound in the ideal case: occ 4 B + g1 1+1 B + g2 1+1 B + occ 4 B + tdf 4+4 B = 20 B per voxel with
// 8-bit intermediates):
//   x pass: the occupancy of 32 consecutive voxels is one ballot word; a thread finds the nearest occupied voxel of
//           its row inside the +-R window with two bit scans on the 96-bit neighbourhood of its warp (3 coalesced loads
//           instead of 2R+1).
//   y / z passes: a block stages (TA + 2R) lines of 128 voxels in shared memory once and produces TA lines from them
//           (1.3 global loads per voxel instead of 2R+1; the window reads hit shared memory, conflict-free because the
//           threads of a warp read consecutive columns of one line).
template <typename T> struct NoneOf;
template <> struct NoneOf<unsigned char> { static constexpr unsigned v = 0xffu; };
template <> struct NoneOf<unsigned short> { static constexpr unsigned v = 0xffffu; };

constexpr int kTdfTileX = 128;   // voxels of a line staged per block
constexpr int kTdfTileA = 32;    // lines produced per block along the pass axis
constexpr int kTdfMaxR = 31;     // one warp-neighbourhood (3 x 32 voxels) covers the x window

template <typename T>
__global__ void __launch_bounds__(256) tdf_pass_x_kernel(const float* __restrict__ occ, T* __restrict__ g1, int dim_x,
                                                         int64_t n, int radius) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (whole warps; n is padded by the guards below)
  const int lane = threadIdx.x & 31;
  const bool in = i < n;
  // occupancy bits of this warp's 32 voxels and of the 32 before / after them
  const unsigned m_prev = __ballot_sync(0xffffffffu, i - 32 >= 0 && i - 32 < n && __ldg(occ + (i - 32)) > 0.0f);
  const unsigned m_cur = __ballot_sync(0xffffffffu, in && __ldg(occ + i) > 0.0f);
  const unsigned m_next = __ballot_sync(0xffffffffu, i + 32 < n && __ldg(occ + (i + 32)) > 0.0f);
  if (!in) return;
  const int x = (int)(i % dim_x);
  const int left = min(radius, x), right = min(radius, dim_x - 1 - x);   // the window stays inside the row
  // bit p of `lo`: voxel (i - 1 - p) is occupied; bit p of `hi`: voxel (i + 1 + p) is occupied
  const unsigned long long w_lo = ((unsigned long long)m_cur << 32) | m_prev;      // bit (32 + lane) = this voxel
  const unsigned long long w_hi = ((unsigned long long)m_next << 32) | m_cur;      // bit lane = this voxel
  unsigned best = NoneOf<T>::v;
  if ((m_cur >> lane) & 1u) {
    best = 0u;
  } else {
    // below: bits (32 + lane - 1) ... (32 + lane - left) of w_lo, nearest first = highest bit first
    unsigned long long below = w_lo << (63 - (31 + lane));       // voxel i - 1 at bit 63
    below &= (left >= 64) ? ~0ull : ~(~0ull >> left);            // keep the top `left` bits
    if (left > 0 && below != 0ull) {
      const unsigned d = (unsigned)__clzll((long long)below) + 1u;
      best = d * d;
    }
    unsigned long long above = w_hi >> (lane + 1);                // voxel i + 1 at bit 0
    above &= (right >= 64) ? ~0ull : ((1ull << right) - 1ull);
    if (right > 0 && above != 0ull) {
      const unsigned d = (unsigned)__ffsll((long long)above);
      best = min(best, d * d);
    }
  }
  g1[i] = (T)best;
}

// out(line a) = min over |da| <= R of da^2 + in(line a + da), lines `stride` elements apart, `len` lines; the other
// two axes are x (contiguous, tiled by kTdfTileX) and `outer` (blockIdx.z; stride `outer_stride`).
template <typename T>
__global__ void __launch_bounds__(kTdfTileX) tdf_pass_axis_kernel(const T* __restrict__ in, T* __restrict__ out, int dim_x,
                                                                  int len, int64_t stride, int64_t outer_stride,
                                                                  int radius) {
  extern __shared__ unsigned char s_raw[];
  T* s = reinterpret_cast<T*>(s_raw);                       // [(kTdfTileA + 2R)][kTdfTileX]
  const int x = blockIdx.x * kTdfTileX + threadIdx.x;
  const int a0 = blockIdx.y * kTdfTileA;
  const int64_t base = (int64_t)blockIdx.z * outer_stride + x;
  const int rows = kTdfTileA + 2 * radius;
  for (int r = 0; r < rows; ++r) {
    const int a = a0 - radius + r;
    unsigned v = NoneOf<T>::v;
    if (x < dim_x && a >= 0 && a < len) v = __ldg(in + base + (int64_t)a * stride);
    s[r * kTdfTileX + threadIdx.x] = (T)v;
  }
  // (every thread reads only its own column: no barrier needed)
  if (x >= dim_x) return;
  for (int t = 0; t < kTdfTileA; ++t) {
    const int a = a0 + t;
    if (a >= len) break;
    unsigned best = NoneOf<T>::v;
    for (int d = -radius; d <= radius; ++d) {
      const unsigned v = s[(t + radius + d) * kTdfTileX + threadIdx.x];
      if (v != NoneOf<T>::v) best = min(best, v + (unsigned)(d * d));
    }
    out[base + (int64_t)a * stride] = (T)min(best, NoneOf<T>::v);
  }
}

// Final pass = the reference thread of voxel_idx (core/demo.cu:29-56) with the two inner loops replaced by g2: one
// block produces kTdfTileA z-planes of a 128-voxel row segment from g2 lines staged in shared memory.
template <typename T>
__global__ void __launch_bounds__(kTdfTileX) tdf_pass_z_kernel(const float* __restrict__ occ, const T* __restrict__ g2,
                                                               float* __restrict__ tdf, int dim_x, int dim_y, int dim_z,
                                                               int radius) {
  extern __shared__ unsigned char s_raw[];
  T* s = reinterpret_cast<T*>(s_raw);
  const int x = blockIdx.x * kTdfTileX + threadIdx.x;
  const int y = blockIdx.z;
  const int z0t = blockIdx.y * kTdfTileA;
  const int64_t plane = (int64_t)dim_x * dim_y;
  const int64_t col = (int64_t)y * dim_x + x;
  const int rows = kTdfTileA + 2 * radius;
  for (int r = 0; r < rows; ++r) {
    const int z = z0t - radius + r;
    unsigned v = NoneOf<T>::v;
    if (x < dim_x && z >= 0 && z < dim_z) v = __ldg(g2 + (int64_t)z * plane + col);
    s[r * kTdfTileX + threadIdx.x] = (T)v;
  }
  if (x >= dim_x) return;
  const float plane_f = __fmul_rn((float)dim_x, (float)dim_y);
  for (int t = 0; t < kTdfTileA; ++t) {
    const int z = z0t + t;
    if (z >= dim_z) break;
    const int64_t i = (int64_t)z * plane + col;
    if (__ldg(occ + i) > 0.0f) {  // core/demo.cu:39-42
      tdf[i] = 1.0f;
      continue;
    }
    // core/demo.cu:33-35, in float like the reference: beyond 2^24 voxels the decomposition can name another column
    const int v = (int)i;
    const int pz = (int)floorf(__fdiv_rn((float)v, plane_f));
    const float rem = __fsub_rn((float)v, __fmul_rn(__fmul_rn((float)pz, (float)dim_x), (float)dim_y));
    const int py = (int)floorf(__fdiv_rn(rem, (float)dim_x));
    const int px = (int)__fsub_rn(rem, __fmul_rn((float)py, (float)dim_x));
    const int zlo = max(0, pz - radius), zhi = min(dim_z, pz + radius + 1);
    unsigned best = NoneOf<T>::v;
    if (px == x && py == y && pz == z) {
      for (int d = -radius; d <= radius; ++d) {     // (lines outside the grid were staged as "none")
        const unsigned g = s[(t + radius + d) * kTdfTileX + threadIdx.x];
        if (g != NoneOf<T>::v) best = min(best, g + (unsigned)(d * d));
      }
    } else if (px >= 0 && px < dim_x && py >= 0 && py < dim_y) {
      const int64_t pcol = (int64_t)py * dim_x + px;
      for (int iz = zlo; iz < zhi; ++iz) {
        const unsigned g = __ldg(g2 + (int64_t)iz * plane + pcol);
        if (g != NoneOf<T>::v) best = min(best, g + (unsigned)((pz - iz) * (pz - iz)));
      }
    } else {
      // decomposed column outside the grid: the reference's clipped loops
      const int x0 = max(0, px - radius), x1 = min(dim_x, px + radius + 1);
      const int y0 = max(0, py - radius), y1 = min(dim_y, py + radius + 1);
      for (int ix = x0; ix < x1; ++ix)
        for (int iy = y0; iy < y1; ++iy)
          for (int iz = zlo; iz < zhi; ++iz)
            if (__ldg(occ + ((int64_t)iz * dim_y + iy) * dim_x + ix) > 0.0f)
              best = min(best, (unsigned)((px - ix) * (px - ix) + (py - iy) * (py - iy) + (pz - iz) * (pz - iz)));
    }
    if (best != NoneOf<T>::v) {
      // core/demo.cu:52-54 at the nearest occupied voxel (the maximum of a monotone function of d2)
      const float dist = __fdiv_rn(__fsqrt_rn((float)best), (float)radius);
      const float val = __fsub_rn(1.0f, dist);
      if (val > tdf[i]) tdf[i] = val;
    }
  }
}

// The brute-force window for radii the bit-parallel x pass does not cover (R > 31).
__global__ void __launch_bounds__(256) tdf_brute_kernel(const float* __restrict__ occ, float* __restrict__ tdf, int dim_x,
                                                        int dim_y, int dim_z, int64_t n, int radius) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (__ldg(occ + i) > 0.0f) { tdf[i] = 1.0f; return; }
  const int v = (int)i;
  const float plane_f = __fmul_rn((float)dim_x, (float)dim_y);
  const int pz = (int)floorf(__fdiv_rn((float)v, plane_f));
  const float rem = __fsub_rn((float)v, __fmul_rn(__fmul_rn((float)pz, (float)dim_x), (float)dim_y));
  const int py = (int)floorf(__fdiv_rn(rem, (float)dim_x));
  const int px = (int)__fsub_rn(rem, __fmul_rn((float)py, (float)dim_x));
  unsigned best = 0xffffffffu;
  for (int ix = max(0, px - radius); ix < min(dim_x, px + radius + 1); ++ix)
    for (int iy = max(0, py - radius); iy < min(dim_y, py + radius + 1); ++iy)
      for (int iz = max(0, pz - radius); iz < min(dim_z, pz + radius + 1); ++iz)
        if (__ldg(occ + ((int64_t)iz * dim_y + iy) * dim_x + ix) > 0.0f)
          best = min(best, (unsigned)((px - ix) * (px - ix) + (py - iy) * (py - iy) + (pz - iz) * (pz - iz)));
  if (best != 0xffffffffu) {
    const float val = __fsub_rn(1.0f, __fdiv_rn(__fsqrt_rn((float)best), (float)radius));
    if (val > tdf[i]) tdf[i] = val;
  }
}

template <typename T>
cudaError_t compute_tdf_t(const float* occ, float* tdf, void* scratch, int dim_x, int dim_y, int dim_z, int radius,
                          cudaStream_t s) {
  const int64_t n = (int64_t)dim_x * dim_y * dim_z;
  T* g1 = static_cast<T*>(scratch);
  T* g2 = g1 + n;
  const int64_t plane = (int64_t)dim_x * dim_y;
  const size_t smem = (size_t)(kTdfTileA + 2 * radius) * kTdfTileX * sizeof(T);
  const unsigned xt = (unsigned)((dim_x + kTdfTileX - 1) / kTdfTileX);
  if (dim_y > 65535 || dim_z > 65535) return cudaErrorInvalidConfiguration;
  tdf_pass_x_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(occ, g1, dim_x, n, radius);
  // y pass: lines dim_x apart, outer axis z
  tdf_pass_axis_kernel<T><<<dim3(xt, (unsigned)((dim_y + kTdfTileA - 1) / kTdfTileA), (unsigned)dim_z), kTdfTileX, smem, s>>>(
      g1, g2, dim_x, dim_y, dim_x, plane, radius);
  // z pass + the reference's final expression: lines a plane apart, outer axis y
  tdf_pass_z_kernel<T><<<dim3(xt, (unsigned)((dim_z + kTdfTileA - 1) / kTdfTileA), (unsigned)dim_y), kTdfTileX, smem, s>>>(
      occ, g2, tdf, dim_x, dim_y, dim_z, radius);
  return cudaGetLastError();
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

size_t compute_tdf_scratch_bytes(int dim_x, int dim_y, int dim_z, int radius) {
  const size_t n = (size_t)dim_x * dim_y * dim_z;
  if (radius > kTdfMaxR) return 0;
  return 2 * n * ((3 * radius * radius < 255) ? sizeof(unsigned char) : sizeof(unsigned short));
}

cudaError_t launch_compute_tdf(const float* occ, float* tdf, void* scratch, int dim_x, int dim_y, int dim_z,
                               int radius, cudaStream_t s) {
  if (radius > kTdfMaxR) {
    const int64_t n = (int64_t)dim_x * dim_y * dim_z;
    tdf_brute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(occ, tdf, dim_x, dim_y, dim_z, n, radius);
    return cudaGetLastError();
  }
  if (3 * radius * radius < 255) return compute_tdf_t<unsigned char>(occ, tdf, scratch, dim_x, dim_y, dim_z, radius, s);
  return compute_tdf_t<unsigned short>(occ, tdf, scratch, dim_x, dim_y, dim_z, radius, s);
}

cudaError_t launch_tdf_scatter_points(const float* pts, int64_t n_pts, float* occ, int dim_x, int dim_y, int dim_z,
                                      float ox, float oy, float oz, float voxel_size, cudaStream_t s) {
  if (n_pts <= 0) return cudaSuccess;
  tdf_scatter_points_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, s>>>(pts, n_pts, occ, dim_x, dim_y, dim_z, ox, oy,
                                                                           oz, voxel_size);
  return cudaGetLastError();
}

}  // namespace tsdf
