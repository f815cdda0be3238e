// tsdf_host.cpp -- host half of the C ABI (include/tsdf_fusion.h): pose algebra with the
// reference's rounding, the tsdf.ply / tsdf.bin writers, the text-matrix loader and a 16-bit
// grayscale PNG reader (zlib only; OpenCV is not a dependency).
//
// Build with -ffp-contract=off: multiply_matrix / invert_matrix must round after every operation
// exactly as the reference's host build does (depth-fusion/utils.hpp:89-238), because their
// results feed the Integrate kernel.
#include "tsdf_fusion.h"
#include "tsdf_host.h"

#include <zlib.h>

#include <cmath>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

int fail(int code, const std::string& msg) {
  tsdf_set_error_text(msg.c_str());
  return code;
}

// Adjugate terms of utils.hpp:115-225, one row per output entry, six signed triple products
// each, in the reference's order.  Encoding: bit 15 = subtract, nibbles = indices a, b, c of
// (m[a] * m[b]) * m[c].
const unsigned short kAdjugate[16][6] = {
    {0x05af, 0x85be, 0x896f, 0x097e, 0x0d6b, 0x8d7a},  // adj[0]   utils.hpp:115-120
    {0x81af, 0x01be, 0x092f, 0x893e, 0x8d2b, 0x0d3a},  // adj[1]   utils.hpp:143-148
    {0x016f, 0x817e, 0x852f, 0x053e, 0x0d27, 0x8d36},  // adj[2]   utils.hpp:171-176
    {0x816b, 0x017a, 0x052b, 0x853a, 0x8927, 0x0936},  // adj[3]   utils.hpp:199-204
    {0x84af, 0x04be, 0x086f, 0x887e, 0x8c6b, 0x0c7a},  // adj[4]   utils.hpp:122-127
    {0x00af, 0x80be, 0x882f, 0x083e, 0x0c2b, 0x8c3a},  // adj[5]   utils.hpp:150-155
    {0x806f, 0x007e, 0x042f, 0x843e, 0x8c27, 0x0c36},  // adj[6]   utils.hpp:178-183
    {0x006b, 0x807a, 0x842b, 0x043a, 0x0827, 0x8836},  // adj[7]   utils.hpp:206-211
    {0x049f, 0x84bd, 0x885f, 0x087d, 0x0c5b, 0x8c79},  // adj[8]   utils.hpp:129-134
    {0x809f, 0x00bd, 0x081f, 0x883d, 0x8c1b, 0x0c39},  // adj[9]   utils.hpp:157-162
    {0x005f, 0x807d, 0x841f, 0x043d, 0x0c17, 0x8c35},  // adj[10]  utils.hpp:185-190
    {0x805b, 0x0079, 0x041b, 0x8439, 0x8817, 0x0835},  // adj[11]  utils.hpp:213-218
    {0x849e, 0x04ad, 0x085e, 0x886d, 0x8c5a, 0x0c69},  // adj[12]  utils.hpp:136-141
    {0x009e, 0x80ad, 0x881e, 0x082d, 0x0c1a, 0x8c29},  // adj[13]  utils.hpp:164-169
    {0x805e, 0x006d, 0x041e, 0x842d, 0x8c16, 0x0c25},  // adj[14]  utils.hpp:192-197
    {0x005a, 0x8069, 0x841a, 0x0429, 0x0816, 0x8825},  // adj[15]  utils.hpp:220-225
};

unsigned be32(const unsigned char* p) {
  return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | (unsigned)p[3];
}

int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  if (pa <= pb && pa <= pc) return a;
  return (pb <= pc) ? b : c;
}

}  // namespace

extern "C" {

// utils.hpp:89-109: ((a0*b0 + a1*b1) + a2*b2) + a3*b3 per entry.
void tsdf_multiply_matrix(const float m1[16], const float m2[16], float m_out[16]) {
  float r[16];
  for (int i = 0; i < 16; ++i) {
    const float* a = m1 + (i & ~3);
    const float* b = m2 + (i & 3);
    volatile float acc = a[0] * b[0];
    acc = acc + a[1] * b[4];
    acc = acc + a[2] * b[8];
    acc = acc + a[3] * b[12];
    r[i] = acc;
  }
  memcpy(m_out, r, sizeof r);
}

// utils.hpp:112-238.
int tsdf_invert_matrix(const float m[16], float inv_out[16]) {
  float adj[16];
  for (int i = 0; i < 16; ++i) {
    float acc = 0.f;
    for (int t = 0; t < 6; ++t) {
      const unsigned code = kAdjugate[i][t];
      const float prod = (m[(code >> 8) & 15] * m[(code >> 4) & 15]) * m[code & 15];
      const bool neg = (code & 0x8000u) != 0;
      if (t == 0)
        acc = neg ? -prod : prod;
      else
        acc = neg ? acc - prod : acc + prod;
    }
    adj[i] = acc;
  }
  float det = m[0] * adj[0] + m[1] * adj[4] + m[2] * adj[8] + m[3] * adj[12];  // utils.hpp:227
  if (det == 0) return 0;                                                       // utils.hpp:229
  det = (float)(1.0 / (double)det);                                             // utils.hpp:232
  for (int i = 0; i < 16; ++i) inv_out[i] = adj[i] * det;                       // utils.hpp:234-235
  return 1;
}

// utils.hpp:82-84
void tsdf_depth_from_u16(const uint16_t* depth_mm, int64_t n, float* depth) {
  for (int64_t i = 0; i < n; ++i) {
    float d = (float)depth_mm[i] / 1000.0f;
    if (d > 6.0f) d = 0;
    depth[i] = d;
  }
}

// core/demo.cu:100-126: bounding box of the point cloud -> grid dimensions (with padding) and origin.
int tsdf_tdf_grid_from_points(const float* pts, int64_t n_pts, float voxel_size, int32_t padding, int32_t dims[3],
                              float origin[3]) {
  if (!pts || n_pts <= 0 || !dims || !origin || !(voxel_size > 0)) return fail(TSDF_ERR_ARG, "tsdf_tdf_grid_from_points: bad argument");
  float lo[3] = {pts[0], pts[1], pts[2]}, hi[3] = {pts[0], pts[1], pts[2]};
  for (int64_t i = 0; i < n_pts; ++i)
    for (int a = 0; a < 3; ++a) {
      const float p = pts[3 * i + a];
      if (p < lo[a]) lo[a] = p;
      if (p > hi[a]) hi[a] = p;
    }
  for (int a = 0; a < 3; ++a) {
    dims[a] = (int32_t)(roundf((hi[a] - lo[a]) / voxel_size) + 1 + padding * 2);   // :119-121
    origin[a] = lo[a] - padding * voxel_size + voxel_size / 2;                      // :123-125
  }
  return TSDF_OK;
}

// utils.hpp:26-33 (header, exactly 7 lines: core/demo.cu:76-86 parses it by line count) and
// :50-52 (payload).  The reference opens the file in text mode and fwrite()s 4 bytes at a
// time; the bytes on disk are the same as one bulk write.
int tsdf_write_ply(const char* path, const float* xyz, int64_t n_points) {
  if (!path || n_points < 0 || (n_points > 0 && !xyz)) return fail(TSDF_ERR_ARG, "tsdf_write_ply: bad argument");
  FILE* fp = fopen(path, "w");
  if (!fp) return fail(TSDF_ERR_IO, std::string("cannot open for writing: ") + path);
  fprintf(fp, "ply\nformat binary_little_endian 1.0\nelement vertex %lld\n", (long long)n_points);
  fprintf(fp, "property float x\nproperty float y\nproperty float z\nend_header\n");
  const size_t want = (size_t)n_points;
  const bool ok = (want == 0) || fwrite(xyz, 3 * sizeof(float), want, fp) == want;
  if (fclose(fp) != 0 || !ok) return fail(TSDF_ERR_IO, std::string("short write: ") + path);
  return TSDF_OK;
}

// demo.cu:196-213
int tsdf_write_bin(const char* path, int64_t dim_x, int64_t dim_y, int64_t dim_z, float origin_x, float origin_y,
                   float origin_z, float voxel_size, float trunc_margin, const float* tsdf_host) {
  if (!path || !tsdf_host || dim_x <= 0 || dim_y <= 0 || dim_z <= 0)
    return fail(TSDF_ERR_ARG, "tsdf_write_bin: bad argument");
  FILE* fp = fopen(path, "wb");
  if (!fp) return fail(TSDF_ERR_IO, std::string("cannot open for writing: ") + path);
  const float header[8] = {(float)dim_x, (float)dim_y, (float)dim_z, origin_x, origin_y, origin_z, voxel_size,
                           trunc_margin};
  const size_t n = (size_t)dim_x * (size_t)dim_y * (size_t)dim_z;
  const bool ok = fwrite(header, sizeof(float), 8, fp) == 8 && fwrite(tsdf_host, sizeof(float), n, fp) == n;
  if (fclose(fp) != 0 || !ok) return fail(TSDF_ERR_IO, std::string("short write: ") + path);
  return TSDF_OK;
}

// utils.hpp:60-70 (fscanf "%f" accepts the sample's 3-digit exponents such as 9.07e-001)
int tsdf_load_matrix(const char* path, int32_t rows, int32_t cols, float* out) {
  if (!path || !out || rows <= 0 || cols <= 0) return fail(TSDF_ERR_ARG, "tsdf_load_matrix: bad argument");
  FILE* fp = fopen(path, "r");
  if (!fp) return fail(TSDF_ERR_IO, std::string("cannot open: ") + path);
  int got = 0;
  for (; got < rows * cols; ++got)
    if (fscanf(fp, "%f", &out[got]) != 1) break;
  fclose(fp);
  if (got != rows * cols) return fail(TSDF_ERR_FORMAT, std::string("too few numbers in ") + path);
  return TSDF_OK;
}

// 16-bit grayscale, non-interlaced PNG -> row-major uint16 (what cv::imread(..., UNCHANGED)
// yields for the reference's depth frames, utils.hpp:75).
int tsdf_read_depth_png_u16(const char* path, int32_t im_height, int32_t im_width, uint16_t* depth_mm) {
  if (!path || !depth_mm || im_height <= 0 || im_width <= 0)
    return fail(TSDF_ERR_ARG, "tsdf_read_depth_png: bad argument");
  FILE* fp = fopen(path, "rb");
  if (!fp) return fail(TSDF_ERR_IO, std::string("cannot open: ") + path);
  std::vector<unsigned char> file;
  unsigned char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof buf, fp)) > 0) file.insert(file.end(), buf, buf + got);
  fclose(fp);
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0)
    return fail(TSDF_ERR_FORMAT, std::string("not a PNG file: ") + path);
  std::vector<unsigned char> idat;
  unsigned width = 0, height = 0;
  bool have_ihdr = false, have_end = false;
  for (size_t pos = 8; pos + 12 <= file.size();) {
    const unsigned len = be32(&file[pos]);
    const unsigned char* type = &file[pos + 4];
    if ((size_t)len > file.size() - pos - 12) return fail(TSDF_ERR_FORMAT, std::string("truncated PNG: ") + path);
    const unsigned char* data = &file[pos + 8];
    if (memcmp(type, "IHDR", 4) == 0) {
      if (len != 13) return fail(TSDF_ERR_FORMAT, std::string("bad IHDR: ") + path);
      width = be32(data);
      height = be32(data + 4);
      if (data[8] != 16 || data[9] != 0 || data[10] != 0 || data[11] != 0 || data[12] != 0)
        return fail(TSDF_ERR_FORMAT, std::string("depth PNG must be 16-bit grayscale, non-interlaced: ") + path);
      have_ihdr = true;
    } else if (memcmp(type, "IDAT", 4) == 0) {
      idat.insert(idat.end(), data, data + len);
    } else if (memcmp(type, "IEND", 4) == 0) {
      have_end = true;
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || !have_end || idat.empty()) return fail(TSDF_ERR_FORMAT, std::string("incomplete PNG: ") + path);
  if ((int64_t)width != im_width || (int64_t)height != im_height)
    return fail(TSDF_ERR_FORMAT, std::string("depth PNG is not the expected size: ") + path);
  const size_t stride = (size_t)width * 2;
  std::vector<unsigned char> raw((stride + 1) * height);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
    return fail(TSDF_ERR_FORMAT, std::string("PNG inflate failed: ") + path);
  std::vector<unsigned char> prev(stride, 0), cur(stride);
  for (unsigned r = 0; r < height; ++r) {
    const unsigned char* line = &raw[r * (stride + 1)];
    const int filter = line[0];
    for (size_t i = 0; i < stride; ++i) {
      const int a = (i >= 2) ? cur[i - 2] : 0, b = prev[i], c = (i >= 2) ? prev[i - 2] : 0;
      int pred;
      switch (filter) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: pred = paeth(a, b, c); break;
        default: return fail(TSDF_ERR_FORMAT, std::string("bad PNG filter: ") + path);
      }
      cur[i] = (unsigned char)(line[1 + i] + pred);
    }
    uint16_t* dst = depth_mm + (size_t)r * width;
    for (unsigned x = 0; x < width; ++x) dst[x] = (uint16_t)((cur[2 * x] << 8) | cur[2 * x + 1]);
    prev.swap(cur);
  }
  return TSDF_OK;
}

// utils.hpp:74-86
int tsdf_read_depth_png(const char* path, int32_t im_height, int32_t im_width, float* depth) {
  if (!depth || im_height <= 0 || im_width <= 0) return fail(TSDF_ERR_ARG, "tsdf_read_depth_png: bad argument");
  std::vector<uint16_t> mm((size_t)im_height * im_width);
  const int rc = tsdf_read_depth_png_u16(path, im_height, im_width, mm.data());
  if (rc != TSDF_OK) return rc;
  tsdf_depth_from_u16(mm.data(), (int64_t)mm.size(), depth);
  return TSDF_OK;
}

}  // extern "C"
