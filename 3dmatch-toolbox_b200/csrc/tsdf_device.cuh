// tsdf_device.cuh -- device helpers shared by the fusion kernels (tsdf_kernels.cu: reference-shaped / validation
// kernels; tsdf_stream_kernels.cu: the production work-list kernel).  Numerical contract: every arithmetic step of
// depth-fusion/demo.cu:30-66 is issued with an explicit round-to-nearest intrinsic, so the result is the
// "--fmad=false" evaluation of the reference whatever the compiler flags.  Include inside no namespace.
#pragma once
#include "tsdf_kernels.h"

#include <type_traits>

#ifndef TSDF_MASK_MINB
#define TSDF_MASK_MINB 4
#endif

namespace tsdf {
namespace {

// ------------------------------------------------------------------------------------------------
// Pixel rounding.  demo.cu:46-49 computes `int p = roundf(v)` and keeps 0 <= p < size.
// roundf(v) in [0, size)  <=>  v > -0.5 and trunc(v (+rz) 0.5) < size: adding 0.5 with
// round-toward-zero never crosses an integer that v + 0.5 does not reach exactly, so truncation
// of the sum is round-half-away-from-zero for v >= 0; v in (-0.5, 0) rounds to -0 -> pixel 0.
// NaN fails the first comparison, +inf the second.  Verified over all 2^32 bit patterns by
// selftest_rounding_kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool round_pixel(float v, float size_f, int& pix) {
  float s = __fadd_rz(v, 0.5f);
  pix = __float2int_rz(s);
  return (v > -0.5f) && (s < size_f);
}

struct Intrinsics {
  float fx, fy, cx, cy;
  float lim_x, lim_y;  // 0.5 - pixel_round_margin(): |u' - round(u')| below this => rounding is certain
};

// demo.cu:43-62 for one voxel, evaluated exactly as the reference does (one rounding per operation).
// Returns dist (in (-1, 1]) when the voxel is to be updated, kNoObservation otherwise.  Kept out of
// line: after the fast classification below only a few percent of the voxels come here.
// BAND_SHORTCUT applies the +-trunc band argument of voxel_classify (2) after the exact pixel has
// been found; with false the function is the plain restatement (validation mode).
constexpr float kNoObservation = 2.0f;
// SANITIZED: `depth` went through sanitize_depth_kernel, where -inf marks a pixel the reference rejects and +inf
// marks a NaN pixel.  The reference does NOT reject NaN (demo.cu:53: both comparisons are false; :57 likewise) and
// ends with fmin(1.0f, NaN) = 1.0f (demo.cu:62), i.e. a dist = 1 observation.
template <bool BAND_SHORTCUT, bool SANITIZED = false>
__device__ __forceinline__ float voxel_observation_exact_inline(float cam_x, float cam_y, float cam_z, float fx, float fy,
                                                                float cx, float cy, const float* __restrict__ depth,
                                                                int im_w, int im_h, float trunc_margin,
                                                                unsigned frame_offset = 0u) {
  // frame_offset: element offset of the frame inside `depth` (a 32-bit sum with the pixel index instead of a 64-bit
  // pointer per frame)
  if (cam_z <= 0.0f) return kNoObservation;                          // demo.cu:43
  const float qx = __fdiv_rn(cam_x, cam_z);
  const float qy = __fdiv_rn(cam_y, cam_z);
  const float u = __fadd_rn(__fmul_rn(fx, qx), cx);                 // demo.cu:46
  const float v = __fadd_rn(__fmul_rn(fy, qy), cy);                 // demo.cu:47
  int px, py;
  const bool in_x = round_pixel(u, (float)im_w, px);
  const bool in_y = round_pixel(v, (float)im_h, py);
  if (!(in_x && in_y)) return kNoObservation;                        // demo.cu:48
  const float d = __ldg(depth + (frame_offset + (unsigned)(py * im_w + px)));   // demo.cu:51
  if (SANITIZED && d == INFINITY) return 1.0f;                       // NaN pixel: demo.cu:53-62 yield dist = 1
  if (d <= 0.0f || d > 6.0f) return kNoObservation;                  // demo.cu:53
  const float e = __fsub_rn(d, cam_z);
  if (BAND_SHORTCUT && trunc_margin > 0.0f) {
    if (e >= trunc_margin) return 1.0f;
    if (e <= -trunc_margin) return kNoObservation;
  }
  // demo.cu:56: (d - z) * sqrtf((1 + qx^2) + qy^2)
  const float ray = __fsqrt_rn(__fadd_rn(__fadd_rn(1.0f, __fmul_rn(qx, qx)), __fmul_rn(qy, qy)));
  const float diff = __fmul_rn(e, ray);
  if (diff <= -trunc_margin) return kNoObservation;                  // demo.cu:57
  return fminf(1.0f, __fdiv_rn(diff, trunc_margin));                 // demo.cu:62
}

// Out-of-line form for the kernels that call it from many places (tsdf_kernels.cu).
template <bool BAND_SHORTCUT, bool SANITIZED = false>
__device__ __noinline__ float voxel_observation_exact(float cam_x, float cam_y, float cam_z, float fx, float fy,
                                                      float cx, float cy, const float* __restrict__ depth,
                                                      int im_w, int im_h, float trunc_margin) {
  return voxel_observation_exact_inline<BAND_SHORTCUT, SANITIZED>(cam_x, cam_y, cam_z, fx, fy, cx, cy, depth, im_w, im_h,
                                                                   trunc_margin);
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ------------------------------------------------------------------------------------------------
// Fast, bit-exact classification of one voxel (branch-free apart from the predicated gather).
//
// (1) Pixel.  u' = cam_x * (fx * rcp(cam_z)) + cx differs from the reference's
//     u = fl(fl(fx * fl(cam_x / cam_z)) + cx) by less than pixel_round_margin() (see there), so
//     whenever u' is further than that from every half-integer, roundf(u) == rint(u') and the two
//     IEEE divisions are not needed.  rint() is taken with the 1.5 * 2^23 trick (no F2I): for
//     |u'| < 2^22 the low mantissa bits of u' + 1.5*2^23 are rint(u'); for larger |u'| the
//     difference (u' + M) - M is not within lim of u' or the integer is far outside the image.
// (2) Distance.  ray = sqrtf(1 + qx^2 + qy^2) >= 1 exactly, hence with e = fl(d - cam_z):
//       e >=  trunc  =>  diff = fl(e * ray) >=  trunc  =>  dist = fmin(1, diff / trunc) = 1
//       e <= -trunc  =>  diff <= -trunc                =>  voxel skipped (demo.cu:57)
//     and only the +-trunc band around the surface needs ray and the division.
// Everything that is not certain is sent to voxel_observation_exact().
// ------------------------------------------------------------------------------------------------
constexpr float kRintMagic = 12582912.0f;  // 1.5 * 2^23
constexpr int kRintMagicBits = 0x4B400000;

// SANITIZED: the depth frame was rewritten by sanitize_depth_kernel (invalid pixels = -inf), so the
// validity test of demo.cu:53 is implied by the band comparisons (e = -inf fails both).
template <bool SANITIZED>
__device__ __forceinline__ void voxel_classify(float cam_x, float cam_y, float cam_z, const Intrinsics& K,
                                               const float* __restrict__ depth, int im_w, int im_h,
                                               float trunc_margin, bool& free_space, bool& need_exact) {
  const float r = rcp_approx(cam_z);
  const float ua = fmaf(cam_x, r * K.fx, K.cx);
  const float va = fmaf(cam_y, r * K.fy, K.cy);
  const float tu = ua + kRintMagic, tv = va + kRintMagic;
  const bool certain = (fabsf(ua - (tu - kRintMagic)) < K.lim_x) & (fabsf(va - (tv - kRintMagic)) < K.lim_y);
  const unsigned ix = (unsigned)(__float_as_int(tu) - kRintMagicBits);
  const unsigned iy = (unsigned)(__float_as_int(tv) - kRintMagicBits);
  const bool pos = cam_z > 0.0f;                                   // demo.cu:43 (NaN: rejected at :48 there)
  const bool inb = (ix < (unsigned)im_w) & (iy < (unsigned)im_h);  // demo.cu:48
  float d = SANITIZED ? -INFINITY : 0.0f;
  if (pos & certain & inb) d = __ldg(depth + (iy * (unsigned)im_w + ix));  // demo.cu:51
  const float e = __fsub_rn(d, cam_z);
  if (SANITIZED) {
    free_space = e >= trunc_margin;
    need_exact = pos & (!certain | (fabsf(e) < trunc_margin));
  } else {
    const bool valid = !((d <= 0.0f) | (d > 6.0f));                // demo.cu:53 (a NaN pixel is NOT rejected there)
    free_space = valid & (e >= trunc_margin);
    need_exact = pos & (!certain | (valid & !(e >= trunc_margin) & !(e <= -trunc_margin)));  // band, or NaN
  }
}

// voxel_classify<true> + the pending-count / exact-bit bookkeeping of integrate_pending_kernel as one PTX
// block: identical arithmetic, but the comparison results stay in predicate registers and drive a
// predicated gather, a predicated add (pend += pend_inc when the voxel is free space) and a predicated
// or (exact |= exact_bit when it needs voxel_observation_exact) -- nvcc's C++ lowering of the same logic
// spends ~8 more ALU-pipe instructions per voxel on selects (measured: 43.8 -> 36.1 SASS instructions per
// voxel, +6.5 % throughput).  pend_inc / exact_bit are compile-time constants after unrolling.
__device__ __forceinline__ void voxel_classify_ptx(float cam_x, float cam_y, float cam_z, const Intrinsics& K,
                                                   const float* __restrict__ depth, int im_w, int im_h,
                                                   float trunc_margin, unsigned pend_inc, unsigned exact_bit,
                                                   unsigned& pend, unsigned& exact) {
  asm("{\n\t"
      ".reg .f32 r, ua, va, tu, tv, du, dv, d, e;\n\t"
      ".reg .s32 ix, iy, idx;\n\t"
      ".reg .u64 addr;\n\t"
      ".reg .pred pc, pi, pp, pg, pf, pb, pe;\n\t"
      "rcp.approx.ftz.f32 r, %4;\n\t"
      "mul.f32 ua, r, %5;\n\t"
      "fma.rn.f32 ua, %2, ua, %7;\n\t"
      "mul.f32 va, r, %6;\n\t"
      "fma.rn.f32 va, %3, va, %8;\n\t"
      "add.f32 tu, ua, 0f4B400000;\n\t"
      "add.f32 tv, va, 0f4B400000;\n\t"
      "sub.f32 du, tu, 0f4B400000;\n\t"
      "sub.f32 dv, tv, 0f4B400000;\n\t"
      "sub.f32 du, ua, du;\n\t"
      "sub.f32 dv, va, dv;\n\t"
      "abs.f32 du, du;\n\t"
      "abs.f32 dv, dv;\n\t"
      "setp.lt.f32 pc, du, %9;\n\t"
      "setp.lt.and.f32 pc, dv, %10, pc;\n\t"
      "mov.b32 ix, tu;\n\t"
      "mov.b32 iy, tv;\n\t"
      "sub.s32 ix, ix, 0x4B400000;\n\t"
      "sub.s32 iy, iy, 0x4B400000;\n\t"
      "setp.lt.u32 pi, ix, %11;\n\t"
      "setp.lt.and.u32 pi, iy, %12, pi;\n\t"
      "setp.gt.f32 pp, %4, 0f00000000;\n\t"
      "and.pred pg, pc, pi;\n\t"
      "and.pred pg, pg, pp;\n\t"
      "mad.lo.s32 idx, iy, %11, ix;\n\t"
      "mad.wide.u32 addr, idx, 4, %13;\n\t"
      "mov.f32 d, 0fFF800000;\n\t"
      "@pg ld.global.nc.f32 d, [addr];\n\t"
      "sub.f32 e, d, %4;\n\t"
      "setp.ge.f32 pf, e, %14;\n\t"
      "@pf add.u32 %0, %0, %15;\n\t"
      "abs.f32 e, e;\n\t"
      "setp.lt.f32 pb, e, %14;\n\t"
      "not.pred pe, pc;\n\t"
      "or.pred pe, pe, pb;\n\t"
      "and.pred pe, pe, pp;\n\t"
      "@pe or.b32 %1, %1, %16;\n\t"
      "}"
      : "+r"(pend), "+r"(exact)
      : "f"(cam_x), "f"(cam_y), "f"(cam_z), "f"(K.fx), "f"(K.fy), "f"(K.cx), "f"(K.cy), "f"(K.lim_x), "f"(K.lim_y),
        "r"(im_w), "r"(im_h), "l"(depth), "f"(trunc_margin), "r"(pend_inc), "r"(exact_bit));
}

// The same block for "interior" tiles: the frame mask stage has shown that every voxel of the warp's tile is
// in front of the camera and projects at least 1.5 px inside the image, so the bounds test (demo.cu:48) and
// the cam_z > 0 test (demo.cu:43) are known to pass and are dropped (5 fewer ALU-pipe instructions).
__device__ __forceinline__ void voxel_classify_ptx_interior(float cam_x, float cam_y, float cam_z, const Intrinsics& K,
                                                            const float* __restrict__ depth, int im_w,
                                                            float trunc_margin, unsigned pend_inc, unsigned exact_bit,
                                                            unsigned& pend, unsigned& exact) {
  asm("{\n\t"
      ".reg .f32 r, ua, va, tu, tv, du, dv, d, e;\n\t"
      ".reg .s32 ix, iy, idx;\n\t"
      ".reg .u64 addr;\n\t"
      ".reg .pred pc, pf, pb, pe;\n\t"
      "rcp.approx.ftz.f32 r, %4;\n\t"
      "mul.f32 ua, r, %5;\n\t"
      "fma.rn.f32 ua, %2, ua, %7;\n\t"
      "mul.f32 va, r, %6;\n\t"
      "fma.rn.f32 va, %3, va, %8;\n\t"
      "add.f32 tu, ua, 0f4B400000;\n\t"
      "add.f32 tv, va, 0f4B400000;\n\t"
      "sub.f32 du, tu, 0f4B400000;\n\t"
      "sub.f32 dv, tv, 0f4B400000;\n\t"
      "sub.f32 du, ua, du;\n\t"
      "sub.f32 dv, va, dv;\n\t"
      "abs.f32 du, du;\n\t"
      "abs.f32 dv, dv;\n\t"
      "setp.lt.f32 pc, du, %9;\n\t"
      "setp.lt.and.f32 pc, dv, %10, pc;\n\t"
      "mov.b32 ix, tu;\n\t"
      "mov.b32 iy, tv;\n\t"
      "sub.s32 ix, ix, 0x4B400000;\n\t"
      "sub.s32 iy, iy, 0x4B400000;\n\t"
      "mad.lo.s32 idx, iy, %11, ix;\n\t"
      "mad.wide.u32 addr, idx, 4, %12;\n\t"
      "mov.f32 d, 0fFF800000;\n\t"
      "@pc ld.global.nc.f32 d, [addr];\n\t"
      "sub.f32 e, d, %4;\n\t"
      "setp.ge.f32 pf, e, %13;\n\t"
      "@pf add.u32 %0, %0, %14;\n\t"
      "abs.f32 e, e;\n\t"
      "setp.lt.f32 pb, e, %13;\n\t"
      "not.pred pe, pc;\n\t"
      "or.pred pe, pe, pb;\n\t"
      "@pe or.b32 %1, %1, %15;\n\t"
      "}"
      : "+r"(pend), "+r"(exact)
      : "f"(cam_x), "f"(cam_y), "f"(cam_z), "f"(K.fx), "f"(K.fy), "f"(K.cx), "f"(K.cy), "f"(K.lim_x), "f"(K.lim_y),
        "r"(im_w), "l"(depth), "f"(trunc_margin), "r"(pend_inc), "r"(exact_bit));
}

// ------------------------------------------------------------------------------------------------
// Packed classification (sm_100a FFMA2 / FADD2): the same classification as voxel_classify_ptx for TWO
// x-adjacent voxels at once.  Blackwell issues fp32 add/mul/fma on register pairs as one instruction
// (add/mul/fma.rn.f32x2), with scalar operands broadcast for free, so the projection costs half the
// issue slots -- the kernel is issue bound, not FMA-pipe bound.  Arithmetic per lane is identical to the
// scalar block where exactness matters (cam_z and e = d - cam_z are single IEEE adds); the pixel
// coordinate only has to stay inside pixel_round_margin(), which any contraction ptxas applies to the
// f32x2 multiply-adds can only tighten.  Two further instruction savings:
//   * the gather index is formed in floating point: fma(rint(v'), W, u' + M) = M + iy*W + ix exactly
//     (M = 1.5 * 2^23, the sum stays below 2^24), so its bit pattern is 0x4B400000 + iy*W + ix and one
//     mad.wide with a pre-biased base pointer gives the address;
//   * bounds (non-interior tiles) are tested on the rounded float coordinate: |rint(u') - (W-1)/2| <=
//     (W-1)/2, exact for integers below 2^22 and false for NaN / huge values.
// sx2/sy2/sz2 hold (M0*t0 + M4*t1) of the two voxels, zx/zy/zz the plane's M8*t2 products.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

template <bool INTERIOR>
__device__ __forceinline__ void voxel_classify_pair(f32x2_t sx2, f32x2_t sy2, f32x2_t sz2, float zx, float zy, float zz,
                                                    const Intrinsics& K, float im_w_f, float half_w, float half_h,
                                                    const void* depth_biased, float trunc_margin, unsigned pend_inc0,
                                                    unsigned pend_inc1, unsigned exact_bit0, unsigned exact_bit1,
                                                    unsigned& pend, unsigned& exact) {
  if constexpr (INTERIOR) {
    asm("{\n\t"
        ".reg .b64 bz, cxx, cyy, czz, rr, kk, ua, va, tu, tv, ru, rv, du, dv, fi, dd, ee, bm, bk;\n\t"
        ".reg .f32 z0, z1, r0, r1, a0, a1, b0, b1, i0, i1, d0, d1, e0, e1, mg;\n\t"
        ".reg .b32 n0, n1;\n\t"
        ".reg .u64 ad0, ad1;\n\t"
        ".reg .pred pc0, pc1, pf0, pf1, pb0, pb1;\n\t"
        "mov.b64 bz, {%5, %5};\n\t"
        "add.rn.f32x2 cxx, %2, bz;\n\t"
        "mov.b64 bz, {%6, %6};\n\t"
        "add.rn.f32x2 cyy, %3, bz;\n\t"
        "mov.b64 bz, {%7, %7};\n\t"
        "add.rn.f32x2 czz, %4, bz;\n\t"
        "mov.b64 {z0, z1}, czz;\n\t"
        "rcp.approx.ftz.f32 r0, z0;\n\t"
        "rcp.approx.ftz.f32 r1, z1;\n\t"
        "mov.b64 rr, {r0, r1};\n\t"
        "mov.b64 bk, {%8, %8};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%10, %10};\n\t"
        "fma.rn.f32x2 ua, cxx, kk, bk;\n\t"
        "mov.b64 bk, {%9, %9};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%11, %11};\n\t"
        "fma.rn.f32x2 va, cyy, kk, bk;\n\t"
        "mov.f32 mg, 0f4B400000;\n\t"
        "mov.b64 bm, {mg, mg};\n\t"
        "add.rn.f32x2 tu, ua, bm;\n\t"
        "add.rn.f32x2 tv, va, bm;\n\t"
        "sub.rn.f32x2 ru, tu, bm;\n\t"
        "sub.rn.f32x2 rv, tv, bm;\n\t"
        "sub.rn.f32x2 du, ua, ru;\n\t"
        "sub.rn.f32x2 dv, va, rv;\n\t"
        "mov.b64 bk, {%14, %14};\n\t"
        "fma.rn.f32x2 fi, rv, bk, tu;\n\t"
        "mov.b64 {a0, a1}, du;\n\t"
        "mov.b64 {b0, b1}, dv;\n\t"
        "abs.f32 a0, a0;\n\t"
        "abs.f32 a1, a1;\n\t"
        "abs.f32 b0, b0;\n\t"
        "abs.f32 b1, b1;\n\t"
        "setp.lt.f32 pc0, a0, %12;\n\t"
        "setp.lt.f32 pc1, a1, %12;\n\t"
        "setp.lt.and.f32 pc0, b0, %13, pc0;\n\t"
        "setp.lt.and.f32 pc1, b1, %13, pc1;\n\t"
        "mov.b64 {i0, i1}, fi;\n\t"
        "mov.b32 n0, i0;\n\t"
        "mov.b32 n1, i1;\n\t"
        "mad.wide.u32 ad0, n0, 4, %15;\n\t"
        "mad.wide.u32 ad1, n1, 4, %15;\n\t"
        "ld.global.nc.f32 d0, [ad0];\n\t"
        "ld.global.nc.f32 d1, [ad1];\n\t"
        "mov.b64 dd, {d0, d1};\n\t"
        "sub.rn.f32x2 ee, dd, czz;\n\t"
        "mov.b64 {e0, e1}, ee;\n\t"
        "setp.ge.and.f32 pf0, e0, %16, pc0;\n\t"   // the gather is unconditional here: an uncertain pixel
        "setp.ge.and.f32 pf1, e1, %16, pc1;\n\t"   // must not count (the exact path will)
        "@pf0 add.u32 %0, %0, %17;\n\t"
        "@pf1 add.u32 %0, %0, %18;\n\t"
        "abs.f32 e0, e0;\n\t"
        "abs.f32 e1, e1;\n\t"
        "setp.lt.f32 pb0, e0, %16;\n\t"
        "setp.lt.f32 pb1, e1, %16;\n\t"
        "not.pred pc0, pc0;\n\t"
        "not.pred pc1, pc1;\n\t"
        "or.pred pb0, pb0, pc0;\n\t"
        "or.pred pb1, pb1, pc1;\n\t"
        "@pb0 or.b32 %1, %1, %19;\n\t"
        "@pb1 or.b32 %1, %1, %20;\n\t"
        "}"
        : "+r"(pend), "+r"(exact)
        : "l"(sx2), "l"(sy2), "l"(sz2), "f"(zx), "f"(zy), "f"(zz), "f"(K.fx), "f"(K.fy), "f"(K.cx), "f"(K.cy),
          "f"(K.lim_x), "f"(K.lim_y), "f"(im_w_f), "l"(depth_biased), "f"(trunc_margin), "r"(pend_inc0),
          "r"(pend_inc1), "r"(exact_bit0), "r"(exact_bit1));
  } else {
    asm("{\n\t"
        ".reg .b64 bz, cxx, cyy, czz, rr, kk, ua, va, tu, tv, ru, rv, du, dv, fi, dd, ee, bm, bk, qu, qv;\n\t"
        ".reg .f32 z0, z1, r0, r1, a0, a1, b0, b1, i0, i1, d0, d1, e0, e1, mg, g0, g1, h0, h1;\n\t"
        ".reg .b32 n0, n1;\n\t"
        ".reg .u64 ad0, ad1;\n\t"
        ".reg .pred pc0, pc1, pg0, pg1, pp0, pp1, pf0, pf1, pb0, pb1;\n\t"
        "mov.b64 bz, {%5, %5};\n\t"
        "add.rn.f32x2 cxx, %2, bz;\n\t"
        "mov.b64 bz, {%6, %6};\n\t"
        "add.rn.f32x2 cyy, %3, bz;\n\t"
        "mov.b64 bz, {%7, %7};\n\t"
        "add.rn.f32x2 czz, %4, bz;\n\t"
        "mov.b64 {z0, z1}, czz;\n\t"
        "rcp.approx.ftz.f32 r0, z0;\n\t"
        "rcp.approx.ftz.f32 r1, z1;\n\t"
        "mov.b64 rr, {r0, r1};\n\t"
        "mov.b64 bk, {%8, %8};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%10, %10};\n\t"
        "fma.rn.f32x2 ua, cxx, kk, bk;\n\t"
        "mov.b64 bk, {%9, %9};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%11, %11};\n\t"
        "fma.rn.f32x2 va, cyy, kk, bk;\n\t"
        "mov.f32 mg, 0f4B400000;\n\t"
        "mov.b64 bm, {mg, mg};\n\t"
        "add.rn.f32x2 tu, ua, bm;\n\t"
        "add.rn.f32x2 tv, va, bm;\n\t"
        "sub.rn.f32x2 ru, tu, bm;\n\t"
        "sub.rn.f32x2 rv, tv, bm;\n\t"
        "sub.rn.f32x2 du, ua, ru;\n\t"
        "sub.rn.f32x2 dv, va, rv;\n\t"
        "mov.b64 bk, {%14, %14};\n\t"
        "fma.rn.f32x2 fi, rv, bk, tu;\n\t"
        "mov.b64 bk, {%21, %21};\n\t"
        "sub.rn.f32x2 qu, ru, bk;\n\t"
        "mov.b64 bk, {%22, %22};\n\t"
        "sub.rn.f32x2 qv, rv, bk;\n\t"
        "mov.b64 {a0, a1}, du;\n\t"
        "mov.b64 {b0, b1}, dv;\n\t"
        "mov.b64 {g0, g1}, qu;\n\t"
        "mov.b64 {h0, h1}, qv;\n\t"
        "abs.f32 a0, a0;\n\t"
        "abs.f32 a1, a1;\n\t"
        "abs.f32 b0, b0;\n\t"
        "abs.f32 b1, b1;\n\t"
        "abs.f32 g0, g0;\n\t"
        "abs.f32 g1, g1;\n\t"
        "abs.f32 h0, h0;\n\t"
        "abs.f32 h1, h1;\n\t"
        "setp.lt.f32 pc0, a0, %12;\n\t"
        "setp.lt.f32 pc1, a1, %12;\n\t"
        "setp.lt.and.f32 pc0, b0, %13, pc0;\n\t"
        "setp.lt.and.f32 pc1, b1, %13, pc1;\n\t"
        "setp.gt.f32 pp0, z0, 0f00000000;\n\t"
        "setp.gt.f32 pp1, z1, 0f00000000;\n\t"
        "setp.le.and.f32 pg0, g0, %21, pc0;\n\t"
        "setp.le.and.f32 pg1, g1, %21, pc1;\n\t"
        "setp.le.and.f32 pg0, h0, %22, pg0;\n\t"
        "setp.le.and.f32 pg1, h1, %22, pg1;\n\t"
        "and.pred pg0, pg0, pp0;\n\t"
        "and.pred pg1, pg1, pp1;\n\t"
        "mov.b64 {i0, i1}, fi;\n\t"
        "mov.b32 n0, i0;\n\t"
        "mov.b32 n1, i1;\n\t"
        "mad.wide.u32 ad0, n0, 4, %15;\n\t"
        "mad.wide.u32 ad1, n1, 4, %15;\n\t"
        "mov.f32 d0, 0fFF800000;\n\t"
        "mov.f32 d1, 0fFF800000;\n\t"
        "@pg0 ld.global.nc.f32 d0, [ad0];\n\t"
        "@pg1 ld.global.nc.f32 d1, [ad1];\n\t"
        "mov.b64 dd, {d0, d1};\n\t"
        "sub.rn.f32x2 ee, dd, czz;\n\t"
        "mov.b64 {e0, e1}, ee;\n\t"
        "setp.ge.f32 pf0, e0, %16;\n\t"
        "setp.ge.f32 pf1, e1, %16;\n\t"
        "@pf0 add.u32 %0, %0, %17;\n\t"
        "@pf1 add.u32 %0, %0, %18;\n\t"
        "abs.f32 e0, e0;\n\t"
        "abs.f32 e1, e1;\n\t"
        "setp.lt.f32 pb0, e0, %16;\n\t"
        "setp.lt.f32 pb1, e1, %16;\n\t"
        "not.pred pc0, pc0;\n\t"
        "not.pred pc1, pc1;\n\t"
        "or.pred pb0, pb0, pc0;\n\t"
        "or.pred pb1, pb1, pc1;\n\t"
        "and.pred pb0, pb0, pp0;\n\t"
        "and.pred pb1, pb1, pp1;\n\t"
        "@pb0 or.b32 %1, %1, %19;\n\t"
        "@pb1 or.b32 %1, %1, %20;\n\t"
        "}"
        : "+r"(pend), "+r"(exact)
        : "l"(sx2), "l"(sy2), "l"(sz2), "f"(zx), "f"(zy), "f"(zz), "f"(K.fx), "f"(K.fy), "f"(K.cx), "f"(K.cy),
          "f"(K.lim_x), "f"(K.lim_y), "f"(im_w_f), "l"(depth_biased), "f"(trunc_margin), "r"(pend_inc0),
          "r"(pend_inc1), "r"(exact_bit0), "r"(exact_bit1), "f"(half_w), "f"(half_h));
  }
}

#ifdef TSDF_VALIDATION
// Diagnostic twin of voxel_classify_pair (validation library only): the depth gathers carry an L2 evict_last cache
// hint (createpolicy.fractional.L2::evict_last), the per-access form of an L2-persisting access-policy window, for
// the A/B measurement in profiles/ (BASELINE north_star (2)).  Identical arithmetic.
template <bool INTERIOR>
__device__ __forceinline__ void voxel_classify_pair_hint(f32x2_t sx2, f32x2_t sy2, f32x2_t sz2, float zx, float zy, float zz,
                                                    const Intrinsics& K, float im_w_f, float half_w, float half_h,
                                                    const void* depth_biased, float trunc_margin, unsigned pend_inc0,
                                                    unsigned pend_inc1, unsigned exact_bit0, unsigned exact_bit1,
                                                    unsigned& pend, unsigned& exact, unsigned long long policy) {
  if constexpr (INTERIOR) {
    asm("{\n\t"
        ".reg .b64 bz, cxx, cyy, czz, rr, kk, ua, va, tu, tv, ru, rv, du, dv, fi, dd, ee, bm, bk;\n\t"
        ".reg .f32 z0, z1, r0, r1, a0, a1, b0, b1, i0, i1, d0, d1, e0, e1, mg;\n\t"
        ".reg .b32 n0, n1;\n\t"
        ".reg .u64 ad0, ad1;\n\t"
        ".reg .pred pc0, pc1, pf0, pf1, pb0, pb1;\n\t"
        "mov.b64 bz, {%5, %5};\n\t"
        "add.rn.f32x2 cxx, %2, bz;\n\t"
        "mov.b64 bz, {%6, %6};\n\t"
        "add.rn.f32x2 cyy, %3, bz;\n\t"
        "mov.b64 bz, {%7, %7};\n\t"
        "add.rn.f32x2 czz, %4, bz;\n\t"
        "mov.b64 {z0, z1}, czz;\n\t"
        "rcp.approx.ftz.f32 r0, z0;\n\t"
        "rcp.approx.ftz.f32 r1, z1;\n\t"
        "mov.b64 rr, {r0, r1};\n\t"
        "mov.b64 bk, {%8, %8};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%10, %10};\n\t"
        "fma.rn.f32x2 ua, cxx, kk, bk;\n\t"
        "mov.b64 bk, {%9, %9};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%11, %11};\n\t"
        "fma.rn.f32x2 va, cyy, kk, bk;\n\t"
        "mov.f32 mg, 0f4B400000;\n\t"
        "mov.b64 bm, {mg, mg};\n\t"
        "add.rn.f32x2 tu, ua, bm;\n\t"
        "add.rn.f32x2 tv, va, bm;\n\t"
        "sub.rn.f32x2 ru, tu, bm;\n\t"
        "sub.rn.f32x2 rv, tv, bm;\n\t"
        "sub.rn.f32x2 du, ua, ru;\n\t"
        "sub.rn.f32x2 dv, va, rv;\n\t"
        "mov.b64 bk, {%14, %14};\n\t"
        "fma.rn.f32x2 fi, rv, bk, tu;\n\t"
        "mov.b64 {a0, a1}, du;\n\t"
        "mov.b64 {b0, b1}, dv;\n\t"
        "abs.f32 a0, a0;\n\t"
        "abs.f32 a1, a1;\n\t"
        "abs.f32 b0, b0;\n\t"
        "abs.f32 b1, b1;\n\t"
        "setp.lt.f32 pc0, a0, %12;\n\t"
        "setp.lt.f32 pc1, a1, %12;\n\t"
        "setp.lt.and.f32 pc0, b0, %13, pc0;\n\t"
        "setp.lt.and.f32 pc1, b1, %13, pc1;\n\t"
        "mov.b64 {i0, i1}, fi;\n\t"
        "mov.b32 n0, i0;\n\t"
        "mov.b32 n1, i1;\n\t"
        "mad.wide.u32 ad0, n0, 4, %15;\n\t"
        "mad.wide.u32 ad1, n1, 4, %15;\n\t"
        "ld.global.nc.L2::cache_hint.f32 d0, [ad0], %21;\n\t"
        "ld.global.nc.L2::cache_hint.f32 d1, [ad1], %21;\n\t"
        "mov.b64 dd, {d0, d1};\n\t"
        "sub.rn.f32x2 ee, dd, czz;\n\t"
        "mov.b64 {e0, e1}, ee;\n\t"
        "setp.ge.and.f32 pf0, e0, %16, pc0;\n\t"   // the gather is unconditional here: an uncertain pixel
        "setp.ge.and.f32 pf1, e1, %16, pc1;\n\t"   // must not count (the exact path will)
        "@pf0 add.u32 %0, %0, %17;\n\t"
        "@pf1 add.u32 %0, %0, %18;\n\t"
        "abs.f32 e0, e0;\n\t"
        "abs.f32 e1, e1;\n\t"
        "setp.lt.f32 pb0, e0, %16;\n\t"
        "setp.lt.f32 pb1, e1, %16;\n\t"
        "not.pred pc0, pc0;\n\t"
        "not.pred pc1, pc1;\n\t"
        "or.pred pb0, pb0, pc0;\n\t"
        "or.pred pb1, pb1, pc1;\n\t"
        "@pb0 or.b32 %1, %1, %19;\n\t"
        "@pb1 or.b32 %1, %1, %20;\n\t"
        "}"
        : "+r"(pend), "+r"(exact)
        : "l"(sx2), "l"(sy2), "l"(sz2), "f"(zx), "f"(zy), "f"(zz), "f"(K.fx), "f"(K.fy), "f"(K.cx), "f"(K.cy),
          "f"(K.lim_x), "f"(K.lim_y), "f"(im_w_f), "l"(depth_biased), "f"(trunc_margin), "r"(pend_inc0),
          "r"(pend_inc1), "r"(exact_bit0), "r"(exact_bit1), "l"(policy));
  } else {
    asm("{\n\t"
        ".reg .b64 bz, cxx, cyy, czz, rr, kk, ua, va, tu, tv, ru, rv, du, dv, fi, dd, ee, bm, bk, qu, qv;\n\t"
        ".reg .f32 z0, z1, r0, r1, a0, a1, b0, b1, i0, i1, d0, d1, e0, e1, mg, g0, g1, h0, h1;\n\t"
        ".reg .b32 n0, n1;\n\t"
        ".reg .u64 ad0, ad1;\n\t"
        ".reg .pred pc0, pc1, pg0, pg1, pp0, pp1, pf0, pf1, pb0, pb1;\n\t"
        "mov.b64 bz, {%5, %5};\n\t"
        "add.rn.f32x2 cxx, %2, bz;\n\t"
        "mov.b64 bz, {%6, %6};\n\t"
        "add.rn.f32x2 cyy, %3, bz;\n\t"
        "mov.b64 bz, {%7, %7};\n\t"
        "add.rn.f32x2 czz, %4, bz;\n\t"
        "mov.b64 {z0, z1}, czz;\n\t"
        "rcp.approx.ftz.f32 r0, z0;\n\t"
        "rcp.approx.ftz.f32 r1, z1;\n\t"
        "mov.b64 rr, {r0, r1};\n\t"
        "mov.b64 bk, {%8, %8};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%10, %10};\n\t"
        "fma.rn.f32x2 ua, cxx, kk, bk;\n\t"
        "mov.b64 bk, {%9, %9};\n\t"
        "mul.rn.f32x2 kk, rr, bk;\n\t"
        "mov.b64 bk, {%11, %11};\n\t"
        "fma.rn.f32x2 va, cyy, kk, bk;\n\t"
        "mov.f32 mg, 0f4B400000;\n\t"
        "mov.b64 bm, {mg, mg};\n\t"
        "add.rn.f32x2 tu, ua, bm;\n\t"
        "add.rn.f32x2 tv, va, bm;\n\t"
        "sub.rn.f32x2 ru, tu, bm;\n\t"
        "sub.rn.f32x2 rv, tv, bm;\n\t"
        "sub.rn.f32x2 du, ua, ru;\n\t"
        "sub.rn.f32x2 dv, va, rv;\n\t"
        "mov.b64 bk, {%14, %14};\n\t"
        "fma.rn.f32x2 fi, rv, bk, tu;\n\t"
        "mov.b64 bk, {%21, %21};\n\t"
        "sub.rn.f32x2 qu, ru, bk;\n\t"
        "mov.b64 bk, {%22, %22};\n\t"
        "sub.rn.f32x2 qv, rv, bk;\n\t"
        "mov.b64 {a0, a1}, du;\n\t"
        "mov.b64 {b0, b1}, dv;\n\t"
        "mov.b64 {g0, g1}, qu;\n\t"
        "mov.b64 {h0, h1}, qv;\n\t"
        "abs.f32 a0, a0;\n\t"
        "abs.f32 a1, a1;\n\t"
        "abs.f32 b0, b0;\n\t"
        "abs.f32 b1, b1;\n\t"
        "abs.f32 g0, g0;\n\t"
        "abs.f32 g1, g1;\n\t"
        "abs.f32 h0, h0;\n\t"
        "abs.f32 h1, h1;\n\t"
        "setp.lt.f32 pc0, a0, %12;\n\t"
        "setp.lt.f32 pc1, a1, %12;\n\t"
        "setp.lt.and.f32 pc0, b0, %13, pc0;\n\t"
        "setp.lt.and.f32 pc1, b1, %13, pc1;\n\t"
        "setp.gt.f32 pp0, z0, 0f00000000;\n\t"
        "setp.gt.f32 pp1, z1, 0f00000000;\n\t"
        "setp.le.and.f32 pg0, g0, %21, pc0;\n\t"
        "setp.le.and.f32 pg1, g1, %21, pc1;\n\t"
        "setp.le.and.f32 pg0, h0, %22, pg0;\n\t"
        "setp.le.and.f32 pg1, h1, %22, pg1;\n\t"
        "and.pred pg0, pg0, pp0;\n\t"
        "and.pred pg1, pg1, pp1;\n\t"
        "mov.b64 {i0, i1}, fi;\n\t"
        "mov.b32 n0, i0;\n\t"
        "mov.b32 n1, i1;\n\t"
        "mad.wide.u32 ad0, n0, 4, %15;\n\t"
        "mad.wide.u32 ad1, n1, 4, %15;\n\t"
        "mov.f32 d0, 0fFF800000;\n\t"
        "mov.f32 d1, 0fFF800000;\n\t"
        "@pg0 ld.global.nc.L2::cache_hint.f32 d0, [ad0], %23;\n\t"
        "@pg1 ld.global.nc.L2::cache_hint.f32 d1, [ad1], %23;\n\t"
        "mov.b64 dd, {d0, d1};\n\t"
        "sub.rn.f32x2 ee, dd, czz;\n\t"
        "mov.b64 {e0, e1}, ee;\n\t"
        "setp.ge.f32 pf0, e0, %16;\n\t"
        "setp.ge.f32 pf1, e1, %16;\n\t"
        "@pf0 add.u32 %0, %0, %17;\n\t"
        "@pf1 add.u32 %0, %0, %18;\n\t"
        "abs.f32 e0, e0;\n\t"
        "abs.f32 e1, e1;\n\t"
        "setp.lt.f32 pb0, e0, %16;\n\t"
        "setp.lt.f32 pb1, e1, %16;\n\t"
        "not.pred pc0, pc0;\n\t"
        "not.pred pc1, pc1;\n\t"
        "or.pred pb0, pb0, pc0;\n\t"
        "or.pred pb1, pb1, pc1;\n\t"
        "and.pred pb0, pb0, pp0;\n\t"
        "and.pred pb1, pb1, pp1;\n\t"
        "@pb0 or.b32 %1, %1, %19;\n\t"
        "@pb1 or.b32 %1, %1, %20;\n\t"
        "}"
        : "+r"(pend), "+r"(exact)
        : "l"(sx2), "l"(sy2), "l"(sz2), "f"(zx), "f"(zy), "f"(zz), "f"(K.fx), "f"(K.fy), "f"(K.cx), "f"(K.cy),
          "f"(K.lim_x), "f"(K.lim_y), "f"(im_w_f), "l"(depth_biased), "f"(trunc_margin), "r"(pend_inc0),
          "r"(pend_inc1), "r"(exact_bit0), "r"(exact_bit1), "f"(half_w), "f"(half_h), "l"(policy));
  }
}

#endif  // TSDF_VALIDATION

// demo.cu:63-66.  (t*w + 1)/(w + 1) with t == 1 is exactly 1 (numerator and denominator are
// the same rounded sum), so the division is skipped for untouched free-space voxels.
__device__ __forceinline__ void voxel_update(float& t, float& w, float dist) {
  const float w_new = __fadd_rn(w, 1.0f);
  if (!(dist == 1.0f && t == 1.0f)) t = __fdiv_rn(__fadd_rn(__fmul_rn(t, w), dist), w_new);
  w = w_new;
}

template <int VEC>
struct Row;
template <>
struct Row<4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <>
struct Row<1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldcs(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { __stcs(p, v[0]); }
};

// Conservative frustum test on camera-frame points (any rounding; the margins absorb it).
// A set of points whose convex hull is to be culled must ALL satisfy one of the five
// half-space conditions; each condition is linear in position, so testing the hull's corners
// covers every voxel inside it.
struct CullAccum {
  bool behind, left, right, top, bottom;
  float cz_min, cz_max, u_min, u_max, v_min, v_max;  // for the occlusion test (valid when cz_min > 0)
  float d_far;  // set by occluded(): the largest depth the hull can project to (+inf when the pyramid was not consulted)
  __device__ __forceinline__ CullAccum()
      : behind(true), left(true), right(true), top(true), bottom(true), cz_min(INFINITY), cz_max(-INFINITY),
        u_min(INFINITY), u_max(-INFINITY), v_min(INFINITY), v_max(-INFINITY), d_far(INFINITY) {}
  // PT: any struct with the cull fields of IntegrateParams (cull_z/gx/gy, im_w/h, mip_w/off, trunc_margin)
  template <class PT>
  __device__ __forceinline__ void add(float cx, float cy, float cz, const PT& P, const Intrinsics& K) {
    cz_min = fminf(cz_min, cz);
    cz_max = fmaxf(cz_max, cz);
    const float inv = __fdividef(1.0f, cz);
    const float u = fmaf(K.fx * cx, inv, K.cx), v = fmaf(K.fy * cy, inv, K.cy);
    u_min = fminf(u_min, u); u_max = fmaxf(u_max, u);
    v_min = fminf(v_min, v); v_max = fmaxf(v_max, v);
    // cam_z <= 0 for every voxel (demo.cu:43)
    behind = behind && (cz < -P.cull_z);
    // u = fx*cx/cz + cx0 < -2.5 (left) or > im_w + 1.5 (right) for every voxel with cam_z > 0
    left   = left   && (fmaf(K.fx, cx, (K.cx + 2.5f) * cz) < -P.cull_gx);
    right  = right  && (fmaf(K.fx, cx, (K.cx - (float)P.im_w - 1.5f) * cz) > P.cull_gx);
    top    = top    && (fmaf(K.fy, cy, (K.cy + 2.5f) * cz) < -P.cull_gy);
    bottom = bottom && (fmaf(K.fy, cy, (K.cy - (float)P.im_h - 1.5f) * cz) > P.cull_gy);
  }
  __device__ __forceinline__ bool culled() const { return behind || left || right || top || bottom; }

  // Hierarchical-Z test: true when every voxel of the hull lies at least trunc_margin BEHIND every valid
  // depth pixel it can project to, i.e. the reference would skip all of them at demo.cu:57 (or :53).
  //  * the hull is in front of the camera (cz_min > margin), so its projection is inside the bounding box of
  //    the projected corners; the box is widened by 1.5 px for rounding to the nearest pixel and the
  //    approximate divisions;
  //  * `mips` holds, per frame, the maxima of the sanitized depth (invalid = -inf) over aligned blocks of
  //    8..128 px; the finest level that covers the box with <= 12 blocks is read at every block it touches;
  //  * e = fl(d - cam_z) <= d_max - (cz_min - cull_z) <= -trunc * (1 + 1e-4)  =>  diff = fl(e * ray) <= -trunc
  //    because ray >= 1.
  // True when every voxel of the hull is in front of the camera and projects >= 1.5 px inside the image.
  template <class PT>
  __device__ __forceinline__ bool interior(const PT& P) const {
    return (cz_min > P.cull_z + 1e-3f) && (u_min >= 1.5f) && (u_max <= (float)P.im_w - 2.5f) && (v_min >= 1.5f) &&
           (v_max <= (float)P.im_h - 2.5f);
  }

  // Planning heuristic (never used for culling): after occluded() returned false, does the hull reach the +-trunc band
  // of the FARTHEST surface it can project to?  Those tiles pay for exact evaluations on top of the classification.
  template <class PT>
  __device__ __forceinline__ bool near_far_surface(const PT& P) const {
    return d_far < 1e30f && cz_max + P.trunc_margin >= d_far;
  }

  template <class PT>
  __device__ __forceinline__ bool occluded(const PT& P, const float* __restrict__ mips) {
    if (!(cz_min > P.cull_z + 1e-3f)) return false;
    const float fw = (float)P.im_w, fh = (float)P.im_h;
    if (!(u_min > -1e6f && u_max < 1e6f && v_min > -1e6f && v_max < 1e6f)) return false;
    const int x0 = max(0, (int)floorf(u_min - 1.5f)), x1 = min(P.im_w - 1, (int)ceilf(u_max + 1.5f));
    const int y0 = max(0, (int)floorf(v_min - 1.5f)), y1 = min(P.im_h - 1, (int)ceilf(v_max + 1.5f));
    (void)fw; (void)fh;
    if (x1 < x0 || y1 < y0) return true;  // projects entirely outside the image (demo.cu:48)
    // the FINEST level whose block grid over the box has at most kMaxBlocks blocks: the maximum is then taken over little
    // more than the box itself (a warp tile of the 1024^3 grid projects to ~30 x 6 px: 5 x 2 blocks of 8 px; round 1 / 2
    // read 2 x 2 blocks of the level covering the larger extent, 128 x 128 px here, whose maximum over a sloped wall
    // lies centimetres behind the depth under the tile -- 11 % more (tile, frame) pairs were kept for nothing)
    constexpr int kMaxBlocks = 12;
    int level = 0, tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;
    for (; level < 5; ++level) {           // index into levels 3..7 (block sizes 8..128)
      const int shift = 3 + level;
      tx0 = x0 >> shift; tx1 = x1 >> shift; ty0 = y0 >> shift; ty1 = y1 >> shift;
      if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) <= kMaxBlocks) break;
    }
    if (level == 5) return false;
    const int w_l = P.mip_w[level];
    const float* m = mips + P.mip_off[level];
    float d_max = -INFINITY;
    // (a fully unrolled form with predicated loads, all in flight together, was measured slower: twice the registers)
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) d_max = fmaxf(d_max, __ldg(m + ty * w_l + tx));
    d_far = d_max;
    return d_max + P.trunc_margin * 1.0001f + P.cull_z <= cz_min;
  }
};

__device__ __forceinline__ void apply_pending(float& t, float& w, unsigned c) {
  if (c == 0u) return;
  if (t == 1.0f) {
    w = __fadd_rn(w, (float)c);
  } else {
    for (unsigned i = 0; i < c; ++i) {
      const float w_new = __fadd_rn(w, 1.0f);
      t = __fdiv_rn(__fadd_rn(__fmul_rn(t, w), 1.0f), w_new);
      w = w_new;
    }
  }
}

// Frame-vs-tile test of the K-frame kernels: can frame f touch any voxel of the warp tile whose lowest corner is
// (x_lo, y_lo, z_lo) [global indices], and does the whole tile project well inside the image?  Frustum: five
// half-space tests on the 8 corners; occlusion: hierarchical-Z against the frame's depth max-pyramid.
template <int VEC, int ZT>
__device__ __forceinline__ void tile_frame_test(const IntegrateParams& P, const Intrinsics& K, int f, int x_lo, int y_lo,
                                                int rows_per_warp, int z_lo, bool& keep, bool& inside) {
  const FramePose& ps = P.pose[f];
  const int x_hi = min(P.dim_x - 1, x_lo + ((VEC << P.lx_shift) - 1));
  const int y_hi = min(P.dim_y - 1, y_lo + rows_per_warp - 1);
  const int z_hi = z_lo + ZT - 1;
  CullAccum acc;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float bx = fmaf((float)((c & 1) ? x_hi : x_lo), P.voxel_size, P.ox) - ps.tx;
    const float by = fmaf((float)((c & 2) ? y_hi : y_lo), P.voxel_size, P.oy) - ps.ty;
    const float bz = fmaf((float)((c & 4) ? z_hi : z_lo), P.voxel_size, P.oz) - ps.tz;
    acc.add(fmaf(ps.m0, bx, fmaf(ps.m4, by, ps.m8 * bz)), fmaf(ps.m1, bx, fmaf(ps.m5, by, ps.m9 * bz)),
            fmaf(ps.m2, bx, fmaf(ps.m6, by, ps.m10 * bz)), P, K);
  }
  keep = !acc.culled() && (y_lo < P.dim_y);
  if (keep && P.cull_level >= 2 && P.mips != nullptr && P.trunc_margin > 0.0f)
    keep = !acc.occluded(P, P.mips + (int64_t)f * P.mip_stride);
  inside = keep && acc.interior(P);
}

// The same test for an arbitrary voxel box [x_lo, x_hi] x [y_lo, y_hi] x [z_lo, z_hi] (global indices, inclusive)
// against one frame; PT carries the cull fields (cull_level, cull_z/gx/gy, im_w/h, mip_w/off, trunc_margin, ox/oy/oz,
// voxel_size).  mips_f: the frame's depth max-pyramid or nullptr.
template <class PT>
__device__ __forceinline__ void box_frame_test(const PT& P, const Intrinsics& K, const FramePose& ps,
                                               const float* __restrict__ mips_f, int x_lo, int x_hi, int y_lo, int y_hi,
                                               int z_lo, int z_hi, bool& keep, bool& inside, bool* near = nullptr) {
  CullAccum acc;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float bx = fmaf((float)((c & 1) ? x_hi : x_lo), P.voxel_size, P.ox) - ps.tx;
    const float by = fmaf((float)((c & 2) ? y_hi : y_lo), P.voxel_size, P.oy) - ps.ty;
    const float bz = fmaf((float)((c & 4) ? z_hi : z_lo), P.voxel_size, P.oz) - ps.tz;
    acc.add(fmaf(ps.m0, bx, fmaf(ps.m4, by, ps.m8 * bz)), fmaf(ps.m1, bx, fmaf(ps.m5, by, ps.m9 * bz)),
            fmaf(ps.m2, bx, fmaf(ps.m6, by, ps.m10 * bz)), P, K);
  }
  keep = !acc.culled();
  if (keep && P.cull_level >= 2 && mips_f != nullptr && P.trunc_margin > 0.0f) keep = !acc.occluded(P, mips_f);
  inside = keep && acc.interior(P);
  if (near != nullptr) *near = keep && acc.near_far_surface(P);
}

}  // namespace
}  // namespace tsdf
