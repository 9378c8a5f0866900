// geometry.cuh -- the epipolar masks of matchFeaturesConstrained / matchFeaturesDoubleConstrained,
// written with explicit rounding (__f*_rn intrinsics are never contracted) so that every
// intermediate equals what the reference's device code produces.  The sequences were
// read off the SASS nvcc 12.9 emits for the reference sources at sm_100a:
//   getEpipolarEndpoints   ref: src/MatchFactory.cu:1241-1277
//   rotatePoint            ref: src/matrix_util.cu:314-327 (+ matrixMulVector :269-282)
//   dotProduct(float3)     ref: src/cuda_vec_util.cu:1216-1218   -> FMUL(y) FFMA(x) FFMA(z)
//   multiply(float4[3],..) ref: src/matrix_util.cu:58-62         -> FMUL(y) FFMA(x) FFMA(z) FFMA(w)
//   mask tests             ref: src/MatchFactory.cu:2241-2254, :1613-1627
// This file must be compiled WITHOUT --use_fast_math (precise sinf/cosf/div/sqrt).
#pragma once
#include "matcher.cuh"

namespace b200 {

struct Segment {           // per-query state of the DoubleConstrained mask
  float lo_x, hi_x;        // left.x - eps, right.x + eps
  float left_x, left_y, slope;
  float top_lo, bot_hi;    // top - eps, bottom + eps (vertical segments only)
  int   vertical;
};
struct EpiLine { float ex, ey, ez; };   // F * (x, y, 1)

__device__ __forceinline__ float dot3_ref(float ax, float ay, float az, float bx, float by, float bz) {
  return __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by)));
}

__device__ __forceinline__ void rotate_point_ref(float x, float y, float z, const float* ang, float& ox, float& oy, float& oz) {
  const float cx = cosf(ang[0]), sx = sinf(ang[0]);
  const float cy = cosf(ang[1]), sy = sinf(ang[1]);
  const float cz = cosf(ang[2]), sz = sinf(ang[2]);
  const float czsy = __fmul_rn(cz, sy), szsy = __fmul_rn(sz, sy);
  const float m00 = __fmul_rn(cz, cy);
  const float m01 = __fmaf_rn(-sz, cx, __fmul_rn(czsy, sx));
  const float m02 = __fmaf_rn(sz, sx, __fmul_rn(czsy, cx));
  const float m10 = __fmul_rn(sz, cy);
  const float m11 = __fmaf_rn(cz, cx, __fmul_rn(szsy, sx));
  const float m12 = __fmaf_rn(-cz, sx, __fmul_rn(szsy, cx));
  const float m20 = -sy;
  const float m21 = __fmul_rn(cy, sx);
  const float m22 = __fmul_rn(cy, cx);
  ox = __fmaf_rn(m02, z, __fmaf_rn(m01, y, __fmaf_rn(m00, x, 0.0f)));
  oy = __fmaf_rn(m12, z, __fmaf_rn(m11, y, __fmaf_rn(m10, x, 0.0f)));
  oz = __fmaf_rn(m22, z, __fmaf_rn(m21, y, __fmaf_rn(m20, x, 0.0f)));
}

__device__ __forceinline__ float row4_ref(const float* r, float X, float Y, float Z) {
  return __fmaf_rn(r[3], 1.0f, __fmaf_rn(r[2], Z, __fmaf_rn(r[0], X, __fmul_rn(r[1], Y))));
}

// p = (p1.x, p1.y, p2.x, p2.y)
__device__ __forceinline__ float4 epipolar_endpoints_ref(const MaskParams& mp, float2 loc) {
  const RefCamera& c = mp.qcam;
  const float v0x = __fmul_rn(c.dpix[0], __fmaf_rn(static_cast<float>(c.size[0]), -0.5f, loc.x));
  const float v0y = __fmul_rn(c.dpix[1], __fmaf_rn(static_cast<float>(c.size[1]), -0.5f, loc.y));
  float vx, vy, vz;
  rotate_point_ref(v0x, v0y, c.foc, c.cam_rot, vx, vy, vz);
  const float cx = __fadd_rn(c.cam_pos[0], c.ecef_offset[0]);
  const float cy = __fadd_rn(c.cam_pos[1], c.ecef_offset[1]);
  const float cz = __fadd_rn(c.cam_pos[2], c.ecef_offset[2]);
  const float a = dot3_ref(vx, vy, vz, vx, vy, vz);
  const float hb = dot3_ref(vx, vy, vz, cx, cy, cz);
  const float b = __fadd_rn(hb, hb);
  const double cc = static_cast<double>(dot3_ref(cx, cy, cz, cx, cy, cz));
  const double rmax = __dadd_rn(6384.4, static_cast<double>(mp.delta));     // EARTH_MAX_KM_FROM_CENT + delta
  const double rmin = __dadd_rn(6356.77, -static_cast<double>(mp.delta));   // EARTH_MIN_KM_FROM_CENT - delta
  const float c1 = __double2float_rn(__fma_rn(-rmax, rmax, cc));
  const float c2 = __double2float_rn(__fma_rn(-rmin, rmin, cc));
  const float bb = __fmul_rn(b, b), a4 = __fmul_rn(a, 4.0f), a2 = __fadd_rn(a, a);
  float out[4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float disc = __fmaf_rn(k == 0 ? c1 : c2, -a4, bb);
    const float t = __fdiv_rn(__fadd_rn(-b, -__fsqrt_rn(disc)), a2);
    const float X = __fadd_rn(__fmul_rn(t, vx), cx);
    const float Y = __fadd_rn(__fmul_rn(t, vy), cy);
    const float Z = __fadd_rn(__fmul_rn(t, vz), cz);
    const float x = row4_ref(mp.P + 0, X, Y, Z), y = row4_ref(mp.P + 4, X, Y, Z), z = row4_ref(mp.P + 8, X, Y, Z);
    out[2 * k] = __fdiv_rn(x, z);
    out[2 * k + 1] = __fdiv_rn(y, z);
  }
  return make_float4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ Segment make_segment_ref(float4 p, float eps) {
  Segment s;
  float lx, ly, rx, ry;
  if (p.x < p.z) { lx = p.x; ly = p.y; rx = p.z; ry = p.w; }
  else           { lx = p.z; ly = p.w; rx = p.x; ry = p.y; }
  s.vertical = (lx == rx);
  float top = 0.0f, bottom = 0.0f;
  s.slope = 0.0f;
  if (s.vertical) {
    if (p.y < p.w) { top = p.y; bottom = p.w; } else { top = p.w; bottom = p.y; }
  } else {
    s.slope = __fdiv_rn(__fadd_rn(ly, -ry), __fadd_rn(lx, -rx));
  }
  s.left_x = lx; s.left_y = ly;
  s.lo_x = __fadd_rn(lx, -eps); s.hi_x = __fadd_rn(rx, eps);
  s.top_lo = __fadd_rn(top, -eps); s.bot_hi = __fadd_rn(bottom, eps);
  return s;
}

// true = the target is skipped.  NaN operands compare false everywhere => the target passes (reference quirk).
__device__ __forceinline__ bool segment_skips(const Segment& s, float tx, float ty, float eps) {
  if (tx < s.lo_x || tx > s.hi_x) return true;
  if (s.vertical) return (s.top_lo > ty) || (s.bot_hi < ty);
  const float yl = __fmaf_rn(__fadd_rn(tx, -s.left_x), s.slope, s.left_y);
  return fabsf(__fadd_rn(yl, -ty)) > eps;
}
// hot-loop form for non-vertical segments
__device__ __forceinline__ bool segment_skips_nv(const Segment& s, float tx, float ty, float eps) {
  const float yl = __fmaf_rn(__fadd_rn(tx, -s.left_x), s.slope, s.left_y);
  return (tx < s.lo_x) | (tx > s.hi_x) | (fabsf(__fadd_rn(yl, -ty)) > eps);
}

__device__ __forceinline__ EpiLine make_epiline_ref(const float* F, float2 loc) {
  EpiLine e;
  e.ex = __fadd_rn(__fmaf_rn(F[0], loc.x, __fmul_rn(F[1], loc.y)), F[2]);
  e.ey = __fadd_rn(__fmaf_rn(F[3], loc.x, __fmul_rn(F[4], loc.y)), F[5]);
  e.ez = __fadd_rn(__fmaf_rn(F[6], loc.x, __fmul_rn(F[7], loc.y)), F[8]);
  return e;
}
__device__ __forceinline__ bool epiline_skips(const EpiLine& e, float tx, float ty, float eps) {
  const float pq = __fdiv_rn(__fmaf_rn(e.ex, tx, e.ez), e.ey);
  return fabsf(__fadd_rn(ty, pq)) > eps;
}

// One object for all three mask kinds, built once per query.
struct QueryMask {
  Segment seg;
  EpiLine line;
};
template <int kMask>
__device__ __forceinline__ QueryMask make_query_mask(const MaskParams& mp, float2 loc) {
  QueryMask m;
  if (kMask == SSRLCV_B200_MASK_DOUBLE) m.seg = make_segment_ref(epipolar_endpoints_ref(mp, loc), mp.eps);
  if (kMask == SSRLCV_B200_MASK_FUNDAMENTAL) m.line = make_epiline_ref(mp.F, loc);
  return m;
}
template <int kMask>
__device__ __forceinline__ bool mask_skips(const QueryMask& m, float tx, float ty, float eps) {
  if (kMask == SSRLCV_B200_MASK_DOUBLE) return segment_skips(m.seg, tx, ty, eps);
  if (kMask == SSRLCV_B200_MASK_FUNDAMENTAL) return epiline_skips(m.line, tx, ty, eps);
  return false;
}

// true = EVERY target whose location lies in the box (x0, x1, y0, y1) = (b.x, b.y, b.z, b.w) is skipped by this query's mask,
// decided from the box alone and exactly conservative -- no tolerance needed: every rounded operation of the per-target tests
// above is monotone in the target's coordinates (tx - left_x, the fma, yl - ty; ex * tx + ez, the division, ty + pq), so over
// the box the tested quantity ranges between its values at the box's corners.  Any NaN (in the box or in the query's state)
// makes every comparison false: not culled, as a NaN target passes the per-target test.  An empty box (+inf, -inf, ...) culls.
template <int kMask>
__device__ __forceinline__ bool box_skips(const QueryMask& m, float4 b, float eps) {
  if (kMask == SSRLCV_B200_MASK_DOUBLE) {
    const Segment& s = m.seg;
    if (b.y < s.lo_x || b.x > s.hi_x) return true;            // all of it left / right of the segment
    if (s.vertical) return (s.top_lo > b.w) || (s.bot_hi < b.z);
    const float yl0 = __fmaf_rn(__fadd_rn(b.x, -s.left_x), s.slope, s.left_y);
    const float yl1 = __fmaf_rn(__fadd_rn(b.y, -s.left_x), s.slope, s.left_y);
    if (yl0 != yl0 || yl1 != yl1) return false;
    const float ya = fminf(yl0, yl1), yb = fmaxf(yl0, yl1);
    return (__fadd_rn(ya, -b.w) > eps) || (__fadd_rn(yb, -b.z) < -eps);   // the band passes entirely below / above the box
  }
  if (kMask == SSRLCV_B200_MASK_FUNDAMENTAL) {
    const EpiLine& e = m.line;
    const float p0 = __fdiv_rn(__fmaf_rn(e.ex, b.x, e.ez), e.ey);
    const float p1 = __fdiv_rn(__fmaf_rn(e.ex, b.y, e.ez), e.ey);
    if (p0 != p0 || p1 != p1) return false;
    const float pa = fminf(p0, p1), pb = fmaxf(p0, p1);
    return (__fadd_rn(b.z, pa) > eps) || (__fadd_rn(b.w, pb) < -eps);
  }
  return false;
}

// the mask state of lane `src`, on every lane of the warp
template <int kMask>
__device__ __forceinline__ QueryMask shfl_query_mask(const QueryMask& m, int src) {
  QueryMask r;
  if (kMask == SSRLCV_B200_MASK_DOUBLE) {
    r.seg.lo_x = __shfl_sync(0xffffffffu, m.seg.lo_x, src); r.seg.hi_x = __shfl_sync(0xffffffffu, m.seg.hi_x, src);
    r.seg.left_x = __shfl_sync(0xffffffffu, m.seg.left_x, src); r.seg.left_y = __shfl_sync(0xffffffffu, m.seg.left_y, src);
    r.seg.slope = __shfl_sync(0xffffffffu, m.seg.slope, src); r.seg.top_lo = __shfl_sync(0xffffffffu, m.seg.top_lo, src);
    r.seg.bot_hi = __shfl_sync(0xffffffffu, m.seg.bot_hi, src); r.seg.vertical = __shfl_sync(0xffffffffu, m.seg.vertical, src);
  }
  if (kMask == SSRLCV_B200_MASK_FUNDAMENTAL) {
    r.line.ex = __shfl_sync(0xffffffffu, m.line.ex, src); r.line.ey = __shfl_sync(0xffffffffu, m.line.ey, src);
    r.line.ez = __shfl_sync(0xffffffffu, m.line.ez, src);
  }
  return r;
}

// The per-query mask state is computed once per call by a prologue kernel (the trigonometry of getEpipolarEndpoints is
// ~10^3 instructions) and stored as two float4 per query; the matcher kernels only load it.
template <int kMask>
__device__ __forceinline__ void store_query_mask(float4* dst, const QueryMask& m) {
  if (kMask == SSRLCV_B200_MASK_DOUBLE) {
    dst[0] = make_float4(m.seg.lo_x, m.seg.hi_x, m.seg.left_x, m.seg.left_y);
    dst[1] = make_float4(m.seg.slope, m.seg.top_lo, m.seg.bot_hi, __int_as_float(m.seg.vertical));
  } else {
    dst[0] = make_float4(m.line.ex, m.line.ey, m.line.ez, 0.f);
    dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int kMask>
__device__ __forceinline__ QueryMask load_query_mask(const float4* src) {
  QueryMask m;
  if (kMask == SSRLCV_B200_MASK_DOUBLE) {
    const float4 a = src[0], b = src[1];
    m.seg.lo_x = a.x; m.seg.hi_x = a.y; m.seg.left_x = a.z; m.seg.left_y = a.w;
    m.seg.slope = b.x; m.seg.top_lo = b.y; m.seg.bot_hi = b.z; m.seg.vertical = __float_as_int(b.w);
  }
  if (kMask == SSRLCV_B200_MASK_FUNDAMENTAL) {
    const float4 a = src[0];
    m.line.ex = a.x; m.line.ey = a.y; m.line.ez = a.z;
  }
  return m;
}

// The same load for a bookkeeping path that must not keep the state in registers across the matcher's hot loop (smax_tc.cu kLean):
// volatile, so that the compiler cannot hoist it.
template <int kMask>
__device__ __forceinline__ QueryMask load_query_mask_rare(const float4* src) {
  QueryMask m;
  float4 a, b;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src) : "memory");
  if (kMask == SSRLCV_B200_MASK_DOUBLE) {
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + 1) : "memory");
    m.seg.lo_x = a.x; m.seg.hi_x = a.y; m.seg.left_x = a.z; m.seg.left_y = a.w;
    m.seg.slope = b.x; m.seg.top_lo = b.y; m.seg.bot_hi = b.z; m.seg.vertical = __float_as_int(b.w);
  }
  if (kMask == SSRLCV_B200_MASK_FUNDAMENTAL) { m.line.ex = a.x; m.line.ey = a.y; m.line.ez = a.z; }
  return m;
}

}  // namespace b200
