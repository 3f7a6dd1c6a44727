// Decode3DBbox + RotatedIOU (+ fused batch max) for sm_100a.
//
// Reference behaviour reproduced (files under /root/reference):
//   operator_cxx/contrib/decode_3d_bbox-inl.h:64-277   per-point box decode
//   operator_cxx/contrib/rotated_iou-inl.h:49-523      all-pairs rotated IoU, box types 5/7/8
//   operator_py/batch_rotated_iou.py:11-68             per-image IoU -> sanitise -> row max
//
// Compiled with -fmad=false: the polygon clipping takes many sign decisions on differences of
// products; keeping every operation individually rounded (as the reference's host build does)
// keeps those decisions identical, so only atan2f/sinf/cosf/expf last-ulp differences remain.
//
// Roofline notes: decode is HBM-bound (84 B/box, AoS staged through shared memory so
// all global traffic is coalesced); rotated IoU is ALU/SFU-latency bound -- an exact AABB
// rejection (valid for convex quads, see quad_is_convex) removes ~99% of the clipping work.
#include <math.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// Decode
// ------------------------------------------------------------------------------------------
constexpr int DEC_THREADS = 256;

struct F2 {
  float x, y;
};

__device__ __forceinline__ F2 rot2(float x, float y, float s, float c) {
  F2 r;
  r.x = x * c - y * s;
  r.y = x * s + y * c;
  return r;
}

template <bool IS_BIN>
__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const float* __restrict__ delta, const float* __restrict__ pc, float* __restrict__ out,
              int64_t n_total) {
  constexpr int D = IS_BIN ? 7 : 8;
  constexpr int DP = 9;   // padded row strides: 9 and 11 are odd -> conflict-free per-box access
  constexpr int OP = 11;
  __shared__ float sd[DEC_THREADS * DP];
  __shared__ float sp[DEC_THREADS * 3];
  __shared__ float so[DEC_THREADS * OP];
  const int64_t base = (int64_t)blockIdx.x * DEC_THREADS;
  const int nloc = (int)min((int64_t)DEC_THREADS, n_total - base);
  const int t = threadIdx.x;
  // coalesced staging of the AoS rows
  for (int e = t; e < nloc * D; e += DEC_THREADS) sd[(e / D) * DP + (e % D)] = __ldg(delta + base * D + e);
  for (int e = t; e < nloc * 3; e += DEC_THREADS) sp[e] = __ldg(pc + base * 3 + e);
  __syncthreads();
  if (t < nloc) {
    const float* d = sd + t * DP;
    const float px = sp[t * 3 + 0], py = sp[t * 3 + 1];
    const float az = atan2f(py, px);
    float sa, ca;
    sa = sinf(az);
    ca = cosf(az);
    float dx, dy, width, length, height, z0, yaw;
    if (IS_BIN) {  // decode_3d_bbox-inl.h:82-125
      dx = d[0];
      dy = d[1];
      width = expf(d[3]);
      length = expf(d[4]);
      height = expf(d[5]);
      const float cz = sp[t * 3 + 2] + d[2];
      z0 = cz - height / 2.0f;
      yaw = d[6] + az;
    } else {  // :186-235
      dx = d[0] * fabsf(d[0]);
      dy = d[1] * fabsf(d[1]);
      width = expf(d[2]);
      length = expf(d[3]);
      height = expf(d[7]);
      z0 = d[6];
      yaw = atan2f(d[5], d[4]) + az;
    }
    const float cx = px + (dx * ca - dy * sa);
    const float cy = py + (dx * sa + dy * ca);
    const float s = sinf(yaw), c = cosf(yaw);
    const float hl = 0.5f * length, hw = 0.5f * width;
    const F2 A = rot2(hl, -hw, s, c), B = rot2(-hl, -hw, s, c), C = rot2(-hl, hw, s, c),
             Dd = rot2(hl, hw, s, c);
    float* o = so + t * OP;
    o[0] = A.x + cx;  o[1] = A.y + cy;
    o[2] = B.x + cx;  o[3] = B.y + cy;
    o[4] = C.x + cx;  o[5] = C.y + cy;
    o[6] = Dd.x + cx; o[7] = Dd.y + cy;
    o[8] = z0;
    o[9] = z0 + height;
  }
  __syncthreads();
  for (int e = t; e < nloc * 10; e += DEC_THREADS) out[base * 10 + e] = so[(e / 10) * OP + (e % 10)];
}

// ------------------------------------------------------------------------------------------
// Rotated IoU
// ------------------------------------------------------------------------------------------
constexpr float R_EPS = 1e-8f;

__device__ __forceinline__ float mmin(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float mmax(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float smin(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return a < b ? b : a; }

__device__ __forceinline__ bool rel_equal(float a, float b) {
  return fabsf(__fdiv_rn(a - b, mmin(a, b))) < R_EPS;
}
__device__ __forceinline__ bool within(float lo, float hi, float v) {
  return (lo < v || rel_equal(lo, v)) && (hi > v || rel_equal(hi, v));
}

struct Quad {
  float x[4], y[4];
};

struct Poly {
  float px[24], py[24];
  int n;
  float sx, sy;
  __device__ __forceinline__ void push(float x, float y) {
    sx = sx + x;
    sy = sy + y;
    if (n < 24) {
      px[n] = x;
      py[n] = y;
    }
    n++;
  }
};

// segment (p0->p1) x (q0->q1), rotated_iou-inl.h:131-172
__device__ __forceinline__ bool seg_isect(float p1x, float p1y, float p0x, float p0y, float q1x,
                                          float q1y, float q0x, float q0y, float* ox, float* oy) {
  const bool touch = mmin(p0x, p1x) <= mmax(q0x, q1x) && mmin(q0x, q1x) <= mmax(p0x, p1x) &&
                     mmin(p0y, p1y) <= mmax(q0y, q1y) && mmin(q0y, q1y) <= mmax(p0y, p1y);
  if (!touch) return false;
  const float A1 = p1y - p0y, B1 = p0x - p1x, C1 = A1 * p0x + B1 * p0y;
  const float A2 = q1y - q0y, B2 = q0x - q1x, C2 = A2 * q0x + B2 * q0y;
  const float det = A1 * B2 - A2 * B1;
  if (rel_equal(det, 0.0f)) return false;
  const float x = __fdiv_rn(B2 * C1 - B1 * C2, det);
  const float y = __fdiv_rn(A1 * C2 - A2 * C1, det);
  const bool on1 = within(smin(p0x, p1x), smax(p0x, p1x), x) && within(smin(p0y, p1y), smax(p0y, p1y), y);
  const bool on2 = within(smin(q0x, q1x), smax(q0x, q1x), x) && within(smin(q0y, q1y), smax(q0y, q1y), y);
  if (on1 && on2) {
    *ox = x;
    *oy = y;
    return true;
  }
  return false;
}

__device__ __forceinline__ void edge_crossings(const Quad& a, const Quad& b, Poly& P) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int i1 = (i + 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int j1 = (j + 1) & 3;
      float x, y;
      if (seg_isect(a.x[i1], a.y[i1], a.x[i], a.y[i], b.x[j1], b.y[j1], b.x[j], b.y[j], &x, &y))
        P.push(x, y);
    }
  }
}

// centroid, angular bubble sort (angles evaluated once per vertex: point_cmp is a pure function of
// vertex and centre, rotated_iou-inl.h:187-192), triangle-fan area :441-463
__device__ float poly_area(Poly& P) {
  const int cnt = P.n < 24 ? P.n : 24;
  const float cx = __fdiv_rn(P.sx, (float)P.n), cy = __fdiv_rn(P.sy, (float)P.n);
  float ang[24];
  for (int i = 0; i < cnt; ++i) ang[i] = atan2f(P.py[i] - cy, P.px[i] - cx);
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        float t = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = t;
        t = P.px[i]; P.px[i] = P.px[i + 1]; P.px[i + 1] = t;
        t = P.py[i]; P.py[i] = P.py[i + 1]; P.py[i + 1] = t;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ax = P.px[k] - P.px[0], ay = P.py[k] - P.py[0];
    const float bx = P.px[k + 1] - P.px[0], by = P.py[k + 1] - P.py[0];
    area += ax * by - ay * bx;
  }
  return fabsf(area) / 2.0f;
}

__device__ __forceinline__ bool in_quad(const Quad& q, float px, float py) {  // :113-128
  int flag = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const float pos = (q.x[j] - q.x[i]) * (py - q.y[i]) - (q.y[j] - q.y[i]) * (px - q.x[i]);
    const int s = pos >= 0.0f;
    if (flag == -1) flag = s;
    else if (flag != s) return false;
  }
  return true;
}

__device__ __forceinline__ float quad_area(const Quad& q) {  // iou_bev_8pts :482-487
  float s = (q.x[1] - q.x[0]) * (q.y[2] - q.y[0]) - (q.y[1] - q.y[0]) * (q.x[2] - q.x[0]);
  s += (q.x[2] - q.x[0]) * (q.y[3] - q.y[0]) - (q.y[2] - q.y[0]) * (q.x[3] - q.x[0]);
  return fabsf(s) / 2.0f;
}

// Strictly convex, non-degenerate quad?  For such quads a point outside the axis-aligned bounding
// box fails check_in_box2d_8pts and no edge pair passes check_rect_cross, so AABB-disjoint pairs
// give overlap 0 -> IoU +0 exactly as the full evaluation would.  Anything else (bow-ties, NaNs,
// zero-area padding boxes are fine: they are tiny squares) takes the full path.
__device__ __forceinline__ bool quad_is_convex(const Quad& q) {
  int pos = 0, neg = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3, k = (i + 2) & 3;
    const float z = (q.x[j] - q.x[i]) * (q.y[k] - q.y[j]) - (q.y[j] - q.y[i]) * (q.x[k] - q.x[j]);
    pos += z > 0.f;
    neg += z < 0.f;
  }
  return pos == 4 || neg == 4;
}

struct Box2 {  // quad + cached bbox / area / convexity
  Quad q;
  float minx, maxx, miny, maxy;
  float area;
  int convex;
};

__device__ __forceinline__ void finish_box(Box2& b) {
  b.minx = fminf(fminf(b.q.x[0], b.q.x[1]), fminf(b.q.x[2], b.q.x[3]));
  b.maxx = fmaxf(fmaxf(b.q.x[0], b.q.x[1]), fmaxf(b.q.x[2], b.q.x[3]));
  b.miny = fminf(fminf(b.q.y[0], b.q.y[1]), fminf(b.q.y[2], b.q.y[3]));
  b.maxy = fmaxf(fmaxf(b.q.y[0], b.q.y[1]), fmaxf(b.q.y[2], b.q.y[3]));
  // fminf/fmaxf drop NaNs: make NaN corners defeat the rejection test
  bool finite = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) finite = finite && isfinite(b.q.x[i]) && isfinite(b.q.y[i]);
  b.convex = finite && quad_is_convex(b.q);
}

__device__ __forceinline__ bool aabb_disjoint(const Box2& a, const Box2& b) {
  return a.maxx < b.minx || b.maxx < a.minx || a.maxy < b.miny || b.maxy < a.miny;
}

__device__ float iou_quads(const Box2& a, const Box2& b) {  // iou_bev_8pts :478-493
  if (a.area < R_EPS || b.area < R_EPS) return 0.f;
  if (a.convex && b.convex && aabb_disjoint(a, b)) return 0.f;
  Poly P;
  P.n = 0;
  P.sx = 0.f;
  P.sy = 0.f;
  edge_crossings(a.q, b.q, P);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (in_quad(a.q, b.q.x[k], b.q.y[k])) P.push(b.q.x[k], b.q.y[k]);
    if (in_quad(b.q, a.q.x[k], a.q.y[k])) P.push(a.q.x[k], a.q.y[k]);
  }
  const float so = poly_area(P);
  return __fdiv_rn(so, fmaxf(a.area + b.area - so, R_EPS));
}

// ---- box types 5 (x,y,w,h,angle) and 7 (x,y,z,w,l,h,angle) : :195-386, :467-475, :496-507 ----
struct Rect {
  float cx, cy, w, h, ang;  // w,h = footprint extents along the box axes
  Quad q;                   // rotated corners
  float minx, maxx, miny, maxy;
  int plain;  // w>0, h>0, all finite -> AABB rejection is exact
};

__device__ __forceinline__ void make_rect(Rect& r) {
  const float c = cosf(r.ang), s = sinf(r.ang);
  const float lx[4] = {r.cx - r.w / 2, r.cx + r.w / 2, r.cx + r.w / 2, r.cx - r.w / 2};
  const float ly[4] = {r.cy - r.h / 2, r.cy - r.h / 2, r.cy + r.h / 2, r.cy + r.h / 2};
  bool finite = isfinite(r.ang);
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // rotate_around_center :175-184
    r.q.x[k] = (lx[k] - r.cx) * c + (ly[k] - r.cy) * s + r.cx;
    r.q.y[k] = -(lx[k] - r.cx) * s + (ly[k] - r.cy) * c + r.cy;
    finite = finite && isfinite(r.q.x[k]) && isfinite(r.q.y[k]);
  }
  r.minx = fminf(fminf(r.q.x[0], r.q.x[1]), fminf(r.q.x[2], r.q.x[3]));
  r.maxx = fmaxf(fmaxf(r.q.x[0], r.q.x[1]), fmaxf(r.q.x[2], r.q.x[3]));
  r.miny = fminf(fminf(r.q.y[0], r.q.y[1]), fminf(r.q.y[2], r.q.y[3]));
  r.maxy = fmaxf(fmaxf(r.q.y[0], r.q.y[1]), fmaxf(r.q.y[2], r.q.y[3]));
  r.plain = finite && r.w > 0.f && r.h > 0.f;
}

__device__ __forceinline__ bool in_rect(const Rect& r, float px, float py) {  // :81-110
  const float c = cosf(-r.ang), s = sinf(-r.ang);
  const float rx = (px - r.cx) * c + (py - r.cy) * s + r.cx;
  const float ry = -(px - r.cx) * s + (py - r.cy) * c + r.cy;
  return rx >= r.cx - r.w / 2 && rx <= r.cx + r.w / 2 && ry >= r.cy - r.h / 2 && ry <= r.cy + r.h / 2;
}

__device__ float overlap_rects(const Rect& a, const Rect& b) {
  if (a.plain && b.plain &&
      (a.maxx < b.minx || b.maxx < a.minx || a.maxy < b.miny || b.maxy < a.miny))
    return 0.f;
  Poly P;
  P.n = 0;
  P.sx = 0.f;
  P.sy = 0.f;
  edge_crossings(a.q, b.q, P);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (in_rect(a, b.q.x[k], b.q.y[k])) P.push(b.q.x[k], b.q.y[k]);
    if (in_rect(b, a.q.x[k], a.q.y[k])) P.push(a.q.x[k], a.q.y[k]);
  }
  return poly_area(P);
}

__device__ __forceinline__ void load_rect5(const float* p, Rect& r) {
  r.cx = p[0]; r.cy = p[1]; r.w = p[2]; r.h = p[3]; r.ang = p[4];
  make_rect(r);
}
__device__ __forceinline__ void load_rect7(const float* p, Rect& r, float* z, float* hh) {
  r.cx = p[0]; r.cy = p[1]; r.w = p[3]; r.h = p[4]; r.ang = p[6];
  *z = p[2];
  *hh = p[5];
  make_rect(r);
}

__device__ __forceinline__ float iou_rect5(const Rect& a, const Rect& b) {
  const float sa = a.w * a.h, sb = b.w * b.h;
  if (sa < R_EPS || sb < R_EPS) return 0.f;
  const float so = overlap_rects(a, b);
  return __fdiv_rn(so, fmaxf(sa + sb - so, R_EPS));
}
__device__ __forceinline__ float iou_rect7(const Rect& a, float az, float ah, const Rect& b, float bz,
                                           float bh) {
  const float sa = a.w * a.h * ah, sb = b.w * b.h * bh;
  if (sa < R_EPS || sb < R_EPS) return 0.f;
  const float so = overlap_rects(a, b);
  const float ho = mmax(0.0f, mmin(az + ah / 2.0f, bz + bh / 2.0f) - mmax(az - ah / 2.0f, bz - bh / 2.0f));
  return __fdiv_rn(so * ho, fmaxf(sa + sb - so * ho, R_EPS));
}

__device__ __forceinline__ void load_box8(const float* p, Box2& b) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    b.q.x[k] = p[2 * k];
    b.q.y[k] = p[2 * k + 1];
  }
  b.area = quad_area(b.q);
  finish_box(b);
}

// All-pairs matrix: lanes run along boxes2 so the (n1,n2) row-major output is written coalesced.
constexpr int IOU_TJ = 128;  // boxes2 per block (= threads)
constexpr int IOU_TI = 32;   // boxes1 per block

template <int T>
__global__ void __launch_bounds__(IOU_TJ)
rotated_iou_kernel(const float* __restrict__ b1, const float* __restrict__ b2, float* __restrict__ out,
                   int64_t n1, int64_t n2) {
  __shared__ float s1[IOU_TI * T];
  const int64_t i0 = (int64_t)blockIdx.y * IOU_TI;
  const int64_t j = (int64_t)blockIdx.x * IOU_TJ + threadIdx.x;
  const int ni = (int)min((int64_t)IOU_TI, n1 - i0);
  for (int e = threadIdx.x; e < ni * T; e += IOU_TJ) s1[e] = __ldg(b1 + i0 * T + e);
  __syncthreads();
  if (j >= n2) return;
  float mine[T];
#pragma unroll
  for (int k = 0; k < T; ++k) mine[k] = __ldg(b2 + j * T + k);
  if (T == 8) {
    Box2 bj;
    load_box8(mine, bj);
    for (int i = 0; i < ni; ++i) {
      Box2 bi;
      load_box8(s1 + i * T, bi);
      out[(i0 + i) * n2 + j] = iou_quads(bi, bj);
    }
  } else if (T == 5) {
    Rect rj;
    load_rect5(mine, rj);
    for (int i = 0; i < ni; ++i) {
      Rect ri;
      load_rect5(s1 + i * T, ri);
      out[(i0 + i) * n2 + j] = iou_rect5(ri, rj);
    }
  } else {
    Rect rj;
    float zj, hj;
    load_rect7(mine, rj, &zj, &hj);
    for (int i = 0; i < ni; ++i) {
      Rect ri;
      float zi, hi;
      load_rect7(s1 + i * T, ri, &zi, &hi);
      out[(i0 + i) * n2 + j] = iou_rect7(ri, zi, hi, rj, zj, hj);
    }
  }
}

// Fused batch_rotated_iou: one thread per proposal, GT boxes of the image staged in shared memory
// with their bbox/area/convexity precomputed; sanitise + running max in registers.
constexpr int BMAX_THREADS = 128;
constexpr int BMAX_GCHUNK = 256;

__device__ __forceinline__ float sanitise(float v) {  // batch_rotated_iou.py:43-46
  return (isnan(v) || isinf(v) || v > 1.0f || v < 0.0f) ? 0.f : v;
}

__global__ void __launch_bounds__(BMAX_THREADS)
batch_iou_max_bev_kernel(const float* __restrict__ prop, const float* __restrict__ gt,
                         float* __restrict__ out, int64_t N, int G) {
  __shared__ Box2 sg[BMAX_GCHUNK];
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * BMAX_THREADS + threadIdx.x;
  Box2 me;
  const bool active = n < N;
  if (active) {
    const float* p = prop + ((int64_t)b * N + n) * 10;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(p + k);
    load_box8(v, me);
  }
  float best = -INFINITY;
  for (int g0 = 0; g0 < G; g0 += BMAX_GCHUNK) {
    const int ng = min(BMAX_GCHUNK, G - g0);
    __syncthreads();
    for (int g = threadIdx.x; g < ng; g += BMAX_THREADS) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(gt + ((int64_t)b * G + g0 + g) * 8 + k);
      Box2 t;
      load_box8(v, t);
      sg[g] = t;
    }
    __syncthreads();
    if (active) {
      for (int g = 0; g < ng; ++g) {
        const float v = sanitise(iou_quads(me, sg[g]));
        best = v > best ? v : best;
      }
    }
  }
  if (active) out[(int64_t)b * N + n] = best;
}

// '3d' mode: to_box_type_7 (batch_rotated_iou.py:51-68) + yaw negation (:35-36) + iou_3d.
__global__ void __launch_bounds__(BMAX_THREADS)
batch_iou_max_3d_kernel(const float* __restrict__ prop, const float* __restrict__ gt,
                        float* __restrict__ out, int64_t N, int G) {
  __shared__ float sg[BMAX_GCHUNK * 7];
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * BMAX_THREADS + threadIdx.x;
  const bool active = n < N;
  Rect me;
  float mz = 0.f, mh = 0.f;
  if (active) {
    const float* p = prop + ((int64_t)b * N + n) * 10;
    float v[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) v[k] = __ldg(p + k);
    float b7[7];
    b7[0] = (((v[0] + v[2]) + v[4]) + v[6]) / 4.0f;
    b7[1] = (((v[1] + v[3]) + v[5]) + v[7]) / 4.0f;
    b7[2] = (v[8] + v[9]) / 2.0f;
    const float l0 = v[0] - v[2], l1 = v[1] - v[3], w0 = v[2] - v[4], w1 = v[3] - v[5];
    b7[3] = sqrtf(l0 * l0 + l1 * l1);
    b7[4] = sqrtf(w0 * w0 + w1 * w1);
    b7[5] = v[9] - v[8];
    b7[6] = -1.0f * atan2f(v[1] - v[3], v[0] - v[2]);
    load_rect7(b7, me, &mz, &mh);
  }
  float best = -INFINITY;
  for (int g0 = 0; g0 < G; g0 += BMAX_GCHUNK) {
    const int ng = min(BMAX_GCHUNK, G - g0);
    __syncthreads();
    for (int e = threadIdx.x; e < ng * 7; e += BMAX_THREADS) {
      float v = __ldg(gt + ((int64_t)b * G + g0) * 7 + e);
      if (e % 7 == 6) v = -1.0f * v;
      sg[e] = v;
    }
    __syncthreads();
    if (active) {
      for (int g = 0; g < ng; ++g) {
        Rect rg;
        float gz, gh;
        load_rect7(sg + g * 7, rg, &gz, &gh);
        const float v = sanitise(iou_rect7(me, mz, mh, rg, gz, gh));
        best = v > best ? v : best;
      }
    }
  }
  if (active) out[(int64_t)b * N + n] = best;
}

}  // namespace

extern "C" {

int rd_decode_3d_bbox(const float* delta, const float* pc, float* out, int64_t n_total, int is_bin,
                      rd_stream_t stream) {
  RD_REQUIRE(n_total >= 0, "rd_decode_3d_bbox: negative n_total");
  if (n_total == 0) return 0;
  RD_REQUIRE(delta && pc && out, "rd_decode_3d_bbox: null pointer");
  if (rd_check_device()) return 1;
  const int64_t blocks = (n_total + DEC_THREADS - 1) / DEC_THREADS;
  RD_REQUIRE(blocks <= 0x7fffffffLL, "rd_decode_3d_bbox: n_total too large");
  if (is_bin)
    decode_kernel<true><<<(unsigned)blocks, DEC_THREADS, 0, rd::as_stream(stream)>>>(delta, pc, out, n_total);
  else
    decode_kernel<false><<<(unsigned)blocks, DEC_THREADS, 0, rd::as_stream(stream)>>>(delta, pc, out, n_total);
  rd::count_launch();
  return rd::check_launch("rd_decode_3d_bbox");
}

int rd_rotated_iou(const float* boxes1, const float* boxes2, float* ious, int64_t n1, int64_t n2,
                   int box_type, rd_stream_t stream) {
  RD_REQUIRE(box_type == 5 || box_type == 7 || box_type == 8,
             "rd_rotated_iou: box_type must be 5, 7 or 8 (got %d)", box_type);
  RD_REQUIRE(n1 >= 0 && n2 >= 0, "rd_rotated_iou: negative size");
  if (n1 == 0 || n2 == 0) return 0;
  RD_REQUIRE(boxes1 && boxes2 && ious, "rd_rotated_iou: null pointer");
  if (rd_check_device()) return 1;
  const int64_t gx = (n2 + IOU_TJ - 1) / IOU_TJ, gy = (n1 + IOU_TI - 1) / IOU_TI;
  // gridDim.y is limited to 65535: fold the overflow into several launches
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    const int64_t ny = gy - y0 < max_gy ? gy - y0 : max_gy;
    dim3 grid((unsigned)gx, (unsigned)ny);
    const float* p1 = boxes1 + y0 * IOU_TI * box_type;
    float* po = ious + y0 * IOU_TI * n2;
    const int64_t rows = n1 - y0 * IOU_TI;
    if (box_type == 8)
      rotated_iou_kernel<8><<<grid, IOU_TJ, 0, rd::as_stream(stream)>>>(p1, boxes2, po, rows, n2);
    else if (box_type == 5)
      rotated_iou_kernel<5><<<grid, IOU_TJ, 0, rd::as_stream(stream)>>>(p1, boxes2, po, rows, n2);
    else
      rotated_iou_kernel<7><<<grid, IOU_TJ, 0, rd::as_stream(stream)>>>(p1, boxes2, po, rows, n2);
    rd::count_launch();
  }
  return rd::check_launch("rd_rotated_iou");
}

int rd_batch_rotated_iou_max(const float* proposal, const float* gt, float* out, int B, int64_t N,
                             int G, int iou_type, rd_stream_t stream) {
  RD_REQUIRE(iou_type == 0 || iou_type == 1, "rd_batch_rotated_iou_max: iou_type must be 0 (bev) or 1 (3d)");
  RD_REQUIRE(B >= 0 && N >= 0 && G >= 1, "rd_batch_rotated_iou_max: bad sizes B=%d G=%d", B, G);
  if (B == 0 || N == 0) return 0;
  RD_REQUIRE(proposal && gt && out, "rd_batch_rotated_iou_max: null pointer");
  RD_REQUIRE(B <= 65535, "rd_batch_rotated_iou_max: B too large");
  if (rd_check_device()) return 1;
  dim3 grid((unsigned)((N + BMAX_THREADS - 1) / BMAX_THREADS), (unsigned)B);
  if (iou_type == 0)
    batch_iou_max_bev_kernel<<<grid, BMAX_THREADS, 0, rd::as_stream(stream)>>>(proposal, gt, out, N, G);
  else
    batch_iou_max_3d_kernel<<<grid, BMAX_THREADS, 0, rd::as_stream(stream)>>>(proposal, gt, out, N, G);
  rd::count_launch();
  return rd::check_launch("rd_batch_rotated_iou_max");
}

}  // extern "C"
