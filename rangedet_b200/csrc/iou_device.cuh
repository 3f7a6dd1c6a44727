// Rotated-IoU device functions shared by decode_iou.cu (all-pairs / batch-max kernels) and rpn_loss.cu
// (fused decode -> IoU target -> loss).  Reference: operator_cxx/contrib/rotated_iou-inl.h:49-523.
// Translation units including this header MUST be compiled with -fmad=false (see build.py): the
// clipping takes sign decisions on differences of products and follows the reference's host rounding.
#pragma once
#include <math.h>

namespace rd_iou {

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
static __device__ float poly_area(Poly& P) {
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

static __device__ float iou_quads(const Box2& a, const Box2& b) {  // iou_bev_8pts :478-493
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

static __device__ float overlap_rects(const Rect& a, const Rect& b) {
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


}  // namespace rd_iou
