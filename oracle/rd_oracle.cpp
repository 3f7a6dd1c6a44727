// oracle/rd_oracle.cpp -- CPU restatement of the RangeDet post-process ops.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; nothing under
// rangedet_b200/ (the product) may, and the product never falls back to it.
//
// Each function cites the reference file:line (under /root/reference) it follows.
// Parity of this restatement is PINNED against the reference's own C++ compiled
// from source (oracle/_ref/librd_ref.so, see oracle/build_ref.py) by
// tests/test_oracle_pinning.py and by the golden vectors in tests/golden/.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared  (no -ffast-math, no
// -march=native: every float operation is individually rounded, as in the
// reference's x86-64 build, CMakeLists.txt:18-19).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <limits>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Decode3DBbox     operator_cxx/contrib/decode_3d_bbox-inl.h:169-277 (8-dim) and :64-167 (bin)
// ---------------------------------------------------------------------------------------------
struct V2 {
  float x, y;
};

inline V2 rot(V2 p, float s, float c) {  // Point::rotate, decode_3d_bbox-inl.h:48-53
  V2 r;
  r.x = p.x * c - p.y * s;
  r.y = p.x * s + p.y * c;
  return r;
}

inline void write_corners(float* o, float cx, float cy, float length, float width, float s, float c,
                          float z0, float height) {
  // decode_3d_bbox-inl.h:244-274: A(+l/2,-w/2) B(-l/2,-w/2) C(-l/2,+w/2) D(+l/2,+w/2)
  const float hl = 0.5f * length, hw = 0.5f * width;
  V2 A = rot({hl, -hw}, s, c), B = rot({-hl, -hw}, s, c), C = rot({-hl, hw}, s, c),
     D = rot({hl, hw}, s, c);
  o[0] = A.x + cx; o[1] = A.y + cy;
  o[2] = B.x + cx; o[3] = B.y + cy;
  o[4] = C.x + cx; o[5] = C.y + cy;
  o[6] = D.x + cx; o[7] = D.y + cy;
  o[8] = z0;
  o[9] = z0 + height;
}

inline void decode8(const float* d, const float* p, float* o) {
  const float px = p[0], py = p[1];
  const float az = atan2f(py, px);                       // :179
  const float ca = cosf(az), sa = sinf(az);              // :205-206
  const float dx = d[0] * fabsf(d[0]);                   // :212
  const float dy = d[1] * fabsf(d[1]);                   // :213
  const float width = expf(d[2]), length = expf(d[3]), height = expf(d[7]);  // :219-221
  const float dxl = dx * ca - dy * sa;                   // :223
  const float dyl = dx * sa + dy * ca;                   // :224
  const float cx = px + dxl, cy = py + dyl;              // :230-231
  const float yaw = atan2f(d[5], d[4]) + az;             // :234-235  (sin=d[5], cos=d[4])
  write_corners(o, cx, cy, length, width, sinf(yaw), cosf(yaw), d[6], height);
}

inline void decode7bin(const float* d, const float* p, float* o) {
  const float px = p[0], py = p[1], pz = p[2];
  const float az = atan2f(py, px);                                   // :75
  const float ca = cosf(az), sa = sinf(az);                          // :100-101
  const float width = expf(d[3]), length = expf(d[4]), height = expf(d[5]);  // :107-109
  const float dxl = d[0] * ca - d[1] * sa;                           // :111
  const float dyl = d[0] * sa + d[1] * ca;                           // :112
  const float cx = px + dxl, cy = py + dyl, cz = pz + d[2];          // :118-120
  const float z0 = cz - height / 2.0f;                               // :121
  const float yaw = d[6] + az;                                       // :124-125
  write_corners(o, cx, cy, length, width, sinf(yaw), cosf(yaw), z0, height);
}

// ---------------------------------------------------------------------------------------------
// RotatedIOU       operator_cxx/contrib/rotated_iou-inl.h:49-523
// ---------------------------------------------------------------------------------------------
const float R_EPS = 1e-8f;  // :21
inline float mmin(float a, float b) { return a < b ? a : b; }  // MACRO_MIN :23
inline float mmax(float a, float b) { return a > b ? a : b; }  // MACRO_MAX :22
inline float smin(float a, float b) { return b < a ? b : a; }  // std::min
inline float smax(float a, float b) { return a < b ? b : a; }  // std::max

inline bool rel_equal(float a, float b) {  // isEqual :51-53  (NaN for a==b==0 -> false)
  return fabsf((a - b) / mmin(a, b)) < R_EPS;
}

inline int seg_rects_touch(V2 p1, V2 p2, V2 q1, V2 q2) {  // check_rect_cross :69-78
  return mmin(p1.x, p2.x) <= mmax(q1.x, q2.x) && mmin(q1.x, q2.x) <= mmax(p1.x, p2.x) &&
         mmin(p1.y, p2.y) <= mmax(q1.y, q2.y) && mmin(q1.y, q2.y) <= mmax(p1.y, p2.y);
}

inline bool within(float lo, float hi, float v) {  // one axis of `online` :155-158
  return (lo < v || rel_equal(lo, v)) && (hi > v || rel_equal(hi, v));
}

inline int seg_intersection(V2 p1, V2 p0, V2 q1, V2 q0, V2* ans) {  // intersection :131-172
  if (!seg_rects_touch(p0, p1, q0, q1)) return 0;
  const float A1 = p1.y - p0.y, B1 = p0.x - p1.x, C1 = A1 * p0.x + B1 * p0.y;
  const float A2 = q1.y - q0.y, B2 = q0.x - q1.x, C2 = A2 * q0.x + B2 * q0.y;
  const float det = A1 * B2 - A2 * B1;
  if (rel_equal(det, 0.0f)) return 0;  // never true (x/0 or 0/0), kept for fidelity :149
  const float x = (B2 * C1 - B1 * C2) / det;
  const float y = (A1 * C2 - A2 * C1) / det;
  const bool on1 = within(smin(p0.x, p1.x), smax(p0.x, p1.x), x) &&
                   within(smin(p0.y, p1.y), smax(p0.y, p1.y), y);
  const bool on2 = within(smin(q0.x, q1.x), smax(q0.x, q1.x), x) &&
                   within(smin(q0.y, q1.y), smax(q0.y, q1.y), y);
  if (on1 && on2) {
    ans->x = x;
    ans->y = y;
    return 1;
  }
  return 0;
}

inline int in_quad(const float* q, V2 p) {  // check_in_box2d_8pts :113-128
  int flag = -1;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) % 4;
    const float pos = (q[2 * j] - q[2 * i]) * (p.y - q[2 * i + 1]) -
                      (q[2 * j + 1] - q[2 * i + 1]) * (p.x - q[2 * i]);
    const int s = pos >= 0.0f;
    if (flag == -1) flag = s;
    else if (flag != s) return 0;
  }
  return 1;
}

inline int in_rot_rect(float cx, float cy, float w, float h, float ang, V2 p) {
  // check_in_box2d :81-94 (w,h = box[2],box[3]) and check_in_box2d_xyzwlh :97-110 (box[3],box[4])
  const float c = cosf(-ang), s = sinf(-ang);
  const float rx = (p.x - cx) * c + (p.y - cy) * s + cx;
  const float ry = -(p.x - cx) * s + (p.y - cy) * c + cy;
  return rx >= cx - w / 2 && rx <= cx + w / 2 && ry >= cy - h / 2 && ry <= cy + h / 2;
}

// Shared tail of box_overlap_{xywh,xyzwlh,8pts}: centroid, atan2 bubble sort, fan area.
// (:266-288, :363-385, :441-463).  24 slots instead of the reference's 16 (its worst legal
// case, two identical boxes, is exactly 16; beyond that the reference is UB).
struct Poly {
  V2 pt[24];
  int n;
  V2 sum;
};

inline float poly_area(Poly& P) {
  const int cnt = P.n;
  V2 c;
  c.x = P.sum.x / cnt;  // cnt==0 -> NaN centre, loops below do not run
  c.y = P.sum.y / cnt;
  float ang[24];
  for (int i = 0; i < cnt; ++i) ang[i] = atan2f(P.pt[i].y - c.y, P.pt[i].x - c.x);  // point_cmp :187
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        std::swap(ang[i], ang[i + 1]);
        std::swap(P.pt[i], P.pt[i + 1]);
      }
  float area = 0;
  for (int k = 0; k < cnt - 1; ++k) {
    const V2 a = {P.pt[k].x - P.pt[0].x, P.pt[k].y - P.pt[0].y};
    const V2 b = {P.pt[k + 1].x - P.pt[0].x, P.pt[k + 1].y - P.pt[0].y};
    area += a.x * b.y - a.y * b.x;
  }
  return fabsf(area) / 2.0f;
}

inline void poly_push(Poly& P, V2 v) {
  P.sum.x = P.sum.x + v.x;
  P.sum.y = P.sum.y + v.y;
  if (P.n < 24) P.pt[P.n] = v;
  P.n++;
}

inline void edge_crossings(const V2* a, const V2* b, Poly& P) {  // :415-425
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      V2 x;
      if (seg_intersection(a[i + 1], a[i], b[j + 1], b[j], &x)) poly_push(P, x);
    }
}

inline float overlap_8pts(const float* qa, const float* qb) {  // box_overlap_8pts :389-464
  V2 a[5], b[5];
  for (int k = 0; k < 4; ++k) {
    a[k] = {qa[2 * k], qa[2 * k + 1]};
    b[k] = {qb[2 * k], qb[2 * k + 1]};
  }
  a[4] = a[0];
  b[4] = b[0];
  Poly P;
  P.n = 0;
  P.sum = {0, 0};
  edge_crossings(a, b, P);
  for (int k = 0; k < 4; ++k) {  // :428-439
    if (in_quad(qa, b[k])) poly_push(P, b[k]);
    if (in_quad(qb, a[k])) poly_push(P, a[k]);
  }
  return poly_area(P);
}

inline void rot_corner(float cx, float cy, float c, float s, V2* p) {  // rotate_around_center :175-184
  const float nx = (p->x - cx) * c + (p->y - cy) * s + cx;
  const float ny = -(p->x - cx) * s + (p->y - cy) * c + cy;
  p->x = nx;
  p->y = ny;
}

// box_overlap_xywh :195-289 / box_overlap_xyzwlh :292-386.  (cx,cy,w,h,ang) per box.
inline float overlap_rects(float ax, float ay, float aw, float ah, float aa, float bx, float by,
                           float bw, float bh, float ba) {
  V2 a[5], b[5];
  a[0] = {ax - aw / 2, ay - ah / 2}; a[1] = {ax + aw / 2, ay - ah / 2};
  a[2] = {ax + aw / 2, ay + ah / 2}; a[3] = {ax - aw / 2, ay + ah / 2};
  b[0] = {bx - bw / 2, by - bh / 2}; b[1] = {bx + bw / 2, by - bh / 2};
  b[2] = {bx + bw / 2, by + bh / 2}; b[3] = {bx - bw / 2, by + bh / 2};
  const float ac = cosf(aa), as = sinf(aa), bc = cosf(ba), bs = sinf(ba);
  for (int k = 0; k < 4; ++k) {
    rot_corner(ax, ay, ac, as, &a[k]);
    rot_corner(bx, by, bc, bs, &b[k]);
  }
  a[4] = a[0];
  b[4] = b[0];
  Poly P;
  P.n = 0;
  P.sum = {0, 0};
  edge_crossings(a, b, P);
  for (int k = 0; k < 4; ++k) {
    if (in_rot_rect(ax, ay, aw, ah, aa, b[k])) poly_push(P, b[k]);
    if (in_rot_rect(bx, by, bw, bh, ba, a[k])) poly_push(P, a[k]);
  }
  return poly_area(P);
}

inline float quad_area2(const float* q) {  // iou_bev_8pts :482-487
  float s = (q[2] - q[0]) * (q[5] - q[1]) - (q[3] - q[1]) * (q[4] - q[0]);
  s += (q[4] - q[0]) * (q[7] - q[1]) - (q[5] - q[1]) * (q[6] - q[0]);
  return fabsf(s) / 2.0f;
}

inline float pair_iou(const float* a, const float* b, int box_type) {  // Map :510-522
  if (box_type == 8) {  // iou_bev_8pts :478-493
    const float sa = quad_area2(a), sb = quad_area2(b);
    if (sa < R_EPS || sb < R_EPS) return 0.f;
    const float so = overlap_8pts(a, b);
    return so / fmaxf(sa + sb - so, R_EPS);
  } else if (box_type == 5) {  // iou_bev :467-475
    const float sa = a[2] * a[3], sb = b[2] * b[3];
    if (sa < R_EPS || sb < R_EPS) return 0.f;
    const float so = overlap_rects(a[0], a[1], a[2], a[3], a[4], b[0], b[1], b[2], b[3], b[4]);
    return so / fmaxf(sa + sb - so, R_EPS);
  } else {  // 7: iou_3d :496-507
    const float sa = a[3] * a[4] * a[5], sb = b[3] * b[4] * b[5];
    if (sa < R_EPS || sb < R_EPS) return 0.f;
    const float so = overlap_rects(a[0], a[1], a[3], a[4], a[6], b[0], b[1], b[3], b[4], b[6]);
    const float ho = mmax(0.0f, mmin(a[2] + a[5] / 2.0f, b[2] + b[5] / 2.0f) -
                                    mmax(a[2] - a[5] / 2.0f, b[2] - b[5] / 2.0f));
    return so * ho / fmaxf(sa + sb - so * ho, R_EPS);
  }
}

// ---------------------------------------------------------------------------------------------
// wnms_4c          operator_cxx/src_cxx/nms.h:32-250 (OverlapChecker), :252-307 (BBoxHash),
//                  :452-577 (wnms_4c), :781-794 (point4_wnms_4c)
// ---------------------------------------------------------------------------------------------
const float W_EPS = 1e-5f;  // nms.h:35
struct HLine {
  float ax, ay, bx, by, ang;
};

inline int sgn(float k) {  // dblcmp :48-52
  if (fabsf(k) < W_EPS) return 0;
  return k > 0 ? 1 : -1;
}
inline float tri(float x0, float y0, float x1, float y1, float x2, float y2) {  // multi :54-56
  return (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);
}
inline bool line_before(const HLine& l1, const HLine& l2) {  // cmp :58-64
  const int d = sgn(l1.ang - l2.ang);
  if (!d) return sgn(tri(l1.ax, l1.ay, l2.ax, l2.ay, l2.bx, l2.by)) > 0;
  return d < 0;
}
inline void meet(const HLine& l1, const HLine& l2, float* px, float* py) {  // getIntersect :74-83
  const float A1 = l1.by - l1.ay, B1 = l1.ax - l1.bx;
  const float C1 = (l1.bx - l1.ax) * l1.ay - (l1.by - l1.ay) * l1.ax;
  const float A2 = l2.by - l2.ay, B2 = l2.ax - l2.bx;
  const float C2 = (l2.bx - l2.ax) * l2.ay - (l2.by - l2.ay) * l2.ax;
  *px = (C2 * B1 - C1 * B2) / (A1 * B2 - A2 * B1);
  *py = (C1 * A2 - C2 * A1) / (A1 * B2 - A2 * B1);
}
inline bool outside(const HLine& l0, const HLine& l1, const HLine& l2) {  // judge :85-90
  float x, y;
  meet(l1, l2, &x, &y);
  return sgn(tri(x, y, l0.ax, l0.ay, l0.bx, l0.by)) > 0;
}

struct Checker {
  float px[16], py[16];
  HLine l[16];
  int dq[16];

  float fan_area(int s, int e) const {  // getArea :151-166
    if (e - s < 3) return 0;
    float area = 0;
    for (int i = s + 1; i < e - 1; ++i) area += tri(px[s], py[s], px[i], py[i], px[i + 1], py[i + 1]);
    if (area < 0) area = -area;
    return area / 2;
  }
  void load(const float* box, int s) {  // readRec :186-194
    for (int k = 0; k < 4; ++k) {
      px[s + k] = box[2 * k];
      py[s + k] = box[2 * k + 1];
    }
    const bool tag = ((px[s + 1] - px[s]) * (py[s + 2] - py[s]) -
                      (px[s + 2] - px[s]) * (py[s + 1] - py[s])) > 0;  // checkClockwise :92-94
    if (tag) {
      std::swap(px[s], px[s + 3]); std::swap(py[s], py[s + 3]);
      std::swap(px[s + 1], px[s + 2]); std::swap(py[s + 1], py[s + 2]);
    }
  }
  void set_line(int i, int a, int b) {  // addLine :66-72
    l[i].ax = px[a]; l[i].ay = py[a]; l[i].bx = px[b]; l[i].by = py[b];
    l[i].ang = atan2f(py[b] - py[a], px[b] - px[a]);
  }
  // libstdc++ std::sort for n<=16 == __insertion_sort (bits/stl_algo.h); `line_before` is not a
  // strict weak order, so the exact move sequence matters (SURVEY hard-part 4).
  void sort_lines(int n) {
    for (int i = 1; i < n; ++i) {
      const HLine v = l[i];
      if (line_before(v, l[0])) {
        for (int k = i; k > 0; --k) l[k] = l[k - 1];
        l[0] = v;
      } else {
        int k = i;
        while (line_before(v, l[k - 1])) {
          l[k] = l[k - 1];
          --k;
        }
        l[k] = v;
      }
    }
  }
  int half_plane_polygon() {  // HalfPlaneIntersect :96-149, returns pn
    const int n = 8;
    sort_lines(n);
    int i, j;
    for (i = 0, j = 0; i < n; i++)
      if (sgn(l[i].ang - l[j].ang) > 0) l[++j] = l[i];
    const int t = j + 1;
    dq[0] = 0;
    dq[1] = 1;
    int top = 1, bot = 0;
    for (i = 2; i < t; i++) {
      while (top > bot && outside(l[i], l[dq[top]], l[dq[top - 1]])) top--;
      while (top > bot && outside(l[i], l[dq[bot]], l[dq[bot + 1]])) bot++;
      dq[++top] = i;
    }
    while (top > bot && outside(l[dq[bot]], l[dq[top]], l[dq[top - 1]])) top--;
    while (top > bot && outside(l[dq[top]], l[dq[bot]], l[dq[bot + 1]])) bot++;
    dq[++top] = dq[bot];
    int pn = 8;
    for (i = bot; i < top; i++, pn++) meet(l[dq[i + 1]], l[dq[i]], &px[pn], &py[pn]);
    return pn;
  }
  float overlap(const float* box1, const float* box2, bool is3d) {  // single_overlap :195-249
    float h1 = -1, h2 = -1, oh = -1;
    if (is3d) {
      h1 = box1[10];
      h2 = box2[10];
      const float bot1 = box1[9], top1 = bot1 + box1[10];  // getOverlapHeight :172-184
      const float bot2 = box2[9], top2 = bot2 + box2[10];
      const float mt = (top1 > top2) ? top2 : top1;
      const float mb = (bot1 > bot2) ? bot1 : bot2;
      oh = mt - mb;
      if (!(oh > 0)) oh = 0;
    }
    load(box2, 0);
    float area2 = fan_area(0, 4);
    memset(dq, 0, sizeof(dq));
    load(box1, 4);
    for (int z = 0; z < 4; ++z) {
      set_line(z, z, (z + 1) % 4);
      set_line(z + 4, z + 4, (z + 1) % 4 + 4);
    }
    float area1 = fan_area(4, 8);
    const int pn = half_plane_polygon();
    float inter = fan_area(8, pn);
    if (is3d) {
      inter *= oh;
      area1 *= h1;
      area2 *= h2;
    }
    return inter / (area1 + area2 - inter);
  }
};

// BBoxHash::getHash :268-291 -- cell keys i*100+j of the box's AABB (with the reference's
// numeric_limits<float>::min() initialiser for the maxima, :270-273).
void hash_keys(const float* b, float scale, std::vector<int>* keys) {
  keys->clear();
  const float fmin_pos = std::numeric_limits<float>::min(), fmax_v = std::numeric_limits<float>::max();
  float mn0 = fmax_v, mn1 = fmax_v, mx0 = fmin_pos, mx1 = fmin_pos;
  for (int i = 0; i < 4; ++i) {
    mn0 = std::min(mn0, b[2 * i]);
    mn1 = std::min(mn1, b[2 * i + 1]);
    mx0 = std::max(mx0, b[2 * i]);
    mx1 = std::max(mx1, b[2 * i + 1]);
  }
  const int16_t x0 = int16_t(std::floor(mn0 / scale)), y0 = int16_t(std::floor(mn1 / scale));
  const int16_t x1 = int16_t(std::ceil(mx0 / scale)), y1 = int16_t(std::ceil(mx1 / scale));
  for (int i = x0; i < x1; ++i)
    for (int j = y0; j < y1; ++j) keys->push_back(i * 100 + j);
}


// ---------------------------------------------------------------------------------------------
// NMS3D            operator_cxx/contrib/nms_3d.cu:30-534  (GPU-only op in the reference: the CPU
//                  FCompute is LOG(FATAL), nms_3d.cc:11-18 -> nothing to compile here; restated)
// ---------------------------------------------------------------------------------------------
namespace n3 {
const float EPS3 = 1e-8f;  // :27
struct Pt { float x, y; };
inline float cross2(Pt a, Pt b) { return a.x * b.y - a.y * b.x; }                        // :51-53
inline float cross3(Pt p1, Pt p2, Pt p0) {                                                // :62-64
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
inline int rect_cross(Pt p1, Pt p2, Pt q1, Pt q2) {                                       // :66-72
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
inline int in_box(const float* box, Pt P) {                                               // check_in_box3d_anotherway :93-150
  const float MARGIN = -1e-2f;
  const Pt A = {box[0], box[1]}, B = {box[2], box[3]}, C = {box[4], box[5]}, D = {box[6], box[7]};
  const Pt AB = {B.x - A.x, B.y - A.y}, BC = {C.x - B.x, C.y - B.y}, CD = {D.x - C.x, D.y - C.y}, DA = {A.x - D.x, A.y - D.y};
  const float cw = cross2(AB, BC);
  const Pt PA = {A.x - P.x, A.y - P.y};
  if (cross2(PA, AB) * cw < MARGIN) return 0;
  const Pt PB = {B.x - P.x, B.y - P.y};
  if (cross2(PB, BC) * cw < MARGIN) return 0;
  const Pt PC = {C.x - P.x, C.y - P.y};
  if (cross2(PC, CD) * cw < MARGIN) return 0;
  const Pt PD = {D.x - P.x, D.y - P.y};
  if (cross2(PD, DA) * cw < MARGIN) return 0;
  return 1;
}
inline int isect(Pt p1, Pt p0, Pt q1, Pt q0, Pt* ans) {                                   // intersection :152-181
  if (!rect_cross(p0, p1, q0, q1)) return 0;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > EPS3) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float Dd = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / Dd;
    ans->y = (a1 * c0 - a0 * c1) / Dd;
  }
  return 1;
}
inline float area_of(const float* b) {                                                    // get_area :193-198
  const float e1 = (b[0] - b[2]) * (b[0] - b[2]) + (b[1] - b[3]) * (b[1] - b[3]);
  const float e2 = (b[4] - b[2]) * (b[4] - b[2]) + (b[5] - b[3]) * (b[5] - b[3]);
  return sqrtf(e1 * e2);
}
inline float overlap(const float* a, const float* b) {                                    // box_overlap :218-336
  Pt ca[5], cb[5];
  for (int k = 0; k < 4; ++k) { ca[k] = {a[2 * k], a[2 * k + 1]}; cb[k] = {b[2 * k], b[2 * k + 1]}; }
  ca[4] = ca[0]; cb[4] = cb[0];
  Pt pts[24], ctr = {0, 0};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      Pt x;
      if (isect(ca[i + 1], ca[i], cb[j + 1], cb[j], &x)) { ctr.x += x.x; ctr.y += x.y; if (cnt < 24) pts[cnt] = x; cnt++; }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; if (cnt < 24) pts[cnt] = cb[k]; cnt++; }
    if (in_box(b, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; if (cnt < 24) pts[cnt] = ca[k]; cnt++; }
  }
  ctr.x /= cnt; ctr.y /= cnt;
  const int n = cnt < 24 ? cnt : 24;
  float ang[24];
  for (int i = 0; i < n; ++i) ang[i] = atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x);
  for (int j = 0; j < n - 1; ++j)
    for (int i = 0; i < n - j - 1; ++i)
      if (ang[i] > ang[i + 1]) { std::swap(ang[i], ang[i + 1]); std::swap(pts[i], pts[i + 1]); }
  float area = 0;
  for (int k = 0; k < n - 1; ++k) {
    const Pt u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
    area += cross2(u, v);
  }
  return fabsf(area) / 2.0f;
}
inline float iou_bev3d(const float* a, const float* b) {                                  // iou_bev :342-368
  const float ha = a[9] - a[8], hb = b[9] - b[8];
  float oh = fminf(a[9], b[9]) - fmaxf(a[8], b[8]);
  if (oh < 0) oh = 0;
  const float va = area_of(a) * ha, vb = area_of(b) * hb;
  const float vo = overlap(a, b) * oh;
  return vo / fmaxf(va + vb - vo, EPS3);
}
inline float iou_normal(const float* a, const float* b) {                                 // :370-378
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  const float inter = w * h;
  const float Sa = (a[2] - a[0]) * (a[3] - a[1]), Sb = (b[2] - b[0]) * (b[3] - b[1]);
  return inter / fmaxf(Sa + Sb - inter, EPS3);
}
}  // namespace n3

}  // namespace

extern "C" {

// Decode3DBboxForward :279-305 -- output pre-zeroed (:297), one Map per (b,n).
void orc_decode_3d_bbox(const float* delta, const float* pc, float* out, long n_total, int is_bin) {
  memset(out, 0, sizeof(float) * 10 * n_total);
  for (long i = 0; i < n_total; ++i) {
    if (is_bin) decode7bin(delta + 7 * i, pc + 3 * i, out + 10 * i);
    else decode8(delta + 8 * i, pc + 3 * i, out + 10 * i);
  }
}

// RotatedIOUForward rotated_iou-inl.h:525-547 -- (N1,N2) matrix, pre-filled -1 (:543).
void orc_rotated_iou(const float* b1, const float* b2, float* out, long n1, long n2, int box_type) {
  for (long i = 0; i < n1; ++i)
    for (long j = 0; j < n2; ++j) {
      float v = -1.f;
      if (box_type == 5 || box_type == 7 || box_type == 8)
        v = pair_iou(b1 + i * box_type, b2 + j * box_type, box_type);
      out[i * n2 + j] = v;
    }
}

// BatchRotatedIOU.forward / get_iou  operator_py/batch_rotated_iou.py:11-49, to_box_type_7 :51-68.
// proposal (B,N,10), gt (B,G,8) for 'bev' or (B,G,7) for '3d'; out (B,N).
void orc_batch_rotated_iou_max(const float* prop, const float* gt, float* out, int B, long N, int G,
                               int is3d) {
  for (int b = 0; b < B; ++b)
    for (long n = 0; n < N; ++n) {
      const float* p = prop + ((long)b * N + n) * 10;
      float box7[7];
      if (is3d) {  // to_box_type_7 :51-68 (float32 numpy arithmetic), then yaw negated :35
        const float cx = (((p[0] + p[2]) + p[4]) + p[6]) / 4.0f;
        const float cy = (((p[1] + p[3]) + p[5]) + p[7]) / 4.0f;
        const float cz = (p[8] + p[9]) / 2.0f;
        const float l0 = p[0] - p[2], l1 = p[1] - p[3];
        const float w0 = p[2] - p[4], w1 = p[3] - p[5];
        box7[0] = cx; box7[1] = cy; box7[2] = cz;
        box7[3] = sqrtf(l0 * l0 + l1 * l1);
        box7[4] = sqrtf(w0 * w0 + w1 * w1);
        box7[5] = p[9] - p[8];
        box7[6] = -1.0f * atan2f(p[1] - p[3], p[0] - p[2]);
      }
      float best = -INFINITY;
      for (int g = 0; g < G; ++g) {
        float v;
        if (is3d) {
          float g7[7];
          memcpy(g7, gt + ((long)b * G + g) * 7, sizeof(g7));
          g7[6] = -1.0f * g7[6];  // :36
          v = pair_iou(box7, g7, 7);
        } else {
          v = pair_iou(p, gt + ((long)b * G + g) * 8, 8);
        }
        if (isnan(v) || isinf(v) || v > 1.0f || v < 0.0f) v = 0.f;  // :43-46
        if (v > best) best = v;                                       // :47
      }
      out[(long)b * N + n] = best;
    }
}

float orc_single_overlap(const float* box1, const float* box2, int is3d) {
  Checker c;
  memset(&c, 0, sizeof(c));
  return c.overlap(box1, box2, is3d != 0);
}

// point4_wnms_4c :781-794 + wnms_4c :452-577.  Scores must be distinct (unstable std::sort :791).
// out_dets: (K,12) = 11 merged dims + original score; keep_inds index the INPUT array.  Returns K.
int orc_wnms_4c(const float* dets, int n, float thresh, float thresh_vote, int is3d, int hash_scale,
                float* out_dets, int* keep_inds) {
  if (n == 0) return 0;  // :464-466
  const int D = 12;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int i, int j) { return dets[i * D + 11] > dets[j * D + 11]; });

  // BBoxHash::createBBoxMap :255-267
  std::unordered_map<int, std::vector<int>> cell;
  std::vector<std::vector<int>> keys(n);
  for (int i = 0; i < n; ++i) {
    hash_keys(dets + (long)i * D, (float)hash_scale, &keys[i]);
    for (int k : keys[i]) cell[k].push_back(i);
  }
  std::vector<char> suppressed(n, 0);
  std::vector<int> stamp(n, -1);
  std::vector<int> nb;
  std::vector<float> nbyaw;
  Checker chk;
  memset(&chk, 0, sizeof(chk));
  int K = 0;
  for (int _i = 0; _i < n; ++_i) {
    const int i = order[_i];
    if (suppressed[i]) continue;
    nb.clear();
    nb.push_back(i);  // :499
    for (int k : keys[i])  // getFilterResult :292-302
      for (int j : cell[k]) stamp[j] = _i;
    for (int _j = _i + 1; _j < n; ++_j) {
      const int j = order[_j];
      if (suppressed[j]) continue;
      if (stamp[j] != _i) continue;
      const float ovr = chk.overlap(dets + (long)i * D, dets + (long)j * D, is3d != 0);
      if (ovr >= thresh) suppressed[j] = 1;     // :511
      if (ovr > thresh_vote) nb.push_back(j);   // :513
    }
    // median yaw :527-540
    float median;
    const float yaw_i = dets[(long)i * D + 8];
    if (nb.size() <= 2) {
      median = yaw_i;
    } else {
      nbyaw.clear();
      for (int j : nb) nbyaw.push_back(dets[(long)j * D + 8]);
      if (nb.size() % 2 == 0) nbyaw.push_back(yaw_i);
      std::sort(nbyaw.begin(), nbyaw.end());
      median = nbyaw[nbyaw.size() / 2];
    }
    float s1[11], s3[11];
    for (int k = 0; k < 11; ++k) s1[k] = s3[k] = 0.f;
    for (int j : nb) {
      const float dy = fabsf(dets[(long)j * D + 8] - median);
      if ((double)fmodf(dy, float(2 * 3.1415926)) >= 0.3) continue;  // :542
      const float p = dets[(long)j * D + 11];
      for (int k = 0; k < 11; ++k) {  // :555-567
        s1[k] += p * dets[(long)j * D + k];
        s3[k] += p;
      }
    }
    for (int k = 0; k < 11; ++k) out_dets[(long)K * D + k] = s1[k] / s3[k];  // :570-572
    out_dets[(long)K * D + 11] = dets[(long)i * D + 11];                      // :573
    keep_inds[K] = i;
    ++K;
  }
  return K;
}

float orc_nms3d_iou(const float* box_a, const float* box_b, int normal_iou) {
  return normal_iou ? n3::iou_normal(box_a, box_b) : n3::iou_bev3d(box_a, box_b);
}

// NMS3DForward<gpu> nms_3d.cu:470-534 + nms_kernel_3d :380-434 + prepare_output_kernel_3d :436-468.
// boxes (B,N,10) sorted by score; keep_idx (B,max_keep) filled -1; boxes_out (B,max_keep,10) filled 0.
void orc_nms3d(const float* boxes, int B, int N, float thr, int max_keep, int normal_iou, int* keep_idx,
               float* boxes_out) {
  for (long i = 0; i < (long)B * max_keep; ++i) keep_idx[i] = -1;
  memset(boxes_out, 0, sizeof(float) * (size_t)B * max_keep * 10);
  std::vector<char> removed(N);
  for (int b = 0; b < B; ++b) {
    const float* bx = boxes + (long)b * N * 10;
    std::fill(removed.begin(), removed.end(), 0);
    int kept = 0;
    for (int i = 0; i < N; ++i) {
      if (kept >= max_keep) break;
      if (removed[i]) continue;
      for (int k = 0; k < 10; ++k) boxes_out[((long)b * max_keep + kept) * 10 + k] = bx[i * 10 + k];
      keep_idx[(long)b * max_keep + kept] = i;
      ++kept;
      for (int j = i + 1; j < N; ++j) {
        if (removed[j]) continue;  // (the bitmask ORs all rows; already-removed columns stay removed)
        const float v = normal_iou ? n3::iou_normal(bx + i * 10, bx + j * 10) : n3::iou_bev3d(bx + i * 10, bx + j * 10);
        if (v > thr) removed[j] = 1;
      }
    }
  }
}

// ---- training-target assignment ---------------------------------------------------------------------------
// assign3D_v2, operator_cxx/src_cxx/assigner.h:11-87 (Eigen + pybind11; Eigen is not available here, so this is a
// restatement -- "parity unpinned" for this piece).  Distances are SQUARED norms compared with radius / max_dist
// as given (:46-50); the squared norm is accumulated in index order like Eigen's scalar loop over a dynamic row.
void orc_assign3d_v2(const float* pc, const float* bbox, const float* center, const float* radius, const float* mask,
                     const float* nlz, float max_x, float min_x, float max_y, float min_y, float max_z, float min_z,
                     float max_dist, long N, int M, int* result) {
  std::vector<float> dist(M > 0 ? M : 1);
  for (long i = 0; i < N; ++i) {
    result[i] = -1;
    if (mask[i] < 0.5f || nlz[i] > 0.f) continue;
    const float px = pc[i * 3], py = pc[i * 3 + 1], pz = pc[i * 3 + 2];
    if (px < min_x || px > max_x) continue;
    if (py < min_y || py > max_y) continue;
    if (pz < min_z || pz > max_z) continue;
    float min_d = INFINITY;
    for (int j = 0; j < M; ++j) {
      const float d0 = center[j * 3] - px, d1 = center[j * 3 + 1] - py, d2 = center[j * 3 + 2] - pz;
      float acc = d0 * d0;
      acc = acc + d1 * d1;
      acc = acc + d2 * d2;
      dist[j] = acc;
      if (acc < min_d) min_d = acc;
    }
    if (min_d > max_dist) continue;
    for (int j = 0; j < M; ++j) {
      const float* q = bbox + (long)j * 24;
      const float ax = q[0], ay = q[1], az = q[2], bx = q[3], by = q[4], cx = q[6], cy = q[7], dx = q[9], dy = q[10],
                  ez = q[14];
      if (dist[j] > radius[j]) continue;
      if (pz <= az || pz >= ez) continue;
      if (px < ax && px < bx && px < cx && px < dx) continue;
      if (py < ay && py < by && py < cy && py < dy) continue;
      if (px > ax && px > bx && px > cx && px > dx) continue;
      if (py > ay && py > by && py > cy && py > dy) continue;
      const float bpx = px - bx, bpy = py - by;
      if ((ax - bx) * bpx + (ay - by) * bpy <= 0.f) continue;
      if ((cx - bx) * bpx + (cy - by) * bpy <= 0.f) continue;
      const float dpx = px - dx, dpy = py - dy;
      if ((ax - dx) * dpx + (ay - dy) * dpy <= 0.f) continue;
      if ((cx - dx) * dpx + (cy - dy) * dpy <= 0.f) continue;
      result[i] = j;
      break;
    }
  }
}

// get_point_num, assigner.h:89-109: float indices in, float counts out, -1 where the index is negative.
void orc_get_point_num(const float* inds, long N, float* out) {
  const int MAX_BOX_NUM = 500;
  std::vector<float> cnt(MAX_BOX_NUM, 0.f);
  for (long i = 0; i < N; ++i) {
    if (inds[i] < 0) continue;
    const int k = (int)inds[i];
    if (k < MAX_BOX_NUM) cnt[k] += 1;
  }
  for (long i = 0; i < N; ++i) {
    out[i] = -1.f;
    if (inds[i] < 0) continue;
    const int k = (int)inds[i];
    if (k < MAX_BOX_NUM) out[i] = cnt[k];
  }
}

}  // extern "C"
