// Weighted NMS (processing_cxx.wnms_4c) for sm_100a.
//
// Reference: /root/reference operator_cxx/src_cxx/nms.h
//   :781-794 point4_wnms_4c   sort by score (descending), call wnms_4c
//   :452-577 wnms_4c          greedy suppression + neighbourhood voting + weighted box merge
//   :195-249 single_overlap   half-plane-intersection IoU (:96-149) with fuzzy EPS=1e-5 compares
//   :252-307 BBoxHash         100 m hash cells used as a pre-filter
//
// The keep indices must be BIT-EXACT, so the geometry is reproduced operation for operation:
//   * compiled with -fmad=false, IEEE division (every float op individually rounded, like the
//     reference's x86-64 build, CMakeLists.txt:18-19)
//   * atan2f is the fdlibm algorithm glibc 2.39 uses (checked bit-identical on 2e8 inputs)
//   * std::sort on 8 lines == libstdc++ __insertion_sort; the comparator (nms.h:58-64) is not a
//     strict weak order, so the exact move sequence is mirrored
//
// Parallel structure (the greedy scan is inherently sequential in the kept boxes):
//   1. CUB radix sort of scores (descending)                                   [device-wide]
//   2. rank-ordered box copy, AABBs, hash-cell ranges, uniform-grid binning     [device-wide]
//   3. greedy scan: ONE CTA walks the kept boxes in rank order; for each it gathers the
//      unsuppressed later boxes from the 3x3 grid neighbourhood whose AABB overlaps and which share
//      a hash key (an exact superset of the pairs with non-zero IoU), evaluates the IoUs in
//      parallel (one pair per thread), sets suppression bits in a shared-memory bitmap and appends
//      the voting neighbours.  Only rows of KEPT boxes are ever evaluated.
//   4. merge: one thread per kept box (neighbour sort by rank, median yaw, weighted mean).
// Which pairs must be evaluated: the reference evaluates every (kept i, unsuppressed later j) pair
// that shares a hash cell.  Its half-plane intersection is only well behaved when the 8 edge
// directions are distinct: for AABB-disjoint boxes it then returns 0 / NaN (no effect), BUT when
// the two rectangles are parallel within EPS=1e-5 rad (mod pi/2) the duplicate-angle removal
// (nms.h:105-107) leaves 4 lines and the routine returns the area of the GAP between the boxes --
// e.g. ovr = 1.3 for two boxes 23 m apart, which the reference then suppresses.  An exhaustive
// scan of 1.3e9 disjoint pairs found no other source of non-zero overlap.  So the candidate set is
//     (AABB-near  OR  edge direction equal within 1e-4 rad mod pi/2)  AND  shared hash key,
// and the quirk is reproduced bit for bit.
#include <cub/cub.cuh>
#include <float.h>
#include <math.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"

namespace wn {

constexpr int D = 12;
constexpr int NCELL_MAX = 1 << 20;
constexpr int GRID_DIM_MAX = 1024;
constexpr int NT = 512;         // threads of the greedy CTA (113 registers per thread)
constexpr int LCAP = 4 * NT;    // candidate list capacity
constexpr float AABB_MARGIN = 0.05f;
constexpr float HALF_PI = 1.5707963f;
constexpr float THETA_TOL = 1e-4f;  // >> EPS (1e-5) + float noise of the four edge angles

// ---------------------------------------------------------------------------------------------
// glibc-compatible atan2f (fdlibm e_atan2f.c / s_atanf.c)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float atanf_fdlibm(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const int hx = __float_as_int(x);
  const int ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {
    if (ix > 0x7f800000) return x + x;
    return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = __fdiv_rn(2.0f * x - 1.0f, 2.0f + x); }
      else { id = 1; x = __fdiv_rn(x - 1.0f, x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = __fdiv_rn(x - 1.5f, 1.0f + 1.5f * x); }
      else { id = 3; x = __fdiv_rn(-1.0f, x); }
    }
  }
  const float z = x * x, w = z * z;
  const float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f +
                   w * (6.6610731184e-02f + w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
  const float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f +
                   w * (-5.8335702866e-02f + w * -3.6531571299e-02f))));
  if (id < 0) return x - x * (s1 + s2);
  const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return hx < 0 ? -r : r;
}

__device__ float atan2f_glibc(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
              pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  const int hx = __float_as_int(x), hy = __float_as_int(y);
  const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return atanf_fdlibm(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_fdlibm(fabsf(__fdiv_rn(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return __int_as_float(__float_as_int(z) ^ 0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// ---------------------------------------------------------------------------------------------
// OverlapChecker::single_overlap (nms.h:32-250)
// ---------------------------------------------------------------------------------------------
constexpr float W_EPS = 1e-5f;
struct HLine {
  float ax, ay, bx, by, ang;
};
__device__ __forceinline__ int sgn(float k) {
  if (fabsf(k) < W_EPS) return 0;
  return k > 0.f ? 1 : -1;
}
__device__ __forceinline__ float tri(float x0, float y0, float x1, float y1, float x2, float y2) {
  return (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);
}
__device__ __forceinline__ bool line_before(const HLine& l1, const HLine& l2) {
  const int d = sgn(l1.ang - l2.ang);
  if (!d) return sgn(tri(l1.ax, l1.ay, l2.ax, l2.ay, l2.bx, l2.by)) > 0;
  return d < 0;
}
__device__ __forceinline__ void meet(const HLine& l1, const HLine& l2, float* px, float* py) {
  const float A1 = l1.by - l1.ay, B1 = l1.ax - l1.bx;
  const float C1 = (l1.bx - l1.ax) * l1.ay - (l1.by - l1.ay) * l1.ax;
  const float A2 = l2.by - l2.ay, B2 = l2.ax - l2.bx;
  const float C2 = (l2.bx - l2.ax) * l2.ay - (l2.by - l2.ay) * l2.ax;
  *px = __fdiv_rn(C2 * B1 - C1 * B2, A1 * B2 - A2 * B1);
  *py = __fdiv_rn(C1 * A2 - C2 * A1, A1 * B2 - A2 * B1);
}
__device__ __forceinline__ bool outside(const HLine& l0, const HLine& l1, const HLine& l2) {
  float x, y;
  meet(l1, l2, &x, &y);
  return sgn(tri(x, y, l0.ax, l0.ay, l0.bx, l0.by)) > 0;
}

struct Checker {
  float px[16], py[16];
  HLine l[8];
  int dq[16];

  __device__ float fan_area(int s, int e) const {
    if (e - s < 3) return 0.f;
    float area = 0.f;
    for (int i = s + 1; i < e - 1; ++i) area += tri(px[s], py[s], px[i], py[i], px[i + 1], py[i + 1]);
    if (area < 0.f) area = -area;
    return area / 2;
  }
  __device__ void load(const float* box, int s) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      px[s + k] = box[2 * k];
      py[s + k] = box[2 * k + 1];
    }
    const bool tag = ((px[s + 1] - px[s]) * (py[s + 2] - py[s]) - (px[s + 2] - px[s]) * (py[s + 1] - py[s])) > 0.f;
    if (tag) {
      float t;
      t = px[s]; px[s] = px[s + 3]; px[s + 3] = t;
      t = py[s]; py[s] = py[s + 3]; py[s + 3] = t;
      t = px[s + 1]; px[s + 1] = px[s + 2]; px[s + 2] = t;
      t = py[s + 1]; py[s + 1] = py[s + 2]; py[s + 2] = t;
    }
  }
  __device__ void set_line(int i, int a, int b) {
    l[i].ax = px[a]; l[i].ay = py[a]; l[i].bx = px[b]; l[i].by = py[b];
    l[i].ang = atan2f_glibc(py[b] - py[a], px[b] - px[a]);
  }
  __device__ void sort_lines() {  // libstdc++ __insertion_sort over 8 elements
    for (int i = 1; i < 8; ++i) {
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
  __device__ int half_plane_polygon() {
    const int n = 8;
    sort_lines();
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
  __device__ float overlap(const float* box1, const float* box2, bool is3d) {
    float h1 = -1.f, h2 = -1.f, oh = -1.f;
    if (is3d) {
      h1 = box1[10];
      h2 = box2[10];
      const float bot1 = box1[9], top1 = bot1 + box1[10];
      const float bot2 = box2[9], top2 = bot2 + box2[10];
      const float mt = (top1 > top2) ? top2 : top1;
      const float mb = (bot1 > bot2) ? bot1 : bot2;
      oh = mt - mb;
      if (!(oh > 0.f)) oh = 0.f;
    }
    load(box2, 0);
    float area2 = fan_area(0, 4);
#pragma unroll
    for (int k = 0; k < 16; ++k) dq[k] = 0;
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
    return __fdiv_rn(inter, area1 + area2 - inter);
  }
};

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
struct GridParams {
  float ox, oy, inv_s;
  int nx, ny;
};
struct Counters {
  int kept;        // K
  int nb_total;    // neighbour entries
  int nb_overflow;
  unsigned min_x, min_y, max_x, max_y, max_ext;  // order-preserving encodings of floats
  int cand_overflow;   // the candidate lists of the parallel path do not fit their reservation: use the sequential scan
};

struct Layout {
  size_t keys_in, keys_out, idx_in, order, bx, aabb, hashr, cell_id, cell_id_s, rank_in, cell_rank,
      cell_start, cell_end, params, counters, kept_rank, nb_start, nb_list, theta, theta_s, theta_rank,
      aabb_c, hash_c, th_range, cand_cnt, cand_off, cand_j, cand_flag, cub_temp, total;
  size_t cub_bytes, nb_cap, cand_cap;
};

__host__ inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline Layout make_layout(int n) {
  Layout L;
  size_t o = 0;
  const size_t N = (size_t)(n > 0 ? n : 1);
  auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
  L.keys_in = take(N * 4); L.keys_out = take(N * 4); L.idx_in = take(N * 4); L.order = take(N * 4);
  L.bx = take(N * D * 4); L.aabb = take(N * 16); L.hashr = take(N * 8);
  L.cell_id = take(N * 4); L.cell_id_s = take(N * 4); L.rank_in = take(N * 4); L.cell_rank = take(N * 4);
  L.cell_start = take((size_t)(NCELL_MAX + 1) * 4); L.cell_end = take((size_t)(NCELL_MAX + 1) * 4);
  L.params = take(sizeof(GridParams)); L.counters = take(sizeof(Counters));
  L.kept_rank = take(N * 4); L.nb_start = take((N + 1) * 4);
  L.nb_cap = N + 4096;
  L.nb_list = take(L.nb_cap * 4);
  L.theta = take(N * 4); L.theta_s = take(N * 4); L.theta_rank = take(N * 4);
  L.aabb_c = take(N * 16); L.hash_c = take(N * 8);   // AABB / hash range in CELL order (aligned with cell_rank)
  L.th_range = take(N * 6 * 4);                      // per box: [first, last) of its three theta segments
  // parallel path: candidate pairs (i, j > i) of ALL boxes, 320 per box on average (cfg-3: 100 k boxes on 150 m x 150 m
  // have ~220 later boxes with overlapping AABBs each); beyond that the sequential scan runs.  5 bytes per pair.
  L.cand_cap = N * 320 + 65536;
  L.cand_cnt = take((N + 1) * 4); L.cand_off = take((N + 1) * 4);
  L.cand_j = take(L.cand_cap * 4); L.cand_flag = take(L.cand_cap);
  L.cub_bytes = (size_t)(16u << 20) + N * 32;
  L.cub_temp = take(L.cub_bytes);
  L.total = o;
  return L;
}

// order-preserving float <-> unsigned
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void init_kernel(const float* __restrict__ dets, int n, float* keys, int* idx, Counters* ctr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    ctr->kept = 0; ctr->nb_total = 0; ctr->nb_overflow = 0; ctr->cand_overflow = 0;
    ctr->min_x = 0xffffffffu; ctr->min_y = 0xffffffffu; ctr->max_x = 0u; ctr->max_y = 0u; ctr->max_ext = 0u;
  }
  if (i < n) {
    keys[i] = dets[(size_t)i * D + 11];
    idx[i] = i;
  }
}

// rank-ordered copies + AABB + BBoxHash cell ranges (nms.h:268-291) + global bounds
__global__ void prep_kernel(const float* __restrict__ dets, const int* __restrict__ order, int n,
                            float hash_scale, float* __restrict__ bx, float4* __restrict__ aabb,
                            short4* __restrict__ hashr, float* __restrict__ theta, Counters* ctr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int i = order[r];
  float v[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    v[k] = dets[(size_t)i * D + k];
    bx[(size_t)r * D + k] = v[k];
  }
  float mnx = FLT_MAX, mny = FLT_MAX, mxx = FLT_MIN, mxy = FLT_MIN;  // reference initialisers (:270-273)
  float tx0 = v[0], tx1 = v[0], ty0 = v[1], ty1 = v[1];              // true AABB
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float x = v[2 * k], y = v[2 * k + 1];
    mnx = x < mnx ? x : mnx;  // std::min(a,b) = b<a ? b : a
    mny = y < mny ? y : mny;
    mxx = mxx < x ? x : mxx;  // std::max(a,b) = a<b ? b : a
    mxy = mxy < y ? y : mxy;
    tx0 = fminf(tx0, x); tx1 = fmaxf(tx1, x);
    ty0 = fminf(ty0, y); ty1 = fmaxf(ty1, y);
  }
  short4 hr;
  hr.x = (short)__float2int_rz(floorf(__fdiv_rn(mnx, hash_scale)));
  hr.y = (short)__float2int_rz(floorf(__fdiv_rn(mny, hash_scale)));
  hr.z = (short)__float2int_rz(ceilf(__fdiv_rn(mxx, hash_scale)));
  hr.w = (short)__float2int_rz(ceilf(__fdiv_rn(mxy, hash_scale)));
  hashr[r] = hr;
  {  // direction of the first edge, folded into [0, pi/2): parallel-rectangle detector
    float a = atan2f_glibc(v[3] - v[1], v[2] - v[0]);
    a = a - floorf(a * 0.63661977f) * HALF_PI;
    if (!(a >= 0.f)) a = 0.f;          // also catches NaN
    if (a >= HALF_PI) a = 0.f;
    theta[r] = a;
  }
  aabb[r] = make_float4(tx0 - AABB_MARGIN, ty0 - AABB_MARGIN, tx1 + AABB_MARGIN, ty1 + AABB_MARGIN);
  if (isfinite(tx0) && isfinite(tx1) && isfinite(ty0) && isfinite(ty1)) {
    atomicMin(&ctr->min_x, f2ord(tx0));
    atomicMin(&ctr->min_y, f2ord(ty0));
    atomicMax(&ctr->max_x, f2ord(tx0));
    atomicMax(&ctr->max_y, f2ord(ty0));
    const float ext = fmaxf(tx1 - tx0, ty1 - ty0) + 2.0f * AABB_MARGIN;
    atomicMax(&ctr->max_ext, f2ord(ext));
  }
}

__global__ void grid_setup_kernel(const Counters* ctr, GridParams* gp) {
  float ox = ord2f(ctr->min_x), oy = ord2f(ctr->min_y);
  float ex = ord2f(ctr->max_x), ey = ord2f(ctr->max_y);
  float s = ord2f(ctr->max_ext);
  if (!(ctr->min_x <= ctr->max_x)) { ox = oy = 0.f; ex = ey = 0.f; s = 1.f; }  // no finite box
  if (!(s > 1e-3f)) s = 1e-3f;
  const float rx = ex - ox, ry = ey - oy;
  // cell size >= max AABB extent so every AABB-overlapping pair lies within a 3x3 neighbourhood
  if (rx / s > (float)(GRID_DIM_MAX - 2)) s = rx / (float)(GRID_DIM_MAX - 2);
  if (ry / s > (float)(GRID_DIM_MAX - 2)) s = ry / (float)(GRID_DIM_MAX - 2);
  gp->ox = ox; gp->oy = oy; gp->inv_s = 1.0f / s;
  gp->nx = min(GRID_DIM_MAX, (int)floorf(rx / s) + 1);
  gp->ny = min(GRID_DIM_MAX, (int)floorf(ry / s) + 1);
}

__device__ __forceinline__ void cell_of(const GridParams& gp, float x0, float y0, int* cx, int* cy) {
  int ix = (int)floorf((x0 + AABB_MARGIN - gp.ox) * gp.inv_s);
  int iy = (int)floorf((y0 + AABB_MARGIN - gp.oy) * gp.inv_s);
  if (!(ix >= 0)) ix = 0;  // also catches NaN
  if (!(iy >= 0)) iy = 0;
  *cx = ix >= gp.nx ? gp.nx - 1 : ix;
  *cy = iy >= gp.ny ? gp.ny - 1 : iy;
}

__global__ void cell_kernel(const float4* __restrict__ aabb, int n, const GridParams* gpp,
                            int* __restrict__ cell_id, int* __restrict__ rank_in) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const GridParams gp = *gpp;
  int cx, cy;
  cell_of(gp, aabb[r].x, aabb[r].y, &cx, &cy);
  cell_id[r] = cy * gp.nx + cx;
  rank_in[r] = r;
}

__global__ void cell_bounds_kernel(const int* __restrict__ cell_id_s, int n, int* cell_start, int* cell_end) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int c = cell_id_s[e];
  if (e == 0 || cell_id_s[e - 1] != c) cell_start[c] = e;
  if (e == n - 1 || cell_id_s[e + 1] != c) cell_end[c] = e + 1;
}

// do boxes a and b share a BBoxHash key i*100+j ?  (getFilterResult nms.h:292-302; keys can alias)
__device__ bool share_key(short4 a, short4 b) {
  // fast path: plain 2-D range overlap implies a shared key
  if (a.x < b.z && b.x < a.z && a.y < b.w && b.y < a.w) return true;
  // aliasing (i*100+j == i'*100+j') is only possible when a j-range spans >= 100 or leaves (-100,100)
  const int wa = a.w - a.y, wb = b.w - b.y;
  if (wa <= 0 || wb <= 0 || a.z <= a.x || b.z <= b.x) return false;
  for (int i = a.x; i < a.z; ++i)
    for (int j = a.y; j < a.w; ++j) {
      const int key = i * 100 + j;
      for (int i2 = b.x; i2 < b.z; ++i2) {
        const int j2 = key - i2 * 100;
        if (j2 >= b.y && j2 < b.w) return true;
      }
    }
  return false;
}

// Cell-ordered copies of what the candidate test reads, so that the greedy scan issues independent loads
// (cell_rank[e], aabb_c[e], hash_c[e]) instead of a dependent chain through the rank.
__global__ void cell_gather_kernel(const int* __restrict__ cell_rank, const float4* __restrict__ aabb,
                                   const short4* __restrict__ hashr, int n, float4* __restrict__ aabb_c,
                                   short4* __restrict__ hash_c) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int r = cell_rank[e];
  aabb_c[e] = aabb[r];
  hash_c[e] = hashr[r];
}

// The ranges of the theta-sorted array that hold the boxes parallel to box r (its own direction +- THETA_TOL, with
// the wrap at 0 / pi/2): six binary searches per box, done here for ALL boxes in parallel instead of by the single
// greedy CTA for every kept box (34-100 dependent loads on its critical path).  Empty segment: first == last.
__global__ void theta_range_kernel(const float* __restrict__ theta, const float* __restrict__ theta_s, int n,
                                   int* __restrict__ th_range) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float th = theta[r];
  for (int seg = 0; seg < 3; ++seg) {
    float lo, hi;
    bool on = true;
    if (seg == 0) { lo = th - THETA_TOL; hi = th + THETA_TOL; }
    else if (seg == 1) { lo = th - THETA_TOL + HALF_PI; hi = HALF_PI + 1.f; on = th < THETA_TOL; }
    else { lo = -1.f; hi = th + THETA_TOL - HALF_PI; on = th + THETA_TOL > HALF_PI; }
    int first = 0, last = 0;
    if (on) {
      int a = 0, b = n;  // lower_bound(lo)
      while (a < b) { const int mid = (a + b) >> 1; if (theta_s[mid] < lo) a = mid + 1; else b = mid; }
      first = a;
      b = n;             // upper_bound(hi)
      while (a < b) { const int mid = (a + b) >> 1; if (theta_s[mid] <= hi) a = mid + 1; else b = mid; }
      last = a;
    }
    th_range[(size_t)r * 6 + 2 * seg] = first;
    th_range[(size_t)r * 6 + 2 * seg + 1] = last;
  }
}

constexpr int NSEG = 12;   // candidate segments of one kept box: 9 grid cells + 3 theta ranges

// ---------------------------------------------------------------------------------------------------------------
// Parallel path.  The reference's loop (nms.h:497-517) evaluates box i against later boxes only when i is kept, one i
// after the other.  Which later boxes i WOULD suppress / take as voting neighbours does not depend on that order, so
// the expensive part -- the half-plane-intersection IoU of every candidate pair -- is computed for ALL boxes at once,
// one warp per box; what stays sequential is a walk over the kept boxes that only flips bitmap bits:
//   cand_count_kernel   candidates of box i = later boxes in its 12 segments that pass the AABB / hash tests
//   (exclusive scan)    -> cand_off
//   cand_eval_kernel    compact them (rank order within a segment), evaluate overlap(box_i, box_j): flag bit 0
//                       ovr >= thresh (suppress, nms.h:511), bit 1 ovr > thresh_vote (neighbour, :513)
//   resolve_kernel      one CTA: next unsuppressed rank i is kept; for its candidates j still unsuppressed: set the
//                       suppress bit / append the neighbour -- exactly the state the reference sees when it reaches i
// Same candidate set, same overlap(box_i, box_j) argument order, same suppression semantics as the sequential scan:
// keep indices and neighbour sets are identical (the merge kernel restores rank order of the neighbours).
// ---------------------------------------------------------------------------------------------------------------
struct SegInfo {
  int start[NSEG];
  int off[NSEG + 1];
};

// segments of box ri, filled by lanes 0..11 of the calling warp and broadcast through shuffles
__device__ __forceinline__ void warp_segments(int ri, const GridParams& gp, const float4& ai, const int* __restrict__ cell_start,
                                              const int* __restrict__ cell_end, const int* __restrict__ th_range, int lane,
                                              int& my_start, int& my_len) {
  int cx, cy;
  cell_of(gp, ai.x, ai.y, &cx, &cy);
  int s0 = 0, len = 0;
  if (lane < 9) {
    const int yy = cy + lane / 3 - 1, xx = cx + lane % 3 - 1;
    if (yy >= 0 && yy < gp.ny && xx >= 0 && xx < gp.nx) {
      const int c = yy * gp.nx + xx;
      s0 = cell_start[c];
      len = cell_end[c] - s0;
    }
  } else if (lane < NSEG) {
    s0 = th_range[(size_t)ri * 6 + 2 * (lane - 9)];
    len = th_range[(size_t)ri * 6 + 2 * (lane - 9) + 1] - s0;
  }
  my_start = s0;
  my_len = len > 0 ? len : 0;
}

// candidate test of entry e of segment q for box ri; returns the candidate's rank or -1
__device__ __forceinline__ int cand_test(int q, int e, int ri, const float4& ai, const short4& hi_key,
                                         const int* __restrict__ cell_rank, const float4* __restrict__ aabb_c,
                                         const short4* __restrict__ hash_c, const int* __restrict__ theta_rank,
                                         const float4* __restrict__ aabb, const short4* __restrict__ hashr) {
  if (q < 9) {
    const int rb = cell_rank[e];
    if (rb <= ri) return -1;
    const float4 ab = aabb_c[e];
    const bool near = !(ai.z < ab.x || ab.z < ai.x || ai.w < ab.y || ab.w < ai.y);
    return (near && share_key(hi_key, hash_c[e])) ? rb : -1;
  }
  const int rb = theta_rank[e];
  if (rb <= ri) return -1;
  const float4 ab = aabb[rb];
  const bool near = !(ai.z < ab.x || ab.z < ai.x || ai.w < ab.y || ab.w < ai.y);
  return (!near && share_key(hi_key, hashr[rb])) ? rb : -1;
}

constexpr int CAND_WARPS = 8;   // boxes per CTA of the two candidate kernels

template <bool EVAL>
__global__ void __launch_bounds__(CAND_WARPS * 32)
cand_kernel(const float* __restrict__ bx, const float4* __restrict__ aabb, const short4* __restrict__ hashr,
            const int* __restrict__ cell_rank, const int* __restrict__ cell_start, const int* __restrict__ cell_end,
            const float4* __restrict__ aabb_c, const short4* __restrict__ hash_c, const int* __restrict__ th_range,
            const int* __restrict__ theta_rank, const GridParams* gpp, int n, float thresh, float thresh_vote, int is3d,
            int* __restrict__ cand_cnt, const int* __restrict__ cand_off, int* __restrict__ cand_j,
            unsigned char* __restrict__ cand_flag, int cand_cap, Counters* ctr) {
  const int lane = threadIdx.x & 31;
  const int ri = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
  if (ri >= n) return;
  const GridParams gp = *gpp;
  const float4 ai = aabb[ri];
  const short4 hi_key = hashr[ri];
  int my_start, my_len;
  warp_segments(ri, gp, ai, cell_start, cell_end, th_range, lane, my_start, my_len);
  int base = 0;
  if (EVAL) {
    base = cand_off[ri];
    if (ctr->cand_overflow || cand_off[n] > cand_cap) {   // uniform: the host falls back to the sequential scan
      if (ri == 0 && lane == 0) ctr->cand_overflow = 1;
      return;
    }
  }
  int count = 0;
  for (int q = 0; q < NSEG; ++q) {
    const int s0 = __shfl_sync(0xffffffffu, my_start, q), len = __shfl_sync(0xffffffffu, my_len, q);
    for (int b0 = 0; b0 < len; b0 += 32) {
      const int e = b0 + lane;
      const int rb = e < len ? cand_test(q, s0 + e, ri, ai, hi_key, cell_rank, aabb_c, hash_c, theta_rank, aabb, hashr) : -1;
      const unsigned hit = __ballot_sync(0xffffffffu, rb >= 0);
      if (EVAL && rb >= 0) cand_j[base + count + __popc(hit & ((1u << lane) - 1u))] = rb;
      count += __popc(hit);
    }
  }
  if (!EVAL) {
    if (lane == 0) cand_cnt[ri] = count;
    return;
  }
  __syncwarp();
  float bi[D];
#pragma unroll
  for (int k = 0; k < D; ++k) bi[k] = bx[(size_t)ri * D + k];
  for (int k = lane; k < count; k += 32) {   // one pair per lane
    const int rb = cand_j[base + k];
    float bj[D];
#pragma unroll
    for (int u = 0; u < D; ++u) bj[u] = bx[(size_t)rb * D + u];
    Checker chk;
    const float ovr = chk.overlap(bi, bj, is3d != 0);
    cand_flag[base + k] = (unsigned char)((ovr >= thresh ? 1 : 0) | (ovr > thresh_vote ? 2 : 0));
  }
}

// ONE CTA.  Dynamic shared memory: suppression bitmap.  Per kept box: two loads for its candidate range, one coalesced
// sweep over (rank, flags) -- no geometry on this critical path any more.
__global__ void __launch_bounds__(512, 1)
resolve_kernel(const int* __restrict__ cand_off, const int* __restrict__ cand_j, const unsigned char* __restrict__ cand_flag,
               int n, int* __restrict__ kept_rank, int* __restrict__ nb_start, int* __restrict__ nb_list, int nb_cap,
               Counters* ctr) {
  extern __shared__ unsigned bitmap[];
  __shared__ int s_cur, s_nb;
  const int t = threadIdx.x;
  const int nwords = (n + 31) >> 5;
  if (ctr->cand_overflow) return;
  for (int w = t; w < nwords; w += blockDim.x) bitmap[w] = 0u;
  if (t == 0) { s_cur = 0; s_nb = 0; }
  int K = 0;
  __syncthreads();
  while (true) {
    const int ri = s_cur;
    if (ri < 0 || ri >= n) break;
    const int nb0 = s_nb;
    if (t == 0) {
      kept_rank[K] = ri;
      nb_start[K] = nb0 < nb_cap ? nb0 : nb_cap;
    }
    const int c0 = cand_off[ri], c1 = cand_off[ri + 1];
    __syncthreads();   // everyone has read s_cur / s_nb
    for (int k = c0 + t; k < c1; k += blockDim.x) {
      const unsigned f = cand_flag[k];
      if (f == 0u) continue;
      const int rb = cand_j[k];
      // state of j when the reference reaches box i: bits set by EARLIER kept boxes (a candidate occurs once per box, so
      // the bits set in this sweep belong to other candidates and are never read here)
      if ((bitmap[rb >> 5] >> (rb & 31)) & 1u) continue;                 // suppressed before: skipped (nms.h:503)
      if (f & 1u) atomicOr(&bitmap[rb >> 5], 1u << (rb & 31));           // nms.h:511
      if (f & 2u) {                                                      // nms.h:513
        const int pos = atomicAdd(&s_nb, 1);
        if (pos < nb_cap) nb_list[pos] = rb;
        else ctr->nb_overflow = 1;
      }
    }
    __syncthreads();
    ++K;
    if (t < 32) {   // next unsuppressed rank > ri
      int found = -1;
      const int w0 = (ri + 1) >> 5;
      const unsigned first_mask = ~0u << ((ri + 1) & 31);
      for (int wb = w0; wb < nwords && found < 0; wb += 32) {
        const int w = wb + t;
        unsigned free_bits = 0u;
        if (w < nwords) {
          free_bits = ~bitmap[w];
          if (w == w0) free_bits &= first_mask;
          if (w == nwords - 1 && (n & 31)) free_bits &= (1u << (n & 31)) - 1u;
        }
        if (ri + 1 >= n) free_bits = 0u;
        const unsigned ball = __ballot_sync(0xffffffffu, free_bits != 0u);
        if (ball) {
          const int src = __ffs(ball) - 1;
          const unsigned fb = __shfl_sync(0xffffffffu, free_bits, src);
          found = ((wb + src) << 5) + (__ffs(fb) - 1);
        }
      }
      if (t == 0) s_cur = found;
    }
    __syncthreads();
  }
  if (t == 0) {
    nb_start[K] = s_nb < nb_cap ? s_nb : nb_cap;
    ctr->kept = K;
    ctr->nb_total = s_nb;
  }
}

struct GreedySmem {
  int seg_start[NSEG];     // first entry of each segment (cell-sorted resp. theta-sorted array)
  int seg_off[NSEG + 1];   // exclusive prefix of the segment lengths: one flat index space for the gather
  float box_i[D];
  float4 aabb_i;
  short4 hash_i;
  int cand[LCAP];
  int ncand;
  int cur;       // rank being processed (-1: done)
};

// ONE CTA.  Dynamic shared memory: suppression bitmap, ceil(n/32) words.
__global__ void __launch_bounds__(NT, 1)
greedy_kernel(const float* __restrict__ bx, const float4* __restrict__ aabb, const short4* __restrict__ hashr,
              const int* __restrict__ cell_rank, const int* __restrict__ cell_start,
              const int* __restrict__ cell_end, const float4* __restrict__ aabb_c,
              const short4* __restrict__ hash_c, const int* __restrict__ th_range,
              const int* __restrict__ theta_rank,
              const GridParams* gpp, int n, float thresh,
              float thresh_vote, int is3d, int* __restrict__ kept_rank, int* __restrict__ nb_start,
              int* __restrict__ nb_list, int nb_cap, Counters* ctr) {
  extern __shared__ unsigned bitmap[];
  __shared__ GreedySmem S;
  const int t = threadIdx.x;
  const int nwords = (n + 31) >> 5;
  for (int w = t; w < nwords; w += NT) bitmap[w] = 0u;
  if (t == 0) { S.cur = 0; S.ncand = 0; }
  const GridParams gp = *gpp;
  int K = 0, nb_total = 0;  // meaningful in thread 0
  __syncthreads();

  auto evaluate = [&]() {  // IoU of box_i against every gathered candidate, one pair per thread
    const int nc = S.ncand;
    for (int e = t; e < nc; e += NT) {
      const int rb = S.cand[e];
      float bj[D];
#pragma unroll
      for (int k = 0; k < D; ++k) bj[k] = bx[(size_t)rb * D + k];
      Checker chk;
      const float ovr = chk.overlap(S.box_i, bj, is3d != 0);
      if (ovr >= thresh) atomicOr(&bitmap[rb >> 5], 1u << (rb & 31));   // nms.h:511
      if (ovr > thresh_vote) {                                           // nms.h:513
        const int pos = atomicAdd(&ctr->nb_total, 1);
        if (pos < nb_cap) nb_list[pos] = rb;
        else ctr->nb_overflow = 1;
      }
    }
  };

  while (true) {
    const int ri = S.cur;
    if (ri < 0 || ri >= n) break;
    // stage box i
    if (t >= 32 && t < 32 + D) S.box_i[t - 32] = bx[(size_t)ri * D + (t - 32)];
    if (t == 64) S.aabb_i = aabb[ri];
    if (t == 65) S.hash_i = hashr[ri];
    if (t == 0) {  // K / nb_total live in thread 0 only
      kept_rank[K] = ri;
      nb_start[K] = nb_total;
    }
    __syncthreads();
    const float4 ai = S.aabb_i;
    const short4 hi_key = S.hash_i;
    int cx, cy;
    cell_of(gp, ai.x, ai.y, &cx, &cy);
    // The candidates of box i live in up to 12 segments: the 3x3 grid cells around it (entries in cell order) and the
    // three precomputed ranges of the theta-sorted array (parallel boxes at any distance).  Their bounds are fetched
    // by 12 threads at once and laid end to end, so the gather below is ONE flat loop whose loads are independent.
    if (t < NSEG) {
      int s0 = 0, len = 0;
      if (t < 9) {
        const int yy = cy + t / 3 - 1, xx = cx + t % 3 - 1;
        if (yy >= 0 && yy < gp.ny && xx >= 0 && xx < gp.nx) {
          const int c = yy * gp.nx + xx;
          s0 = cell_start[c];
          len = cell_end[c] - s0;
        }
      } else {
        s0 = th_range[(size_t)ri * 6 + 2 * (t - 9)];
        len = th_range[(size_t)ri * 6 + 2 * (t - 9) + 1] - s0;
      }
      S.seg_start[t] = s0;
      S.seg_off[t + 1] = len > 0 ? len : 0;
    }
    __syncthreads();
    if (t == 0) {
      S.seg_off[0] = 0;
      for (int q = 0; q < NSEG; ++q) S.seg_off[q + 1] += S.seg_off[q];
    }
    __syncthreads();
    const int total = S.seg_off[NSEG];
    for (int base = 0; base < total; base += NT) {
      const int f = base + t;
      if (f < total) {
        int q = 0;
#pragma unroll
        for (int u = 1; u < NSEG; ++u) q += (f >= S.seg_off[u]) ? 1 : 0;   // segment of flat index f
        const int e = S.seg_start[q] + (f - S.seg_off[q]);
        if (q < 9) {   // grid cells: near boxes
          const int rb = cell_rank[e];
          const float4 ab = aabb_c[e];
          const short4 hb = hash_c[e];
          if (rb > ri && !((bitmap[rb >> 5] >> (rb & 31)) & 1u)) {
            const bool near = !(ai.z < ab.x || ab.z < ai.x || ai.w < ab.y || ab.w < ai.y);
            if (near && share_key(hi_key, hb)) {
              const int pos = atomicAdd(&S.ncand, 1);
              S.cand[pos] = rb;
            }
          }
        } else {       // theta ranges: parallel boxes that are NOT near (the near ones came through the grid)
          const int rb = theta_rank[e];
          if (rb > ri && !((bitmap[rb >> 5] >> (rb & 31)) & 1u)) {
            const float4 ab = aabb[rb];
            const bool near = !(ai.z < ab.x || ab.z < ai.x || ai.w < ab.y || ab.w < ai.y);
            if (!near && share_key(hi_key, hashr[rb])) {
              const int pos = atomicAdd(&S.ncand, 1);
              S.cand[pos] = rb;
            }
          }
        }
      }
      __syncthreads();
      if (S.ncand > LCAP - NT) {  // uniform decision
        evaluate();
        __syncthreads();
        if (t == 0) S.ncand = 0;
        __syncthreads();
      }
    }
    __syncthreads();
    evaluate();
    __syncthreads();
    if (t == 0) {
      S.ncand = 0;
      K++;
      nb_total = atomicAdd(&ctr->nb_total, 0);  // all appends of this box are complete (barrier above)
    }
    // next unsuppressed rank > ri  (warp 0)
    if (t < 32) {
      int found = -1;
      int w0 = (ri + 1) >> 5;
      const unsigned first_mask = ~0u << ((ri + 1) & 31);
      for (int wb = w0; wb < nwords && found < 0; wb += 32) {
        const int w = wb + t;
        unsigned free_bits = 0u;
        if (w < nwords) {
          free_bits = ~bitmap[w];
          if (w == w0) free_bits &= first_mask;
          if (w == nwords - 1 && (n & 31)) free_bits &= (1u << (n & 31)) - 1u;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, free_bits != 0u);
        if (ball) {
          const int src = __ffs(ball) - 1;
          const unsigned fb = __shfl_sync(0xffffffffu, free_bits, src);
          found = ((wb + src) << 5) + (__ffs(fb) - 1);
        }
      }
      if (t == 0) S.cur = found;
    }
    __syncthreads();
  }
  if (t == 0) {
    nb_start[K] = nb_total < nb_cap ? nb_total : nb_cap;
    ctr->kept = K;
  }
}

// one thread per kept box: voting + weighted merge (nms.h:518-574)
__global__ void merge_kernel(const float* __restrict__ bx, const int* __restrict__ order,
                             const int* __restrict__ kept_rank, const int* __restrict__ nb_start,
                             int* __restrict__ nb_list, const Counters* ctr, float* __restrict__ out_dets,
                             int* __restrict__ keep_inds) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ctr->kept) return;
  const int ri = kept_rank[k];
  const int s = nb_start[k], e = nb_start[k + 1];
  // neighbours were appended in arbitrary order within the box's segment: restore rank order
  for (int a = s + 1; a < e; ++a) {
    const int v = nb_list[a];
    int b = a - 1;
    while (b >= s && nb_list[b] > v) {
      nb_list[b + 1] = nb_list[b];
      --b;
    }
    nb_list[b + 1] = v;
  }
  const int m = e - s + 1;  // neighbourhood size including the box itself (:499)
  const float yaw_i = bx[(size_t)ri * D + 8];
  float median = yaw_i;
  if (m > 2) {  // :531-540  sorted(list)[len/2], list = neighbour yaws (+ yaw_i once more if m even)
    const int len = m + ((m % 2 == 0) ? 1 : 0);
    const int want = len / 2;
    // element q of the multiset: q = 0 -> yaw_i, 1..m-1 -> neighbours, m -> yaw_i (extra copy)
    auto val = [&](int q) -> float {
      if (q == 0 || q == m) return yaw_i;
      return bx[(size_t)nb_list[s + q - 1] * D + 8];
    };
    for (int q = 0; q < len; ++q) {
      const float v = val(q);
      int lt = 0, eq = 0;
      for (int u = 0; u < len; ++u) {
        const float x = val(u);
        lt += x < v;
        eq += x == v;
      }
      if (lt <= want && want < lt + eq) {
        median = v;
        break;
      }
    }
  }
  float s1[11], s3[11];
#pragma unroll
  for (int q = 0; q < 11; ++q) s1[q] = s3[q] = 0.f;
  for (int q = 0; q < m; ++q) {
    const int rj = q == 0 ? ri : nb_list[s + q - 1];
    const float* bj = bx + (size_t)rj * D;
    const float dy = fabsf(bj[8] - median);
    if ((double)fmodf(dy, 6.2831852f) >= 0.3) continue;  // :542  float(2 * 3.1415926) == 6.2831852f
    const float p = bj[11];
#pragma unroll
    for (int u = 0; u < 11; ++u) {
      s1[u] += p * bj[u];
      s3[u] += p;
    }
  }
#pragma unroll
  for (int u = 0; u < 11; ++u) out_dets[(size_t)k * D + u] = __fdiv_rn(s1[u], s3[u]);
  out_dets[(size_t)k * D + 11] = bx[(size_t)ri * D + 11];
  keep_inds[k] = order[ri];
}

}  // namespace wn

extern "C" {

size_t rd_wnms_4c_workspace_bytes(int n) {
  if (n < 0) return 0;
  return wn::make_layout(n).total;
}

int rd_wnms_4c(const float* dets, int n, float thresh, float thresh_vote, int is_3d, int hash_scale,
               float* out_dets, int32_t* keep_inds, int* out_count, void* workspace,
               size_t workspace_bytes, rd_stream_t stream) {
  using namespace wn;
  RD_REQUIRE(out_count != nullptr, "rd_wnms_4c: out_count is null");
  *out_count = 0;
  RD_REQUIRE(n >= 0, "rd_wnms_4c: negative n");
  if (n == 0) return 0;  // nms.h:464-466
  RD_REQUIRE(dets && out_dets && keep_inds, "rd_wnms_4c: null pointer");
  RD_REQUIRE(hash_scale != 0, "rd_wnms_4c: hash_scale must be non-zero");
  const Layout L = make_layout(n);
  RD_REQUIRE(workspace && workspace_bytes >= L.total, "rd_wnms_4c: workspace too small (%zu < %zu)",
             workspace_bytes, L.total);
  if (rd_check_device()) return 1;
  const size_t bitmap_bytes = (size_t)((n + 31) / 32) * 4;
  RD_REQUIRE(bitmap_bytes <= 200 * 1024, "rd_wnms_4c: n=%d exceeds the shared-memory bitmap capacity (1.6M boxes)", n);
  cudaStream_t st = rd::as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  float* keys_in = (float*)(ws + L.keys_in);
  float* keys_out = (float*)(ws + L.keys_out);
  int* idx_in = (int*)(ws + L.idx_in);
  int* order = (int*)(ws + L.order);
  float* bx = (float*)(ws + L.bx);
  float4* aabb = (float4*)(ws + L.aabb);
  short4* hashr = (short4*)(ws + L.hashr);
  int* cell_id = (int*)(ws + L.cell_id);
  int* cell_id_s = (int*)(ws + L.cell_id_s);
  int* rank_in = (int*)(ws + L.rank_in);
  int* cell_rank = (int*)(ws + L.cell_rank);
  int* cell_start = (int*)(ws + L.cell_start);
  int* cell_end = (int*)(ws + L.cell_end);
  GridParams* gp = (GridParams*)(ws + L.params);
  Counters* ctr = (Counters*)(ws + L.counters);
  int* kept_rank = (int*)(ws + L.kept_rank);
  int* nb_start = (int*)(ws + L.nb_start);
  int* nb_list = (int*)(ws + L.nb_list);
  float* theta = (float*)(ws + L.theta);
  float* theta_s = (float*)(ws + L.theta_s);
  int* theta_rank = (int*)(ws + L.theta_rank);
  float4* aabb_c = (float4*)(ws + L.aabb_c);
  short4* hash_c = (short4*)(ws + L.hash_c);
  int* th_range = (int*)(ws + L.th_range);
  void* cub_temp = ws + L.cub_temp;

  const int TB = 256, nb = (n + TB - 1) / TB;
  init_kernel<<<nb, TB, 0, st>>>(dets, n, keys_in, idx_in, ctr);
  rd::count_launch();
  {
    size_t need = 0;
    RD_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, keys_in, keys_out, idx_in, order, n, 0, 32, st));
    RD_REQUIRE(need <= L.cub_bytes, "rd_wnms_4c: CUB temp storage %zu exceeds reserved %zu", need, L.cub_bytes);
    size_t tb = L.cub_bytes;
    RD_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_temp, tb, keys_in, keys_out, idx_in, order, n, 0, 32, st));
    rd::count_launch(3);
  }
  prep_kernel<<<nb, TB, 0, st>>>(dets, order, n, (float)hash_scale, bx, aabb, hashr, theta, ctr);
  grid_setup_kernel<<<1, 1, 0, st>>>(ctr, gp);
  cell_kernel<<<nb, TB, 0, st>>>(aabb, n, gp, cell_id, rank_in);
  rd::count_launch(3);
  {
    size_t need = 0;
    RD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, cell_id, cell_id_s, rank_in, cell_rank, n, 0, 21, st));
    RD_REQUIRE(need <= L.cub_bytes, "rd_wnms_4c: CUB temp storage %zu exceeds reserved %zu", need, L.cub_bytes);
    size_t tb = L.cub_bytes;
    RD_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, tb, cell_id, cell_id_s, rank_in, cell_rank, n, 0, 21, st));
    rd::count_launch(3);
  }
  {  // boxes sorted by folded edge direction (rank_in still holds 0..n-1)
    size_t tb = L.cub_bytes;
    RD_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, tb, theta, theta_s, rank_in, theta_rank, n, 0, 32, st));
    rd::count_launch(3);
  }
  RD_CUDA(cudaMemsetAsync(cell_start, 0, (size_t)(NCELL_MAX + 1) * 4, st));
  RD_CUDA(cudaMemsetAsync(cell_end, 0, (size_t)(NCELL_MAX + 1) * 4, st));
  cell_bounds_kernel<<<nb, TB, 0, st>>>(cell_id_s, n, cell_start, cell_end);
  cell_gather_kernel<<<nb, TB, 0, st>>>(cell_rank, aabb, hashr, n, aabb_c, hash_c);
  theta_range_kernel<<<nb, TB, 0, st>>>(theta, theta_s, n, th_range);
  rd::count_launch(3);
  // Parallel path (default for n >= 2048): adjacency of all boxes in parallel, then the light sequential walk.
  // RD_WNMS_PARALLEL=0 forces the single-CTA scan; it is also the fallback when the candidate lists overflow.
  static const int par_env = [] { const char* e = getenv("RD_WNMS_PARALLEL"); return e ? atoi(e) : -1; }();
  bool parallel = par_env < 0 ? n >= 2048 : par_env != 0;
  int* cand_cnt = (int*)(ws + L.cand_cnt);
  int* cand_off = (int*)(ws + L.cand_off);
  int* cand_j = (int*)(ws + L.cand_j);
  unsigned char* cand_flag = (unsigned char*)(ws + L.cand_flag);
  auto run_greedy = [&]() -> int {
    RD_CUDA(rd::smem_optin(greedy_kernel, bitmap_bytes));
    greedy_kernel<<<1, NT, bitmap_bytes, st>>>(bx, aabb, hashr, cell_rank, cell_start, cell_end, aabb_c, hash_c, th_range,
                                               theta_rank, gp, n, thresh, thresh_vote, is_3d, kept_rank, nb_start, nb_list,
                                               (int)L.nb_cap, ctr);
    rd::count_launch();
    return rd::check_launch("rd_wnms_4c(greedy)");
  };
  if (parallel) {
    const int nbw = (n + CAND_WARPS - 1) / CAND_WARPS;
    cand_kernel<false><<<nbw, CAND_WARPS * 32, 0, st>>>(bx, aabb, hashr, cell_rank, cell_start, cell_end, aabb_c, hash_c, th_range,
                                                        theta_rank, gp, n, thresh, thresh_vote, is_3d, cand_cnt, cand_off, cand_j,
                                                        cand_flag, (int)L.cand_cap, ctr);
    RD_CUDA(cudaMemsetAsync(cand_cnt + n, 0, 4, st));
    {
      size_t need = 0;
      RD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, cand_cnt, cand_off, n + 1, st));
      RD_REQUIRE(need <= L.cub_bytes, "rd_wnms_4c: CUB temp storage %zu exceeds reserved %zu", need, L.cub_bytes);
      size_t tb = L.cub_bytes;
      RD_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp, tb, cand_cnt, cand_off, n + 1, st));
    }
    cand_kernel<true><<<nbw, CAND_WARPS * 32, 0, st>>>(bx, aabb, hashr, cell_rank, cell_start, cell_end, aabb_c, hash_c, th_range,
                                                       theta_rank, gp, n, thresh, thresh_vote, is_3d, cand_cnt, cand_off, cand_j,
                                                       cand_flag, (int)L.cand_cap, ctr);
    RD_CUDA(rd::smem_optin(resolve_kernel, bitmap_bytes));
    resolve_kernel<<<1, 512, bitmap_bytes, st>>>(cand_off, cand_j, cand_flag, n, kept_rank, nb_start, nb_list, (int)L.nb_cap, ctr);
    rd::count_launch(5);
    if (rd::check_launch("rd_wnms_4c(parallel)")) return 1;
  } else {
    if (run_greedy()) return 1;
  }
  Counters h;
  RD_CUDA(cudaMemcpyAsync(&h, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  RD_CUDA(cudaStreamSynchronize(st));
  if (parallel && h.cand_overflow) {   // more than 320 candidate pairs per box on average: the sequential scan needs no lists
    if (run_greedy()) return 1;
    RD_CUDA(cudaMemcpyAsync(&h, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    RD_CUDA(cudaStreamSynchronize(st));
  }
  RD_REQUIRE(!h.nb_overflow, "rd_wnms_4c: neighbour list overflow (thresh > thresh_vote with very dense boxes)");
  if (h.kept > 0) {
    merge_kernel<<<(h.kept + 127) / 128, 128, 0, st>>>(bx, order, kept_rank, nb_start, nb_list, ctr, out_dets, keep_inds);
    rd::count_launch();
    if (rd::check_launch("rd_wnms_4c(merge)")) return 1;
    RD_CUDA(cudaStreamSynchronize(st));
  }
  *out_count = h.kept;
  return 0;
}

}  // extern "C"
