// Mean squared distance of every point to its 3 nearest neighbours — the initialisation statistic of the surfel scales.
// Behavioural contract: submodules/simple-knn/simple_knn.cu:134-183 (updateKBest<3>, boxMeanDist) and spatial.cu:15-26
// (distCUDA2): exact 3 nearest neighbours under squared L2, the point itself excluded BY INDEX (a coincident point is a
// neighbour at distance 0), result (best0 + best1 + best2) / 3 with the three distances in ascending order, FLT_MAX
// standing in for neighbours that do not exist (P < 4).
//
// Design (not the reference's): the reference sorts the points along a Morton curve, boxes every 1024 of them and lets
// each thread scan whole 1024-point boxes.  Here the sorted points are gathered once into a contiguous float4 array
// (coalesced from then on) and boxed at two levels — leaves of 32 points, groups of 32 leaves.  One warp owns one leaf
// of 32 neighbouring queries: it tests 32 group boxes, then 32 leaf boxes, per instruction against the box of its
// queries and visits only the leaves that can still hold a neighbour closer than the largest "current third best" of
// its lanes; a visited leaf is staged in shared memory and read back as broadcast LDS.128.  Exact: a skipped leaf lies
// farther from every query of the warp than that query's current third best (bound shrunk by 1e-5 relative).
#include "raster_common.cuh"
#include "knn.cuh"

#include <cfloat>

namespace d2gs {

namespace {
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned int f2o(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// sorted points (x, y, z, original index) + leaf boxes; one warp per leaf, one CTA of 32 warps per group
__global__ void __launch_bounds__(1024) knn_build_kernel(int P, const float* __restrict__ xyz, const int* __restrict__ order,
                                                         float4* __restrict__ sp, float4* __restrict__ leaf_box,
                                                         float4* __restrict__ group_box) {
  __shared__ float s_lo[32][3], s_hi[32][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int leaf = blockIdx.x * 32 + warp;
  const int nleaf = (P + 31) >> 5;
  const int j = leaf * 32 + lane;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (leaf < nleaf) {
    float4 p = make_float4(1e30f, 1e30f, 1e30f, __int_as_float(-1));   // padding: farther than any real point
    if (j < P) {
      const int i = order[j];
      p = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __int_as_float(i));
    }
    sp[j] = p;
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      // NaN coordinates must not be boxed away: they make the leaf's box infinite (always visited)
      const bool nan = (j < P) && !(c[a] == c[a]);
      const unsigned int ulo = (j < P && !nan) ? f2o(c[a]) : 0xffffffffu, uhi = (j < P && !nan) ? f2o(c[a]) : 0u;
      lo[a] = o2f(__reduce_min_sync(FULL, ulo));
      hi[a] = o2f(__reduce_max_sync(FULL, uhi));
      if (__any_sync(FULL, nan)) { lo[a] = -INFINITY; hi[a] = INFINITY; }
    }
    if (lane == 0) {
      leaf_box[2 * (size_t)leaf] = make_float4(lo[0], lo[1], lo[2], 0.f);
      leaf_box[2 * (size_t)leaf + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { s_lo[warp][a] = lo[a]; s_hi[warp][a] = hi[a]; }   // empty leaves: (+inf, -inf), neutral
  }
  __syncthreads();
  if (warp == 0) {
    float glo[3], ghi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      glo[a] = o2f(__reduce_min_sync(FULL, f2o(s_lo[lane][a])));
      ghi[a] = o2f(__reduce_max_sync(FULL, f2o(s_hi[lane][a])));
    }
    if (lane == 0) {
      group_box[2 * (size_t)blockIdx.x] = make_float4(glo[0], glo[1], glo[2], 0.f);
      group_box[2 * (size_t)blockIdx.x + 1] = make_float4(ghi[0], ghi[1], ghi[2], 0.f);
    }
  }
}

// simple_knn.cu:134-146: keep the three smallest distances in ascending order
__device__ __forceinline__ void keep3(float best[3], float dist) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (best[k] > dist) { const float t = best[k]; best[k] = dist; dist = t; }
  }
}

__global__ void __launch_bounds__(256) knn_search_kernel(int P, const float4* __restrict__ sp, const float4* __restrict__ leaf_box,
                                                         const float4* __restrict__ group_box, float* __restrict__ out) {
  __shared__ float4 s_leaf[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nleaf = (P + 31) >> 5, ngroup = (nleaf + 31) >> 5;
  const int own = blockIdx.x * 8 + warp;
  if (own >= nleaf) return;
  const int j = own * 32 + lane;
  const bool live = j < P;
  const int jq = live ? j : P - 1;                 // idle lanes of the last leaf shadow the last point
  const float4 q = sp[jq];
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};

  float qlo[3], qhi[3];
  {
    const float c[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      qlo[a] = o2f(__reduce_min_sync(FULL, f2o(c[a])));
      qhi[a] = o2f(__reduce_max_sync(FULL, f2o(c[a])));
    }
  }
  // lower bound of the squared distance between the warp's query box and a box
  auto bound = [&](const float4 lo, const float4 hi) {
    const float dx = fmaxf(fmaxf(lo.x - qhi[0], qlo[0] - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - qhi[1], qlo[1] - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - qhi[2], qlo[2] - hi.z), 0.f);
    const float b = (dx * dx + dy * dy + dz * dz) * 0.99999f;
    return (b == b) ? b : 0.f;                     // NaN: never rejected
  };
  // all 32 points of leaf L against every lane's query
  auto visit = [&](int L) {
    __syncwarp();
    s_leaf[warp][lane] = sp[(size_t)L * 32 + lane];
    __syncwarp();
#pragma unroll 4
    for (int t = 0; t < 32; t++) {
      const float4 c = s_leaf[warp][t];
      const float dx = __fsub_rn(c.x, q.x), dy = __fsub_rn(c.y, q.y), dz = __fsub_rn(c.z, q.z);
      // simple_knn.cu:138: d.x*d.x + d.y*d.y + d.z*d.z as nvcc contracts it
      const float dist = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      if (L * 32 + t != jq) keep3(best, dist);      // the point itself is excluded by position, as in the reference
    }
  };
  auto warp_worst = [&]() { return __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(best[2]))); };   // best[2] >= 0 (NaN sorts last: visit all)

  // own leaf and its neighbours on the curve first: they almost always hold the three nearest
  const int first = max(own - 1, 0), last = min(own + 1, nleaf - 1);
  for (int L = first; L <= last; L++) visit(L);
  float wmax = warp_worst();

  // 32 group boxes per instruction, then the 32 leaf boxes of every group that survives
  for (int gbase = 0; gbase < ngroup; gbase += 32) {
    const int g = gbase + lane;
    float gb = INFINITY;
    if (g < ngroup) gb = bound(group_box[2 * (size_t)g], group_box[2 * (size_t)g + 1]);
    uint32_t gc = __ballot_sync(FULL, gb <= wmax);
    while (gc) {
      const int gl = __ffs(gc) - 1;
      gc &= gc - 1;
      if (__shfl_sync(FULL, gb, gl) > wmax) continue;          // wmax tightened since the ballot
      const int L = (gbase + gl) * 32 + lane;
      float lb = INFINITY;
      if (L < nleaf && (L < first || L > last)) lb = bound(leaf_box[2 * (size_t)L], leaf_box[2 * (size_t)L + 1]);
      uint32_t lc = __ballot_sync(FULL, lb <= wmax);
      while (lc) {
        const int ll = __ffs(lc) - 1;
        lc &= lc - 1;
        if (__shfl_sync(FULL, lb, ll) > wmax) continue;
        visit((gbase + gl) * 32 + ll);
        wmax = warp_worst();
      }
    }
  }
  if (live) out[__float_as_int(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;   // simple_knn.cu:182
}
}  // namespace

void knn_mean_dist2_launch(int P, const float* xyz, const int* order, float4* sp, float4* leaf_box, float4* group_box,
                           float* out, cudaStream_t s) {
  if (P <= 0) return;
  const int nleaf = (P + 31) / 32, ngroup = (nleaf + 31) / 32;
  knn_build_kernel<<<ngroup, 1024, 0, s>>>(P, xyz, order, sp, leaf_box, group_box);
  knn_search_kernel<<<(nleaf + 7) / 8, 256, 0, s>>>(P, sp, leaf_box, group_box, out);
}

}  // namespace d2gs
