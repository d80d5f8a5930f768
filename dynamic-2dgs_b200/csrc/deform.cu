// Node-controlled deformation on sm_100a: brute-force K-nearest control nodes in the (3+hyper)-D embedding,
// radial-basis weights, and the blend of per-node MLP outputs onto the surfels — one fused kernel each way.
// Behavioural contract: utils/time_utils.py:934-967 (cal_nn_weight; knn_points = squared L2, K smallest, ascending),
// :1145-1157 (local-frame translation), :1190-1193 (rotation / scaling residuals), :115-132 (quaternion_to_matrix with
// the 2/|q|^2 factor, quaternion NOT normalised).  The reference materialises P*K*{3,4,9} temporaries across ~25
// eager kernels; here nothing of size P*K*c is written except the (idx, dist, weight) triple the backward needs.
#include "raster_common.cuh"
#include "deform.cuh"

namespace d2gs {

extern int g_deform_bwd_smem;
extern int g_knn_filter;
constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_K = 8;
constexpr int MAX_D = 3 + 16;   // 3 spatial + up to 16 hyper coordinates

// rotation matrix rows from an un-normalised quaternion (r,i,j,k)
__device__ __forceinline__ void quat_to_matrix_raw(const float q[4], float R[9]) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);     R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r);     R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r);     R[7] = two_s * (j * k + i * r);     R[8] = 1 - two_s * (i * i + j * j);
}

// VJP of quat_to_matrix_raw: given dL/dR (row-major 9) returns dL/dq
__device__ __forceinline__ void quat_to_matrix_raw_vjp(const float q[4], const float dR[9], float dq[4]) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float n2 = r * r + i * i + j * j + k * k;
  const float ts = 2.0f / n2;
  // R = I + ts * A(q), with A the quadratic forms below; dL/dts = <dR, A>, dts/dq = -2*ts/n2 * q
  const float A[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r,
                      i * j + k * r, -(i * i + k * k), j * k - i * r,
                      i * k - j * r, j * k + i * r, -(i * i + j * j)};
  float dts = 0.f;
#pragma unroll
  for (int t = 0; t < 9; t++) dts += dR[t] * A[t];
  const float c = -ts / n2 * 2.0f * dts;   // multiplies q
  // ts * dA/dq contracted with dR
  const float gr = ts * (-k * dR[1] + j * dR[2] + k * dR[3] - i * dR[5] - j * dR[6] + i * dR[7]);
  const float gi = ts * (j * dR[1] + k * dR[2] + j * dR[3] - 2 * i * dR[4] - r * dR[5] + k * dR[6] + r * dR[7] - 2 * i * dR[8]);
  const float gj = ts * (-2 * j * dR[0] + i * dR[1] + r * dR[2] + i * dR[3] + k * dR[5] - r * dR[6] + k * dR[7] - 2 * j * dR[8]);
  const float gk = ts * (-2 * k * dR[0] - r * dR[1] + i * dR[2] + r * dR[3] - 2 * k * dR[4] + j * dR[5] + i * dR[6] + j * dR[7]);
  dq[0] = gr + c * r; dq[1] = gi + c * i; dq[2] = gj + c * j; dq[3] = gk + c * k;
}

__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct DeformFwdP {
  int P, M, K, D, hyper;
  const float* xyz; const float* feature; int fstride;
  const float* nodes; const float* radius_log; const float* weight_logit;
  const float* trans; const float* rot; const float* scale; const float* local_rot;
  const float* mask;
  int64_t* nn_idx; float* nn_dist; float* nn_weight;
  float* d_xyz; float* d_rot; float* d_scale;
  int st_t, st_r, st_s, st_l;   // row strides of the node attribute tables (3,4,2,4 or the packed MLP output width)
  const int* order;             // optional: thread t handles surfel order[t] (spatially coherent warps)
  int filter;                   // 1: warp-level candidate filter (below); 0: every node is visited (same results)
};

// K nearest control nodes of every surfel + the radial-basis blend of the node outputs, one thread per surfel.
//
// The K-set is an unordered "replace the current worst" set keyed by (distance, node index): a candidate enters iff it
// is lexicographically smaller than the current worst, so the result does not depend on the order — or on the subset —
// in which nodes are offered, as long as every node that belongs to the final set is offered once.  Every step is a
// predicated select: lanes that insert at different nodes do not serialise a sorted-insertion ladder.  Ties resolve to
// the lower node index like a stable sort (pytorch3d's knn_points).
//
// Warp-level candidate filter.  With the Morton processing order the 32 surfels of a warp sit in a small box B.  A node
// whose spatial distance to B exceeds the largest "current worst" of the 32 lanes (wmax) cannot enter any lane's set
// (the hyper coordinates only add to the distance), so the warp
//   1. gives every lane the node nearest to B among "its" nodes (slots lane, lane+32, ...) and visits those 32 nodes:
//      every lane now holds K real candidates and wmax is finite;
//   2. walks the node table in chunks of 32: lane l bounds the distance of node 32c+l to B (one LDS.128 + 12 ALU),
//      a ballot collects the nodes with bound <= wmax, the warp visits exactly those, wmax is refreshed per chunk.
// At C3 (512 uniform nodes, K=4) a warp evaluates distances for ~70 nodes instead of 512.  The bound is shrunk by 1e-5
// relative — far above the fp32 error of either side — so a rejected node is strictly farther than every lane's worst;
// nn_idx / nn_dist are bit-identical to the exhaustive search (filter = 0; test_knn_candidate_filter_changes_nothing).
// Distances accumulate coordinate by coordinate with fused multiply-adds in storage order, in both modes.
//
// Persistent warps: the node table (slot = node index, [M][4*NQ] floats) is staged once per CTA; warp w of the grid
// handles surfels 32*w.., stepping by the number of warps, with no barrier after the staging.
template <int K, int NQ>   // NQ = round_up(D, 4) / 4 coordinate quads
__global__ void __launch_bounds__(256) deform_fwd_kernel(DeformFwdP a) {
  extern __shared__ float4 s_node4[];   // [MP][NQ]
  constexpr int DP = 4 * NQ;
  const int D = a.D;
  const int nstride = 3 + a.hyper;
  const int MP = (a.M + 31) & ~31;       // whole chunks of 32 slots; padding slots are infinitely far
  {
    float* s_node = reinterpret_cast<float*>(s_node4);
    for (int t = threadIdx.x; t < MP * DP; t += blockDim.x) {
      const int m = t / DP, d = t - m * DP;
      float v = 0.f;
      if (d < D) v = (m < a.M) ? a.nodes[(size_t)m * nstride + d] : 1e30f;
      s_node[t] = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int nch = MP >> 5;
  for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < a.P; base += nwarps * 32) {
    const int t_raw = base + lane;
    const bool live = t_raw < a.P;
    const int t_lin = live ? t_raw : a.P - 1;                      // idle lanes of the last warp shadow the last surfel
    const int i = a.order ? __ldg(a.order + t_lin) : t_lin;

    float q[DP];
#pragma unroll
    for (int d = 0; d < DP; d++) {
      float v = 0.f;
      if (d < 3) v = a.xyz[3 * (size_t)i + d];
      else if (d < D) v = a.feature[(size_t)i * a.fstride + (d - 3)];
      q[d] = v;
    }

    float bd[K];
    int bi[K];
#pragma unroll
    for (int k = 0; k < K; k++) { bd[k] = INFINITY; bi[k] = 0x7ffffff0 - k; }
    float wd = INFINITY;   // current worst (largest (dist, idx)) of the set, its index and its slot
    int wi = 0x7ffffff0, ws = 0;
    auto offer = [&](float dist, int m) {
      if (dist < wd || (dist == wd && m < wi)) {
#pragma unroll
        for (int k = 0; k < K; k++) {
          const bool sel = (k == ws);
          bd[k] = sel ? dist : bd[k];
          bi[k] = sel ? m : bi[k];
        }
        wd = bd[0]; ws = 0; wi = bi[0];
#pragma unroll
        for (int k = 1; k < K; k++) {
          const bool worse = (bd[k] > wd) || (bd[k] == wd && bi[k] > wi);
          wd = worse ? bd[k] : wd;
          wi = worse ? bi[k] : wi;
          ws = worse ? k : ws;
        }
      }
    };
    // squared distance of the lane's query to the node in slot `s` (warp-uniform: broadcast LDS.128), offered to the set
    auto visit = [&](int s) {
      const float4* n4 = s_node4 + (size_t)s * NQ;
      float dist = 0.f;
      {
        const float4 v = n4[0];
        float df = __fsub_rn(q[0], v.x); dist = __fmaf_rn(df, df, dist);
        df = __fsub_rn(q[1], v.y); dist = __fmaf_rn(df, df, dist);
        df = __fsub_rn(q[2], v.z); dist = __fmaf_rn(df, df, dist);
        df = __fsub_rn(q[3], v.w); dist = __fmaf_rn(df, df, dist);
      }
      if (NQ > 1) {
        // the running sum only grows: a node whose prefix (x, y, z, h0) already exceeds the lane's worst cannot enter
        if (!(dist <= wd)) return;
#pragma unroll
        for (int c = 1; c < NQ; c++) {
          const float4 v = n4[c];
          float df = __fsub_rn(q[4 * c], v.x); dist = __fmaf_rn(df, df, dist);
          df = __fsub_rn(q[4 * c + 1], v.y); dist = __fmaf_rn(df, df, dist);
          df = __fsub_rn(q[4 * c + 2], v.z); dist = __fmaf_rn(df, df, dist);
          df = __fsub_rn(q[4 * c + 3], v.w); dist = __fmaf_rn(df, df, dist);
        }
      }
      offer(dist, s < a.M ? s : 0x7fffffff);
    };

    if (!a.filter) {
      for (int s = 0; s < MP; s++) visit(s);
    } else {
      // the warp's query box (ordered-integer min/max: one REDUX each)
      float qlo[3], qhi[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const unsigned int u = float_to_ordered(q[c]);
        qlo[c] = ordered_to_float(__reduce_min_sync(FULL, u));
        qhi[c] = ordered_to_float(__reduce_max_sync(FULL, u));
      }
      // lower bound of the squared distance between any query of the warp and the node in slot s
      auto bound = [&](int s) {
        const float4 v = s_node4[(size_t)s * NQ];
        const float dx = fmaxf(fmaxf(qlo[0] - v.x, v.x - qhi[0]), 0.f);
        const float dy = fmaxf(fmaxf(qlo[1] - v.y, v.y - qhi[1]), 0.f);
        const float dz = fmaxf(fmaxf(qlo[2] - v.z, v.z - qhi[2]), 0.f);
        float b = (dx * dx + dy * dy + dz * dz) * 0.99999f;
        return (b == b) ? b : 0.f;          // NaN (a NaN node or query): never rejected
      };
      // 1. every lane nominates the nearest of its own nodes
      float nb = INFINITY;
      int nch_sel = 0;
      for (int c = 0; c < nch; c++) {
        const float b = bound(32 * c + lane);
        if (b < nb) { nb = b; nch_sel = c; }
      }
      const int nominated = 32 * nch_sel + lane;
#pragma unroll 1
      for (int l = 0; l < 32; l++) visit(__shfl_sync(FULL, nominated, l));
      float wmax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(wd)));   // wd >= 0 or +inf: bit order = value order
      // 2. chunks of 32 nodes: visit those that can still matter
#pragma unroll 1
      for (int c = 0; c < nch; c++) {
        const float b = bound(32 * c + lane);
        uint32_t cand = __ballot_sync(FULL, b <= wmax && c != nch_sel);       // the nominated node was visited above
        if (!cand) continue;
        while (cand) {
          const int l = __ffs(cand) - 1;
          cand &= cand - 1;
          visit(32 * c + l);
        }
        wmax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(wd)));
      }
    }
    // order the K survivors by (distance, index): tiny odd-even transposition network
#pragma unroll
    for (int pass = 0; pass < K; pass++) {
#pragma unroll
      for (int k = (pass & 1); k + 1 < K; k += 2) {
        const bool sw = (bd[k + 1] < bd[k]) || (bd[k + 1] == bd[k] && bi[k + 1] < bi[k]);
        const float td = bd[k]; const int ti = bi[k];
        bd[k] = sw ? bd[k + 1] : bd[k]; bi[k] = sw ? bi[k + 1] : bi[k];
        bd[k + 1] = sw ? td : bd[k + 1]; bi[k + 1] = sw ? ti : bi[k + 1];
      }
    }
    // a query with NaN/Inf coordinates accepts no node: its set keeps the initial sentinels; point them at node k
#pragma unroll
    for (int k = 0; k < K; k++) bi[k] = ((unsigned int)bi[k] < (unsigned int)a.M) ? bi[k] : k;

    // radial-basis weights
    float w[K];
    float wsum = 0.f;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int m = bi[k];
      const float r = expf(__ldg(a.radius_log + m));
      float wk = expf(-bd[k] / (2 * (r * r)));
      if (a.weight_logit) wk = wk * (1.0f / (1.0f + expf(-__ldg(a.weight_logit + m))));
      wk = wk + 1e-7f;
      w[k] = wk;
      wsum += wk;
    }
    const float x0 = q[0], x1 = q[1], x2 = q[2];
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int m = bi[k];
      const float wk = w[k] / wsum;
      const float tr0 = __ldg(a.trans + a.st_t * m), tr1 = __ldg(a.trans + a.st_t * m + 1), tr2 = __ldg(a.trans + a.st_t * m + 2);
      if (a.local_rot) {
        float lq[4] = {__ldg(a.local_rot + a.st_l * m) + 1.0f, __ldg(a.local_rot + a.st_l * m + 1), __ldg(a.local_rot + a.st_l * m + 2),
                       __ldg(a.local_rot + a.st_l * m + 3)};
        float R[9];
        quat_to_matrix_raw(lq, R);
        const float4 nv = s_node4[(size_t)m * NQ];
        const float n0 = nv.x, n1 = nv.y, n2 = nv.z;
        const float e0 = x0 - n0, e1 = x1 - n1, e2 = x2 - n2;
        const float A0 = (R[0] * e0 + R[1] * e1 + R[2] * e2) + n0 + tr0;
        const float A1 = (R[3] * e0 + R[4] * e1 + R[5] * e2) + n1 + tr1;
        const float A2 = (R[6] * e0 + R[7] * e1 + R[8] * e2) + n2 + tr2;
        t0 += A0 * wk; t1 += A1 * wk; t2 += A2 * wk;
      } else {
        t0 += tr0 * wk; t1 += tr1 * wk; t2 += tr2 * wk;
      }
      r0 += __ldg(a.rot + a.st_r * m) * wk; r1 += __ldg(a.rot + a.st_r * m + 1) * wk;
      r2 += __ldg(a.rot + a.st_r * m + 2) * wk; r3 += __ldg(a.rot + a.st_r * m + 3) * wk;
      s0 += __ldg(a.scale + a.st_s * m) * wk; s1 += __ldg(a.scale + a.st_s * m + 1) * wk;
      if (live) {
        if (a.nn_idx) a.nn_idx[(size_t)i * K + k] = m;
        if (a.nn_dist) a.nn_dist[(size_t)i * K + k] = bd[k];
        if (a.nn_weight) a.nn_weight[(size_t)i * K + k] = wk;
      }
    }
    if (live) {
      if (a.local_rot) { t0 -= x0; t1 -= x1; t2 -= x2; }
      const float mk = a.mask ? a.mask[i] : 1.0f;
      a.d_xyz[3 * (size_t)i] = t0 * mk; a.d_xyz[3 * (size_t)i + 1] = t1 * mk; a.d_xyz[3 * (size_t)i + 2] = t2 * mk;
      a.d_rot[4 * (size_t)i] = r0 * mk; a.d_rot[4 * (size_t)i + 1] = r1 * mk; a.d_rot[4 * (size_t)i + 2] = r2 * mk;
      a.d_rot[4 * (size_t)i + 3] = r3 * mk;
      a.d_scale[2 * (size_t)i] = s0 * mk; a.d_scale[2 * (size_t)i + 1] = s1 * mk;
    }
  }
}

struct DeformBwdP {
  int P, M, K, D, hyper;
  const float* xyz; const float* feature; int fstride;
  const float* nodes; const float* radius_log; const float* weight_logit;
  const float* trans; const float* rot; const float* scale; const float* local_rot;
  const float* mask;
  const int64_t* nn_idx; const float* nn_dist; const float* nn_weight;
  const float* g_xyz; const float* g_rot; const float* g_scale;
  float* d_trans; float* d_rot; float* d_scale; float* d_local_rot; float* d_nodes; float* d_radius_log;
  float* d_weight_logit;
  float* d_feature; float* d_mask;
  int use_smem;   // 1: per-CTA shared accumulators for the node gradients
  int st_t, st_r, st_s, st_l;
  const int* order;
};

// per-node gradient row inside the accumulator: [0..2] trans [3..6] rot [7..8] scale [9..12] local_rot
// [13] radius_log [14] weight_logit [15..15+hyper) hyper coordinates
constexpr int NG_FIXED = 15;

__global__ void __launch_bounds__(256) deform_bwd_kernel(DeformBwdP a) {
  extern __shared__ float s_acc[];   // M * (NG_FIXED + hyper) when use_smem
  const int NG = NG_FIXED + a.hyper;
  const int nstride = 3 + a.hyper;
  if (a.use_smem) {
    for (int t = threadIdx.x; t < a.M * NG; t += blockDim.x) s_acc[t] = 0.f;
    __syncthreads();
  }
  auto add = [&](int m, int c, float v) {
    if (a.use_smem) { atomicAdd(&s_acc[m * NG + c], v); return; }
    if (c < 3) atomicAdd(a.d_trans + a.st_t * m + c, v);
    else if (c < 7) atomicAdd(a.d_rot + a.st_r * m + (c - 3), v);
    else if (c < 9) atomicAdd(a.d_scale + a.st_s * m + (c - 7), v);
    else if (c < 13) { if (a.d_local_rot) atomicAdd(a.d_local_rot + a.st_l * m + (c - 9), v); }
    else if (c == 13) atomicAdd(a.d_radius_log + m, v);
    else if (c == 14) { if (a.d_weight_logit) atomicAdd(a.d_weight_logit + m, v); }
    else atomicAdd(a.d_nodes + (size_t)m * nstride + 3 + (c - NG_FIXED), v);
  };

  const int K = a.K;
  for (int t_lin = blockIdx.x * blockDim.x + threadIdx.x; t_lin < a.P; t_lin += gridDim.x * blockDim.x) {
    const int i = a.order ? __ldg(a.order + t_lin) : t_lin;
    const float mk = a.mask ? a.mask[i] : 1.0f;
    const float x0 = a.xyz[3 * (size_t)i], x1 = a.xyz[3 * (size_t)i + 1], x2 = a.xyz[3 * (size_t)i + 2];
    const float gx0 = a.g_xyz[3 * (size_t)i], gx1 = a.g_xyz[3 * (size_t)i + 1], gx2 = a.g_xyz[3 * (size_t)i + 2];
    const float gr0 = a.g_rot[4 * (size_t)i], gr1 = a.g_rot[4 * (size_t)i + 1], gr2 = a.g_rot[4 * (size_t)i + 2],
                gr3 = a.g_rot[4 * (size_t)i + 3];
    const float gs0 = a.g_scale[2 * (size_t)i], gs1 = a.g_scale[2 * (size_t)i + 1];
    const float Gx0 = gx0 * mk, Gx1 = gx1 * mk, Gx2 = gx2 * mk;
    const float Gr0 = gr0 * mk, Gr1 = gr1 * mk, Gr2 = gr2 * mk, Gr3 = gr3 * mk;
    const float Gs0 = gs0 * mk, Gs1 = gs1 * mk;

    float dw[MAX_K], wk[MAX_K];
    int idx[MAX_K];
    float sum_w_dw = 0.f;
    float acc_x0 = 0.f, acc_x1 = 0.f, acc_x2 = 0.f, acc_r0 = 0.f, acc_r1 = 0.f, acc_r2 = 0.f, acc_r3 = 0.f, acc_s0 = 0.f,
          acc_s1 = 0.f;   // un-masked blended outputs, for the mask gradient
#pragma unroll
    for (int k = 0; k < MAX_K; k++) {
      dw[k] = 0.f; wk[k] = 0.f; idx[k] = 0;
      if (k < K) {
        const int m = (int)a.nn_idx[(size_t)i * K + k];
        const float w = a.nn_weight[(size_t)i * K + k];
        idx[k] = m; wk[k] = w;
        const float tr0 = __ldg(a.trans + a.st_t * m), tr1 = __ldg(a.trans + a.st_t * m + 1), tr2 = __ldg(a.trans + a.st_t * m + 2);
        float A0 = tr0, A1 = tr1, A2 = tr2;
        if (a.local_rot) {
          float lq[4] = {__ldg(a.local_rot + a.st_l * m) + 1.0f, __ldg(a.local_rot + a.st_l * m + 1), __ldg(a.local_rot + a.st_l * m + 2),
                         __ldg(a.local_rot + a.st_l * m + 3)};
          float R[9];
          quat_to_matrix_raw(lq, R);
          const float n0 = __ldg(a.nodes + (size_t)m * nstride), n1 = __ldg(a.nodes + (size_t)m * nstride + 1),
                      n2 = __ldg(a.nodes + (size_t)m * nstride + 2);
          const float e0 = x0 - n0, e1 = x1 - n1, e2 = x2 - n2;
          A0 = (R[0] * e0 + R[1] * e1 + R[2] * e2) + n0 + tr0;
          A1 = (R[3] * e0 + R[4] * e1 + R[5] * e2) + n1 + tr1;
          A2 = (R[6] * e0 + R[7] * e1 + R[8] * e2) + n2 + tr2;
          // dL/dR = (w G_x) (x - node)^T  -> quaternion
          const float wg0 = w * Gx0, wg1 = w * Gx1, wg2 = w * Gx2;
          const float dR[9] = {wg0 * e0, wg0 * e1, wg0 * e2, wg1 * e0, wg1 * e1, wg1 * e2, wg2 * e0, wg2 * e1, wg2 * e2};
          float dq[4];
          quat_to_matrix_raw_vjp(lq, dR, dq);
          add(m, 9, dq[0]); add(m, 10, dq[1]); add(m, 11, dq[2]); add(m, 12, dq[3]);
        }
        const float rr0 = __ldg(a.rot + a.st_r * m), rr1 = __ldg(a.rot + a.st_r * m + 1), rr2 = __ldg(a.rot + a.st_r * m + 2),
                    rr3 = __ldg(a.rot + a.st_r * m + 3);
        const float ss0 = __ldg(a.scale + a.st_s * m), ss1 = __ldg(a.scale + a.st_s * m + 1);
        add(m, 0, w * Gx0); add(m, 1, w * Gx1); add(m, 2, w * Gx2);
        add(m, 3, w * Gr0); add(m, 4, w * Gr1); add(m, 5, w * Gr2); add(m, 6, w * Gr3);
        add(m, 7, w * Gs0); add(m, 8, w * Gs1);
        const float d = Gx0 * A0 + Gx1 * A1 + Gx2 * A2 + Gr0 * rr0 + Gr1 * rr1 + Gr2 * rr2 + Gr3 * rr3 + Gs0 * ss0 + Gs1 * ss1;
        dw[k] = d;
        sum_w_dw += w * d;
        acc_x0 += w * A0; acc_x1 += w * A1; acc_x2 += w * A2;
        acc_r0 += w * rr0; acc_r1 += w * rr1; acc_r2 += w * rr2; acc_r3 += w * rr3;
        acc_s0 += w * ss0; acc_s1 += w * ss1;
      }
    }
    if (a.d_mask) {
      if (a.local_rot) { acc_x0 -= x0; acc_x1 -= x1; acc_x2 -= x2; }
      a.d_mask[i] = gx0 * acc_x0 + gx1 * acc_x1 + gx2 * acc_x2 + gr0 * acc_r0 + gr1 * acc_r1 + gr2 * acc_r2 + gr3 * acc_r3 +
                    gs0 * acc_s0 + gs1 * acc_s1;
    }
    // un-normalised weights: u_k = e_k*sigma_k + 1e-7, w_k = u_k / S.  S is recovered from any (w,u) pair.
    float dq_h[MAX_D - 3];
#pragma unroll
    for (int d = 0; d < MAX_D - 3; d++) dq_h[d] = 0.f;
    float S = 0.f;
    {
      const int m = idx[0];
      const float r = expf(__ldg(a.radius_log + m));
      float e = expf(-a.nn_dist[(size_t)i * K] / (2 * (r * r)));
      const float sg = a.weight_logit ? 1.0f / (1.0f + expf(-__ldg(a.weight_logit + m))) : 1.0f;
      S = (e * sg + 1e-7f) / wk[0];
    }
#pragma unroll
    for (int k = 0; k < MAX_K; k++) {
      if (k < K) {
        const int m = idx[k];
        const float du = (dw[k] - sum_w_dw) / S;
        const float r = expf(__ldg(a.radius_log + m));
        const float dist = a.nn_dist[(size_t)i * K + k];
        const float inv2r2 = 1.0f / (2 * (r * r));
        const float e = expf(-dist * inv2r2);
        const float sg = a.weight_logit ? 1.0f / (1.0f + expf(-__ldg(a.weight_logit + m))) : 1.0f;
        const float de = du * sg;
        if (a.weight_logit) add(m, 14, du * e * sg * (1.0f - sg));
        const float dd = -de * e * inv2r2;                 // dL/d dist
        add(m, 13, de * e * dist * 2.0f * inv2r2);          // dL/d log r  (= de * e * dist / r^2)
        if (a.hyper > 0 && a.feature) {
#pragma unroll
          for (int d = 0; d < MAX_D - 3; d++) {
            if (d < a.hyper) {
              const float df = a.feature[(size_t)i * a.fstride + d] - __ldg(a.nodes + (size_t)m * nstride + 3 + d);
              const float gq = dd * 2.0f * df;
              dq_h[d] += gq;
              add(m, NG_FIXED + d, -gq);
            }
          }
        }
      }
    }
    if (a.d_feature) {
      for (int d = 0; d < a.fstride; d++) a.d_feature[(size_t)i * a.fstride + d] = (d < a.hyper) ? dq_h[d < MAX_D - 3 ? d : 0] : 0.f;
    }
  }

  if (a.use_smem) {
    __syncthreads();
    for (int t = threadIdx.x; t < a.M * NG; t += blockDim.x) {
      const float v = s_acc[t];
      if (v == 0.f) continue;
      const int m = t / NG, c = t - m * NG;
      if (c < 3) atomicAdd(a.d_trans + a.st_t * m + c, v);
      else if (c < 7) atomicAdd(a.d_rot + a.st_r * m + (c - 3), v);
      else if (c < 9) atomicAdd(a.d_scale + a.st_s * m + (c - 7), v);
      else if (c < 13) { if (a.d_local_rot) atomicAdd(a.d_local_rot + a.st_l * m + (c - 9), v); }
      else if (c == 13) atomicAdd(a.d_radius_log + m, v);
      else if (c == 14) { if (a.d_weight_logit) atomicAdd(a.d_weight_logit + m, v); }
      else atomicAdd(a.d_nodes + (size_t)m * nstride + 3 + (c - NG_FIXED), v);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// backward for spatially coherent warps (a processing order is given)
// ---------------------------------------------------------------------------------------------------------------
// Sum 32 per-lane values across the warp with 31 shuffles: afterwards lane L holds the warp total of g[L] in g[0].
__device__ __forceinline__ void warp_transpose_reduce32(float (&g)[32], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const float send = b4 ? g[i] : g[i + 16];
    const float keep = b4 ? g[i + 16] : g[i];
    g[i] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float send = b3 ? g[i] : g[i + 8];
    const float keep = b3 ? g[i + 8] : g[i];
    g[i] = keep + __shfl_xor_sync(FULL, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float send = b2 ? g[i] : g[i + 4];
    const float keep = b2 ? g[i + 4] : g[i];
    g[i] = keep + __shfl_xor_sync(FULL, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float send = b1 ? g[i] : g[i + 2];
    const float keep = b1 ? g[i + 2] : g[i];
    g[i] = keep + __shfl_xor_sync(FULL, send, 2);
  }
  {
    const float send = b0 ? g[0] : g[1];
    const float keep = b0 ? g[1] : g[0];
    g[0] = keep + __shfl_xor_sync(FULL, send, 1);
  }
}

// Neighbouring surfels share their control nodes, so the warp walks over the DISTINCT nodes its 32 x K pairs refer to
// (about 9 with a Morton order): every lane contributes the gradient row of its pair with that node (or zeros), one
// 32-value transposed butterfly sums the rows, and lanes 0..NG-1 issue ONE global reduction each.  92 shared-memory CAS
// loops per surfel become ~9 x NG fire-and-forget REDs per warp, and the node attributes are warp-uniform loads.
// A warp whose surfels are not coherent (more than MAX_DISTINCT different nearest nodes) falls back to per-lane REDs.
// HMAX: capacity of the hyper-coordinate loops (8 when hyper_dim <= 8 — the trainer's setting — else 16): the loops are
// fully unrolled and predicated, so half the capacity is half their instructions and registers.
template <int K, int HMAX>
__global__ void __launch_bounds__(256, 3) deform_bwd_coherent_kernel(DeformBwdP a) {
  constexpr int MAX_DISTINCT = 12;
  const int NG = NG_FIXED + a.hyper;
  const int nstride = 3 + a.hyper;
  const int lane = threadIdx.x & 31;
  // destination of gradient component `lane` of node m: flush_base + flush_stride * m (resolved once, not per flush)
  float* flush_base = nullptr;
  int flush_stride = 0;
  {
    const int c = lane;
    if (c < 3) { flush_base = a.d_trans + c; flush_stride = a.st_t; }
    else if (c < 7) { flush_base = a.d_rot + (c - 3); flush_stride = a.st_r; }
    else if (c < 9) { flush_base = a.d_scale + (c - 7); flush_stride = a.st_s; }
    else if (c < 13) { if (a.d_local_rot) { flush_base = a.d_local_rot + (c - 9); flush_stride = a.st_l; } }
    else if (c == 13) { flush_base = a.d_radius_log; flush_stride = 1; }
    else if (c == 14) { if (a.d_weight_logit) { flush_base = a.d_weight_logit; flush_stride = 1; } }
    else if (c < NG) { flush_base = a.d_nodes + 3 + (c - NG_FIXED); flush_stride = nstride; }
  }
  auto dst_of = [&](int m, int c) -> float* {
    if (c < 3) return a.d_trans + a.st_t * m + c;
    if (c < 7) return a.d_rot + a.st_r * m + (c - 3);
    if (c < 9) return a.d_scale + a.st_s * m + (c - 7);
    if (c < 13) return a.d_local_rot ? a.d_local_rot + a.st_l * m + (c - 9) : nullptr;
    if (c == 13) return a.d_radius_log + m;
    if (c == 14) return a.d_weight_logit ? a.d_weight_logit + m : nullptr;
    return a.d_nodes + (size_t)m * nstride + 3 + (c - NG_FIXED);
  };
  const int t_lin = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t_lin < a.P;
  const int i = live ? (a.order ? __ldg(a.order + t_lin) : t_lin) : 0;

  float x[3] = {0.f, 0.f, 0.f}, Gx[3] = {0.f, 0.f, 0.f}, Gr[4] = {0.f, 0.f, 0.f, 0.f}, Gs[2] = {0.f, 0.f};
  float fh[HMAX];
#pragma unroll
  for (int d = 0; d < HMAX; d++) fh[d] = 0.f;
  int idx[K];
  float wk[K], du[K], dist[K];
#pragma unroll
  for (int k = 0; k < K; k++) { idx[k] = -1; wk[k] = 0.f; du[k] = 0.f; dist[k] = 0.f; }

  if (live) {
    const float mk = a.mask ? a.mask[i] : 1.0f;
    float g_x[3], g_r[4], g_s[2];
#pragma unroll
    for (int c = 0; c < 3; c++) { x[c] = a.xyz[3 * (size_t)i + c]; g_x[c] = a.g_xyz[3 * (size_t)i + c]; Gx[c] = g_x[c] * mk; }
#pragma unroll
    for (int c = 0; c < 4; c++) { g_r[c] = a.g_rot[4 * (size_t)i + c]; Gr[c] = g_r[c] * mk; }
#pragma unroll
    for (int c = 0; c < 2; c++) { g_s[c] = a.g_scale[2 * (size_t)i + c]; Gs[c] = g_s[c] * mk; }
    const bool hyp = a.hyper > 0 && a.feature;
    if (hyp) {
#pragma unroll
      for (int d = 0; d < HMAX; d++)
        if (d < a.hyper) fh[d] = a.feature[(size_t)i * a.fstride + d];
    }
    // pass 1: blended outputs per neighbour -> dL/dw_k, the mask gradient
    float dw[K];
    float sum_w_dw = 0.f;
    float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int m = (int)a.nn_idx[(size_t)i * K + k];
      const float w = a.nn_weight[(size_t)i * K + k];
      idx[k] = m; wk[k] = w; dist[k] = a.nn_dist[(size_t)i * K + k];
      const float tr0 = __ldg(a.trans + a.st_t * m), tr1 = __ldg(a.trans + a.st_t * m + 1), tr2 = __ldg(a.trans + a.st_t * m + 2);
      float A0 = tr0, A1 = tr1, A2 = tr2;
      if (a.local_rot) {
        float lq[4] = {__ldg(a.local_rot + a.st_l * m) + 1.0f, __ldg(a.local_rot + a.st_l * m + 1), __ldg(a.local_rot + a.st_l * m + 2),
                       __ldg(a.local_rot + a.st_l * m + 3)};
        float R[9];
        quat_to_matrix_raw(lq, R);
        const float n0 = __ldg(a.nodes + (size_t)m * nstride), n1 = __ldg(a.nodes + (size_t)m * nstride + 1),
                    n2 = __ldg(a.nodes + (size_t)m * nstride + 2);
        const float e0 = x[0] - n0, e1 = x[1] - n1, e2 = x[2] - n2;
        A0 = (R[0] * e0 + R[1] * e1 + R[2] * e2) + n0 + tr0;
        A1 = (R[3] * e0 + R[4] * e1 + R[5] * e2) + n1 + tr1;
        A2 = (R[6] * e0 + R[7] * e1 + R[8] * e2) + n2 + tr2;
      }
      const float rr0 = __ldg(a.rot + a.st_r * m), rr1 = __ldg(a.rot + a.st_r * m + 1), rr2 = __ldg(a.rot + a.st_r * m + 2),
                  rr3 = __ldg(a.rot + a.st_r * m + 3);
      const float ss0 = __ldg(a.scale + a.st_s * m), ss1 = __ldg(a.scale + a.st_s * m + 1);
      const float d = Gx[0] * A0 + Gx[1] * A1 + Gx[2] * A2 + Gr[0] * rr0 + Gr[1] * rr1 + Gr[2] * rr2 + Gr[3] * rr3 + Gs[0] * ss0 + Gs[1] * ss1;
      dw[k] = d;
      sum_w_dw += w * d;
      acc[0] += w * A0; acc[1] += w * A1; acc[2] += w * A2;
      acc[3] += w * rr0; acc[4] += w * rr1; acc[5] += w * rr2; acc[6] += w * rr3;
      acc[7] += w * ss0; acc[8] += w * ss1;
    }
    if (a.d_mask) {
      if (a.local_rot) { acc[0] -= x[0]; acc[1] -= x[1]; acc[2] -= x[2]; }
      a.d_mask[i] = g_x[0] * acc[0] + g_x[1] * acc[1] + g_x[2] * acc[2] + g_r[0] * acc[3] + g_r[1] * acc[4] + g_r[2] * acc[5] +
                    g_r[3] * acc[6] + g_s[0] * acc[7] + g_s[1] * acc[8];
    }
    // un-normalised weights: u_k = e_k*sigma_k + 1e-7, w_k = u_k / S.  S is recovered from the nearest (w,u) pair.
    float S;
    {
      const int m = idx[0];
      const float r = expf(__ldg(a.radius_log + m));
      const float e = expf(-dist[0] / (2 * (r * r)));
      const float sg = a.weight_logit ? 1.0f / (1.0f + expf(-__ldg(a.weight_logit + m))) : 1.0f;
      S = (e * sg + 1e-7f) / wk[0];
    }
#pragma unroll
    for (int k = 0; k < K; k++) du[k] = (dw[k] - sum_w_dw) / S;
    // gradient of the hyper coordinates of the QUERY (per surfel, no reduction)
    if (a.d_feature) {
      float dq_h[HMAX];
#pragma unroll
      for (int d = 0; d < HMAX; d++) dq_h[d] = 0.f;
      if (hyp) {
#pragma unroll
        for (int k = 0; k < K; k++) {
          const int m = idx[k];
          const float r = expf(__ldg(a.radius_log + m));
          const float inv2r2 = 1.0f / (2 * (r * r));
          const float e = expf(-dist[k] * inv2r2);
          const float sg = a.weight_logit ? 1.0f / (1.0f + expf(-__ldg(a.weight_logit + m))) : 1.0f;
          const float dd = -(du[k] * sg) * e * inv2r2;
#pragma unroll
          for (int d = 0; d < HMAX; d++)
            if (d < a.hyper) dq_h[d] += dd * 2.0f * (fh[d] - __ldg(a.nodes + (size_t)m * nstride + 3 + d));
        }
      }
      for (int d = 0; d < a.fstride; d++) a.d_feature[(size_t)i * a.fstride + d] = (d < a.hyper) ? dq_h[d < HMAX ? d : 0] : 0.f;
    }
  }

  // gradient row of the pair (this lane, node mm) with pair scalars (w, du, dist); all-zero scalars give an all-zero row
  auto pair_row = [&](int mm, float w, float duk, float dk, float (&g)[32]) {
#pragma unroll
    for (int c = 0; c < 32; c++) g[c] = 0.f;
    g[0] = w * Gx[0]; g[1] = w * Gx[1]; g[2] = w * Gx[2];
    g[3] = w * Gr[0]; g[4] = w * Gr[1]; g[5] = w * Gr[2]; g[6] = w * Gr[3];
    g[7] = w * Gs[0]; g[8] = w * Gs[1];
    if (a.local_rot) {
      float lq[4] = {__ldg(a.local_rot + a.st_l * mm) + 1.0f, __ldg(a.local_rot + a.st_l * mm + 1), __ldg(a.local_rot + a.st_l * mm + 2),
                     __ldg(a.local_rot + a.st_l * mm + 3)};
      const float e0 = x[0] - __ldg(a.nodes + (size_t)mm * nstride), e1 = x[1] - __ldg(a.nodes + (size_t)mm * nstride + 1),
                  e2 = x[2] - __ldg(a.nodes + (size_t)mm * nstride + 2);
      const float wg0 = g[0], wg1 = g[1], wg2 = g[2];
      const float dR[9] = {wg0 * e0, wg0 * e1, wg0 * e2, wg1 * e0, wg1 * e1, wg1 * e2, wg2 * e0, wg2 * e1, wg2 * e2};
      float dq[4];
      quat_to_matrix_raw_vjp(lq, dR, dq);
      g[9] = dq[0]; g[10] = dq[1]; g[11] = dq[2]; g[12] = dq[3];
    }
    const float r = expf(__ldg(a.radius_log + mm));
    const float inv2r2 = 1.0f / (2 * (r * r));
    const float e = expf(-dk * inv2r2);
    const float sg = a.weight_logit ? 1.0f / (1.0f + expf(-__ldg(a.weight_logit + mm))) : 1.0f;
    const float de = duk * sg;
    g[13] = de * e * dk * 2.0f * inv2r2;
    g[14] = a.weight_logit ? duk * e * sg * (1.0f - sg) : 0.f;
    if (a.hyper > 0 && a.feature) {
      const float dd = -de * e * inv2r2;
#pragma unroll
      for (int d = 0; d < HMAX; d++)
        if (d < a.hyper) g[NG_FIXED + d] = -(dd * 2.0f * (fh[d] - __ldg(a.nodes + (size_t)mm * nstride + 3 + d)));
    }
  };

  // coherent?  count the distinct nearest nodes of the warp
  const uint32_t same = __match_any_sync(FULL, idx[0]);
  const int distinct = __popc(__ballot_sync(FULL, (__ffs(same) - 1) == lane));
  if (distinct > MAX_DISTINCT) {
    if (live) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        float g[32];
        pair_row(idx[k], wk[k], du[k], dist[k], g);
#pragma unroll
        for (int c = 0; c < 32; c++) {
          if (c < NG && g[c] != 0.f) { float* q = dst_of(idx[k], c); if (q) atomicAdd(q, g[c]); }
        }
      }
    }
    return;
  }
  uint32_t todo = live ? ((1u << K) - 1u) : 0u;
  while (true) {
    const uint32_t have = __ballot_sync(FULL, todo != 0u);
    if (have == 0u) break;
    const int leader = __ffs(have) - 1;
    int first_m = -1;
#pragma unroll
    for (int k = K - 1; k >= 0; k--)
      if ((todo >> k) & 1u) first_m = idx[k];
    const int mm = __shfl_sync(FULL, first_m, leader);
    float w = 0.f, duk = 0.f, dk = 0.f;
    uint32_t bit = 0u;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const bool hit = ((todo >> k) & 1u) && idx[k] == mm;
      w = hit ? wk[k] : w; duk = hit ? du[k] : duk; dk = hit ? dist[k] : dk;
      bit = hit ? (1u << k) : bit;
    }
    todo &= ~bit;
    float g[32];
    pair_row(mm, w, duk, dk, g);
    warp_transpose_reduce32(g, lane);
    if (flush_base != nullptr && g[0] != 0.f) atomicAdd(flush_base + (size_t)flush_stride * mm, g[0]);
  }
}

int deform_forward_launch(const DeformFwdHost& h, cudaStream_t s, const char** err) {
  if (h.K < 1 || h.K > MAX_K || h.K > h.M) { *err = "K must be in [1, min(8, M)]"; return -1; }
  if (h.hyper < 0 || h.hyper > MAX_D - 3) { *err = "hyper_dim must be <= 16"; return -1; }
  if (h.P == 0) return 0;
  DeformFwdP a{};
  a.P = h.P; a.M = h.M; a.K = h.K;
  a.hyper = h.hyper;                       // stride of the node table is always 3 + hyper_dim
  a.D = h.feature ? 3 + h.hyper : 3;       // the query is 3-D when no hyper feature is given
  a.xyz = h.xyz; a.feature = h.feature; a.fstride = h.fstride;
  a.nodes = h.nodes; a.radius_log = h.radius_log; a.weight_logit = h.weight_logit;
  a.trans = h.trans; a.rot = h.rot; a.scale = h.scale; a.local_rot = h.local_rot; a.mask = h.mask;
  a.nn_idx = h.nn_idx; a.nn_dist = h.nn_dist; a.nn_weight = h.nn_weight;
  a.d_xyz = h.d_xyz; a.d_rot = h.d_rot; a.d_scale = h.d_scale;
  a.st_t = h.attr_stride > 0 ? h.attr_stride : 3; a.st_r = h.attr_stride > 0 ? h.attr_stride : 4;
  a.st_s = h.attr_stride > 0 ? h.attr_stride : 2; a.st_l = h.attr_stride > 0 ? h.attr_stride : 4;
  a.order = h.order;
  a.filter = g_knn_filter;
  const int nq = (a.D + 3) / 4;
  const size_t smem = sizeof(float4) * (size_t)((a.M + 31) & ~31) * nq;
  if (smem > 200 * 1024) { *err = "node table exceeds shared memory (M*(3+hyper) floats > 200 KB)"; return -1; }
  // persistent warps: as many CTAs as are resident at once (the node table is staged once per CTA), capped by the work
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int chunks = (h.P + 255) / 256;
#define D2GS_KNN_LAUNCH(KK, QQ)                                                                                  \
  do {                                                                                                          \
    cudaFuncSetAttribute(deform_fwd_kernel<KK, QQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    int per_sm = 1;                                                                                             \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, deform_fwd_kernel<KK, QQ>, 256, smem);               \
    const int grid = min(chunks, sms * max(per_sm, 1));                                                         \
    deform_fwd_kernel<KK, QQ><<<grid, 256, smem, s>>>(a);                                                        \
  } while (0)
#define D2GS_KNN_Q(KK)                                                                     \
  switch (nq) {                                                                           \
    case 1: D2GS_KNN_LAUNCH(KK, 1); break;                                                \
    case 2: D2GS_KNN_LAUNCH(KK, 2); break;                                                \
    case 3: D2GS_KNN_LAUNCH(KK, 3); break;                                                \
    case 4: D2GS_KNN_LAUNCH(KK, 4); break;                                                \
    default: D2GS_KNN_LAUNCH(KK, 5); break;                                               \
  }
  switch (h.K) {
    case 1: D2GS_KNN_Q(1) break;
    case 2: D2GS_KNN_Q(2) break;
    case 3: D2GS_KNN_Q(3) break;
    case 4: D2GS_KNN_Q(4) break;
    case 5: D2GS_KNN_Q(5) break;
    case 6: D2GS_KNN_Q(6) break;
    case 7: D2GS_KNN_Q(7) break;
    default: D2GS_KNN_Q(8) break;
  }
#undef D2GS_KNN_Q
#undef D2GS_KNN_LAUNCH
  return 0;
}

int deform_backward_launch(const DeformBwdHost& h, cudaStream_t s, const char** err) {
  DeformBwdP a{};
  a.P = h.P; a.M = h.M; a.K = h.K; a.hyper = h.hyper; a.D = 3 + h.hyper;
  a.xyz = h.xyz; a.feature = h.feature; a.fstride = h.fstride;
  a.nodes = h.nodes; a.radius_log = h.radius_log; a.weight_logit = h.weight_logit;
  a.trans = h.trans; a.rot = h.rot; a.scale = h.scale; a.local_rot = h.local_rot; a.mask = h.mask;
  a.nn_idx = h.nn_idx; a.nn_dist = h.nn_dist; a.nn_weight = h.nn_weight;
  a.g_xyz = h.g_xyz; a.g_rot = h.g_rot; a.g_scale = h.g_scale;
  a.d_trans = h.d_trans; a.d_rot = h.d_rot; a.d_scale = h.d_scale; a.d_local_rot = h.d_local_rot;
  a.d_nodes = h.d_nodes; a.d_radius_log = h.d_radius_log; a.d_weight_logit = h.d_weight_logit;
  a.d_feature = h.d_feature; a.d_mask = h.d_mask;
  a.st_t = h.attr_stride > 0 ? h.attr_stride : 3; a.st_r = h.attr_stride > 0 ? h.attr_stride : 4;
  a.st_s = h.attr_stride > 0 ? h.attr_stride : 2; a.st_l = h.attr_stride > 0 ? h.attr_stride : 4;
  if (h.K < 1 || h.K > MAX_K) { *err = "K must be in [1, 8]"; return -1; }
  if (h.hyper < 0 || h.hyper > MAX_D - 3) { *err = "hyper_dim must be <= 16"; return -1; }
  if (h.P == 0) return 0;
  a.order = h.order;
  if (h.order && NG_FIXED + h.hyper <= 32) {
    const int grid = (h.P + 255) / 256;
    switch (h.K) {
      case 1: if (a.hyper <= 8) deform_bwd_coherent_kernel<1, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<1, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      case 2: if (a.hyper <= 8) deform_bwd_coherent_kernel<2, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<2, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      case 3: if (a.hyper <= 8) deform_bwd_coherent_kernel<3, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<3, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      case 4: if (a.hyper <= 8) deform_bwd_coherent_kernel<4, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<4, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      case 5: if (a.hyper <= 8) deform_bwd_coherent_kernel<5, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<5, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      case 6: if (a.hyper <= 8) deform_bwd_coherent_kernel<6, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<6, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      case 7: if (a.hyper <= 8) deform_bwd_coherent_kernel<7, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<7, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
      default: if (a.hyper <= 8) deform_bwd_coherent_kernel<8, 8><<<grid, 256, 0, s>>>(a); else deform_bwd_coherent_kernel<8, MAX_D - 3><<<grid, 256, 0, s>>>(a); break;
    }
    return 0;
  }
  const size_t smem = sizeof(float) * (size_t)h.M * (NG_FIXED + h.hyper);
  a.use_smem = g_deform_bwd_smem && smem <= 200 * 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = (h.P + 255) / 256;
  if (a.use_smem) {
    cudaFuncSetAttribute(deform_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int per_sm = smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1);
    blocks = min(blocks, sms * per_sm);
  }
  deform_bwd_kernel<<<blocks, 256, a.use_smem ? smem : 0, s>>>(a);
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// processing order: 30-bit Morton keys of the surfel centres inside their bounding box (sorted by the caller with CUB)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) order_bbox_kernel(int P, const float* __restrict__ xyz, unsigned int* __restrict__ bbox) {
  unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float v = xyz[3 * (size_t)i + c];
      if (v == v && fabsf(v) <= 3.0e38f) {   // NaN / inf centres do not stretch the box
        const unsigned int u = float_to_ordered(v);
        lo[c] = min(lo[c], u); hi[c] = max(hi[c], u);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    lo[c] = __reduce_min_sync(FULL, lo[c]);
    hi[c] = __reduce_max_sync(FULL, hi[c]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) { atomicMin(bbox + c, lo[c]); atomicMax(bbox + 3 + c, hi[c]); }
  }
}
__device__ __forceinline__ unsigned int spread10(unsigned int v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void __launch_bounds__(256) order_keys_kernel(int P, const float* __restrict__ xyz, const unsigned int* __restrict__ bbox,
                                                         unsigned int* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  unsigned int q[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float lo = ordered_to_float(bbox[c]), hi = ordered_to_float(bbox[3 + c]);
    const float v = xyz[3 * (size_t)i + c];
    const float ext = hi - lo;
    float u = (ext > 0.f && v == v) ? (v - lo) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    q[c] = min(1023u, (unsigned int)(u * 1024.f));
  }
  keys[i] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
  vals[i] = i;
}
void deform_order_keys_launch(int P, const float* xyz, unsigned int* bbox, unsigned int* keys, int* vals, cudaStream_t s) {
  if (P <= 0) return;
  cudaMemsetAsync(bbox, 0xff, 3 * sizeof(unsigned int), s);       // running minima (ordered-uint encoding)
  cudaMemsetAsync(bbox + 3, 0x00, 3 * sizeof(unsigned int), s);   // running maxima
  order_bbox_kernel<<<min((P + 255) / 256, 592), 256, 0, s>>>(P, xyz, bbox);
  order_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, xyz, bbox, keys, vals);
}

}  // namespace d2gs
