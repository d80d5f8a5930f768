// Fused deformation MLP (DeformNetwork, utils/time_utils.py:310-453) for sm_100a: the whole network — positional
// embeddings, timenet, the 8x256 trunk with its skip connection and the output heads — is ONE kernel forward and two
// kernels backward instead of ~60 + ~120 eager launches (13-14 cuBLAS SGEMMs on M = 512 rows each way).
//
// Why SIMT fp32 and not tcgen05: the reference tolerance is 1e-4 relative against fp32 GEMMs; TF32/BF16 tensor-core
// inputs miss it, a 3xTF32 split would need 128-row UMMA tiles (4 CTAs at M=512) for 0.5 GFLOP of work, and the
// kernel is bound by streaming the 2.1 MB of weights through each SM, not by FMA rate.
//
// Layout: every CTA owns ROWS=4 rows; activations live in shared memory as [k][ROWS] so one broadcast LDS.128 feeds
// 4 FMAs; thread n owns output feature n and streams W^T[k][n] (transposed once per step by mlp_transpose_kernel)
// with coalesced 128-B warp loads.  Backward (dAct) streams W[n][k] in its native layout, thread k owning input k.
// dW/db are a separate kernel tiled over (layer, 64x64) with the reduction over rows.
#include "raster_common.cuh"
#include "mlp.cuh"

namespace d2gs {

constexpr int MW = 256;          // trunk width
constexpr int MD = 8;            // trunk depth
constexpr int SKIP = 4;          // concat [x_emb, t_emb, h] after layer SKIP
constexpr int EX = 63;           // 3 + 3*2*10
constexpr int ROWS = 4;
constexpr int INP_LD = 96;       // padded leading dimension of the [x_emb, t_feat] block (<= 93 used)

// ---------------------------------------------------------------------------------------------------------------
// weight transpose: W (N x K, row-major) -> W^T (K x NP), NP = N rounded up to 4
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mlp_transpose_kernel(MlpLayers L) {
  __shared__ float tile[32][33];
  const int layer = blockIdx.z;
  if (layer >= L.count) return;
  const MlpLayer& ly = L.layer[layer];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  if (n0 >= ly.N || k0 >= ly.K) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    tile[r][tx] = (n < ly.N && k < ly.K) ? ly.W[(size_t)n * ly.K + k] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    if (k < ly.K && n < ly.NP) L.wt[ly.wt_off + (size_t)k * ly.NP + n] = (n < ly.N) ? tile[tx][r] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// out[n] (for the CTA's ROWS rows) = bias[n] + sum_k in[k] * Wt[k][n];   in: shared [K][ROWS]
template <bool RELU>
__device__ __forceinline__ void dense(const float* __restrict__ wt, int NP, const float* __restrict__ bias, int N,
                                      const float4* s_in_a, int Ka, const float4* s_in_b, int Kb, float4* s_out,
                                      float* g_out, int ld_out, int row0, int rows, int tid) {
  for (int n = tid; n < N; n += 256) {
    const float b = bias ? __ldg(bias + n) : 0.f;
    float a0 = b, a1 = b, a2 = b, a3 = b;
    const float* w = wt + n;
    int k = 0;
#pragma unroll 1
    for (; k + 8 <= Ka; k += 8) {
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) wv[u] = __ldg(w + (size_t)(k + u) * NP);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const float4 x = s_in_a[k + u];
        a0 = fmaf(wv[u], x.x, a0); a1 = fmaf(wv[u], x.y, a1); a2 = fmaf(wv[u], x.z, a2); a3 = fmaf(wv[u], x.w, a3);
      }
    }
    for (; k < Ka; k++) {
      const float wv = __ldg(w + (size_t)k * NP);
      const float4 x = s_in_a[k];
      a0 = fmaf(wv, x.x, a0); a1 = fmaf(wv, x.y, a1); a2 = fmaf(wv, x.z, a2); a3 = fmaf(wv, x.w, a3);
    }
    if (Kb > 0) {
      const float* w2 = w + (size_t)Ka * NP;
      int kk = 0;
#pragma unroll 1
      for (; kk + 8 <= Kb; kk += 8) {
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) wv[u] = __ldg(w2 + (size_t)(kk + u) * NP);
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float4 x = s_in_b[kk + u];
          a0 = fmaf(wv[u], x.x, a0); a1 = fmaf(wv[u], x.y, a1); a2 = fmaf(wv[u], x.z, a2); a3 = fmaf(wv[u], x.w, a3);
        }
      }
      for (; kk < Kb; kk++) {
        const float wv = __ldg(w2 + (size_t)kk * NP);
        const float4 x = s_in_b[kk];
        a0 = fmaf(wv, x.x, a0); a1 = fmaf(wv, x.y, a1); a2 = fmaf(wv, x.z, a2); a3 = fmaf(wv, x.w, a3);
      }
    }
    if (RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
    if (s_out) s_out[n] = make_float4(a0, a1, a2, a3);
    if (g_out) {
      if (rows > 0) g_out[(size_t)(row0 + 0) * ld_out + n] = a0;
      if (rows > 1) g_out[(size_t)(row0 + 1) * ld_out + n] = a1;
      if (rows > 2) g_out[(size_t)(row0 + 2) * ld_out + n] = a2;
      if (rows > 3) g_out[(size_t)(row0 + 3) * ld_out + n] = a3;
    }
  }
}

__global__ void __launch_bounds__(256) mlp_fwd_kernel(MlpFwd a) {
  __shared__ float4 s_inp[INP_LD];       // [x_emb (63), t feature (Tt)]
  __shared__ float4 s_te[32];            // time embedding (Et <= 21)
  __shared__ float4 s_h[2][MW];          // ping-pong hidden / timenet hidden
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * ROWS;
  const int rows = min(ROWS, a.rows - row0);
  const MlpLayers& L = a.layers;
  const int Et = a.Et, Tt = a.Tt;

  // positional embeddings: [v, sin(2^0 v), cos(2^0 v), ..., sin(2^(F-1) v), cos(2^(F-1) v)] per input dimension block
  for (int e = tid; e < EX + Et; e += 256) {
    float v[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      float val = 0.f;
      if (r < rows) {
        if (e < EX) {
          const int blk = e / 3, c = e - 3 * blk;          // block 0 = identity, then (sin, cos) pairs per frequency
          const float x = a.x[(size_t)(row0 + r) * 3 + c];
          if (blk == 0) val = x;
          else { const float f = exp2f((float)((blk - 1) >> 1)); val = ((blk - 1) & 1) ? cosf(x * f) : sinf(x * f); }
        } else {
          const int blk = e - EX;
          const float t = a.t[(size_t)(row0 + r) * a.t_stride];
          if (blk == 0) val = t;
          else { const float f = exp2f((float)((blk - 1) >> 1)); val = ((blk - 1) & 1) ? cosf(t * f) : sinf(t * f); }
        }
      }
      v[r] = val;
    }
    const float4 p = make_float4(v[0], v[1], v[2], v[3]);
    if (e < EX) s_inp[e] = p;
    else {
      s_te[e - EX] = p;
      if (!a.has_timenet) s_inp[e] = p;
    }
    if (a.save_inp) {
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        if (r < rows) {
          if (e < EX) a.save_inp[(size_t)(row0 + r) * INP_LD + e] = v[r];
          else { a.save_te[(size_t)(row0 + r) * 32 + (e - EX)] = v[r]; if (!a.has_timenet) a.save_inp[(size_t)(row0 + r) * INP_LD + e] = v[r]; }
        }
    }
  }
  __syncthreads();
  int li = 0;
  if (a.has_timenet) {
    const MlpLayer& t1 = L.layer[li++];
    dense<true>(L.wt + t1.wt_off, t1.NP, t1.b, t1.N, s_te, t1.K, nullptr, 0, s_h[0], a.save_th, MW, row0, rows, tid);
    __syncthreads();
    const MlpLayer& t2 = L.layer[li++];
    dense<false>(L.wt + t2.wt_off, t2.NP, t2.b, t2.N, s_h[0], t2.K, nullptr, 0, s_inp + EX, a.save_inp ? a.save_inp + EX : nullptr,
                 INP_LD, row0, rows, tid);
    __syncthreads();
  }
  const int in0 = EX + Tt;
  int cur = 0;
  for (int l = 0; l < MD; l++) {
    const MlpLayer& ly = L.layer[li++];
    float* save = a.save_h ? a.save_h + (size_t)l * a.rows * MW : nullptr;
    if (l == 0) dense<true>(L.wt + ly.wt_off, ly.NP, ly.b, ly.N, s_inp, in0, nullptr, 0, s_h[cur], save, MW, row0, rows, tid);
    else if (l == SKIP + 1) dense<true>(L.wt + ly.wt_off, ly.NP, ly.b, ly.N, s_inp, in0, s_h[cur ^ 1], MW, s_h[cur], save, MW, row0, rows, tid);
    else dense<true>(L.wt + ly.wt_off, ly.NP, ly.b, ly.N, s_h[cur ^ 1], MW, nullptr, 0, s_h[cur], save, MW, row0, rows, tid);
    __syncthreads();
    cur ^= 1;
  }
  // heads: concatenated outputs [warp 3 | scaling 2 | rotation 4 | local 4 | opacity 1] -> (rows, NH)
  const MlpLayer& hd = L.layer[li];
  dense<false>(L.wt + hd.wt_off, hd.NP, hd.b, hd.N, s_h[cur ^ 1], MW, nullptr, 0, nullptr, a.out, a.NH, row0, rows, tid);
}

// ---------------------------------------------------------------------------------------------------------------
// backward, activation gradients: thread k owns input feature k; g_in[k] = sum_n G[n] * W[n][k]
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 back_dense(const float* __restrict__ W, int K, int N, const float4* s_g, int k) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* w = W + k;
  int n = 0;
#pragma unroll 1
  for (; n + 8 <= N; n += 8) {
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) wv[u] = __ldg(w + (size_t)(n + u) * K);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const float4 g = s_g[n + u];
      a0 = fmaf(wv[u], g.x, a0); a1 = fmaf(wv[u], g.y, a1); a2 = fmaf(wv[u], g.z, a2); a3 = fmaf(wv[u], g.w, a3);
    }
  }
  for (; n < N; n++) {
    const float wv = __ldg(w + (size_t)n * K);
    const float4 g = s_g[n];
    a0 = fmaf(wv, g.x, a0); a1 = fmaf(wv, g.y, a1); a2 = fmaf(wv, g.z, a2); a3 = fmaf(wv, g.w, a3);
  }
  return make_float4(a0, a1, a2, a3);
}

__device__ __forceinline__ void store_rows(float* g, int ld, int row0, int rows, int col, float4 v) {
  if (rows > 0) g[(size_t)(row0 + 0) * ld + col] = v.x;
  if (rows > 1) g[(size_t)(row0 + 1) * ld + col] = v.y;
  if (rows > 2) g[(size_t)(row0 + 2) * ld + col] = v.z;
  if (rows > 3) g[(size_t)(row0 + 3) * ld + col] = v.w;
}
__device__ __forceinline__ float4 load_rows(const float* g, int ld, int row0, int rows, int col) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rows > 0) v.x = g[(size_t)(row0 + 0) * ld + col];
  if (rows > 1) v.y = g[(size_t)(row0 + 1) * ld + col];
  if (rows > 2) v.z = g[(size_t)(row0 + 2) * ld + col];
  if (rows > 3) v.w = g[(size_t)(row0 + 3) * ld + col];
  return v;
}
__device__ __forceinline__ float4 relu_mask(float4 g, float4 h) {
  return make_float4(h.x > 0.f ? g.x : 0.f, h.y > 0.f ? g.y : 0.f, h.z > 0.f ? g.z : 0.f, h.w > 0.f ? g.w : 0.f);
}

__global__ void __launch_bounds__(256) mlp_bwd_act_kernel(MlpBwd a) {
  __shared__ float4 s_g[2][MW];      // pre-activation gradient of the current layer (ping-pong)
  __shared__ float4 s_gh[16];        // head gradients
  __shared__ float4 s_gt[32];        // gradient of the time feature (Tt <= 30)
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * ROWS;
  const int rows = min(ROWS, a.rows - row0);
  const MlpLayers& L = a.layers;
  const int Tt = a.Tt, in0 = EX + Tt;
  const int first_trunk = a.has_timenet ? 2 : 0;
  if (tid < 16) s_gh[tid] = tid < a.NH ? load_rows(a.g_out, a.NH, row0, rows, tid) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < 32) s_gt[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // heads -> h7
  const MlpLayer& hd = L.layer[first_trunk + MD];
  int cur = 0;
  {
    float4 g = back_dense(hd.W, MW, hd.N, s_gh, tid);
    const float4 h = load_rows(a.save_h + (size_t)(MD - 1) * a.rows * MW, MW, row0, rows, tid);
    g = relu_mask(g, h);
    s_g[cur][tid] = g;
    store_rows(a.G + (size_t)(MD - 1) * a.rows * MW, MW, row0, rows, tid, g);
  }
  __syncthreads();
  for (int l = MD - 1; l >= 1; l--) {
    const MlpLayer& ly = L.layer[first_trunk + l];     // input of layer l is h_{l-1} (plus [x,t] for l == SKIP+1)
    const int K = ly.K;
    const int hoff = (l == SKIP + 1) ? in0 : 0;       // column offset of the h block inside the layer input
    float4 g = back_dense(ly.W, K, MW, s_g[cur], hoff + tid);
    const float4 h = load_rows(a.save_h + (size_t)(l - 1) * a.rows * MW, MW, row0, rows, tid);
    g = relu_mask(g, h);
    if (l == SKIP + 1 && tid < Tt) {   // gradient into the time feature through the skip input
      const float4 gt = back_dense(ly.W, K, MW, s_g[cur], EX + tid);
      s_gt[tid] = gt;
    }
    s_g[cur ^ 1][tid] = g;
    store_rows(a.G + (size_t)(l - 1) * a.rows * MW, MW, row0, rows, tid, g);
    __syncthreads();
    cur ^= 1;
  }
  // layer 0 input: only the time feature needs a gradient (node positions are detached)
  {
    const MlpLayer& l0 = L.layer[first_trunk];
    if (tid < Tt) {
      const float4 gt = back_dense(l0.W, l0.K, MW, s_g[cur], EX + tid);
      const float4 o = s_gt[tid];
      s_gt[tid] = make_float4(o.x + gt.x, o.y + gt.y, o.z + gt.z, o.w + gt.w);
    }
  }
  __syncthreads();
  if (a.has_timenet) {
    if (tid < Tt) store_rows(a.g_tfeat, 32, row0, rows, tid, s_gt[tid]);
    const MlpLayer& t2 = L.layer[1];
    float4 g = back_dense(t2.W, MW, t2.N, s_gt, tid);
    const float4 h = load_rows(a.save_th, MW, row0, rows, tid);
    g = relu_mask(g, h);
    store_rows(a.G_t1, MW, row0, rows, tid, g);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, weight gradients: dW[n][k] = sum_r G[r][n] * in[r][k],  db[n] = sum_r G[r][n]
// one CTA per (job, 64x64 tile); 256 threads, 4x4 outputs each; rows streamed 16 at a time through shared memory
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mlp_bwd_w_kernel(MlpWJobs J, int rows) {
  __shared__ float sG[16][64 + 4];
  __shared__ float sA[16][64 + 4];
  int b = blockIdx.x, j = 0;
  while (j < J.count && b >= J.job[j].tiles) { b -= J.job[j].tiles; j++; }
  if (j >= J.count) return;
  const MlpWJob& jb = J.job[j];
  const int Ktot = jb.Ka + jb.Kb;
  const int tk = (Ktot + 63) / 64;
  const int n0 = (b / tk) * 64, k0 = (b % tk) * 64;
  const int tid = threadIdx.x, tn = tid >> 4, tkx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int q = 0; q < 4; q++) acc[i][q] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = 0; r0 < rows; r0 += 16) {
    for (int e = tid; e < 16 * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      const int rr = r0 + r;
      const int n = n0 + c, k = k0 + c;
      sG[r][c] = (rr < rows && n < jb.N) ? jb.G[(size_t)rr * jb.ldG + n] : 0.f;
      float v = 0.f;
      if (rr < rows && k < Ktot) v = (k < jb.Ka) ? jb.A[(size_t)rr * jb.ldA + k] : jb.B[(size_t)rr * jb.ldB + (k - jb.Ka)];
      sA[r][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) {
      float g[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; i++) g[i] = sG[r][tn * 4 + i];
#pragma unroll
      for (int q = 0; q < 4; q++) x[q] = sA[r][tkx * 4 + q];
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) acc[i][q] = fmaf(g[i], x[q], acc[i][q]);
        if (k0 == 0 && tkx == 0) bacc[i] += g[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int n = n0 + tn * 4 + i;
    if (n >= jb.N) continue;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = k0 + tkx * 4 + q;
      if (k < Ktot) jb.dW[(size_t)n * Ktot + k] = acc[i][q];
    }
    if (k0 == 0 && tkx == 0 && jb.db) jb.db[n] = bacc[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
void mlp_launch_transpose(const MlpLayers& L, cudaStream_t s) {
  int maxN = 0, maxK = 0;
  for (int i = 0; i < L.count; i++) { maxN = max(maxN, L.layer[i].N); maxK = max(maxK, L.layer[i].K); }
  dim3 grid((maxK + 31) / 32, (maxN + 31) / 32, L.count);
  mlp_transpose_kernel<<<grid, 256, 0, s>>>(L);
}
void mlp_launch_forward(const MlpFwd& a, cudaStream_t s) {
  if (a.rows <= 0) return;
  mlp_fwd_kernel<<<(a.rows + ROWS - 1) / ROWS, 256, 0, s>>>(a);
}
void mlp_launch_backward(const MlpBwd& a, const MlpWJobs& J, cudaStream_t s) {
  if (a.rows <= 0) return;
  mlp_bwd_act_kernel<<<(a.rows + ROWS - 1) / ROWS, 256, 0, s>>>(a);
  int tiles = 0;
  for (int j = 0; j < J.count; j++) tiles += J.job[j].tiles;
  mlp_bwd_w_kernel<<<tiles, 256, 0, s>>>(J, a.rows);
}

}  // namespace d2gs
