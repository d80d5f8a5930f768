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
// weight streaming: TMA bulk copies (cp.async.bulk, global -> shared) completing on mbarriers
// ---------------------------------------------------------------------------------------------------------------
// Every CTA needs all 2.1 MB of weights once per pass.  A ninth warp (one elected lane) streams them through a ring of
// shared-memory stages with 1-D bulk copies; the eight consumer warps wait on the stage's "full" mbarrier, run their
// FMAs out of shared memory and release the stage through its "empty" mbarrier.  The producer runs ahead across layer
// boundaries, so the next layer's first chunks land while the current layer is reduced and stored.  The copy engine
// keeps the L2 -> SM pipe busy regardless of what the math warps are doing (the LDG version of this kernel was bound
// by load latency: 8 warps x 4 LDG.128 in flight per SM).
#ifndef D2GS_MLP_KGROUPS
#define D2GS_MLP_KGROUPS 4
#endif
constexpr int KG = D2GS_MLP_KGROUPS;         // reduction groups: thread (group, quad) sums the k (or n) with index % KG == group
constexpr int CONSUMERS = 64 * KG;           // math warps: 2 per group (one 116-KB CTA per SM: more warps = more latency hidden)
constexpr int MLP_THREADS = CONSUMERS + 32;  // + 1 producer warp
#ifndef D2GS_MLP_STAGES
#define D2GS_MLP_STAGES 4
#endif
constexpr int STAGES = D2GS_MLP_STAGES;
constexpr int KCH = 32;                      // forward: k-rows of W^T per chunk (32 x NP floats <= 32 KB)
constexpr int NCH = 32;                      // backward: n-rows of W per chunk (32 x K floats <= 44672 B at K = 349)
constexpr int FWD_STAGE_BYTES = KCH * MW * 4;
constexpr int BWD_STAGE_BYTES = 44800;       // >= 32 * 349 * 4, multiple of 128

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 28)) __trap();   // a lost arrival must fail loudly, never hang the device
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }

// Thread-block clusters: the CTAs of a cluster need the SAME weights at the same time, so ONE bulk copy per chunk,
// issued by the cluster's rank-0 CTA with .multicast::cluster, lands in the ring stage of every CTA of the cluster and
// signals each CTA's own "full" mbarrier.  L2 -> SM traffic drops from (CTAs x 2.1 MB) to (clusters x 2.1 MB): at
// 128 CTAs the per-CTA streams added up to 270 MB per pass through an L2 that delivers ~12 TB/s — the floor of the
// un-clustered kernel was ~22 us whatever the math did.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // relaxed: the shared-memory reads of the stage were consumed by the FMAs in front of this instruction, and the stage is
  // overwritten through the async proxy only after rank 0 has OBSERVED the arrival (a .release.cluster arrive measured
  // ~0.3 us each here — 8 warps x 85 chunks of them doubled the kernel)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 28)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}

// Barriers per CTA: full[STAGES] (count 1: the local producer's arrive.expect_tx; the bytes come from rank 0's multicast),
// empty[STAGES] (count 8: this CTA's consumer warps), cempty[STAGES] (count 8 x cluster size; only rank 0's copy is used:
// every consumer warp of the cluster arrives on it remotely, and rank 0's producer waits on it before it overwrites the
// stage in ALL CTAs).
struct Ring {
  uint32_t stage0, stage_bytes, full0, empty0, cempty0;   // shared-window addresses
  uint32_t rank, csize, cempty_rank0;                     // cluster rank / size, rank 0's cempty[0] as a cluster address
  int g;                                                  // running chunk counter (same sequence on both sides)
  __device__ __forceinline__ uint32_t stage(int st) const { return stage0 + (uint32_t)st * stage_bytes; }
  __device__ __forceinline__ uint32_t full(int st) const { return full0 + 8u * st; }
  __device__ __forceinline__ uint32_t empty(int st) const { return empty0 + 8u * st; }
  __device__ __forceinline__ uint32_t cempty(int st) const { return cempty0 + 8u * st; }
};
__device__ __forceinline__ void ring_setup(Ring& r, unsigned char* stages, uint32_t stage_bytes, uint64_t* bars, int tid) {
  r.stage0 = smem_addr(stages); r.stage_bytes = stage_bytes;
  r.full0 = smem_addr(bars); r.empty0 = smem_addr(bars + STAGES); r.cempty0 = smem_addr(bars + 2 * STAGES); r.g = 0;
  r.rank = cluster_rank(); r.csize = cluster_size();
  r.cempty_rank0 = map_to_rank(r.cempty0, 0);
  if (tid == 0) {
    for (int i = 0; i < STAGES; i++) {
      mbar_init(r.full(i), 1); mbar_init(r.empty(i), CONSUMERS / 32); mbar_init(r.cempty(i), (CONSUMERS / 32) * r.csize);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  cluster_sync_all();      // every CTA's barriers are initialised before any remote arrive / multicast can reach them
}
// producer side (one thread per CTA): one chunk
__device__ __forceinline__ void ring_push(Ring& r, const void* src, uint32_t bytes) {
  const int st = r.g % STAGES; const uint32_t par = (uint32_t)(r.g / STAGES) & 1u;
  mbar_wait(r.empty(st), par ^ 1u);              // this CTA's consumers have left the stage
  mbar_expect_tx(r.full(st), bytes);             // arm the local "full" barrier for the multicast bytes
  if (r.csize == 1) {
    bulk_g2s(r.stage(st), src, bytes, r.full(st));
  } else if (r.rank == 0) {
    mbar_wait_cluster(r.cempty(st), par ^ 1u);   // ... and so have the consumers of every CTA of the cluster
    bulk_g2s_multicast(r.stage(st), src, bytes, r.full(st), (uint16_t)((1u << r.csize) - 1u));
  }
  r.g++;
}
// consumer side: returns the stage index once its bytes have landed; release with ring_pop
__device__ __forceinline__ int ring_front(Ring& r) {
  const int st = r.g % STAGES; const uint32_t par = (uint32_t)(r.g / STAGES) & 1u;
  mbar_wait(r.full(st), par);
  return st;
}
__device__ __forceinline__ void ring_pop(Ring& r, int st, int lane) {
  __syncwarp();
  if (lane == 0) { mbar_arrive(r.empty(st)); if (r.csize > 1) mbar_arrive_remote(r.cempty_rank0 + 8u * st); }
  r.g++;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// out[n] (for the CTA's ROWS rows) = bias[n] + sum_k in[k] * Wt[k][n];   in: shared [K][ROWS]
// Thread (kg, nq) = (tid >> 6, tid & 63) owns outputs 4nq..4nq+3 for the k with k % 4 == kg: one conflict-free LDS.128 of
// the staged W^T chunk and one broadcast LDS.128 of the activations feed 16 FMAs; the four k-groups are summed through
// shared memory.
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void fma16(float4 (&acc)[4], const float4 wv, const float4 xv) {
  acc[0].x = fmaf(wv.x, xv.x, acc[0].x); acc[0].y = fmaf(wv.x, xv.y, acc[0].y); acc[0].z = fmaf(wv.x, xv.z, acc[0].z); acc[0].w = fmaf(wv.x, xv.w, acc[0].w);
  acc[1].x = fmaf(wv.y, xv.x, acc[1].x); acc[1].y = fmaf(wv.y, xv.y, acc[1].y); acc[1].z = fmaf(wv.y, xv.z, acc[1].z); acc[1].w = fmaf(wv.y, xv.w, acc[1].w);
  acc[2].x = fmaf(wv.z, xv.x, acc[2].x); acc[2].y = fmaf(wv.z, xv.y, acc[2].y); acc[2].z = fmaf(wv.z, xv.z, acc[2].z); acc[2].w = fmaf(wv.z, xv.w, acc[2].w);
  acc[3].x = fmaf(wv.w, xv.x, acc[3].x); acc[3].y = fmaf(wv.w, xv.y, acc[3].y); acc[3].z = fmaf(wv.w, xv.z, acc[3].z); acc[3].w = fmaf(wv.w, xv.w, acc[3].w);
}

// sum of the KG partial results of output `col` (pairwise, fixed order)
__device__ __forceinline__ float4 sum_groups(const float4* s_red, int col) {
  float4 p[KG];
#pragma unroll
  for (int g = 0; g < KG; g++) p[g] = s_red[g * MW + col];
#pragma unroll
  for (int st = 1; st < KG; st <<= 1)
#pragma unroll
    for (int g = 0; g + st < KG; g += 2 * st) { p[g].x += p[g + st].x; p[g].y += p[g + st].y; p[g].z += p[g + st].z; p[g].w += p[g + st].w; }
  return p[0];
}

template <bool RELU>
__device__ __forceinline__ void dense(Ring& ring, int NP, const float* __restrict__ bias, int N, const float4* s_in, int K,
                                      float4* s_red /*[4][256]*/, float4* s_out, float* g_out, int ld_out, int row0,
                                      int rows, int tid) {
  const int kg = tid >> 6, nq = tid & 63, lane = tid & 31;
  const int np4 = NP >> 2;
  const bool active = nq < np4;
  float4 acc[4];
#pragma unroll
  for (int j = 0; j < 4; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = 0; k0 < K; k0 += KCH) {
    const int st = ring_front(ring);
    const uint32_t w = ring.stage(st) + 16u * (uint32_t)nq;
    const int kr = min(KCH, K - k0);
    if (active) {
      if (kr == KCH) {
#pragma unroll
        for (int kk = 0; kk < KCH / KG; kk++) {
          const int k = kg + KG * kk;
          fma16(acc, lds128(w + 16u * (uint32_t)(k * np4)), s_in[k0 + k]);
        }
      } else {
        for (int k = kg; k < kr; k += KG) fma16(acc, lds128(w + 16u * (uint32_t)(k * np4)), s_in[k0 + k]);
      }
    }
    ring_pop(ring, st, lane);
  }
#pragma unroll
  for (int j = 0; j < 4; j++) s_red[kg * MW + 4 * nq + j] = acc[j];
  consumer_sync();
  if (tid < N) {
    const float b = bias ? __ldg(bias + tid) : 0.f;
    const float4 ps = sum_groups(s_red, tid);
    float a0 = b + ps.x, a1 = b + ps.y, a2 = b + ps.z, a3 = b + ps.w;
    if (RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
    if (s_out) s_out[tid] = make_float4(a0, a1, a2, a3);
    if (g_out) {
      if (rows > 0) g_out[(size_t)(row0 + 0) * ld_out + tid] = a0;
      if (rows > 1) g_out[(size_t)(row0 + 1) * ld_out + tid] = a1;
      if (rows > 2) g_out[(size_t)(row0 + 2) * ld_out + tid] = a2;
      if (rows > 3) g_out[(size_t)(row0 + 3) * ld_out + tid] = a3;
    }
  }
  consumer_sync();
}

struct FwdSmem {
  float4 x[2][INP_LD + MW];  // each buffer: [x_emb (63) | time feature (Tt)] at 0..in0-1, hidden at in0..in0+255
  float4 te[32];             // time embedding (Et <= 21)
  float4 th[MW];             // timenet hidden
  float4 red[KG * MW];       // partial sums of the k-groups
  uint64_t bars[3 * STAGES];
};
constexpr size_t FWD_SMEM_BYTES = (size_t)STAGES * FWD_STAGE_BYTES + sizeof(FwdSmem);

__global__ void __launch_bounds__(MLP_THREADS) mlp_fwd_kernel(MlpFwd a) {
  extern __shared__ __align__(128) unsigned char mlp_smem[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(mlp_smem + (size_t)STAGES * FWD_STAGE_BYTES);
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * ROWS;
  const int rows = max(0, min(ROWS, a.rows - row0));     // 0 for a CTA that only pads its cluster
  const MlpLayers& L = a.layers;
  const int Et = a.Et, Tt = a.Tt;
  const int in0 = EX + Tt;
  Ring ring;
  ring_setup(ring, mlp_smem, FWD_STAGE_BYTES, S.bars, tid);

  if (tid >= CONSUMERS) {   // producer warp: stream W^T of every layer, in layer order, KCH rows per chunk
    if (tid == CONSUMERS) {
      for (int li = 0; li < L.count; li++) {
        const MlpLayer& ly = L.layer[li];
        const float* src = L.wt + ly.wt_off;
        for (int k0 = 0; k0 < ly.K; k0 += KCH)
          ring_push(ring, src + (size_t)k0 * ly.NP, (uint32_t)(min(KCH, ly.K - k0) * ly.NP) * 4u);
      }
    }
    __syncwarp();
    cluster_sync_all();     // no CTA of the cluster leaves while copies or remote arrivals may still target it
    return;
  }
  float4 (*s_x)[INP_LD + MW] = S.x;
  float4* s_te = S.te; float4* s_th = S.th; float4* s_red = S.red;

  // positional embeddings: [v, sin(2^0 v), cos(2^0 v), ..., sin(2^(F-1) v), cos(2^(F-1) v)] per input dimension block
  for (int e = tid; e < EX + Et; e += CONSUMERS) {
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
    if (e < EX) { s_x[0][e] = p; s_x[1][e] = p; }
    else {
      s_te[e - EX] = p;
      if (!a.has_timenet) { s_x[0][e] = p; s_x[1][e] = p; }
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
  consumer_sync();
  int li = 0;
  if (a.has_timenet) {
    const MlpLayer& t1 = L.layer[li++];
    dense<true>(ring, t1.NP, t1.b, t1.N, s_te, t1.K, s_red, s_th, a.save_th, MW, row0, rows, tid);
    const MlpLayer& t2 = L.layer[li++];
    dense<false>(ring, t2.NP, t2.b, t2.N, s_th, t2.K, s_red, s_x[0] + EX, a.save_inp ? a.save_inp + EX : nullptr,
                 INP_LD, row0, rows, tid);
    if (tid < Tt) s_x[1][EX + tid] = s_x[0][EX + tid];
    consumer_sync();
  }
  int cur = 0;   // layer l writes its hidden block into s_x[cur] + in0 and reads from s_x[cur ^ 1]
  for (int l = 0; l < MD; l++) {
    const MlpLayer& ly = L.layer[li++];
    float* save = a.save_h ? a.save_h + (size_t)l * a.rows * MW : nullptr;
    const float4* in = (l == 0 || l == SKIP + 1) ? s_x[cur ^ 1] : s_x[cur ^ 1] + in0;   // skip layer reads [x, t, h] contiguously
    dense<true>(ring, ly.NP, ly.b, ly.N, in, ly.K, s_red, s_x[cur] + in0, save, MW, row0, rows, tid);
    cur ^= 1;
  }
  // heads: concatenated outputs [warp 3 | scaling 2 | rotation 4 | local 4 | opacity 1] -> (rows, NH)
  const MlpLayer& hd = L.layer[li];
  dense<false>(ring, hd.NP, hd.b, hd.N, s_x[cur ^ 1] + in0, MW, s_red, nullptr, a.out, a.NH, row0, rows, tid);
  cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------------
// backward, activation gradients: g_in[k] = sum_n G[n] * W[n][k], W streamed in its native (N, K) layout, NCH rows a chunk
// ---------------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void fma4(float4& acc, float w, const float4 g) {
  acc.x = fmaf(w, g.x, acc.x); acc.y = fmaf(w, g.y, acc.y); acc.z = fmaf(w, g.z, acc.z); acc.w = fmaf(w, g.w, acc.w);
}

// K == 256: thread (ng, kq) = (tid >> 6, tid & 63) owns inputs 4kq..4kq+3 for the n with n % 4 == ng (one LDS.128 of the
// staged W rows + one broadcast LDS.128 of G feed 16 FMAs); the four n-groups are summed through shared memory and
// thread tid returns g_in[tid].
__device__ __forceinline__ float4 back_dense_k256(Ring& ring, int N, const float4* s_g, float4* s_red, int tid) {
  const int ng = tid >> 6, kq = tid & 63, lane = tid & 31;
  float4 acc[4];
#pragma unroll
  for (int j = 0; j < 4; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n0 = 0; n0 < N; n0 += NCH) {
    const int st = ring_front(ring);
    const uint32_t w = ring.stage(st) + 16u * (uint32_t)kq;
    const int nr = min(NCH, N - n0);
    if (nr == NCH) {
#pragma unroll
      for (int nn = 0; nn < NCH / KG; nn++) {
        const int n = ng + KG * nn;
        fma16(acc, lds128(w + 16u * (uint32_t)(n * (MW / 4))), s_g[n0 + n]);
      }
    } else {
      for (int n = ng; n < nr; n += KG) fma16(acc, lds128(w + 16u * (uint32_t)(n * (MW / 4))), s_g[n0 + n]);
    }
    ring_pop(ring, st, lane);
  }
#pragma unroll
  for (int j = 0; j < 4; j++) s_red[ng * MW + 4 * kq + j] = acc[j];
  consumer_sync();
  const float4 ps = sum_groups(s_red, tid & (MW - 1));      // threads >= 256 read a valid column and discard the value
  consumer_sync();
  return ps;
}
// general K (the skip layer, K = 93 + 256, and layer 0, K = 93; rows are not 16-B aligned): thread tid accumulates input
// column `col_main + tid` (when col_main >= 0) and, for tid < Tt, the time-feature column EX + tid.
__device__ __forceinline__ void back_dense_cols(Ring& ring, int K, int N, const float4* s_g, int col_main, int Tt, int tid,
                                                float4& g_main, float4& g_time) {
  const int lane = tid & 31;
  g_main = make_float4(0.f, 0.f, 0.f, 0.f);
  g_time = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool do_t = tid < Tt;
  for (int n0 = 0; n0 < N; n0 += NCH) {
    const int st = ring_front(ring);
    const uint32_t base = ring.stage(st);
    const int nr = min(NCH, N - n0);
    if (col_main >= 0 && tid < MW) {
      const uint32_t w = base + 4u * (uint32_t)(col_main + tid);
#pragma unroll 8
      for (int n = 0; n < nr; n++) fma4(g_main, lds32(w + 4u * (uint32_t)(n * K)), s_g[n0 + n]);
    }
    if (do_t) {
      const uint32_t w = base + 4u * (uint32_t)(EX + tid);
#pragma unroll 8
      for (int n = 0; n < nr; n++) fma4(g_time, lds32(w + 4u * (uint32_t)(n * K)), s_g[n0 + n]);
    }
    ring_pop(ring, st, lane);
  }
}

struct BwdSmem {
  float4 g[2][MW];      // pre-activation gradient of the current layer (ping-pong)
  float4 gh[16];        // head gradients
  float4 gt[32];        // gradient of the time feature (Tt <= 30)
  float4 red[KG * MW];
  uint64_t bars[3 * STAGES];
};
constexpr size_t BWD_SMEM_BYTES_MLP = (size_t)STAGES * BWD_STAGE_BYTES + sizeof(BwdSmem);

__global__ void __launch_bounds__(MLP_THREADS) mlp_bwd_act_kernel(MlpBwd a) {
  extern __shared__ __align__(128) unsigned char mlp_smem[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(mlp_smem + (size_t)STAGES * BWD_STAGE_BYTES);
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * ROWS;
  const int rows = max(0, min(ROWS, a.rows - row0));     // 0 for a CTA that only pads its cluster
  const MlpLayers& L = a.layers;
  const int Tt = a.Tt, in0 = EX + Tt;
  const int first_trunk = a.has_timenet ? 2 : 0;
  Ring ring;
  ring_setup(ring, mlp_smem, BWD_STAGE_BYTES, S.bars, tid);

  if (tid >= CONSUMERS) {   // producer warp: heads, linear.7 .. linear.0, timenet.2 — W rows in chunks of NCH
    if (tid == CONSUMERS) {
      for (int q = 0; q < MD + 1 + (a.has_timenet ? 1 : 0); q++) {
        const int li = q == 0 ? first_trunk + MD : (q <= MD ? first_trunk + MD - q : 1);
        const MlpLayer& ly = L.layer[li];
        for (int n0 = 0; n0 < ly.N; n0 += NCH)
          ring_push(ring, ly.W + (size_t)n0 * ly.K, (uint32_t)(min(NCH, ly.N - n0) * ly.K) * 4u);
      }
    }
    __syncwarp();
    cluster_sync_all();
    return;
  }
  float4 (*s_g)[MW] = S.g;
  float4* s_gh = S.gh; float4* s_gt = S.gt; float4* s_red = S.red;
  if (tid < 16) s_gh[tid] = tid < a.NH ? load_rows(a.g_out, a.NH, row0, rows, tid) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < 32) s_gt[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  consumer_sync();
  // heads -> h7
  const MlpLayer& hd = L.layer[first_trunk + MD];
  int cur = 0;
  {
    float4 g = back_dense_k256(ring, hd.N, s_gh, s_red, tid);
    if (tid < MW) {
      const float4 h = load_rows(a.save_h + (size_t)(MD - 1) * a.rows * MW, MW, row0, rows, tid);
      g = relu_mask(g, h);
      s_g[cur][tid] = g;
      store_rows(a.G + (size_t)(MD - 1) * a.rows * MW, MW, row0, rows, tid, g);
    }
  }
  consumer_sync();
  for (int l = MD - 1; l >= 1; l--) {
    const MlpLayer& ly = L.layer[first_trunk + l];     // input of layer l is h_{l-1} (plus [x,t] for l == SKIP+1)
    float4 g;
    if (l == SKIP + 1) {                               // K = in0 + 256
      float4 gt;
      back_dense_cols(ring, ly.K, MW, s_g[cur], in0, Tt, tid, g, gt);
      if (tid < Tt) s_gt[tid] = gt;                    // time feature through the skip input
    } else {
      g = back_dense_k256(ring, MW, s_g[cur], s_red, tid);
    }
    if (tid < MW) {
      const float4 h = load_rows(a.save_h + (size_t)(l - 1) * a.rows * MW, MW, row0, rows, tid);
      g = relu_mask(g, h);
      s_g[cur ^ 1][tid] = g;
      store_rows(a.G + (size_t)(l - 1) * a.rows * MW, MW, row0, rows, tid, g);
    }
    consumer_sync();
    cur ^= 1;
  }
  // layer 0 input: only the time feature needs a gradient (node positions are detached)
  {
    const MlpLayer& l0 = L.layer[first_trunk];
    float4 unused, gt;
    back_dense_cols(ring, l0.K, MW, s_g[cur], -1, Tt, tid, unused, gt);
    if (tid < Tt) {
      const float4 o = s_gt[tid];
      s_gt[tid] = make_float4(o.x + gt.x, o.y + gt.y, o.z + gt.z, o.w + gt.w);
    }
  }
  consumer_sync();
  if (a.has_timenet) {
    if (tid < Tt) store_rows(a.g_tfeat, 32, row0, rows, tid, s_gt[tid]);
    const MlpLayer& t2 = L.layer[1];
    float4 g = back_dense_k256(ring, t2.N, s_gt, s_red, tid);
    if (tid < MW) {
      const float4 h = load_rows(a.save_th, MW, row0, rows, tid);
      g = relu_mask(g, h);
      store_rows(a.G_t1, MW, row0, rows, tid, g);
    }
  }
  cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------------
// backward, weight gradients: dW[n][k] = sum_r G[r][n] * in[r][k],  db[n] = sum_r G[r][n]
// one CTA per (job, 64x64 tile); 256 threads, 4x4 outputs each; rows streamed 16 at a time through shared memory
// ---------------------------------------------------------------------------------------------------------------
constexpr int WROWS = 64;   // rows reduced per shared-memory chunk (32 independent loads per thread in flight)
__global__ void __launch_bounds__(256) mlp_bwd_w_kernel(MlpWJobs J, int rows) {
  __shared__ float sG[WROWS][64 + 4];
  __shared__ float sA[WROWS][64 + 4];
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
  const int c = tid & 63, rq = tid >> 6;          // this thread stages column c of rows rq, rq+4, ...
  const int n = n0 + c, k = k0 + c;
  const bool n_ok = n < jb.N, k_ok = k < Ktot;
  const float* gcol = jb.G + n;
  const float* acol = (k < jb.Ka) ? jb.A + k : jb.B + (k - jb.Ka);
  const int lda = (k < jb.Ka) ? jb.ldA : jb.ldB;
  // software pipeline: the global loads of chunk c+1 are in flight while chunk c is multiplied out of shared memory
  float gv[WROWS / 4], av[WROWS / 4];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int u = 0; u < WROWS / 4; u++) {
      const int rr = r0 + rq + 4 * u;
      gv[u] = (n_ok && rr < rows) ? __ldg(gcol + (size_t)rr * jb.ldG) : 0.f;
      av[u] = (k_ok && rr < rows) ? __ldg(acol + (size_t)rr * lda) : 0.f;
    }
  };
  fetch(0);
  for (int r0 = 0; r0 < rows; r0 += WROWS) {
#pragma unroll
    for (int u = 0; u < WROWS / 4; u++) { sG[rq + 4 * u][c] = gv[u]; sA[rq + 4 * u][c] = av[u]; }
    __syncthreads();
    if (r0 + WROWS < rows) fetch(r0 + WROWS);
#pragma unroll 8
    for (int r = 0; r < WROWS; r++) {
      const float4 g4 = *reinterpret_cast<const float4*>(&sG[r][tn * 4]);
      const float4 x4 = *reinterpret_cast<const float4*>(&sA[r][tkx * 4]);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w}, x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) acc[i][q] = fmaf(g[i], x[q], acc[i][q]);
      }
      if (k0 == 0 && tkx == 0) { bacc[0] += g[0]; bacc[1] += g[1]; bacc[2] += g[2]; bacc[3] += g[3]; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int nn = n0 + tn * 4 + i;
    if (nn >= jb.N) continue;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int kk = k0 + tkx * 4 + q;
      if (kk < Ktot) jb.dW[(size_t)nn * Ktot + kk] = acc[i][q];
    }
    if (k0 == 0 && tkx == 0 && jb.db) jb.db[nn] = bacc[i];
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
#ifndef D2GS_MLP_CLUSTER
#define D2GS_MLP_CLUSTER 4
#endif
// grid of ceil(rows / ROWS) CTAs, rounded up to whole clusters (a CTA without rows still takes part in the ring protocol)
// cluster sizes (option "mlp_cluster_fwd" / "mlp_cluster_bwd"): the forward multicasts across 4 CTAs; the backward gains nothing
// from it (profiles/mlp_tensor_core_probe_r2.md) and runs next to the early all-reduce at N > 1, where whole clusters of free
// SMs are scarce — it launches plain CTAs
static int g_mlp_cluster[2] = {D2GS_MLP_CLUSTER, 1};
void mlp_set_cluster(int which, int size) {
  if (size == 1 || size == 2 || size == 4 || size == 8) g_mlp_cluster[which & 1] = size;
}
template <typename Kern, typename Arg>
static void launch_clustered(Kern kern, const Arg& a, int rows, size_t smem, int cluster, cudaStream_t s) {
  const int ctas = (rows + ROWS - 1) / ROWS;
  int cl = cluster;
  while (cl > 1 && ctas < cl) cl >>= 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((ctas + cl - 1) / cl * cl), 1, 1);
  cfg.blockDim = dim3(MLP_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, a);
}
void mlp_launch_forward(const MlpFwd& a, cudaStream_t s) {
  if (a.rows <= 0) return;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM_BYTES); attr = true; }
  launch_clustered(mlp_fwd_kernel, a, a.rows, FWD_SMEM_BYTES, g_mlp_cluster[0], s);
}
void mlp_launch_backward(const MlpBwd& a, const MlpWJobs& J, cudaStream_t s) {
  if (a.rows <= 0) return;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(mlp_bwd_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES_MLP); attr = true; }
  launch_clustered(mlp_bwd_act_kernel, a, a.rows, BWD_SMEM_BYTES_MLP, g_mlp_cluster[1], s);
  int tiles = 0;
  for (int j = 0; j < J.count; j++) tiles += J.job[j].tiles;
  mlp_bwd_w_kernel<<<tiles, 256, 0, s>>>(J, a.rows);
}

}  // namespace d2gs
