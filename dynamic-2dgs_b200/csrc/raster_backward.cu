// Backward kernels of the surfel rasterizer for sm_100a.
// Behavioural contract: DSR/cuda_rasterizer/backward.cu:143-449 (blend), :599-649 (AABB), :451-597 (per surfel),
// :20-139 (SH).  Design differences (results agree to fp32 re-association):
//   * the reference issues up to 16 global float atomics per contributing (pixel, surfel) pair; here the contributing
//     lanes of a warp (8.7 of 32 on average) write their 18 gradient components as compacted rows of a small
//     per-warp shared-memory scratch, 18 lanes sum one column each, the warp totals go into a per-instance
//     shared-memory accumulator, and the CTA issues one global reduction per (tile, instance, component) —
//     R*18 instead of pairs*16.  (The cost scales with the contributing lanes: ~35 instructions per surviving
//     (patch, instance) pair, where a 32-lane shuffle butterfly over 16+2 values cost ~110.);
//   * traversal starts at the deepest instance any pixel of the tile actually consumed (block max of n_contrib)
//     instead of the end of the tile list;
//   * AABB backward, the transMat/normal/SH backward and the clearing of the gradient scratch are one
//     per-surfel kernel; all outputs are fully written, so callers need no zero-fill.
// Two blend kernels live here: blend_bwd_kernel (round 1: a warp visits one surfel per iteration; the description above) and
// blend_bwd_lw_kernel (round 2, the default, option "lane_walk"): every lane walks its own list of prefilter hits — taken from
// the hit masks the forward stored — writes one gradient row per (surfel, pixel) pair, and lanes then sum one (surfel, float4
// of components) each and issue a 16-byte global reduction.  Same float operations per pair; see the comment at that kernel.
#include "raster_common.cuh"

namespace d2gs {

#ifndef D2GS_BWD_BATCH
#define D2GS_BWD_BATCH 64
#endif
constexpr int BWD_BATCH = D2GS_BWD_BATCH;   // instances staged per round
// Each 16x16 tile is worked on by TWO CTAs of 4 warps (rows 0-7 / 8-15, blockIdx.z): barriers then wait for the
// slowest of 4 patches instead of 8, and six small CTAs per SM interleave where three large ones stalled together.
#ifndef D2GS_BWD_WARPS
#define D2GS_BWD_WARPS 4
#endif
constexpr int NWARP = D2GS_BWD_WARPS;            // warps (8x4 pixel patches) per CTA: 4, 2 or 1
constexpr int BWD_THREADS = 32 * NWARP;
constexpr int BWD_Z = 8 / NWARP;                 // CTAs per 16x16 tile (blockIdx.z)
constexpr int BWD_SPT = (D2GS_BWD_BATCH + BWD_THREADS - 1) / BWD_THREADS;   // staging slots per thread
constexpr int ACC_STRIDE = 19;   // 18 components, odd stride keeps the flush free of bank conflicts
constexpr unsigned FULL = 0xffffffffu;
// per-warp reduction scratch: RED_ROWS rows of RED_STRIDE floats.  A row holds one contributing lane's 18 components in
// accumulator-slot order; 80-byte rows put the 16-byte stores of 8 consecutive rows on 8 distinct bank groups.
constexpr int RED_ROWS = 16;
constexpr int RED_STRIDE = 20;
constexpr int RED_COMPS = 18;
// dynamic shared memory of blend_bwd_kernel
constexpr int BWD_WORDS = (BWD_BATCH + 31) / 32;
constexpr size_t BWD_SMEM_Q1 = sizeof(float4) * REC_QUADS * BWD_BATCH;                 // one staging buffer
constexpr size_t BWD_SMEM_Q = 2 * BWD_SMEM_Q1;                                         // double buffered (cp.async)
constexpr size_t BWD_SMEM_ACC = sizeof(float) * NWARP * BWD_BATCH * ACC_STRIDE;        // per-warp private accumulators
constexpr size_t BWD_SMEM_RED = sizeof(float) * NWARP * RED_ROWS * RED_STRIDE;         // per-warp reduction scratch
constexpr size_t BWD_SMEM_BYTES = BWD_SMEM_Q + BWD_SMEM_ACC + BWD_SMEM_RED + sizeof(uint32_t) * (2 * BWD_BATCH + NWARP);

// MUFU.RCP without the range scaling of __fdividef (6 instructions less): for operands far from the denormal range
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void pixel_of_thread_b(int tid, int& lx, int& ly) {
  const int w = tid >> 5, l = tid & 31;
  lx = ((w & 1) << 3) | (l & 7);
  ly = ((w >> 1) << 2) | (l >> 3);
}

#ifndef D2GS_BWD_MINBLOCKS
#define D2GS_BWD_MINBLOCKS 6   // 79 registers, no spills
#endif
__global__ void __launch_bounds__(BWD_THREADS, D2GS_BWD_MINBLOCKS) blend_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float* __restrict__ bg, const SurfelRec* __restrict__ rec, const float* __restrict__ final_Ts,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dothers, float* __restrict__ grad_rec, int cull,
    const uint32_t* __restrict__ tile_order, uint32_t gx) {
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  float* s_acc = reinterpret_cast<float*>(bwd_smem + BWD_SMEM_Q);        // [warp][slot][ACC_STRIDE]
  float* s_red = reinterpret_cast<float*>(bwd_smem + BWD_SMEM_Q + BWD_SMEM_ACC);        // [warp][RED_ROWS][RED_STRIDE]
  uint32_t* s_id = reinterpret_cast<uint32_t*>(bwd_smem + BWD_SMEM_Q + BWD_SMEM_ACC + BWD_SMEM_RED);   // [2][BWD_BATCH]
  uint32_t* s_max = s_id + 2 * BWD_BATCH;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1-D grid, longest tile lists first (see blend_fwd_kernel)
  const uint32_t sub = blockIdx.x % BWD_Z;
  const uint32_t tile = tile_order ? __ldg(tile_order + blockIdx.x / BWD_Z) : blockIdx.x / BWD_Z;
  const uint32_t tile_x = tile % gx, tile_y = tile / gx;
  const int gwarp = warp + NWARP * (int)sub;      // patch index inside the tile (0..7)
  int lx, ly;
  pixel_of_thread_b(tid + BWD_THREADS * (int)sub, lx, ly);
  const uint32_t pix_x = tile_x * TILE_X + lx, pix_y = tile_y * TILE_Y + ly;
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const uint32_t pix_id = W * pix_y + pix_x;
  const size_t HW = (size_t)H * W;
  const float2 pixf = {(float)pix_x + 0.5f, (float)pix_y + 0.5f};
  const uint2 range = ranges[tile];
  const float pcx0 = (float)(tile_x * TILE_X + ((gwarp & 1) << 3)) + 0.5f, pcx1 = pcx0 + 7.0f;
  const float pcy0 = (float)(tile_y * TILE_Y + ((gwarp >> 1) << 2)) + 0.5f, pcy1 = pcy0 + 3.0f;
  const uint32_t sq_base = smem_addr(bwd_smem);
  constexpr uint32_t QS = 16u * BWD_BATCH;   // bytes per staged quad plane
  float* my_acc = s_acc + (size_t)warp * BWD_BATCH * ACC_STRIDE;
  const uint32_t lanes_below = (1u << lane) - 1u;
  const uint32_t red_w = smem_addr(s_red + (size_t)warp * RED_ROWS * RED_STRIDE);   // this warp's reduction scratch
  const uint32_t red_col = red_w + 4u * (uint32_t)lane;                               // column `lane` of its first row
  const uint32_t acc_col = smem_addr(my_acc) + 4u * (uint32_t)lane;                   // component `lane` of accumulator slot 0

  const float T_final = inside ? final_Ts[pix_id] : 0;
  float T = T_final;
  const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;
  const int median_contributor = inside ? (int)n_contrib[pix_id + HW] : 0;

  // deepest list position any pixel of this tile consumed
  {
    const uint32_t m = __reduce_max_sync(FULL, last_contributor);
    if (lane == 0) s_max[warp] = m;
  }
  for (int i = tid; i < NWARP * BWD_BATCH * ACC_STRIDE; i += BWD_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  uint32_t len = 0;
#pragma unroll
  for (int i = 0; i < NWARP; i++) len = max(len, s_max[i]);
  len = min(len, range.y - range.x);
  if (len == 0) return;

  float dL_dpixel[3] = {0.f, 0.f, 0.f}, dL_dnormal2D[3] = {0.f, 0.f, 0.f};
  float dL_ddepth = 0.f, dL_daccum = 0.f, dL_dreg = 0.f, dL_dmedian_depth = 0.f, dL_dmax_dweight = 0.f;
  if (inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix_id];
    dL_ddepth = dL_dothers[0 * HW + pix_id];
    dL_daccum = dL_dothers[1 * HW + pix_id];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dL_dnormal2D[ch] = dL_dothers[(2 + ch) * HW + pix_id];
    dL_dmedian_depth = dL_dothers[5 * HW + pix_id];
    dL_dreg = dL_dothers[6 * HW + pix_id];
    dL_dmax_dweight = dL_dothers[7 * HW + pix_id];
  }
  const float final_D = inside ? final_Ts[pix_id + HW] : 0;
  const float final_D2 = inside ? final_Ts[pix_id + 2 * HW] : 0;
  const float final_A = 1 - T_final;
  const float bg_dot_dpixel = bg[0] * dL_dpixel[0] + bg[1] * dL_dpixel[1] + bg[2] * dL_dpixel[2];

  float Q = T_final * bg_dot_dpixel;   // see "ONE running scalar" below

  const int rounds = (len + BWD_BATCH - 1) / BWD_BATCH;
  // instance ids of this thread's slots in batch bi (back to front: slot t holds list position len-1-(bi*B+t))
  struct Ids { uint32_t v[BWD_SPT]; };
  auto slot_id = [&](int bi) -> Ids {
    Ids r;
    const int nb = min(BWD_BATCH, (int)len - bi * BWD_BATCH);
#pragma unroll
    for (int k = 0; k < BWD_SPT; k++) {
      const int t = tid + k * BWD_THREADS;
      r.v[k] = (bi < rounds && t < nb) ? __ldg(&point_list[range.x + (len - 1 - (uint32_t)(bi * BWD_BATCH + t))]) : 0xffffffffu;
    }
    return r;
  };
  // stage batch bi into buffer buf with cp.async; the ids were fetched one iteration earlier so no load latency is exposed
  auto stage = [&](int bi, int buf, const Ids& ids) {
#pragma unroll
    for (int k = 0; k < BWD_SPT; k++) {
      const uint32_t id = ids.v[k];
      const int t = tid + k * BWD_THREADS;
      if (id != 0xffffffffu) {
        s_id[buf * BWD_BATCH + t] = id;
        const float4* r4 = reinterpret_cast<const float4*>(rec + id);
        const uint32_t dst = sq_base + (uint32_t)buf * (uint32_t)BWD_SMEM_Q1 + ((uint32_t)t << 4);
#pragma unroll
        for (int q = 0; q < REC_QUADS; q++) cp_async16(dst + q * QS, r4 + q);
      }
    }
    cp_async_commit();
  };
  stage(0, 0, slot_id(0));
  Ids pre_id = slot_id(1);
  int remaining = (int)len;
  for (int i = 0; i < rounds; i++, remaining -= BWD_BATCH) {
    const int n = min(BWD_BATCH, remaining);
    const int buf = i & 1;
    // buffer buf^1 was last read in batch i-1, which every warp left before the flush barrier of that batch
    if (i + 1 < rounds) {
      stage(i + 1, buf ^ 1, pre_id);
      pre_id = slot_id(i + 2);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sb0 = sq_base + (uint32_t)buf * (uint32_t)BWD_SMEM_Q1;
    const uint32_t sb1 = sb0 + QS, sb2 = sb0 + 2 * QS, sb3 = sb0 + 3 * QS, sb4 = sb0 + 4 * QS, sb5 = sb0 + 5 * QS;
    const uint32_t* cur_id = s_id + buf * BWD_BATCH;

    // Warp-level compaction (see blend_fwd_kernel): survivors of the cull-box test as a bit mask, visited in staged
    // (back-to-front) order.
    uint32_t keepmask[BWD_WORDS];
#pragma unroll
    for (int w = 0; w < BWD_WORDS; w++) {
      const int jj = w * 32 + lane;
      bool keep = jj < n;
      if (keep && cull) {
        const float4 bb = lds128(sb5 + ((uint32_t)jj << 4));
        keep = !(bb.z < pcx0 || bb.x > pcx1 || bb.w < pcy0 || bb.y > pcy1);
      }
      keepmask[w] = __ballot_sync(FULL, keep);
    }
#pragma unroll
    for (int w = 0; w < BWD_WORDS; w++) {
      uint32_t m = keepmask[w];
      while (m != 0u) {
      const int j = w * 32 + (__ffs(m) - 1);
      m &= m - 1u;
      const uint32_t contributor = len - 1 - (uint32_t)(i * BWD_BATCH + j);   // 0-based list position
      const uint32_t off = (uint32_t)j << 4;
      float g[16];          // written only by contributing lanes
      float gm0 = 0.f, gm1 = 0.f;
      bool contrib = false;
      do {
        if (!inside || contributor >= last_contributor) break;
        const float4 a = lds128(sb0 + off), b = lds128(sb1 + off), c = lds128(sb2 + off);
        const float3 Tu = {a.x, a.y, a.z}, Tv = {a.w, b.x, b.y}, Tw = {b.z, b.w, c.x};
        // same pinned roundings as blend_fwd_kernel: the set of contributing pairs must be the forward's, bit for bit
        const float3 k = {__fmaf_rn(pixf.x, Tw.x, -Tu.x), __fmaf_rn(pixf.x, Tw.y, -Tu.y), __fmaf_rn(pixf.x, Tw.z, -Tu.z)};
        const float3 l = {__fmaf_rn(pixf.y, Tw.x, -Tv.x), __fmaf_rn(pixf.y, Tw.y, -Tv.y), __fmaf_rn(pixf.y, Tw.z, -Tv.z)};
        const float3 p = {__fmaf_rn(k.y, l.z, -__fmul_rn(k.z, l.y)), __fmaf_rn(k.z, l.x, -__fmul_rn(k.x, l.z)),
                          __fmaf_rn(k.x, l.y, -__fmul_rn(k.y, l.x))};
        const float2 d = {__fsub_rn(c.y, pixf.x), __fsub_rn(c.z, pixf.y)};
        const float rho2d = 2.0f * __fmaf_rn(d.x, d.x, __fmul_rn(d.y, d.y));
        if (pair_rejected(p.x, p.y, p.z, rho2d, c.w)) break;   // alpha < 1/255 for certain
        if (p.z == 0.0f) break;
        const float2 s = {__fdiv_rn(p.x, p.z), __fdiv_rn(p.y, p.z)};
        const float rho3d = __fmaf_rn(s.x, s.x, __fmul_rn(s.y, s.y));
        const float rho = fminf(rho3d, rho2d);
        const float c_d = (rho3d <= rho2d) ? __fadd_rn(Tw.z, __fmaf_rn(Tw.x, s.x, __fmul_rn(Tw.y, s.y))) : Tw.z;
        if (c_d < 0.2f) break;
        const float power = -0.5f * rho;
        if (power > 0.0f) break;
        const float G = expf(power);
        const float4 col = lds128(sb4 + off);   // rgb + opacity
        const float opac = col.w;
        const float alpha = fminf(0.99f, __fmul_rn(opac, G));
        if (alpha < 1.0f / 255.0f) break;
        const float4 nrm = lds128(sb3 + off);
        const float normal[3] = {nrm.x, nrm.y, nrm.z};
        const float color[3] = {col.x, col.y, col.z};

        // Back-to-front compositing gradient with ONE running scalar.  With T the transmittance in front of this splat,
        // w = alpha*T its blend weight, E the sum over all output channels of (this splat's channel value) x (upstream
        // gradient of that channel) plus the distortion weight term, and Q = T_final*(bg . dL_dpixel) + sum of w*E over
        // the splats behind:   dL/dalpha = T*E - Q/(1-alpha).
        // (The reference carries per-channel "colour seen behind" recurrences — accum_rec, last_color, last_alpha, ...
        // for 3 colour + depth + alpha + 3 normal channels and last_dL_dT, backward.cu:292-372; dotted with the upstream
        // gradients they all collapse into Q.  Same value up to fp32 re-association.)
        const float inv_1ma = __fdividef(1.0f, 1.f - alpha);   // gradients are compared to 1e-4: 2-ulp reciprocal
        T = T * inv_1ma;
        const float w = alpha * T;
        float E = dL_daccum;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
          E = fmaf(color[ch], dL_dpixel[ch], E);
          E = fmaf(normal[ch], dL_dnormal2D[ch], E);
          g[13 + ch] = w * dL_dpixel[ch];
          g[9 + ch] = w * dL_dnormal2D[ch];
        }
        E = fmaf(c_d, dL_ddepth, E);
        float dL_dz = 0.0f, dL_dweight = 0.f;
        // depth mapped to [0,1]: the reference's float, bit for bit (mapped_depth) — the weight gradient below is a
        // difference of nearly equal terms and amplifies a last-bit change of m_d by ~1e3
        const float inv_d = __fdividef(1.0f, c_d);
        const float m_d = mapped_depth(c_d);
        const float dmd_dd = 0.2004008016032064f * inv_d * inv_d;            // near*far/(far-near) / d^2
        if (contributor == (uint32_t)(median_contributor - 1)) {
          dL_dz += dL_dmedian_depth;
          dL_dweight += dL_dmax_dweight;
        }
        // (final_D2 + m^2 A - 2 m D) * dL_dreg with the reference build's roundings (backward.cu:362)
        dL_dweight = __fmaf_rn(dL_dreg, __fmaf_rn(__fadd_rn(m_d, m_d), -final_D, __fmaf_rn(final_A, __fmul_rn(m_d, m_d), final_D2)), dL_dweight);
        E += dL_dweight;
        const float dL_dalpha = T * E - Q * inv_1ma;
        Q = fmaf(w, E, Q);
        const float dL_dmd = 2.0f * w * __fmaf_rn(final_A, m_d, -final_D) * dL_dreg;
        dL_dz += dL_dmd * dmd_dd;

        const float dL_dG = opac * dL_dalpha;
        dL_dz += w * dL_ddepth;

        if (rho3d <= rho2d) {
          const float2 dL_ds = {dL_dG * -G * s.x + dL_dz * Tw.x, dL_dG * -G * s.y + dL_dz * Tw.y};
          const float inv_pz = __fdividef(1.0f, p.z);
          const float dsx_pz = dL_ds.x * inv_pz, dsy_pz = dL_ds.y * inv_pz;
          const float3 dL_dp = {dsx_pz, dsy_pz, -(dsx_pz * s.x + dsy_pz * s.y)};
          const float3 dL_dk = {l.y * dL_dp.z - l.z * dL_dp.y, l.z * dL_dp.x - l.x * dL_dp.z, l.x * dL_dp.y - l.y * dL_dp.x};
          const float3 dL_dl = {dL_dp.y * k.z - dL_dp.z * k.y, dL_dp.z * k.x - dL_dp.x * k.z, dL_dp.x * k.y - dL_dp.y * k.x};
          g[0] = -dL_dk.x; g[1] = -dL_dk.y; g[2] = -dL_dk.z;
          g[3] = -dL_dl.x; g[4] = -dL_dl.y; g[5] = -dL_dl.z;
          g[6] = pixf.x * dL_dk.x + pixf.y * dL_dl.x + dL_dz * s.x;
          g[7] = pixf.x * dL_dk.y + pixf.y * dL_dl.y + dL_dz * s.y;
          g[8] = pixf.x * dL_dk.z + pixf.y * dL_dl.z + dL_dz;
        } else {
          // low-pass-filter branch: gradient goes to the 2-D centre and to Tw.z
          const float dG_ddelx = -G * 2.0f * d.x;
          const float dG_ddely = -G * 2.0f * d.y;
          gm0 = dL_dG * dG_ddelx;
          gm1 = dL_dG * dG_ddely;
#pragma unroll
          for (int q = 0; q < 8; q++) g[q] = 0.f;
          g[8] = dL_dz;
        }
        g[12] = G * dL_dalpha;
        contrib = true;
      } while (0);

      // Warp reduction over the contributing lanes only: the lane holding the r-th set bit of the ballot writes row r of
      // the scratch (accumulator-slot order: 0..8 transMat | 9,10 mean2D | 11..13 normal | 14 opacity | 15..17 colour),
      // every lane then sums one column over the rows (columns >= 18 are never used) and lanes 0..17 add theirs to this
      // warp's private accumulator of instance j.  Shared-memory addresses are precomputed 32-bit registers.
      const uint32_t cmask = __ballot_sync(FULL, contrib);
      if (cmask != 0u) {
        const int nrow = __popc(cmask);
        const int row = __popc(cmask & lanes_below);
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll 1
        for (int base = 0; base < nrow; base += RED_ROWS) {      // one round unless more than 16 lanes contribute
          if (contrib && (unsigned)(row - base) < (unsigned)RED_ROWS) {
            const uint32_t dst = red_w + (uint32_t)(row - base) * (RED_STRIDE * 4);
            sts128(dst, g[0], g[1], g[2], g[3]);
            sts128(dst + 16, g[4], g[5], g[6], g[7]);
            sts128(dst + 32, g[8], gm0, gm1, g[9]);
            sts128(dst + 48, g[10], g[11], g[12], g[13]);
            sts64(dst + 64, g[14], g[15]);
          }
          __syncwarp();
          const int nr = min(RED_ROWS, nrow - base);
          uint32_t ad = red_col;
          int r = 0;
          for (; r + 4 <= nr; r += 4, ad += 4 * RED_STRIDE * 4) {
            t0 += lds32(ad); t1 += lds32(ad + RED_STRIDE * 4); t2 += lds32(ad + 2 * RED_STRIDE * 4); t3 += lds32(ad + 3 * RED_STRIDE * 4);
          }
          if (r + 2 <= nr) { t0 += lds32(ad); t1 += lds32(ad + RED_STRIDE * 4); ad += 2 * RED_STRIDE * 4; r += 2; }
          if (r < nr) t2 += lds32(ad);
          __syncwarp();
        }
        if (lane < RED_COMPS) {
          const uint32_t a = acc_col + (uint32_t)j * (ACC_STRIDE * 4);   // slot private to this warp: plain read-modify-write
          sts32(a, lds32(a) + ((t0 + t1) + (t2 + t3)));
        }
      }
      }
    }
    __syncthreads();
    // fold the 8 warp-private accumulators and issue one global reduction per (tile, instance, component);
    // two threads per instance, nine components each
    {
      const int half = tid & 1;
      for (int slot = tid >> 1; slot < n; slot += BWD_THREADS / 2) {
        float* dst = grad_rec + (size_t)cur_id[slot] * GRAD_REC_FLOATS;
#pragma unroll
        for (int q = 0; q < 9; q++) {
          const int v = half * 9 + q;
          float val = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < NWARP; w8++) {
            float* a = s_acc + ((size_t)w8 * BWD_BATCH + slot) * ACC_STRIDE + v;
            const float x = *a;
            if (x != 0.f) { val += x; *a = 0.f; }
          }
          if (val != 0.f) atomicAdd(dst + v, val);
        }
      }
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------------------
// Lane-walk variant of the backward blend (default; see blend_fwd_lw_kernel).  Per warp and chunk of survivors:
//   phase 1  all lanes run the exact prefilter for every cull-box survivor (broadcast LDS.128); the ballot of the
//            passing lanes is the survivor's hit mask: it reserves popc(mask) rows of the warp's gradient scratch and
//            each lane keeps ITS hits as bits (ordinal of the survivor in the chunk) of a 64-bit mask;
//   phase 2  every lane walks its own hits back to front (lanes work on different surfels at the same time: ~18-20
//            active lanes instead of ~9), runs the compositing gradient with the running scalars T, Q of its pixel and
//            writes its 18 components as row  rowbase(survivor) + rank(lane in the hit mask);
//   phase 3  per survivor, 18 lanes sum one column each over its rows and issue ONE global reduction per
//            (patch, surfel, component).  No per-warp accumulators, no CTA-wide flush, no second barrier per batch.
// Same float operations per (pixel, surfel) pair as blend_bwd_kernel; only the order of the gradient sums differs.
// ------------------------------------------------------------------------------------------------------------
#ifndef D2GS_BWD_LW_ROWS
#define D2GS_BWD_LW_ROWS 64
#endif
constexpr int LW_ROWS = D2GS_BWD_LW_ROWS;      // gradient rows per warp (80 B each)
constexpr int LW_ORD = 32;                      // survivors per chunk (one 32-bit hit mask per lane)
constexpr size_t LWB_SMEM_Q = 2 * BWD_SMEM_Q1;
constexpr size_t LWB_SMEM_ROWS = sizeof(float) * NWARP * LW_ROWS * RED_STRIDE;
constexpr size_t LWB_SMEM_META = sizeof(uint2) * NWARP * LW_ORD;       // {hit mask, first row | slot << 16}
constexpr size_t LWB_SMEM_BYTES = LWB_SMEM_Q + LWB_SMEM_ROWS + LWB_SMEM_META + sizeof(uint32_t) * (2 * BWD_BATCH + NWARP);

#ifndef D2GS_BWD_LW_MINBLOCKS
#define D2GS_BWD_LW_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(BWD_THREADS, D2GS_BWD_LW_MINBLOCKS) blend_bwd_lw_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float* __restrict__ bg, const SurfelRec* __restrict__ rec, const float* __restrict__ final_Ts,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dothers, float* __restrict__ grad_rec, int cull,
    const uint32_t* __restrict__ tile_order, uint32_t gx, const uint32_t* __restrict__ hit_mask,
    const uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  float* s_rows = reinterpret_cast<float*>(bwd_smem + LWB_SMEM_Q);                          // [warp][LW_ROWS][RED_STRIDE]
  uint2* s_meta = reinterpret_cast<uint2*>(bwd_smem + LWB_SMEM_Q + LWB_SMEM_ROWS);          // [warp][LW_ORD]
  uint32_t* s_id = reinterpret_cast<uint32_t*>(bwd_smem + LWB_SMEM_Q + LWB_SMEM_ROWS + LWB_SMEM_META);   // [2][BWD_BATCH]
  uint32_t* s_max = s_id + 2 * BWD_BATCH;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sub = blockIdx.x % BWD_Z;
  const uint32_t tile = tile_order ? __ldg(tile_order + blockIdx.x / BWD_Z) : blockIdx.x / BWD_Z;
  const uint32_t tile_x = tile % gx, tile_y = tile / gx;
  const int gwarp = warp + NWARP * (int)sub;
  int lx, ly;
  pixel_of_thread_b(tid + BWD_THREADS * (int)sub, lx, ly);
  const uint32_t pix_x = tile_x * TILE_X + lx, pix_y = tile_y * TILE_Y + ly;
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const uint32_t pix_id = W * pix_y + pix_x;
  const size_t HW = (size_t)H * W;
  const float2 pixf = {(float)pix_x + 0.5f, (float)pix_y + 0.5f};
  const uint2 range = ranges[tile];
  const float pcx0 = (float)(tile_x * TILE_X + ((gwarp & 1) << 3)) + 0.5f, pcx1 = pcx0 + 7.0f;
  const float pcy0 = (float)(tile_y * TILE_Y + ((gwarp >> 1) << 2)) + 0.5f, pcy1 = pcy0 + 3.0f;
  const uint32_t sq_base = smem_addr(bwd_smem);
  constexpr uint32_t QS = 16u * BWD_BATCH;
  const uint32_t lanes_below = (1u << lane) - 1u;
  const uint32_t rows_w = smem_addr(s_rows + (size_t)warp * LW_ROWS * RED_STRIDE);
  const uint32_t meta_w = smem_addr(s_meta + (size_t)warp * LW_ORD);
  // hit masks of the forward pass (ballots of the exact prefilter), if the forward of this frame wrote them
  const uint32_t* const hm_row = (hit_mask != nullptr && status != nullptr && __ldg(status + 3) == HIT_MASK_MAGIC)
                                     ? hit_mask + (size_t)range.x * 8u + (size_t)gwarp * (range.y - range.x) : nullptr;

  const float T_final = inside ? final_Ts[pix_id] : 0;
  float T = T_final;
  const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0;
  const int median_contributor = inside ? (int)n_contrib[pix_id + HW] : 0;
  {
    const uint32_t m = __reduce_max_sync(FULL, last_contributor);
    if (lane == 0) s_max[warp] = m;
  }
  __syncthreads();
  uint32_t len = 0;
#pragma unroll
  for (int i = 0; i < NWARP; i++) len = max(len, s_max[i]);
  len = min(len, range.y - range.x);
  if (len == 0) return;

  float dL_dpixel[3] = {0.f, 0.f, 0.f}, dL_dnormal2D[3] = {0.f, 0.f, 0.f};
  float dL_ddepth = 0.f, dL_daccum = 0.f, dL_dreg = 0.f, dL_dmedian_depth = 0.f, dL_dmax_dweight = 0.f;
  if (inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix_id];
    dL_ddepth = dL_dothers[0 * HW + pix_id];
    dL_daccum = dL_dothers[1 * HW + pix_id];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dL_dnormal2D[ch] = dL_dothers[(2 + ch) * HW + pix_id];
    dL_dmedian_depth = dL_dothers[5 * HW + pix_id];
    dL_dreg = dL_dothers[6 * HW + pix_id];
    dL_dmax_dweight = dL_dothers[7 * HW + pix_id];
  }
  const float final_D = inside ? final_Ts[pix_id + HW] : 0;
  const float final_D2 = inside ? final_Ts[pix_id + 2 * HW] : 0;
  const float final_A = 1 - T_final;
  const float bg_dot_dpixel = bg[0] * dL_dpixel[0] + bg[1] * dL_dpixel[1] + bg[2] * dL_dpixel[2];
  float Q = T_final * bg_dot_dpixel;

  const int rounds = (len + BWD_BATCH - 1) / BWD_BATCH;
  struct Ids { uint32_t v[BWD_SPT]; };
  auto slot_id = [&](int bi) -> Ids {
    Ids r;
    const int nb = min(BWD_BATCH, (int)len - bi * BWD_BATCH);
#pragma unroll
    for (int k = 0; k < BWD_SPT; k++) {
      const int t = tid + k * BWD_THREADS;
      r.v[k] = (bi < rounds && t < nb) ? __ldg(&point_list[range.x + (len - 1 - (uint32_t)(bi * BWD_BATCH + t))]) : 0xffffffffu;
    }
    return r;
  };
  auto stage = [&](int bi, int buf, const Ids& ids) {
#pragma unroll
    for (int k = 0; k < BWD_SPT; k++) {
      const uint32_t id = ids.v[k];
      const int t = tid + k * BWD_THREADS;
      if (id != 0xffffffffu) {
        s_id[buf * BWD_BATCH + t] = id;
        const float4* r4 = reinterpret_cast<const float4*>(rec + id);
        const uint32_t dst = sq_base + (uint32_t)buf * (uint32_t)BWD_SMEM_Q1 + ((uint32_t)t << 4);
#pragma unroll
        for (int q = 0; q < REC_QUADS; q++)
          if (q < 5 || hm_row == nullptr) cp_async16(dst + q * QS, r4 + q);   // the cull boxes (q5) are not needed with forward hit masks
      }
    }
    cp_async_commit();
  };
  stage(0, 0, slot_id(0));
  Ids pre_id = slot_id(1);
  int remaining = (int)len;
  for (int i = 0; i < rounds; i++, remaining -= BWD_BATCH) {
    const int n = min(BWD_BATCH, remaining);
    const int buf = i & 1;
    // forward hit masks of this batch's slots (lane L: slots L, L + 32, ...), fetched ahead of the staging barrier
    uint32_t fmask[BWD_WORDS];
    if (hm_row) {
#pragma unroll
      for (int q = 0; q < BWD_WORDS; q++) {
        const int jj = q * 32 + lane;
        fmask[q] = jj < n ? __ldg(hm_row + (len - 1 - (uint32_t)(i * BWD_BATCH + jj))) : 0u;
      }
    }
    // buffer buf^1 was last read in batch i-1: every warp has to be past it before it is refilled
    __syncthreads();
    if (i + 1 < rounds) {
      stage(i + 1, buf ^ 1, pre_id);
      pre_id = slot_id(i + 2);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sb0 = sq_base + (uint32_t)buf * (uint32_t)BWD_SMEM_Q1;
    const uint32_t sb1 = sb0 + QS, sb2 = sb0 + 2 * QS, sb3 = sb0 + 3 * QS, sb4 = sb0 + 4 * QS, sb5 = sb0 + 5 * QS;
    const uint32_t* cur_id = s_id + buf * BWD_BATCH;
    const uint32_t pos0 = len - 1 - (uint32_t)(i * BWD_BATCH);      // list position of staged slot 0 (slot j: pos0 - j)

    uint32_t keepmask[BWD_WORDS];
#pragma unroll
    for (int w = 0; w < BWD_WORDS; w++) {
      const int jj = w * 32 + lane;
      bool keep = jj < n;
      if (hm_row) {
        keep = keep && fmask[w] != 0u;       // the forward wrote 0 for the slots its cull box dropped (same box, same patch)
      } else if (keep && cull) {
        const float4 bb = lds128(sb5 + ((uint32_t)jj << 4));
        keep = !(bb.z < pcx0 || bb.x > pcx1 || bb.w < pcy0 || bb.y > pcy1);
      }
      keepmask[w] = __ballot_sync(FULL, keep);
    }
    int w = 0;
    uint32_t m = keepmask[0];
    while (true) {
      // ---- phase 1: exact prefilter on all lanes, rows reserved per survivor, per-lane hit masks
      uint32_t mine = 0u;
      int ord = 0, rows_used = 0;
      while (ord < LW_ORD) {
        if (m == 0u) {
          if (++w >= BWD_WORDS) break;
#pragma unroll
          for (int q = 1; q < BWD_WORDS; q++) if (w == q) m = keepmask[q];
          continue;
        }
        const int j = w * 32 + (__ffs(m) - 1);
        bool pass = inside && (pos0 - (uint32_t)j) < last_contributor;
        if (hm_row) {
          // the forward's ballot of the exact prefilter replaces ~45 instructions of intersection arithmetic
          uint32_t fm = fmask[0];
#pragma unroll
          for (int q = 1; q < BWD_WORDS; q++) if (w == q) fm = fmask[q];
          fm = __shfl_sync(FULL, fm, j & 31);
          pass = pass && ((fm >> lane) & 1u);
        } else {
        const uint32_t off = (uint32_t)j << 4;
        const float4 a = lds128(sb0 + off), b = lds128(sb1 + off), c = lds128(sb2 + off);
        const float3 k = {__fmaf_rn(pixf.x, b.z, -a.x), __fmaf_rn(pixf.x, b.w, -a.y), __fmaf_rn(pixf.x, c.x, -a.z)};
        const float3 l = {__fmaf_rn(pixf.y, b.z, -a.w), __fmaf_rn(pixf.y, b.w, -b.x), __fmaf_rn(pixf.y, c.x, -b.y)};
        const float3 p = {__fmaf_rn(k.y, l.z, -__fmul_rn(k.z, l.y)), __fmaf_rn(k.z, l.x, -__fmul_rn(k.x, l.z)),
                          __fmaf_rn(k.x, l.y, -__fmul_rn(k.y, l.x))};
        const float2 d = {__fsub_rn(c.y, pixf.x), __fsub_rn(c.z, pixf.y)};
        const float rho2d = 2.0f * __fmaf_rn(d.x, d.x, __fmul_rn(d.y, d.y));
        pass = pass && !pair_rejected(p.x, p.y, p.z, rho2d, c.w) && p.z != 0.0f;
        }
        const uint32_t hm = __ballot_sync(FULL, pass);
        const int nh = __popc(hm);
        if (rows_used + nh > LW_ROWS) break;     // does not fit any more: it opens the next chunk (its bit stays in m)
        m &= m - 1u;
        if (hm == 0u) continue;
        mine |= (pass ? 1u : 0u) << ord;
        if (lane == 0)
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(meta_w + 8u * (uint32_t)ord), "r"(hm),
                       "r"((uint32_t)rows_used | ((uint32_t)j << 16)) : "memory");
        rows_used += nh;
        ord++;
      }
      if (ord > 0) {
        __syncwarp();
        // ---- phase 2: every lane walks its own hits (increasing ordinal = back to front)
        {
          const uint32_t mbase = meta_w;
          while (mine != 0u) {
            const uint32_t o = (uint32_t)(__ffs(mine) - 1);
            mine &= mine - 1u;
            uint32_t hm, rj;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(hm), "=r"(rj) : "r"(mbase + 8u * o));
            const uint32_t j = rj >> 16;
            const uint32_t row = (rj & 0xffffu) + (uint32_t)__popc(hm & lanes_below);
            const uint32_t contributor = pos0 - j;
            const uint32_t off = j << 4;
            float g[16];
            float gm0 = 0.f, gm1 = 0.f;
#pragma unroll
            for (int q = 0; q < 16; q++) g[q] = 0.f;
            do {
              const float4 a = lds128(sb0 + off), b = lds128(sb1 + off), c = lds128(sb2 + off);
              const float3 Tw = {b.z, b.w, c.x};
              const float3 k = {__fmaf_rn(pixf.x, b.z, -a.x), __fmaf_rn(pixf.x, b.w, -a.y), __fmaf_rn(pixf.x, c.x, -a.z)};
              const float3 l = {__fmaf_rn(pixf.y, b.z, -a.w), __fmaf_rn(pixf.y, b.w, -b.x), __fmaf_rn(pixf.y, c.x, -b.y)};
              const float3 p = {__fmaf_rn(k.y, l.z, -__fmul_rn(k.z, l.y)), __fmaf_rn(k.z, l.x, -__fmul_rn(k.x, l.z)),
                                __fmaf_rn(k.x, l.y, -__fmul_rn(k.y, l.x))};
              const float2 d = {__fsub_rn(c.y, pixf.x), __fsub_rn(c.z, pixf.y)};
              const float rho2d = 2.0f * __fmaf_rn(d.x, d.x, __fmul_rn(d.y, d.y));
              const float2 s = {__fdiv_rn(p.x, p.z), __fdiv_rn(p.y, p.z)};
              const float rho3d = __fmaf_rn(s.x, s.x, __fmul_rn(s.y, s.y));
              const float rho = fminf(rho3d, rho2d);
              const float c_d = (rho3d <= rho2d) ? __fadd_rn(Tw.z, __fmaf_rn(Tw.x, s.x, __fmul_rn(Tw.y, s.y))) : Tw.z;
              if (c_d < 0.2f) break;
              const float power = -0.5f * rho;
              if (power > 0.0f) break;
              const float G = expf(power);
              const float4 col = lds128(sb4 + off);
              const float opac = col.w;
              const float alpha = fminf(0.99f, __fmul_rn(opac, G));
              if (alpha < 1.0f / 255.0f) break;
              const float4 nrm = lds128(sb3 + off);
              const float normal[3] = {nrm.x, nrm.y, nrm.z};
              const float color[3] = {col.x, col.y, col.z};
              // compositing gradient with the running scalars T, Q (see blend_bwd_kernel)
              const float inv_1ma = rcp_fast(1.f - alpha);          // 1 - alpha in [0.01, 1]
              T = T * inv_1ma;
              const float wgt = alpha * T;
              float E = dL_daccum;
#pragma unroll
              for (int ch = 0; ch < 3; ch++) {
                E = fmaf(color[ch], dL_dpixel[ch], E);
                E = fmaf(normal[ch], dL_dnormal2D[ch], E);
                g[13 + ch] = wgt * dL_dpixel[ch];
                g[9 + ch] = wgt * dL_dnormal2D[ch];
              }
              E = fmaf(c_d, dL_ddepth, E);
              float dL_dz = 0.0f, dL_dweight = 0.f;
              const float inv_d = rcp_fast(c_d);                     // c_d >= 0.2
              const float m_d = mapped_depth(c_d);
              const float dmd_dd = 0.2004008016032064f * inv_d * inv_d;
              if (contributor == (uint32_t)(median_contributor - 1)) {
                dL_dz += dL_dmedian_depth;
                dL_dweight += dL_dmax_dweight;
              }
              dL_dweight = __fmaf_rn(dL_dreg, __fmaf_rn(__fadd_rn(m_d, m_d), -final_D, __fmaf_rn(final_A, __fmul_rn(m_d, m_d), final_D2)), dL_dweight);
              E += dL_dweight;
              const float dL_dalpha = T * E - Q * inv_1ma;
              Q = fmaf(wgt, E, Q);
              const float dL_dmd = 2.0f * wgt * __fmaf_rn(final_A, m_d, -final_D) * dL_dreg;
              dL_dz += dL_dmd * dmd_dd;
              const float dL_dG = opac * dL_dalpha;
              dL_dz += wgt * dL_ddepth;
              if (rho3d <= rho2d) {
                const float2 dL_ds = {dL_dG * -G * s.x + dL_dz * Tw.x, dL_dG * -G * s.y + dL_dz * Tw.y};
                const float inv_pz = rcp_fast(p.z);
                const float dsx_pz = dL_ds.x * inv_pz, dsy_pz = dL_ds.y * inv_pz;
                const float3 dL_dp = {dsx_pz, dsy_pz, -(dsx_pz * s.x + dsy_pz * s.y)};
                const float3 dL_dk = {l.y * dL_dp.z - l.z * dL_dp.y, l.z * dL_dp.x - l.x * dL_dp.z, l.x * dL_dp.y - l.y * dL_dp.x};
                const float3 dL_dl = {dL_dp.y * k.z - dL_dp.z * k.y, dL_dp.z * k.x - dL_dp.x * k.z, dL_dp.x * k.y - dL_dp.y * k.x};
                g[0] = -dL_dk.x; g[1] = -dL_dk.y; g[2] = -dL_dk.z;
                g[3] = -dL_dl.x; g[4] = -dL_dl.y; g[5] = -dL_dl.z;
                g[6] = pixf.x * dL_dk.x + pixf.y * dL_dl.x + dL_dz * s.x;
                g[7] = pixf.x * dL_dk.y + pixf.y * dL_dl.y + dL_dz * s.y;
                g[8] = pixf.x * dL_dk.z + pixf.y * dL_dl.z + dL_dz;
              } else {
                gm0 = dL_dG * (-G * 2.0f * d.x);
                gm1 = dL_dG * (-G * 2.0f * d.y);
                g[8] = dL_dz;
              }
              g[12] = G * dL_dalpha;
            } while (0);
            // row in accumulator-slot order: 0..8 transMat | 9,10 mean2D | 11..13 normal | 14 opacity | 15..17 colour
            // (a pair the later tests dropped writes a row of zeros: its row was reserved by the prefilter ballot)
            const uint32_t dst = rows_w + row * (RED_STRIDE * 4);
            sts128(dst, g[0], g[1], g[2], g[3]);
            sts128(dst + 16, g[4], g[5], g[6], g[7]);
            sts128(dst + 32, g[8], gm0, gm1, g[9]);
            sts128(dst + 48, g[10], g[11], g[12], g[13]);
            sts64(dst + 64, g[14], g[15]);
          }
        }
        __syncwarp();
        // ---- phase 3: every lane sums ONE (survivor, group of 4 components) over the survivor's rows (LDS.128) and
        // issues one 16-byte global reduction: 5 lanes per survivor, ~6 survivors per pass of the warp
        for (int t = lane; t < 5 * ord; t += 32) {
          const int o = t / 5, gq = t - 5 * o;
          uint32_t hm, rj;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(hm), "=r"(rj) : "r"(meta_w + 8u * (uint32_t)o));
          int nr = __popc(hm);
          uint32_t ad = rows_w + (rj & 0xffffu) * (RED_STRIDE * 4) + 16u * (uint32_t)gq;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (; nr >= 2; nr -= 2, ad += 2 * RED_STRIDE * 4) {
            const float4 u = lds128(ad), v = lds128(ad + RED_STRIDE * 4);
            acc.x += u.x + v.x; acc.y += u.y + v.y; acc.z += u.z + v.z; acc.w += u.w + v.w;
          }
          if (nr) {
            const float4 u = lds128(ad);
            acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
          }
          if (gq == 4) { acc.z = 0.f; acc.w = 0.f; }      // floats 18, 19 of a row are padding (never written)
          if (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f || acc.w != 0.f)
            atomicAdd(reinterpret_cast<float4*>(grad_rec + (size_t)cur_id[rj >> 16] * GRAD_REC_FLOATS) + gq, acc);
        }
        __syncwarp();     // rows and meta are rewritten by the next chunk
      }
      if (w >= BWD_WORDS) break;
    }
  }
}

void launch_blend_bwd(const BwdParams& p, const uint2* ranges, const uint32_t* point_list, const SurfelRec* rec,
                      const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dothers,
                      float* grad_rec, int cull, const uint32_t* tile_order, int lane_walk, const uint32_t* hit_mask,
                      const uint32_t* status, cudaStream_t s) {
  const uint32_t grid = p.gx * p.gy * BWD_Z;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(blend_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES);
    // five CTAs of 37.4 KB (+1 KB each reserved by the system) need a large shared-memory carve-out
    cudaFuncSetAttribute(blend_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(blend_bwd_lw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LWB_SMEM_BYTES);
    cudaFuncSetAttribute(blend_bwd_lw_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  if (lane_walk) {
    blend_bwd_lw_kernel<<<grid, BWD_THREADS, LWB_SMEM_BYTES, s>>>(ranges, point_list, p.W, p.H, p.bg, rec, final_T, n_contrib,
                                                                 dL_dpix, dL_dothers, grad_rec, cull, tile_order, p.gx, hit_mask, status);
    return;
  }
  blend_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM_BYTES, s>>>(ranges, point_list, p.W, p.H, p.bg, rec, final_T, n_contrib, dL_dpix,
                                            dL_dothers, grad_rec, cull, tile_order, p.gx);
}

// ------------------------------------------------------------------------------------------------------------
// per-surfel backward: AABB-centre term, homography -> (mean, scale, quaternion), normal, SH.
// ------------------------------------------------------------------------------------------------------------
// SH coefficients 4c..4c+3 of surfel idx (12 floats) in any of the three layouts: packed (P,16,3) with 16-B aligned
// rows (3 x LDG.128), generic packed (P,M,3), or split DC (P,1,3) + rest (P,M-1,3).  Coefficients >= ncoef read as 0.
__device__ __forceinline__ void load_sh_chunk(const BwdParams& p, int idx, int c, int ncoef, float* sv /*[12]*/) {
  if (p.sh_rest == nullptr && p.M == 16 && ((reinterpret_cast<uintptr_t>(p.shs) & 15) == 0)) {
    const float4* b4 = reinterpret_cast<const float4*>(p.shs + (size_t)idx * 48) + 3 * c;
    if (4 * c < ncoef) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const float4 v = __ldg(b4 + i);
        sv[4 * i] = v.x; sv[4 * i + 1] = v.y; sv[4 * i + 2] = v.z; sv[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; i++) sv[i] = 0.f;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int f = 12 * c + i;   // flat float index inside the surfel's SH block
    float v = 0.f;
    if (f < ncoef * 3) {
      if (p.sh_rest == nullptr) v = __ldg(p.shs + (size_t)idx * p.M * 3 + f);
      else v = (f < 3) ? __ldg(p.shs + (size_t)idx * 3 + f) : __ldg(p.sh_rest + (size_t)idx * (p.M - 1) * 3 + (f - 3));
    }
    sv[i] = v;
  }
}
__device__ __forceinline__ void store_sh_chunk(const BwdParams& p, float* dL_dsh, float* dL_dsh_rest, int idx, int c,
                                               const float* g /*[12]*/) {
  if (dL_dsh == nullptr) return;
  if (dL_dsh_rest == nullptr && p.M == 16 && ((reinterpret_cast<uintptr_t>(dL_dsh) & 15) == 0)) {
    float4* b4 = reinterpret_cast<float4*>(dL_dsh + (size_t)idx * 48) + 3 * c;
#pragma unroll
    for (int i = 0; i < 3; i++) b4[i] = make_float4(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int f = 12 * c + i;
    if (f >= p.M * 3) continue;
    if (dL_dsh_rest == nullptr) dL_dsh[(size_t)idx * p.M * 3 + f] = g[i];
    else if (f < 3) dL_dsh[(size_t)idx * 3 + f] = g[i];
    else dL_dsh_rest[(size_t)idx * (p.M - 1) * 3 + (f - 3)] = g[i];
  }
}

constexpr int SH_WARP_FLOATS = 32 * 45 + 32 * 3;   // per-warp staging of the split SH layout: rest block + DC block

#ifndef D2GS_PREB_MINBLOCKS
#define D2GS_PREB_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(256, D2GS_PREB_MINBLOCKS) preprocess_bwd_kernel(
    BwdParams p, const SurfelRec* __restrict__ rec, const uint8_t* __restrict__ clamped,
    const int* __restrict__ radii, float* __restrict__ grad_rec, float* __restrict__ dL_dmeans2D,
    float* __restrict__ dL_dcolors, float* __restrict__ dL_dopacity, float* __restrict__ dL_dmeans3D,
    float* __restrict__ dL_dtransMat, float* __restrict__ dL_dsh, float* __restrict__ dL_dsh_rest,
    float* __restrict__ dL_dscales, float* __restrict__ dL_drot, float* __restrict__ dL_dscales_raw, const bool staged) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  // Split SH layout (DC (P,1,3) + rest (P,15,3), the trainer's parameters): a warp's 32 surfels own one contiguous
  // 5760-B block of `rest`, so the warp moves it with coalesced 16-B accesses through shared memory — each thread then
  // reads / overwrites its own 45 floats there (stride 45 words: conflict-free) instead of issuing 48 scalar global
  // loads and 48 scalar stores at a 180-B stride.
  extern __shared__ float s_sh[];
  const int lane = threadIdx.x & 31;
  float* wrest = s_sh + (threadIdx.x >> 5) * SH_WARP_FLOATS;
  float* wdc = wrest + 32 * 45;
  const int wbase = idx - lane;                       // first surfel of this warp
  const int nsurf = min(32, p.P - wbase);             // <= 0 for a warp past the end
  // The copy is asynchronous (LDGSTS, no registers): it is in flight while the thread loads its gradient record and
  // projected record and runs the homography part; the wait sits in front of the SH part.
  const bool async_stage = staged && nsurf == 32 && ((reinterpret_cast<uintptr_t>(p.shs) & 15) == 0);
  if (async_stage) {
    const float4* src = reinterpret_cast<const float4*>(p.sh_rest + (size_t)wbase * 45);
    const uint32_t d0 = smem_addr(wrest);
    for (int v = lane; v < 360; v += 32) cp_async16(d0 + 16u * (uint32_t)v, src + v);
    if (lane < 24) cp_async16(smem_addr(wdc) + 16u * (uint32_t)lane, reinterpret_cast<const float4*>(p.shs + (size_t)wbase * 3) + lane);
    cp_async_commit();
  } else if (staged && nsurf > 0) {
    const int nrest = nsurf * 45, nvec = nrest >> 2;
    const float4* src = reinterpret_cast<const float4*>(p.sh_rest + (size_t)wbase * 45);
    for (int v = lane; v < nvec; v += 32) reinterpret_cast<float4*>(wrest)[v] = __ldg(src + v);
    for (int f = 4 * nvec + lane; f < nrest; f += 32) wrest[f] = __ldg(p.sh_rest + (size_t)wbase * 45 + f);
    for (int f = lane; f < nsurf * 3; f += 32) wdc[f] = __ldg(p.shs + (size_t)wbase * 3 + f);
    __syncwarp();
  }
  if (idx < p.P) {
  const bool visible = radii[idx] > 0;
  Activated act;
  if (p.raw && visible) act = activate_surfel(idx, p.means3D, p.d_means3D, p.scales, p.d_scales, p.rotations, p.d_rotations, p.opacities);
  float2 dscale_raw = {0.f, 0.f};
  float dopacity = 0.f;

  float gr[GRAD_REC_FLOATS];
  v3 dmean3D = {0.f, 0.f, 0.f};
  float2 dscale = {0.f, 0.f};
  float4 drot = {0.f, 0.f, 0.f, 0.f};
  float2 dm2d = {0.f, 0.f};
  bool sh_written = false;
#pragma unroll
  for (int i = 0; i < GRAD_REC_FLOATS; i++) gr[i] = 0.f;

  if (visible) {
    float4* g4 = reinterpret_cast<float4*>(grad_rec + (size_t)idx * GRAD_REC_FLOATS);
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const float4 v = g4[i];
      gr[4 * i] = v.x; gr[4 * i + 1] = v.y; gr[4 * i + 2] = v.z; gr[4 * i + 3] = v.w;
      g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // leave the scratch clean for the next frame
    }
    const SurfelRec r = rec[idx];
    const v3 Tu = {r.q0.x, r.q0.y, r.q0.z}, Tv = {r.q0.w, r.q1.x, r.q1.y}, Tw = {r.q1.z, r.q1.w, r.q2.x};

    // (1) centre of the screen-space box -> homography rows
    {
      const v3 sgn = {1.0f, 1.0f, -1.0f};
      const float gx2 = gr[G_M2D], gy2 = gr[G_M2D + 1];
      const float d = dot3(sgn, Tw * Tw);
      const v3 f = sgn * (1.0f / d);
      const v3 dT0 = gx2 * f * Tw;
      const v3 dT1 = gy2 * f * Tw;
      v3 dT3 = gx2 * f * Tu + gy2 * f * Tv;
      const v3 dL_df = (gx2 * Tu * Tw) + (gy2 * Tv * Tw);
      const float dL_dd = dot3(dL_df, f) * (-1.0f / d);
      const v3 dd_dT3 = sgn * Tw * 2.0f;
      dT3 = dT3 + dL_dd * dd_dT3;
      gr[0] += dT0.x; gr[1] += dT0.y; gr[2] += dT0.z;
      gr[3] += dT1.x; gr[4] += dT1.y; gr[5] += dT1.z;
      gr[6] += dT3.x; gr[7] += dT3.y; gr[8] += dT3.z;
      // "projected 2-D gradient" the trainer uses for densification
      const float z = Tw.z;
      dm2d.x = gr[2] * z * (p.focal_x * p.tan_fovx);
      dm2d.y = gr[5] * z * (p.focal_y * p.tan_fovy);
    }

    if (p.transMat_precomp == nullptr) {
      // (2) homography rows -> surfel frame
      const float* m = p.view;
      const m3 Wm = view_rot(m);
      const m3 Wt = transpose3(Wm);
      const float fx = p.focal_x, fy = p.focal_y, cx = p.focal_x * p.tan_fovx, cy = p.focal_y * p.tan_fovy;
      const float2 sc = p.raw ? act.sc : reinterpret_cast<const float2*>(p.scales)[idx];
      const float4 q = p.raw ? act.q : reinterpret_cast<const float4*>(p.rotations)[idx];
      const m3 R = quat_to_rot(q);
      const v3 pw = p.raw ? act.pw : v3{p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
      const v3 p_view = Wm * pw + v3{m[12], m[13], m[14]};
      v3 dM[3];
#pragma unroll
      for (int k = 0; k < 3; k++) dM[k] = {fx * gr[k], fy * gr[3 + k], cx * gr[k] + cy * gr[3 + k] + gr[6 + k]};
      const v3 dRS0 = Wt * dM[0], dRS1 = Wt * dM[1], dpw = Wt * dM[2];
      v3 dtn = Wt * v3{gr[G_NRM], gr[G_NRM + 1], gr[G_NRM + 2]};
      const v3 tn = Wm * R.c2;
      const float cs = dot3(-tn, p_view);
      dtn = dtn * (cs > 0 ? 1.f : -1.f);
      const v3 c0 = dRS0 * sc.x, c1 = dRS1 * sc.y, c2 = dtn;   // columns of dL/dR
      {
        const float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
        const float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
        drot.x = 2.f * (x * (c1.z - c2.y) + y * (c2.x - c0.z) + z * (c0.y - c1.x));
        drot.y = 2.f * (-2.f * x * (c1.y + c2.z) + y * (c0.y + c1.x) + z * (c0.z + c2.x) + w * (c1.z - c2.y));
        drot.z = 2.f * (x * (c0.y + c1.x) - 2.f * y * (c0.x + c2.z) + z * (c1.z + c2.y) + w * (c2.x - c0.z));
        drot.w = 2.f * (x * (c0.z + c2.x) + y * (c1.z + c2.y) - 2.f * z * (c0.x + c1.y) + w * (c0.y - c1.x));
      }
      dscale = {dot3(dRS0, R.c0), dot3(dRS1, R.c1)};
      dmean3D = dpw;
      if (p.raw) {
        // chain through exp (scale), F.normalize (rotation: d/dv of v/|v|) — see activate_surfel
        const float2 ls = reinterpret_cast<const float2*>(p.scales)[idx];
        dscale_raw = {dscale.x * expf(ls.x), dscale.y * expf(ls.y)};
        const float qg = act.q.x * drot.x + act.q.y * drot.y + act.q.z * drot.z + act.q.w * drot.w;
        drot = {(drot.x - act.q.x * qg) / act.qnorm, (drot.y - act.q.y * qg) / act.qnorm,
                (drot.z - act.q.z * qg) / act.qnorm, (drot.w - act.q.w * qg) / act.qnorm};
      }
    }
    dopacity = gr[G_OPA];
    if (p.raw) dopacity *= act.opacity * (1.0f - act.opacity);

  }
  // every lane of the warp (visible or not) waits for ITS copies, then the warp barrier publishes them to the other lanes
  // (a warp in async_stage mode is full, so all 32 lanes are here)
  if (async_stage) { cp_async_wait<0>(); __syncwarp(); }
  if (visible) {
    // (3) colour -> SH coefficients and view direction
    if (p.shs != nullptr) {
      const v3 pw = p.raw ? act.pw : v3{p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
      const v3 campos = {p.campos[0], p.campos[1], p.campos[2]};
      const v3 dir_orig = pw - campos;
      const v3 dir = dir_orig / sqrtf(dot3(dir_orig, dir_orig));
      const int deg = p.D;
      const uint8_t cl = clamped[idx];
      v3 gc = {gr[G_COL], gr[G_COL + 1], gr[G_COL + 2]};
      gc.x *= (cl & 1) ? 0.f : 1.f;
      gc.y *= (cl & 2) ? 0.f : 1.f;
      gc.z *= (cl & 4) ? 0.f : 1.f;
      // Streamed in chunks of 4 coefficients (3 x float4 in, 3 x float4 out) so that no 48-float array stays live:
      //   dL/dsh_k = b_k(dir) * gc,   dL/ddir += grad b_k(dir) * (sh_k . gc)
      // with b_k the real SH basis of the forward pass and grad b_k its analytic derivative w.r.t. the unit direction.
      const float x = dir.x, y = dir.y, z = dir.z;
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      v3 dL_ddir = {0.f, 0.f, 0.f};
      const int ncoef = (deg + 1) * (deg + 1);
      sh_written = true;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float sv[12], go[12];
        if (staged) {
#pragma unroll
          for (int i = 0; i < 12; i++) {
            const int f = 12 * c + i;
            sv[i] = (f < 3) ? wdc[lane * 3 + f] : wrest[lane * 45 + (f - 3)];
          }
        } else {
          load_sh_chunk(p, idx, c, ncoef, sv);
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int k = 4 * c + t;
          float bk, bx, by, bz;
          switch (k) {   // compile-time after unrolling
            case 0: bk = kSH_C0; bx = 0.f; by = 0.f; bz = 0.f; break;
            case 1: bk = -kSH_C1 * y; bx = 0.f; by = -kSH_C1; bz = 0.f; break;
            case 2: bk = kSH_C1 * z; bx = 0.f; by = 0.f; bz = kSH_C1; break;
            case 3: bk = -kSH_C1 * x; bx = -kSH_C1; by = 0.f; bz = 0.f; break;
            case 4: bk = kSH_C2[0] * xy; bx = kSH_C2[0] * y; by = kSH_C2[0] * x; bz = 0.f; break;
            case 5: bk = kSH_C2[1] * yz; bx = 0.f; by = kSH_C2[1] * z; bz = kSH_C2[1] * y; break;
            case 6: bk = kSH_C2[2] * (2.f * zz - xx - yy); bx = kSH_C2[2] * -2.f * x; by = kSH_C2[2] * -2.f * y; bz = kSH_C2[2] * 4.f * z; break;
            case 7: bk = kSH_C2[3] * xz; bx = kSH_C2[3] * z; by = 0.f; bz = kSH_C2[3] * x; break;
            case 8: bk = kSH_C2[4] * (xx - yy); bx = kSH_C2[4] * 2.f * x; by = kSH_C2[4] * -2.f * y; bz = 0.f; break;
            case 9: bk = kSH_C3[0] * y * (3.f * xx - yy); bx = kSH_C3[0] * 6.f * xy; by = kSH_C3[0] * 3.f * (xx - yy); bz = 0.f; break;
            case 10: bk = kSH_C3[1] * xy * z; bx = kSH_C3[1] * yz; by = kSH_C3[1] * xz; bz = kSH_C3[1] * xy; break;
            case 11: bk = kSH_C3[2] * y * (4.f * zz - xx - yy); bx = kSH_C3[2] * -2.f * xy; by = kSH_C3[2] * (-3.f * yy + 4.f * zz - xx); bz = kSH_C3[2] * 8.f * yz; break;
            case 12: bk = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); bx = kSH_C3[3] * -6.f * xz; by = kSH_C3[3] * -6.f * yz; bz = kSH_C3[3] * 3.f * (2.f * zz - xx - yy); break;
            case 13: bk = kSH_C3[4] * x * (4.f * zz - xx - yy); bx = kSH_C3[4] * (-3.f * xx + 4.f * zz - yy); by = kSH_C3[4] * -2.f * xy; bz = kSH_C3[4] * 8.f * xz; break;
            case 14: bk = kSH_C3[5] * z * (xx - yy); bx = kSH_C3[5] * 2.f * xz; by = kSH_C3[5] * -2.f * yz; bz = kSH_C3[5] * (xx - yy); break;
            default: bk = kSH_C3[6] * x * (xx - 3.f * yy); bx = kSH_C3[6] * 3.f * (xx - yy); by = kSH_C3[6] * -6.f * xy; bz = 0.f; break;
          }
          const bool on = k < ncoef;
          go[3 * t] = on ? bk * gc.x : 0.f; go[3 * t + 1] = on ? bk * gc.y : 0.f; go[3 * t + 2] = on ? bk * gc.z : 0.f;
          if (on) {
            const float sg = sv[3 * t] * gc.x + sv[3 * t + 1] * gc.y + sv[3 * t + 2] * gc.z;
            dL_ddir.x += bx * sg; dL_ddir.y += by * sg; dL_ddir.z += bz * sg;
          }
        }
        if (staged) {
#pragma unroll
          for (int i = 0; i < 12; i++) {
            const int f = 12 * c + i;
            if (f < 3) wdc[lane * 3 + f] = go[i]; else wrest[lane * 45 + (f - 3)] = go[i];
          }
        } else {
          store_sh_chunk(p, dL_dsh, dL_dsh_rest, idx, c, go);
        }
      }
      // derivative of v/|v|
      const v3 v = dir_orig, dv = dL_ddir;
      const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
      const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmean3D.x += ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
      dmean3D.y += (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
      dmean3D.z += (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    }
  }

  if (dL_dmeans2D) { dL_dmeans2D[3 * (size_t)idx] = dm2d.x; dL_dmeans2D[3 * (size_t)idx + 1] = dm2d.y; dL_dmeans2D[3 * (size_t)idx + 2] = 0.f; }
  if (dL_dcolors) { dL_dcolors[3 * (size_t)idx] = gr[G_COL]; dL_dcolors[3 * (size_t)idx + 1] = gr[G_COL + 1]; dL_dcolors[3 * (size_t)idx + 2] = gr[G_COL + 2]; }
  if (dL_dopacity) dL_dopacity[idx] = dopacity;
  if (dL_dmeans3D) { dL_dmeans3D[3 * (size_t)idx] = dmean3D.x; dL_dmeans3D[3 * (size_t)idx + 1] = dmean3D.y; dL_dmeans3D[3 * (size_t)idx + 2] = dmean3D.z; }
  if (dL_dtransMat) {
#pragma unroll
    for (int i = 0; i < 9; i++) dL_dtransMat[9 * (size_t)idx + i] = gr[i];
  }
  if (dL_dscales) reinterpret_cast<float2*>(dL_dscales)[idx] = dscale;
  if (dL_dscales_raw) reinterpret_cast<float2*>(dL_dscales_raw)[idx] = dscale_raw;
  if (dL_drot) reinterpret_cast<float4*>(dL_drot)[idx] = drot;
  if (dL_dsh && p.M > 0 && !sh_written) {   // culled surfel: the gradient row is zero
    if (staged) {
#pragma unroll
      for (int f = 0; f < 3; f++) wdc[lane * 3 + f] = 0.f;
#pragma unroll
      for (int f = 0; f < 45; f++) wrest[lane * 45 + f] = 0.f;
    } else {
      float zero12[12];
#pragma unroll
      for (int i = 0; i < 12; i++) zero12[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; c++) store_sh_chunk(p, dL_dsh, dL_dsh_rest, idx, c, zero12);
    }
  }
  }   // idx < P
  if (staged && nsurf > 0) {
    __syncwarp();
    const int nrest = nsurf * 45, nvec = nrest >> 2;
    float4* dst = reinterpret_cast<float4*>(dL_dsh_rest + (size_t)wbase * 45);
    for (int v = lane; v < nvec; v += 32) dst[v] = reinterpret_cast<const float4*>(wrest)[v];
    for (int f = 4 * nvec + lane; f < nrest; f += 32) dL_dsh_rest[(size_t)wbase * 45 + f] = wrest[f];
    for (int f = lane; f < nsurf * 3; f += 32) dL_dsh[(size_t)wbase * 3 + f] = wdc[f];
  }
}

void launch_preprocess_bwd(const BwdParams& p, const SurfelRec* rec, const uint8_t* clamped, const int* radii,
                           float* grad_rec, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                           float* dL_dmeans3D, float* dL_dtransMat, float* dL_dsh, float* dL_dsh_rest,
                           float* dL_dscales, float* dL_drot, float* dL_dscales_raw, cudaStream_t s) {
  if (p.P == 0) return;
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool staged = p.shs && p.sh_rest && p.M == 16 && dL_dsh && dL_dsh_rest && al16(p.sh_rest) && al16(dL_dsh_rest);
  const size_t smem = staged ? sizeof(float) * SH_WARP_FLOATS * 8 : 0;
  preprocess_bwd_kernel<<<(p.P + 255) / 256, 256, smem, s>>>(p, rec, clamped, radii, grad_rec, dL_dmeans2D, dL_dcolors,
                                                            dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dsh_rest,
                                                            dL_dscales, dL_drot, dL_dscales_raw, staged);
}

}  // namespace d2gs
