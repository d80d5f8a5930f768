// 3-D Gaussian (EWA volume splat) rasterizer with colour, depth and alpha outputs — the "diff-gaussian-rasterization"
// variant the reference ships next to the surfel rasterizer (used by render_flow,
// gaussian_renderer/__init__.py:222-337).  Paths below: DGR/ = submodules/diff-gaussian-rasterization/.
// Behavioural contract:
//   DGR/cuda_rasterizer/forward.cu:74-112 (2-D covariance), :117-150 (3-D covariance, quaternion NOT normalised),
//   :153-264 (per-Gaussian stage), :270-376 (per-tile blend: colour + depth + alpha = sum of weights),
//   DGR/cuda_rasterizer/backward.cu:143-270 (conic -> covariance -> mean), :274-338 (scale / rotation), :343-411 (means,
//   depth, SH), :415-578 (blend backward incl. the depth and alpha terms), rasterizer_impl.cu:70-110 (tile instances).
//
// Design (not the reference's): one 48-B record per Gaussian (mean2D, depth, prefilter threshold | conic, opacity |
// rgb) instead of four separate arrays, so the per-tile gather is three 16-B loads of one line; an exact prefilter on
// the exponent (`power < -ln(255 o) - 1e-3` implies alpha < 1/255) skips the exp; the blend backward composes the
// gradient with ONE running scalar instead of the reference's per-channel recurrences (same algebra as the surfel
// kernel), reduces the ten per-Gaussian components over the warp with shuffles and issues one reduction
// per (warp, Gaussian, component) instead of ten atomics per contributing pixel; traversal starts at the deepest list
// position any pixel of the tile consumed; the three per-Gaussian backward kernels of the reference are one kernel.
// Binning (scan, CUB radix sort on 32+bit key bits, ranges) is shared with the surfel path.
#include "raster_common.cuh"
#include "sh_basis.cuh"
#include "gs3d.cuh"

namespace d2gs {
namespace {
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ m3 mul3(const m3& A, const m3& B) { return {A * B.c0, A * B.c1, A * B.c2}; }

__device__ __forceinline__ v3 xform4x3(v3 p, const float* m) {
  return {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
          m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
}
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return (float)((((double)v + 1.0) * S - 1.0) * 0.5); }

// world-space covariance from scale and (raw) quaternion: Sigma = (S R)^T (S R), upper triangle (forward.cu:117-150)
__device__ __forceinline__ void cov3d_from_scale_rot(v3 scale, float mod, float4 q, float* cov) {
  const m3 S = {{mod * scale.x, 0.f, 0.f}, {0.f, mod * scale.y, 0.f}, {0.f, 0.f, mod * scale.z}};
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  const m3 R = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
  const m3 M = mul3(S, R);
  const m3 Sg = mul3(transpose3(M), M);
  cov[0] = Sg.c0.x; cov[1] = Sg.c0.y; cov[2] = Sg.c0.z; cov[3] = Sg.c1.y; cov[4] = Sg.c1.z; cov[5] = Sg.c2.z;
}

// EWA projection of the covariance (forward.cu:74-112).  Returns (a, b, c) of the 2x2 screen covariance with the 0.3
// low-pass already added; T = W J and the clamped view-space mean are handed back for the backward.
struct Ewa { float a, b, c; m3 T; v3 t; float xmul, ymul; };
__device__ __forceinline__ Ewa ewa_project(v3 mean, float fx, float fy, float tanx, float tany, const float* cov3D, const float* view) {
  Ewa e;
  v3 t = xform4x3(mean, view);
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float txtz = t.x / t.z, tytz = t.y / t.z;
  t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
  t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
  e.xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  e.ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  const m3 J = {{fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z)}, {0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z)}, {0.f, 0.f, 0.f}};
  const m3 W = {{view[0], view[4], view[8]}, {view[1], view[5], view[9]}, {view[2], view[6], view[10]}};
  e.T = mul3(W, J);
  const m3 V = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
  const m3 cov = mul3(mul3(transpose3(e.T), transpose3(V)), e.T);
  e.a = cov.c0.x + 0.3f; e.b = cov.c0.y; e.c = cov.c1.y + 0.3f;
  e.t = t;
  return e;
}

// ------------------------------------------------------------------------------------------------------------
// per-Gaussian forward
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) g3_preprocess_fwd_kernel(G3Params p, G3Rec* __restrict__ rec, float* __restrict__ cov3Ds,
                                                                uint8_t* __restrict__ clamped, int* __restrict__ radii,
                                                                uint32_t* __restrict__ tiles_touched, uint4* __restrict__ tile_box) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.P) return;
  radii[idx] = 0;
  tiles_touched[idx] = 0;
  if (tile_box) tile_box[idx] = make_uint4(0u, 0u, 0u, 0u);
  const v3 po = {p.means3D[3 * (size_t)idx], p.means3D[3 * (size_t)idx + 1], p.means3D[3 * (size_t)idx + 2]};
  const v3 p_view = xform4x3(po, p.view);
  if (p_view.z <= 0.2f) {
    if (p.prefiltered) {
      printf("Point is filtered although prefiltered is set. This shouldn't happen!");
      __trap();
    }
    return;
  }
  const float* pm = p.proj;
  const float hx = pm[0] * po.x + pm[4] * po.y + pm[8] * po.z + pm[12];
  const float hy = pm[1] * po.x + pm[5] * po.y + pm[9] * po.z + pm[13];
  const float hw = pm[3] * po.x + pm[7] * po.y + pm[11] * po.z + pm[15];
  const float p_w = 1.0f / (hw + 0.0000001f);
  const float ndc_x = hx * p_w, ndc_y = hy * p_w;

  float* cov3D = cov3Ds + 6 * (size_t)idx;
  if (p.cov3D_precomp != nullptr) {
#pragma unroll
    for (int i = 0; i < 6; i++) cov3D[i] = p.cov3D_precomp[6 * (size_t)idx + i];
  } else {
    const v3 sc = {p.scales[3 * (size_t)idx], p.scales[3 * (size_t)idx + 1], p.scales[3 * (size_t)idx + 2]};
    cov3d_from_scale_rot(sc, p.scale_modifier, reinterpret_cast<const float4*>(p.rotations)[idx], cov3D);
  }
  const Ewa e = ewa_project(po, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D, p.view);
  const float det = e.a * e.c - e.b * e.b;
  if (det == 0.0f) return;
  const float det_inv = 1.f / det;
  const float con_x = e.c * det_inv, con_y = -e.b * det_inv, con_z = e.a * det_inv;
  const float mid = 0.5f * (e.a + e.c);
  const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
  const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
  const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
  const float px = ndc_to_pix(ndc_x, p.W), py = ndc_to_pix(ndc_y, p.H);
  const RectU r = tile_rect(px, py, (int)my_radius, p.gx, p.gy);
  if ((r.x1 - r.x0) * (r.y1 - r.y0) == 0) return;

  v3 rgb;
  uint8_t cl = 0;
  if (p.colors_precomp == nullptr) {
    const v3 campos = {p.campos[0], p.campos[1], p.campos[2]};
    v3 dir = po - campos;
    dir = dir / sqrtf(dot3(dir, dir));
    float sh[48];
    const int ncoef = (p.D + 1) * (p.D + 1);
    const float* base = p.shs + (size_t)idx * p.M * 3;
#pragma unroll
    for (int i = 0; i < 48; i++)
      if (i < ncoef * 3) sh[i] = __ldg(base + i);
    v3 res = eval_sh(p.D, dir, sh);
    res = res + v3{0.5f, 0.5f, 0.5f};
    cl = (res.x < 0 ? 1 : 0) | (res.y < 0 ? 2 : 0) | (res.z < 0 ? 4 : 0);
    rgb = {fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f)};
  } else {
    rgb = {p.colors_precomp[3 * (size_t)idx], p.colors_precomp[3 * (size_t)idx + 1], p.colors_precomp[3 * (size_t)idx + 2]};
  }
  clamped[idx] = cl;
  const float o = p.opacities[idx];
  // exponent below which o*exp(power) < 1/255 with a margin far above the error of expf; o <= 0 never contributes
  const float thr = (o > 0.f) ? -logf(255.f * o) - 1e-3f : ((o <= 0.f) ? INFINITY : NAN);
  G3Rec g;
  g.q0 = make_float4(px, py, p_view.z, thr);
  g.q1 = make_float4(con_x, con_y, con_z, o);
  g.q2 = make_float4(rgb.x, rgb.y, rgb.z, 0.f);
  rec[idx] = g;
  radii[idx] = (int)my_radius;
  tiles_touched[idx] = (r.y1 - r.y0) * (r.x1 - r.x0);
  // packed tile rectangle + depth bits for the per-tile binning (tile_binning.cu), as in the surfel path
  if (tile_box) tile_box[idx] = make_uint4(r.x0 | (r.x1 << 16), r.y0 | (r.y1 << 16), __float_as_uint(p_view.z), 0u);
}

// one (tile | depth) key per tile of the Gaussian's rectangle (rasterizer_impl.cu:70-110)
__global__ void __launch_bounds__(256) g3_duplicate_kernel(int P, const G3Rec* __restrict__ rec, const int* __restrict__ radii,
                                                           const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                                                           uint32_t* __restrict__ vals, uint32_t gx, uint32_t gy) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const int rad = radii[idx];
  if (rad <= 0) return;
  uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
  const float4 q0 = rec[idx].q0;
  const RectU r = tile_rect(q0.x, q0.y, rad, gx, gy);
  const uint32_t dbits = __float_as_uint(q0.z);
  for (uint32_t y = r.y0; y < r.y1; y++)
    for (uint32_t x = r.x0; x < r.x1; x++) {
      keys[off] = ((uint64_t)(y * gx + x) << 32) | dbits;
      vals[off] = (uint32_t)idx;
      off++;
    }
}

// a warp owns a compact 8x4 pixel patch of the 16x16 tile
__device__ __forceinline__ void g3_pixel(int tid, int& lx, int& ly) {
  const int w = tid >> 5, l = tid & 31;
  lx = (w & 1) * 8 + (l & 7);
  ly = (w >> 1) * 4 + (l >> 3);
}

// ------------------------------------------------------------------------------------------------------------
// per-tile blend, forward (forward.cu:270-376)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) g3_blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                           const G3Rec* __restrict__ rec, int W, int H, uint32_t gx,
                                                           const float* __restrict__ bg, float* __restrict__ out_color,
                                                           float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                           uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ status) {
  __shared__ G3Rec s_rec[256];
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % gx, ty = tile / gx;
  int lx, ly;
  g3_pixel(threadIdx.x, lx, ly);
  const uint32_t pxi = tx * TILE_X + lx, pyi = ty * TILE_Y + ly;
  const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
  const float pfx = (float)pxi, pfy = (float)pyi;
  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + 255) / 256;
  bool done = !inside;
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, weight = 0.f, D = 0.f;
  uint32_t contributor = 0, last_contributor = 0;
  int todo = total;
  for (int i = 0; i < rounds; i++, todo -= 256) {
    if (__syncthreads_count(done) == 256) break;
    const int progress = i * 256 + threadIdx.x;
    if (progress < total) s_rec[threadIdx.x] = rec[point_list[range.x + progress]];
    __syncthreads();
    const int nb = min(256, todo);
    for (int j = 0; !done && j < nb; j++) {
      contributor++;
      const float4 q0 = s_rec[j].q0, q1 = s_rec[j].q1;
      const float dx = q0.x - pfx, dy = q0.y - pfy;
      const float power = -0.5f * (q1.x * dx * dx + q1.z * dy * dy) - q1.y * dx * dy;
      if (power > 0.0f) continue;
      if (power < q0.w) continue;                       // alpha < 1/255 for certain: no exp
      const float alpha = fminf(0.99f, q1.w * expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1 - alpha);
      if (test_T < 0.0001f) { done = true; continue; }
      const float4 q2 = s_rec[j].q2;
      const float w = alpha * T;
      C0 += q2.x * w; C1 += q2.y * w; C2 += q2.z * w;
      weight += w;
      D += q0.z * w;
      T = test_T;
      last_contributor = contributor;
    }
  }
  if (inside) {
    const size_t pix = (size_t)pyi * W + pxi, HW = (size_t)H * W;
    n_contrib[pix] = last_contributor;
    // deferred-count mode: a frame whose instance count exceeded the binning capacity is never silently wrong
    const bool poisoned = status != nullptr && status[1] != 0u;
    const float qnan = __int_as_float(0x7fc00000);
    out_color[pix] = poisoned ? qnan : C0 + T * bg[0];
    out_color[HW + pix] = poisoned ? qnan : C1 + T * bg[1];
    out_color[2 * HW + pix] = poisoned ? qnan : C2 + T * bg[2];
    out_alpha[pix] = weight;
    out_depth[pix] = D;
  }
}

// ------------------------------------------------------------------------------------------------------------
// per-tile blend, backward (backward.cu:415-578)
//   pixel = sum_i w_i e_i + T_final b,  w_i = alpha_i T_i,  e_i = c_i . g_c + depth_i g_d + g_a,  b = bg . g_c
//   dL/dalpha_i = T_i e_i - Q_i / (1 - alpha_i),   Q_i = T_final b + sum_{k behind i} w_k e_k        (one running scalar)
// Gradient record per Gaussian (12 floats): [0..1] mean2D (NDC units), [2..4] conic (x, y/2, z), [5] opacity,
// [6..8] colour, [9] depth.
// ------------------------------------------------------------------------------------------------------------
constexpr int G3_GRAD = 12;
__global__ void __launch_bounds__(256) g3_blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                           const G3Rec* __restrict__ rec, int W, int H, uint32_t gx,
                                                           const float* __restrict__ bg, const float* __restrict__ alphas,
                                                           const uint32_t* __restrict__ n_contrib,
                                                           const float* __restrict__ dL_dpix, const float* __restrict__ dL_ddepth,
                                                           const float* __restrict__ dL_dalpha_pix, float* __restrict__ grad) {
  __shared__ G3Rec s_rec[256];
  __shared__ uint32_t s_id[256];
  __shared__ uint32_t s_max;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile % gx, ty = tile / gx;
  int lx, ly;
  g3_pixel(threadIdx.x, lx, ly);
  const uint32_t pxi = tx * TILE_X + lx, pyi = ty * TILE_Y + ly;
  const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
  const float pfx = (float)pxi, pfy = (float)pyi;
  const uint2 range = ranges[tile];
  const size_t pix = (size_t)pyi * W + pxi, HW = (size_t)H * W;
  const float T_final = inside ? (1 - alphas[pix]) : 0;
  const uint32_t last_contributor = inside ? n_contrib[pix] : 0;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f, ga = 0.f;
  if (inside) { g0 = dL_dpix[pix]; g1 = dL_dpix[HW + pix]; g2 = dL_dpix[2 * HW + pix]; gd = dL_ddepth[pix]; ga = dL_dalpha_pix[pix]; }
  // the tile's list is walked back to front starting at the deepest position any pixel consumed
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  {
    const uint32_t m = __reduce_max_sync(FULL, last_contributor);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, m);
  }
  __syncthreads();
  const int total = (int)min(s_max, range.y - range.x);
  const int rounds = (total + 255) / 256;
  float T = T_final;
  float Q = T_final * (bg[0] * g0 + bg[1] * g1 + bg[2] * g2);
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  uint32_t contributor = (uint32_t)total;
  int todo = total;
  for (int i = 0; i < rounds; i++, todo -= 256) {
    __syncthreads();
    const int progress = i * 256 + threadIdx.x;
    if (progress < total) {
      const uint32_t id = point_list[range.x + total - progress - 1];
      s_id[threadIdx.x] = id;
      s_rec[threadIdx.x] = rec[id];
    }
    __syncthreads();
    const int nb = min(256, todo);
    for (int j = 0; j < nb; j++) {
      contributor--;
      const float4 q0 = s_rec[j].q0, q1 = s_rec[j].q1;
      const float dx = q0.x - pfx, dy = q0.y - pfy;
      const float power = -0.5f * (q1.x * dx * dx + q1.z * dy * dy) - q1.y * dx * dy;
      bool on = contributor < last_contributor && !(power > 0.0f) && !(power < q0.w);
      float G = 0.f, alpha = 0.f;
      if (on) {
        G = expf(power);
        alpha = fminf(0.99f, q1.w * G);
        on = !(alpha < 1.0f / 255.0f);
      }
      if (!__any_sync(FULL, on)) continue;
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; k++) v[k] = 0.f;
      if (on) {
        const float4 q2 = s_rec[j].q2;
        T = T / (1.f - alpha);
        const float w = alpha * T;
        const float E = q2.x * g0 + q2.y * g1 + q2.z * g2 + q0.z * gd + ga;
        const float dL_dopa = T * E - Q / (1.f - alpha);
        Q += w * E;
        const float dL_dG = q1.w * dL_dopa;
        const float gdx = G * dx, gdy = G * dy;
        const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
        const float dG_ddely = -gdy * q1.z - gdx * q1.y;
        v[0] = dL_dG * dG_ddelx * ddelx_dx;
        v[1] = dL_dG * dG_ddely * ddely_dy;
        v[2] = -0.5f * gdx * dx * dL_dG;
        v[3] = -0.5f * gdx * dy * dL_dG;
        v[4] = -0.5f * gdy * dy * dL_dG;
        v[5] = G * dL_dopa;
        v[6] = w * g0; v[7] = w * g1; v[8] = w * g2;
        v[9] = w * gd;
      }
#pragma unroll
      for (int k = 0; k < 10; k++) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], s);
      }
      const int lane = threadIdx.x & 31;
      if (lane < 10) {
        float mine = v[0];
#pragma unroll
        for (int k = 1; k < 10; k++) mine = (lane == k) ? v[k] : mine;
        atomicAdd(grad + (size_t)s_id[j] * G3_GRAD + lane, mine);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// per-Gaussian backward: conic -> 2-D covariance -> (3-D covariance, mean); projected mean and depth -> mean;
// colour -> SH and view direction -> mean; 3-D covariance -> scale and quaternion.  Every output element is written.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) g3_preprocess_bwd_kernel(G3Params p, const float* __restrict__ cov3Ds,
                                                                const uint8_t* __restrict__ clamped, const int* __restrict__ radii,
                                                                const float* __restrict__ grad, float* __restrict__ dL_dmeans2D,
                                                                float* __restrict__ dL_dcolors, float* __restrict__ dL_dopacity,
                                                                float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dcov3D,
                                                                float* __restrict__ dL_dsh, float* __restrict__ dL_dscales,
                                                                float* __restrict__ dL_drot) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.P) return;
  const size_t i = (size_t)idx;
  const bool live = radii[idx] > 0;
  float g[G3_GRAD];
#pragma unroll
  for (int k = 0; k < G3_GRAD; k++) g[k] = live ? grad[i * G3_GRAD + k] : 0.f;
  if (dL_dmeans2D) { dL_dmeans2D[3 * i] = g[0]; dL_dmeans2D[3 * i + 1] = g[1]; dL_dmeans2D[3 * i + 2] = 0.f; }
  if (dL_dopacity) dL_dopacity[i] = g[5];
  if (dL_dcolors) { dL_dcolors[3 * i] = g[6]; dL_dcolors[3 * i + 1] = g[7]; dL_dcolors[3 * i + 2] = g[8]; }
  float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  v3 dmean = {0.f, 0.f, 0.f};
  float dscale[3] = {0.f, 0.f, 0.f};
  float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ncoef_all = p.M;
  if (live) {
    const v3 m = {p.means3D[3 * i], p.means3D[3 * i + 1], p.means3D[3 * i + 2]};
    const float* cov3D = cov3Ds + 6 * i;
    // (1) conic -> 2-D covariance (a, b, c) -> 3-D covariance and the view-space mean
    const Ewa e = ewa_project(m, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D, p.view);
    const float a = e.a, b = e.b, c = e.c;
    const float denom = a * c - b * b;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float da = 0.f, db = 0.f, dc = 0.f;
    const v3 u = e.T.c0, w = e.T.c1;      // a = u.V u, b = u.V w, c = w.V w
    if (denom2inv != 0) {
      const float cx = g[2], cy = g[3], cz = g[4];
      da = denom2inv * (-c * c * cx + 2 * b * c * cy + (denom - a * c) * cz);
      dc = denom2inv * (-a * a * cz + 2 * a * b * cy + (denom - a * c) * cx);
      db = denom2inv * 2 * (b * c * cx - (denom + 2 * b * b) * cy + a * b * cz);
      dcov[0] = u.x * u.x * da + u.x * w.x * db + w.x * w.x * dc;
      dcov[3] = u.y * u.y * da + u.y * w.y * db + w.y * w.y * dc;
      dcov[5] = u.z * u.z * da + u.z * w.z * db + w.z * w.z * dc;
      dcov[1] = 2 * u.x * u.y * da + (u.x * w.y + u.y * w.x) * db + 2 * w.x * w.y * dc;
      dcov[2] = 2 * u.x * u.z * da + (u.x * w.z + u.z * w.x) * db + 2 * w.x * w.z * dc;
      dcov[4] = 2 * u.z * u.y * da + (u.y * w.z + u.z * w.y) * db + 2 * w.y * w.z * dc;
    }
    const m3 V = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
    const v3 Vu = V * u, Vw = V * w;
    const v3 du = 2.f * da * Vu + db * Vw;     // dL/du
    const v3 dw = 2.f * dc * Vw + db * Vu;     // dL/dw
    const float* vm = p.view;
    const v3 W0 = {vm[0], vm[4], vm[8]}, W1 = {vm[1], vm[5], vm[9]}, W2 = {vm[2], vm[6], vm[10]};
    const float dJ00 = dot3(W0, du), dJ02 = dot3(W2, du), dJ11 = dot3(W1, dw), dJ12 = dot3(W2, dw);
    const float tz = 1.f / e.t.z, tz2 = tz * tz, tz3 = tz2 * tz;
    const float hx = p.focal_x, hy = p.focal_y;
    const float dtx = e.xmul * -hx * tz2 * dJ02;
    const float dty = e.ymul * -hy * tz2 * dJ12;
    const float dtz = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * e.t.x) * tz3 * dJ02 + (2 * hy * e.t.y) * tz3 * dJ12;
    dmean = {vm[0] * dtx + vm[1] * dty + vm[2] * dtz, vm[4] * dtx + vm[5] * dty + vm[6] * dtz, vm[8] * dtx + vm[9] * dty + vm[10] * dtz};
    // (2) projected mean (NDC gradient) -> mean
    const float* pm = p.proj;
    const float hw = pm[3] * m.x + pm[7] * m.y + pm[11] * m.z + pm[15];
    const float m_w = 1.0f / (hw + 0.0000001f);
    const float mul1 = (pm[0] * m.x + pm[4] * m.y + pm[8] * m.z + pm[12]) * m_w * m_w;
    const float mul2 = (pm[1] * m.x + pm[5] * m.y + pm[9] * m.z + pm[13]) * m_w * m_w;
    dmean.x += (pm[0] * m_w - pm[3] * mul1) * g[0] + (pm[1] * m_w - pm[3] * mul2) * g[1];
    dmean.y += (pm[4] * m_w - pm[7] * mul1) * g[0] + (pm[5] * m_w - pm[7] * mul2) * g[1];
    dmean.z += (pm[8] * m_w - pm[11] * mul1) * g[0] + (pm[9] * m_w - pm[11] * mul2) * g[1];
    // (3) depth -> mean (backward.cu:391-403)
    const float mul3v = vm[2] * m.x + vm[6] * m.y + vm[10] * m.z + vm[14];
    dmean.x += (vm[2] - vm[3] * mul3v) * g[9];
    dmean.y += (vm[6] - vm[7] * mul3v) * g[9];
    dmean.z += (vm[10] - vm[11] * mul3v) * g[9];
    // (4) colour -> SH coefficients and view direction
    if (p.shs != nullptr) {
      const uint8_t cl = clamped[idx];
      const v3 gc = {(cl & 1) ? 0.f : g[6], (cl & 2) ? 0.f : g[7], (cl & 4) ? 0.f : g[8]};
      const v3 campos = {p.campos[0], p.campos[1], p.campos[2]};
      const v3 dir_orig = m - campos;
      const v3 dir = dir_orig / sqrtf(dot3(dir_orig, dir_orig));
      const int ncoef = (p.D + 1) * (p.D + 1);
      const float* base = p.shs + i * p.M * 3;
      v3 ddir = {0.f, 0.f, 0.f};
      for (int k = 0; k < ncoef_all; k++) {
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        if (k < ncoef) {
          float bk, bx, by, bz;
          sh_basis_grad(k, dir.x, dir.y, dir.z, bk, bx, by, bz);
          const float sg = __ldg(base + 3 * k) * gc.x + __ldg(base + 3 * k + 1) * gc.y + __ldg(base + 3 * k + 2) * gc.z;
          ddir = ddir + v3{bx, by, bz} * sg;
          o0 = bk * gc.x; o1 = bk * gc.y; o2 = bk * gc.z;
        }
        if (dL_dsh) { dL_dsh[(i * p.M + k) * 3] = o0; dL_dsh[(i * p.M + k) * 3 + 1] = o1; dL_dsh[(i * p.M + k) * 3 + 2] = o2; }
      }
      // through the normalisation of the direction
      const float sum2 = dot3(dir_orig, dir_orig);
      const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      const v3 v = dir_orig;
      dmean.x += ((sum2 - v.x * v.x) * ddir.x - v.y * v.x * ddir.y - v.z * v.x * ddir.z) * inv32;
      dmean.y += (-v.x * v.y * ddir.x + (sum2 - v.y * v.y) * ddir.y - v.z * v.y * ddir.z) * inv32;
      dmean.z += (-v.x * v.z * ddir.x - v.y * v.z * ddir.y + (sum2 - v.z * v.z) * ddir.z) * inv32;
    }
    // (5) 3-D covariance -> scale and quaternion (backward.cu:274-338): Sigma = M^T M, M = S R
    if (p.scales != nullptr) {
      const float4 q = reinterpret_cast<const float4*>(p.rotations)[idx];
      const float r = q.x, x = q.y, y = q.z, z = q.w;
      const m3 R = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                    {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                    {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
      const v3 s = {p.scale_modifier * p.scales[3 * i], p.scale_modifier * p.scales[3 * i + 1], p.scale_modifier * p.scales[3 * i + 2]};
      const m3 S = {{s.x, 0.f, 0.f}, {0.f, s.y, 0.f}, {0.f, 0.f, s.z}};
      const m3 M = mul3(S, R);
      const m3 dSigma = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                         {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
      const m3 MdS = mul3(M, dSigma);
      const m3 dM = {2.0f * MdS.c0, 2.0f * MdS.c1, 2.0f * MdS.c2};
      const m3 Rt = transpose3(R);
      m3 dMt = transpose3(dM);
      dscale[0] = dot3(Rt.c0, dMt.c0); dscale[1] = dot3(Rt.c1, dMt.c1); dscale[2] = dot3(Rt.c2, dMt.c2);
      dMt.c0 = dMt.c0 * s.x; dMt.c1 = dMt.c1 * s.y; dMt.c2 = dMt.c2 * s.z;
      // entries dMt[col][row]
      const float m00 = dMt.c0.x, m01 = dMt.c0.y, m02 = dMt.c0.z, m10 = dMt.c1.x, m11 = dMt.c1.y, m12 = dMt.c1.z,
                  m20 = dMt.c2.x, m21 = dMt.c2.y, m22 = dMt.c2.z;
      dq.x = 2 * z * (m01 - m10) + 2 * y * (m20 - m02) + 2 * x * (m12 - m21);
      dq.y = 2 * y * (m10 + m01) + 2 * z * (m20 + m02) + 2 * r * (m12 - m21) - 4 * x * (m22 + m11);
      dq.z = 2 * x * (m10 + m01) + 2 * r * (m20 - m02) + 2 * z * (m12 + m21) - 4 * y * (m22 + m00);
      dq.w = 2 * r * (m01 - m10) + 2 * x * (m20 + m02) + 2 * y * (m12 + m21) - 4 * z * (m11 + m00);
    }
  } else if (dL_dsh) {
    for (int k = 0; k < ncoef_all * 3; k++) dL_dsh[i * p.M * 3 + k] = 0.f;
  }
  if (live && p.shs == nullptr && dL_dsh) {
    for (int k = 0; k < ncoef_all * 3; k++) dL_dsh[i * p.M * 3 + k] = 0.f;
  }
  if (dL_dmeans3D) { dL_dmeans3D[3 * i] = dmean.x; dL_dmeans3D[3 * i + 1] = dmean.y; dL_dmeans3D[3 * i + 2] = dmean.z; }
  if (dL_dcov3D) {
#pragma unroll
    for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = dcov[k];
  }
  if (dL_dscales) { dL_dscales[3 * i] = dscale[0]; dL_dscales[3 * i + 1] = dscale[1]; dL_dscales[3 * i + 2] = dscale[2]; }
  if (dL_drot) reinterpret_cast<float4*>(dL_drot)[idx] = dq;
}

}  // namespace

void g3_launch_preprocess_fwd(const G3Params& p, G3Rec* rec, float* cov3Ds, uint8_t* clamped, int* radii, uint32_t* tiles_touched,
                              uint4* tile_box, cudaStream_t s) {
  if (p.P == 0) return;
  g3_preprocess_fwd_kernel<<<(p.P + 255) / 256, 256, 0, s>>>(p, rec, cov3Ds, clamped, radii, tiles_touched, tile_box);
}
void g3_launch_duplicate(int P, const G3Rec* rec, const int* radii, const uint32_t* offsets, uint64_t* keys, uint32_t* vals,
                         uint32_t gx, uint32_t gy, cudaStream_t s) {
  if (P == 0) return;
  g3_duplicate_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, rec, radii, offsets, keys, vals, gx, gy);
}
void g3_launch_blend_fwd(const G3Params& p, const uint2* ranges, const uint32_t* point_list, const G3Rec* rec, float* out_color,
                         float* out_depth, float* out_alpha, uint32_t* n_contrib, const uint32_t* status, cudaStream_t s) {
  g3_blend_fwd_kernel<<<p.gx * p.gy, 256, 0, s>>>(ranges, point_list, rec, p.W, p.H, p.gx, p.bg, out_color, out_depth, out_alpha,
                                                  n_contrib, status);
}
void g3_launch_blend_bwd(const G3Params& p, const uint2* ranges, const uint32_t* point_list, const G3Rec* rec, const float* alphas,
                         const uint32_t* n_contrib, const float* dL_dpix, const float* dL_ddepth, const float* dL_dalpha,
                         float* grad, cudaStream_t s) {
  g3_blend_bwd_kernel<<<p.gx * p.gy, 256, 0, s>>>(ranges, point_list, rec, p.W, p.H, p.gx, p.bg, alphas, n_contrib, dL_dpix,
                                                  dL_ddepth, dL_dalpha, grad);
}
void g3_launch_preprocess_bwd(const G3Params& p, const float* cov3Ds, const uint8_t* clamped, const int* radii, const float* grad,
                              float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D, float* dL_dcov3D,
                              float* dL_dsh, float* dL_dscales, float* dL_drot, cudaStream_t s) {
  if (p.P == 0) return;
  g3_preprocess_bwd_kernel<<<(p.P + 255) / 256, 256, 0, s>>>(p, cov3Ds, clamped, radii, grad, dL_dmeans2D, dL_dcolors, dL_dopacity,
                                                             dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drot);
}

}  // namespace d2gs
