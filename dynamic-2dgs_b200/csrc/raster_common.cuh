// Shared definitions of the surfel rasterizer kernels (sm_100a).
//
// HBM layout (all offsets 256-B aligned inside the caller-owned workspaces):
//   geometry workspace : SurfelRec[P]  (5 x float4 = 80 B, one contiguous record per surfel so the per-tile
//                        gather of the blend kernels touches one or two 128-B lines per instance)
//                        | clamped u8[P] (bit c set = colour channel c clamped at 0)
//                        | tiles_touched u32[P] | point_offsets u32[P] | scan scratch
//   image workspace    : ranges uint2[tiles] | final_T,dist1,dist2 f32[3*H*W] | n_contrib,median u32[2*H*W]
//   binning workspace  : keys_unsorted u64[R] | keys_sorted u64[R] | values_unsorted u32[R] | point_list u32[R]
//                        | radix-sort scratch
//   grad scratch       : GradRec[P] (20 floats) accumulated by the blend backward, consumed+cleared per surfel
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>
#include <stddef.h>

namespace d2gs {

constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int NUM_CH = 3;
#define D2GS_FILTER_SIZE 0.7071067811865476   // double literal, like the reference macro (auxiliary.h:20)
#define D2GS_NEAR_PLANE 0.2
#define D2GS_FAR_PLANE 100.0

// spherical-harmonics constants
__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                                 -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                                 0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                                 -0.5900435899266435f};

// Projected surfel record, 80 B.
//  q0 = (Tu.x, Tu.y, Tu.z, Tv.x)   q1 = (Tv.y, Tv.z, Tw.x, Tw.y)   q2 = (Tw.z, mean2D.x, mean2D.y, tau)
//  q3 = (normal.x, normal.y, normal.z, depth)                      q4 = (r, g, b, opacity)
//  q5 = (x0, y0, x1, y1): screen box outside which every pixel centre is certainly rejected (see cull_box)
//  tau = 2*ln(255*opacity) + 1e-4: a pixel whose Mahalanobis term exceeds tau has alpha < 1/255 with margin, so the
//  blend kernels can drop it after ~20 instructions (q0..q2 only) with exactly the reference's outcome.
struct __align__(16) SurfelRec {
  float4 q0, q1, q2, q3, q4, q5;
};
static_assert(sizeof(SurfelRec) == 96, "SurfelRec must be 96 bytes");
constexpr int REC_QUADS = 6;

// Per-surfel raster gradients accumulated by the blend backward (80 B).
//  [0..8] dL/dtransMat  [9..10] dL/dmean2D.xy  [11..13] dL/dnormal  [14] dL/dopacity  [15..17] dL/dcolor  [18..19] pad
constexpr int GRAD_REC_FLOATS = 20;
constexpr int G_T = 0, G_M2D = 9, G_NRM = 11, G_OPA = 14, G_COL = 15;

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct GeomLayout {
  size_t rec, clamped, tiles_touched, point_offsets, scan_temp, status, tile_box, total;
  size_t scan_temp_bytes;
};
struct ImgLayout {
  size_t ranges, final_T, n_contrib, tile_count, seg_begin, seg_end, tile_order, total;
};
struct BinLayout {
  size_t keys_unsorted, keys_sorted, vals_unsorted, point_list, sort_temp, hit_mask, total;
  size_t sort_temp_bytes;
};
constexpr int TILE_SORT_MAX_TILES = 24576;   // per-tile binning keeps two counters per tile in shared memory (192 KB at the limit)

GeomLayout geom_layout(int P);
ImgLayout img_layout(int W, int H);
BinLayout bin_layout(int64_t R);

struct RectU { uint32_t x0, y0, x1, y1; };

// Conservative rejection of a (pixel, surfel) pair from the homogeneous hit point p and the screen distance:
// true only if BOTH the ray-splat term (|p.xy|^2 / p.z^2) and the low-pass term (rho2d) are safely above tau,
// i.e. min(rho3d, rho2d) > tau => opacity*exp(-rho/2) < 1/255.  Non-finite inputs never reject (slow path decides).
__device__ __forceinline__ bool pair_rejected(float px, float py, float pz, float rho2d, float tau) {
  const float m2 = px * px + py * py;
  const float z2 = pz * pz;
  return (rho2d > tau) && (m2 > tau * z2) && (m2 < 3.0e38f);
}

// Conservative screen-space box of the pixels that can survive pair_rejected(): the union of
//   (a) the disc |pix - mean2D|^2 <= tau/2 of the low-pass term, and
//   (b) the image of the splat-plane disc u^2+v^2 <= tau.  A pixel column x touches that disc iff the line
//       (Tu - x*Tw).(u,v,1) = 0 comes within sqrt(tau) of the origin, i.e.  A x^2 - 2 B x + C <= 0  with
//       g = (tau, tau, -1), A = g.(Tw*Tw), B = g.(Tu*Tw), C = g.(Tu*Tu)   (tau = 1 gives the reference's 1-sigma box).
// tau is inflated by 1e-3 relative + 1e-3 absolute over the prefilter threshold and the box by 1e-4 relative + 0.01 px, far
// above fp32 evaluation error; when the disc reaches the camera plane (A >= 0 up to a guard band) the box is infinite.
// A NaN anywhere yields a box that never culls.
__device__ __forceinline__ float4 cull_box(const float* T, float mx, float my, float tau_prefilter) {
  const float inf = __int_as_float(0x7f800000);
  if (!(tau_prefilter > 0.f)) {
    // opacity < 1/255: the prefilter already rejects every pair (tau <= 0); NaN keeps the box open
    return (tau_prefilter <= 0.f) ? make_float4(inf, inf, -inf, -inf) : make_float4(-inf, -inf, inf, inf);
  }
  const float tau = tau_prefilter * 1.001f + 1e-3f;
  const float r2 = sqrtf(0.5f * tau) + 0.01f;
  float x0 = mx - r2, x1 = mx + r2, y0 = my - r2, y1 = my + r2;
  const float Axy = T[6] * T[6] + T[7] * T[7];
  const float A = tau * Axy - T[8] * T[8];
  if (!(A < -1e-3f * (T[8] * T[8]))) return make_float4(-inf, -inf, inf, inf);
  const float Bx = tau * (T[0] * T[6] + T[1] * T[7]) - T[2] * T[8];
  const float Cx = tau * (T[0] * T[0] + T[1] * T[1]) - T[2] * T[2];
  const float By = tau * (T[3] * T[6] + T[4] * T[7]) - T[5] * T[8];
  const float Cy = tau * (T[3] * T[3] + T[4] * T[4]) - T[5] * T[5];
  const float ia = 1.0f / A;
  const float cxm = Bx * ia, cym = By * ia;
  const float ex = sqrtf(fmaxf(0.f, Bx * Bx - A * Cx)) * fabsf(ia), ey = sqrtf(fmaxf(0.f, By * By - A * Cy)) * fabsf(ia);
  const float mxg = 1e-4f * (fabsf(cxm) + ex) + 0.01f, myg = 1e-4f * (fabsf(cym) + ey) + 0.01f;
  x0 = fminf(x0, cxm - ex - mxg); x1 = fmaxf(x1, cxm + ex + mxg);
  y0 = fminf(y0, cym - ey - myg); y1 = fmaxf(y1, cym + ey + myg);
  if (!(x0 == x0 && x1 == x1 && y0 == y0 && y1 == y1)) return make_float4(-inf, -inf, inf, inf);
  return make_float4(x0, y0, x1, y1);
}

// 32-bit shared-window address, made opaque (volatile asm) so the compiler keeps it in a register instead of
// re-deriving it from the CTA id in front of every load of the inner loop.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  uint32_t a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
  return a;
}
// 16-byte asynchronous global->shared copy (LDGSTS): the staging of the next batch of records overlaps the blending of
// the current one.
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// 32 x 32 bit-matrix transpose across a warp: lane r holds row r, afterwards lane c holds column c (bit r = old row r's bit c).
// Five block-swap steps (16, 8, 4, 2, 1), one shuffle each.
__device__ __forceinline__ uint32_t transpose_bits32(uint32_t x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
  }
  return x;
}    // cull-box survivors per phase-1/phase-2 round (one or two 32-bit masks per lane)

// Tile rectangle of a surfel (reference: auxiliary.h:64-74).  Float arithmetic and the float->int truncation
// are part of the contract: tile lists must be bit-exact.
__device__ __forceinline__ RectU tile_rect(float px, float py, int max_radius, uint32_t gx, uint32_t gy) {
  RectU r;
  r.x0 = min(gx, (uint32_t)max(0, (int)((px - max_radius) / TILE_X)));
  r.y0 = min(gy, (uint32_t)max(0, (int)((py - max_radius) / TILE_Y)));
  r.x1 = min(gx, (uint32_t)max(0, (int)((px + max_radius + TILE_X - 1) / TILE_X)));
  r.y1 = min(gy, (uint32_t)max(0, (int)((py + max_radius + TILE_Y - 1) / TILE_Y)));
  return r;
}

// ---- minimal column-major 3-vector algebra; operators are component-wise so that the float-op DAG the
// ---- compiler sees (and therefore its FMA contraction) is the one a generic vector library produces.
struct v3 { float x, y, z; };
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ v3 operator-(v3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ v3 operator*(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ v3 operator*(v3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ v3 operator*(float s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ v3 operator/(v3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ float dot3(v3 a, v3 b) {
  v3 t = a * b;
  return t.x + t.y + t.z;
}
__device__ __forceinline__ v3 cross3(v3 a, v3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
struct m3 { v3 c0, c1, c2; };   // columns
// matrix * vector with the roundings of the reference build's glm product (m0*x + m1*y + m2*z compiles to
// fma(m2, z, fma(m0, x, rn(m1*y))) there): pinned, because left to the compiler the SAME source line was contracted one
// way for the second homography column and another way for the first (found at BASELINE sizes in round 2: the first
// column of transMat differed from the reference in the last bits for ~10 % of the surfels).
__device__ __forceinline__ float mad3_glm(float a0, float b0, float a1, float b1, float a2, float b2) {
  return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}
__device__ __forceinline__ v3 operator*(const m3& m, v3 v) {
  return {mad3_glm(m.c0.x, v.x, m.c1.x, v.y, m.c2.x, v.z), mad3_glm(m.c0.y, v.x, m.c1.y, v.y, m.c2.y, v.z),
          mad3_glm(m.c0.z, v.x, m.c1.z, v.y, m.c2.z, v.z)};
}
__device__ __forceinline__ m3 transpose3(const m3& m) {
  return {{m.c0.x, m.c1.x, m.c2.x}, {m.c0.y, m.c1.y, m.c2.y}, {m.c0.z, m.c1.z, m.c2.z}};
}
// world->view rotation block of the (transposed, column-major indexed) view matrix
__device__ __forceinline__ m3 view_rot(const float* v) {
  return {{v[0], v[1], v[2]}, {v[4], v[5], v[6]}, {v[8], v[9], v[10]}};
}
// unit-quaternion (w,x,y,z) -> rotation, columns (reference: auxiliary.h:188-210).
// Every rounding is pinned to what the reference build executes (its SASS shares y*y and z*z between the three diagonal
// entries: 1-2(yy+zz) adds two ROUNDED squares, the other two fuse x*x into the addition), so the homography
// rows — and with them every alpha of the blend — come out bit-identical; left to the compiler, the first column differed
// in the last bit for ~10 % of the surfels (found by the live comparison at BASELINE sizes, round 2).
__device__ __forceinline__ m3 quat_to_rot(float4 q) {
  const float s = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  const float w = __fmul_rn(q.x, s), x = __fmul_rn(q.y, s), y = __fmul_rn(q.z, s), z = __fmul_rn(q.w, s);
  const float wz = __fmul_rn(w, z), wy = __fmul_rn(w, y), wx = __fmul_rn(w, x);
  const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  const float d0 = __fadd_rn(yy, zz), d1 = __fmaf_rn(x, x, zz), d2 = __fmaf_rn(x, x, yy);
  const float xy_p = __fmaf_rn(x, y, wz), xy_m = __fmaf_rn(x, y, -wz);
  const float yz_p = __fmaf_rn(y, z, wx), yz_m = __fmaf_rn(y, z, -wx);
  const float xz_p = __fmaf_rn(x, z, wy), xz_m = __fmaf_rn(x, z, -wy);
  return {{__fsub_rn(1.f, __fadd_rn(d0, d0)), __fadd_rn(xy_p, xy_p), __fadd_rn(xz_m, xz_m)},
          {__fadd_rn(xy_m, xy_m), __fsub_rn(1.f, __fadd_rn(d1, d1)), __fadd_rn(yz_p, yz_p)},
          {__fadd_rn(xz_p, xz_p), __fadd_rn(yz_m, yz_m), __fsub_rn(1.f, __fadd_rn(d2, d2))}};
}

// Depth mapped to [0,1] between the near and far planes for the distortion bookkeeping.  The reference evaluates
//   (FAR*d - FAR*NEAR) / ((FAR-NEAR)*d)   in DOUBLE (its macros are double literals, forward.cu:399, backward.cu:352) and
// rounds to float.  The distortion terms m*m*A + dist2 - 2*m*dist1 cancel to ~1e-3 of their operands, so a last-bit
// difference in m shows up as a 1e-3 relative difference of the distortion map: m has to be the reference's float
// exactly.  Instead of a double division per contributing (pixel, surfel) pair (14 half-rate DFMAs), m = a - b/d with
// a = 100/99.8, b = 20/99.8 is evaluated in float-float arithmetic (error < 2^-44) and rounded once; when the result lies
// within 2^-14 ulp of a rounding boundary — one evaluation in ~10^4 — the double expression itself decides.
static __device__ __noinline__ float mapped_depth_double(float d) {
  return (float)((D2GS_FAR_PLANE * d - D2GS_FAR_PLANE * D2GS_NEAR_PLANE) / ((D2GS_FAR_PLANE - D2GS_NEAR_PLANE) * d));
}
__device__ __forceinline__ float mapped_depth(float d) {
  constexpr double C = D2GS_FAR_PLANE - D2GS_NEAR_PLANE;
  constexpr double A = D2GS_FAR_PLANE / C, B = (D2GS_FAR_PLANE * D2GS_NEAR_PLANE) / C;
  constexpr float a_hi = (float)A, a_lo = (float)(A - (double)a_hi), b_hi = (float)B, b_lo = (float)(B - (double)b_hi);
  const float q_hi = __fdiv_rn(b_hi, d);
  const float r = __fmaf_rn(-q_hi, d, b_hi);                    // exact remainder of the rounded quotient
  const float q_lo = __fdividef(__fadd_rn(r, b_lo), d);
  const float s = __fsub_rn(a_hi, q_hi);                        // a_hi >= q_hi for every depth behind the near plane
  const float e = __fsub_rn(__fsub_rn(a_hi, s), q_hi);          // Fast2Sum: exact rounding error of s
  const float t = __fadd_rn(e, __fsub_rn(a_lo, q_lo));
  const float m = __fadd_rn(s, t);
  // distance of |t| from half an ulp of s, relative to that half ulp
  const float h = __int_as_float((__float_as_int(s) & 0x7f800000) - (24 << 23));
  if (!(fabsf(fabsf(t) - h) > h * 6.1035156e-5f) || !(s > 1e-30f)) return mapped_depth_double(d);
  return m;
}

// ---- launch-side argument blocks ----------------------------------------------------------------------
struct FwdParams {
  int P, D, M, W, H;
  const float* bg;
  const float* means3D;
  const float* shs;
  const float* sh_rest;
  const float* colors_precomp;
  const float* opacities;
  const float* scales;
  const float* rotations;
  const float* transMat_precomp;
  const float* view;
  const float* proj;
  const float* campos;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int prefiltered;
  uint32_t gx, gy;
  int raw;                 // raw-parameter mode: activations + deltas applied in-kernel
  const float* d_means3D;
  const float* d_scales;
  const float* d_rotations;
  uint4* tile_box;         // optional (P): {x0 | x1 << 16, y0 | y1 << 16, depth bits, 0} of the tile rectangle, all zero when culled
  uint32_t* frame_flag;    // optional: status[3], set to `frame_flag_value` by the per-surfel kernel (hit-mask handshake)
  uint32_t frame_flag_value;
};

// Activated surfel parameters from the raw ones, op for op what the eager glue computes
// (exp / sigmoid / F.normalize with eps 1e-12, IEEE division): see DESIGN.md "raw-parameter mode".
struct Activated { v3 pw; float2 sc; float4 q; float opacity; float qnorm; };
__device__ __forceinline__ Activated activate_surfel(int idx, const float* means3D, const float* d_means3D,
                                                     const float* scales, const float* d_scales, const float* rotations,
                                                     const float* d_rotations, const float* opacities) {
  Activated a;
  a.pw = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
  if (d_means3D) { a.pw.x += d_means3D[3 * idx]; a.pw.y += d_means3D[3 * idx + 1]; a.pw.z += d_means3D[3 * idx + 2]; }
  const float2 ls = reinterpret_cast<const float2*>(scales)[idx];
  a.sc = {expf(ls.x), expf(ls.y)};
  if (d_scales) { a.sc.x += d_scales[2 * idx]; a.sc.y += d_scales[2 * idx + 1]; }
  float4 r = reinterpret_cast<const float4*>(rotations)[idx];
  if (d_rotations) {
    const float4 dr = reinterpret_cast<const float4*>(d_rotations)[idx];
    r.x += dr.x; r.y += dr.y; r.z += dr.z; r.w += dr.w;
  }
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.x, r.x), __fmul_rn(r.y, r.y)), __fmul_rn(r.z, r.z)), __fmul_rn(r.w, r.w)));
  a.qnorm = fmaxf(n, 1e-12f);
  a.q = {r.x / a.qnorm, r.y / a.qnorm, r.z / a.qnorm, r.w / a.qnorm};
  a.opacity = opacities ? 1.0f / (1.0f + expf(-opacities[idx])) : 0.f;
  return a;
}

void launch_preprocess_fwd(const FwdParams& p, SurfelRec* rec, uint8_t* clamped, int* radii, uint32_t* tiles_touched,
                           cudaStream_t s);
// `capacity`: number of instance slots behind keys/vals; instances past it are dropped (deferred-count mode reports
// that as an overflow; with the exact count of the synchronous mode nothing is ever dropped)
void launch_duplicate(int P, const SurfelRec* rec, const int* radii, const uint32_t* offsets, uint64_t* keys,
                      uint32_t* vals, uint32_t gx, uint32_t gy, uint32_t capacity, cudaStream_t s);
void launch_ranges(int64_t R, const uint64_t* keys_sorted, uint2* ranges, cudaStream_t s);
// Tile-bucketed binning (tile_binning.cu): count -> scan (ranges, R and overflow flag in status, list of long tiles) ->
// scatter -> per-tile sort.  status: {R, overflow, number of long tiles}
void launch_tile_count(int P, const uint4* tile_box, uint32_t gx, uint32_t gy, uint32_t* tile_count, cudaStream_t s);
void launch_tile_scan(uint32_t tiles, uint32_t capacity, uint32_t* tile_count, uint32_t* seg_begin, uint2* ranges, uint32_t* big_list,
                      uint32_t* status, uint32_t* tile_order, cudaStream_t s);
// tile ids in order of decreasing list length (256 length classes): the blend kernels map their CTAs through it, so the
// hardware's in-order CTA dispatch starts the longest lists first (longest-processing-time-first scheduling) and the
// ~60 % of tiles that are empty cost nothing until the very end.  tile_scan writes it too; this is for the global-sort path.
void launch_tile_order(uint32_t tiles, const uint2* ranges, uint32_t* tile_order, cudaStream_t s);
void launch_tile_scatter(int P, const uint4* tile_box, uint32_t gx, uint32_t gy, const uint32_t* seg_begin,
                         uint32_t* cursor, const uint32_t* status, uint64_t* keys, cudaStream_t s);
void launch_tile_sort(uint32_t tiles, const uint2* ranges, const uint32_t* big_list, const uint32_t* status, uint64_t* keys_in,
                      uint64_t* keys_out, uint32_t* point_list, cudaStream_t s);
void launch_tile_export_keys(uint32_t tiles, const uint2* ranges, const uint64_t* keys, uint64_t* out_keys, uint32_t* out_vals,
                             cudaStream_t s);
// Deferred-count mode (no host readback of the instance count R = offsets[P-1]):
//   pad_keys : status = {R, R > capacity}; keys[R..capacity) = all ones, so a stable sort of all `capacity` slots leaves
//              the R real instances first, in exactly the order a sort of R items produces
//   ranges   : same as launch_ranges with L = min(status[0], capacity) read on the device
void launch_pad_keys(uint32_t capacity, const uint32_t* total, uint64_t* keys, uint32_t* status, cudaStream_t s);
void launch_ranges_deferred(uint32_t capacity, const uint32_t* status, const uint64_t* keys_sorted, uint2* ranges, cudaStream_t s);
// `status` (may be NULL): when status[1] != 0 the binning overflowed and the colour planes are poisoned with NaN
void launch_blend_fwd(const FwdParams& p, const uint2* ranges, const uint32_t* point_list, const SurfelRec* rec,
                      float* final_T, uint32_t* n_contrib, float* out_color, float* out_others, int cull,
                      const uint32_t* status, const uint32_t* tile_order, int lane_walk, uint32_t* hit_mask, cudaStream_t s);
// Hit masks handed from the lane-walk forward to the lane-walk backward: one 32-bit word per (list position, 8x4 patch) —
// the ballot of the exact prefilter over the patch's pixels.  Stored patch-major inside a tile's list segment:
//   hit_mask[8 * range.x + patch * (range.y - range.x) + position]        (only cull-box survivors are written)
// status[3] == HIT_MASK_MAGIC tells the backward that the forward of this frame wrote them.
constexpr uint32_t HIT_MASK_MAGIC = 0x4b53414du;

struct BwdParams {
  int P, D, M, W, H;
  const float* bg;
  const float* means3D;
  const float* shs;
  const float* sh_rest;
  const float* colors_precomp;
  const float* scales;
  const float* rotations;
  const float* transMat_precomp;
  const float* view;
  const float* proj;
  const float* campos;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  uint32_t gx, gy;
  int raw;
  const float* opacities;
  const float* d_means3D;
  const float* d_scales;
  const float* d_rotations;
};
void launch_blend_bwd(const BwdParams& p, const uint2* ranges, const uint32_t* point_list, const SurfelRec* rec,
                      const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dothers,
                      float* grad_rec, int cull, const uint32_t* tile_order, int lane_walk, const uint32_t* hit_mask,
                      const uint32_t* status, cudaStream_t s);
void launch_preprocess_bwd(const BwdParams& p, const SurfelRec* rec, const uint8_t* clamped, const int* radii,
                           float* grad_rec, float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                           float* dL_dmeans3D, float* dL_dtransMat, float* dL_dsh, float* dL_dsh_rest,
                           float* dL_dscales, float* dL_drot, float* dL_dscales_raw, cudaStream_t s);
void launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s);

}  // namespace d2gs
