// Image-space epilogue of render(), fused: one kernel forward, one backward, over the 8 auxiliary planes.
// Behavioural contract: gaussian_renderer/__init__.py:172-207 (alpha, view->world normals, median depth with
// nan_to_num, distortion, pseudo surface normal scaled by the detached alpha) and utils/point_utils.py:9-38
// (unprojection with integer pixel coordinates, central differences, cross product, F.normalize, zero border).
// The reference runs ~40 eager kernels forward (+~60 backward) and rebuilds a CPU meshgrid and two matrix inverses per
// call; here the camera-to-world rotation is the 3x3 adjugate inverse computed in-kernel from the view matrix.
#include "raster_common.cuh"
#include "epilogue.cuh"

namespace d2gs {

struct Cam {
  float A[9];    // world rotation rows applied to view-space normals: out_c = sum_k n_k * A[3*c+k]
  float C[9];    // camera-to-world rotation (row-major)
  float o[3];    // camera centre in world space
  float fx, fy, cx, cy;
};

__device__ __forceinline__ Cam make_cam(const float* m, float fx, float fy, int W, int H) {
  Cam c;
  // view matrix V (column-vector convention) = transpose of the stored tensor: V[r][c] = m[4c + r]
  float R[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) R[3 * r + k] = m[4 * k + r];
  const float t0 = m[12], t1 = m[13], t2 = m[14];
  // inverse of R by the adjugate
  const float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
  const float det = R[0] * c00 + R[1] * c01 + R[2] * c02;
  const float id = 1.0f / det;
  c.C[0] = c00 * id; c.C[1] = (R[2] * R[7] - R[1] * R[8]) * id; c.C[2] = (R[1] * R[5] - R[2] * R[4]) * id;
  c.C[3] = c01 * id; c.C[4] = (R[0] * R[8] - R[2] * R[6]) * id; c.C[5] = (R[2] * R[3] - R[0] * R[5]) * id;
  c.C[6] = c02 * id; c.C[7] = (R[1] * R[6] - R[0] * R[7]) * id; c.C[8] = (R[0] * R[4] - R[1] * R[3]) * id;
  c.o[0] = -(c.C[0] * t0 + c.C[1] * t1 + c.C[2] * t2);
  c.o[1] = -(c.C[3] * t0 + c.C[4] * t1 + c.C[5] * t2);
  c.o[2] = -(c.C[6] * t0 + c.C[7] * t1 + c.C[8] * t2);
  // rend_normal = n_view @ (m[:3,:3]).T  ->  out_c = sum_k n_k * m[4c + k]
#pragma unroll
  for (int cc = 0; cc < 3; cc++)
#pragma unroll
    for (int k = 0; k < 3; k++) c.A[3 * cc + k] = m[4 * cc + k];
  c.fx = fx; c.fy = fy; c.cx = W / 2.0f; c.cy = H / 2.0f;
  return c;
}

__device__ __forceinline__ float clean_depth(float d) {   // torch.nan_to_num(d, 0, 0)
  if (isnan(d)) return 0.f;
  if (isinf(d)) return d > 0 ? 0.f : -3.4028234663852886e38f;
  return d;
}
__device__ __forceinline__ v3 ray_dir(const Cam& c, int x, int y) {
  const float vx = ((float)x - c.cx) / c.fx, vy = ((float)y - c.cy) / c.fy;
  return {c.C[0] * vx + c.C[1] * vy + c.C[2], c.C[3] * vx + c.C[4] * vy + c.C[5], c.C[6] * vx + c.C[7] * vy + c.C[8]};
}
__device__ __forceinline__ v3 point_at(const Cam& c, const float* depth_plane, int W, int x, int y) {
  const float d = clean_depth(__ldg(depth_plane + (size_t)y * W + x));
  const v3 r = ray_dir(c, x, y);
  return {d * r.x + c.o[0], d * r.y + c.o[1], d * r.z + c.o[2]};
}

__global__ void __launch_bounds__(256) epilogue_fwd_kernel(int W, int H, const float* __restrict__ allmap,
                                                           const float* __restrict__ view, float fx, float fy,
                                                           float* __restrict__ alpha, float* __restrict__ rend_normal,
                                                           float* __restrict__ rend_dist, float* __restrict__ depth,
                                                           float* __restrict__ surf_normal, float* __restrict__ surf_point) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const Cam c = make_cam(view, fx, fy, W, H);
  const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
  const float a = allmap[1 * HW + pix];
  const float n0 = allmap[2 * HW + pix], n1 = allmap[3 * HW + pix], n2 = allmap[4 * HW + pix];
  const float* dplane = allmap + 5 * HW;
  alpha[pix] = a;
  rend_dist[pix] = allmap[6 * HW + pix];
#pragma unroll
  for (int cc = 0; cc < 3; cc++) rend_normal[cc * HW + pix] = n0 * c.A[3 * cc] + n1 * c.A[3 * cc + 1] + n2 * c.A[3 * cc + 2];
  const float d = clean_depth(dplane[pix]);
  depth[pix] = d;
  const v3 r = ray_dir(c, x, y);
  surf_point[0 * HW + pix] = d * r.x + c.o[0];
  surf_point[1 * HW + pix] = d * r.y + c.o[1];
  surf_point[2 * HW + pix] = d * r.z + c.o[2];
  v3 nrm = {0.f, 0.f, 0.f};
  if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) {
    const v3 dx = point_at(c, dplane, W, x, y + 1) - point_at(c, dplane, W, x, y - 1);
    const v3 dy = point_at(c, dplane, W, x + 1, y) - point_at(c, dplane, W, x - 1, y);
    const v3 cr = cross3(dx, dy);
    const float n = fmaxf(sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z), 1e-12f);
    nrm = {cr.x / n * a, cr.y / n * a, cr.z / n * a};
  }
  surf_normal[0 * HW + pix] = nrm.x; surf_normal[1 * HW + pix] = nrm.y; surf_normal[2 * HW + pix] = nrm.z;
}

// gradient of the normal at interior pixel (x,y) w.r.t. its two finite differences
__device__ __forceinline__ void normal_vjp(const Cam& c, const float* dplane, const float* alpha_plane,
                                           const float* g_sn, size_t HW, int W, int H, int x, int y, v3& g_dx, v3& g_dy) {
  g_dx = {0.f, 0.f, 0.f}; g_dy = {0.f, 0.f, 0.f};
  if (x < 1 || x > W - 2 || y < 1 || y > H - 2) return;
  const size_t pix = (size_t)y * W + x;
  const float a = __ldg(alpha_plane + pix);
  const v3 go = {__ldg(g_sn + pix) * a, __ldg(g_sn + HW + pix) * a, __ldg(g_sn + 2 * HW + pix) * a};
  const v3 dx = point_at(c, dplane, W, x, y + 1) - point_at(c, dplane, W, x, y - 1);
  const v3 dy = point_at(c, dplane, W, x + 1, y) - point_at(c, dplane, W, x - 1, y);
  const v3 cr = cross3(dx, dy);
  const float n = sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
  v3 gc;
  if (n > 1e-12f) {
    const v3 u = {cr.x / n, cr.y / n, cr.z / n};
    const float ug = u.x * go.x + u.y * go.y + u.z * go.z;
    gc = {(go.x - u.x * ug) / n, (go.y - u.y * ug) / n, (go.z - u.z * ug) / n};
  } else {
    gc = {go.x / 1e-12f, go.y / 1e-12f, go.z / 1e-12f};
  }
  g_dx = cross3(dy, gc);
  g_dy = cross3(gc, dx);
}

__global__ void __launch_bounds__(256) epilogue_bwd_kernel(int W, int H, const float* __restrict__ allmap,
                                                           const float* __restrict__ view, float fx, float fy,
                                                           const float* __restrict__ g_alpha, const float* __restrict__ g_rn,
                                                           const float* __restrict__ g_dist, const float* __restrict__ g_depth,
                                                           const float* __restrict__ g_sn, const float* __restrict__ g_sp,
                                                           float* __restrict__ dA) {
  // The VJP of the normal at a pixel feeds its four neighbours: each CTA evaluates it ONCE per pixel of its 32x8 tile
  // plus a one-pixel halo (340 evaluations for 256 threads) into shared memory, instead of four times per thread.
  __shared__ float s_gdx[3][10][34], s_gdy[3][10][34];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + tx, y = y0 + ty;
  const Cam c = make_cam(view, fx, fy, W, H);
  const size_t HW = (size_t)H * W;
  const float* dplane = allmap + 5 * HW;
  const float* aplane = allmap + 1 * HW;
  if (g_sn) {
    for (int e = threadIdx.x; e < 340; e += 256) {
      const int hy = e / 34, hx = e - hy * 34;
      v3 a, b;
      normal_vjp(c, dplane, aplane, g_sn, HW, W, H, x0 + hx - 1, y0 + hy - 1, a, b);   // zero outside the interior
      s_gdx[0][hy][hx] = a.x; s_gdx[1][hy][hx] = a.y; s_gdx[2][hy][hx] = a.z;
      s_gdy[0][hy][hx] = b.x; s_gdy[1][hy][hx] = b.y; s_gdy[2][hy][hx] = b.z;
    }
    __syncthreads();
  }
  if (x >= W || y >= H) return;
  const size_t pix = (size_t)y * W + x;
  dA[0 * HW + pix] = 0.f;
  dA[1 * HW + pix] = g_alpha ? g_alpha[pix] : 0.f;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  if (g_rn) { r0 = g_rn[pix]; r1 = g_rn[HW + pix]; r2 = g_rn[2 * HW + pix]; }
#pragma unroll
  for (int k = 0; k < 3; k++) dA[(2 + k) * HW + pix] = r0 * c.A[k] + r1 * c.A[3 + k] + r2 * c.A[6 + k];
  dA[6 * HW + pix] = g_dist ? g_dist[pix] : 0.f;
  dA[7 * HW + pix] = 0.f;
  v3 gP = {0.f, 0.f, 0.f};
  if (g_sp) gP = {g_sp[pix], g_sp[HW + pix], g_sp[2 * HW + pix]};
  if (g_sn) {
    // P(y,x) is the "+" end of dx at (y-1) and its "-" end at (y+1); likewise for dy along x.  Same order of additions
    // as the per-thread version: +a(y-1), -a(y+1), +b(x-1), -b(x+1).
    const int hx = tx + 1, hy = ty + 1;
    gP.x = gP.x + s_gdx[0][hy - 1][hx]; gP.y = gP.y + s_gdx[1][hy - 1][hx]; gP.z = gP.z + s_gdx[2][hy - 1][hx];
    gP.x = gP.x - s_gdx[0][hy + 1][hx]; gP.y = gP.y - s_gdx[1][hy + 1][hx]; gP.z = gP.z - s_gdx[2][hy + 1][hx];
    gP.x = gP.x + s_gdy[0][hy][hx - 1]; gP.y = gP.y + s_gdy[1][hy][hx - 1]; gP.z = gP.z + s_gdy[2][hy][hx - 1];
    gP.x = gP.x - s_gdy[0][hy][hx + 1]; gP.y = gP.y - s_gdy[1][hy][hx + 1]; gP.z = gP.z - s_gdy[2][hy][hx + 1];
  }
  const float draw = dplane[pix];
  float gd = g_depth ? g_depth[pix] : 0.f;
  const v3 r = ray_dir(c, x, y);
  gd += gP.x * r.x + gP.y * r.y + gP.z * r.z;
  dA[5 * HW + pix] = (isnan(draw) || isinf(draw)) ? 0.f : gd;
}

void launch_epilogue_fwd(int W, int H, const float* allmap, const float* view, float fx, float fy, float* alpha,
                         float* rend_normal, float* rend_dist, float* depth, float* surf_normal, float* surf_point,
                         cudaStream_t s) {
  dim3 grid((W + 31) / 32, (H + 7) / 8);
  epilogue_fwd_kernel<<<grid, 256, 0, s>>>(W, H, allmap, view, fx, fy, alpha, rend_normal, rend_dist, depth, surf_normal,
                                          surf_point);
}
void launch_epilogue_bwd(int W, int H, const float* allmap, const float* view, float fx, float fy, const float* g_alpha,
                         const float* g_rn, const float* g_dist, const float* g_depth, const float* g_sn, const float* g_sp,
                         float* dA, cudaStream_t s) {
  dim3 grid((W + 31) / 32, (H + 7) / 8);
  epilogue_bwd_kernel<<<grid, 256, 0, s>>>(W, H, allmap, view, fx, fy, g_alpha, g_rn, g_dist, g_depth, g_sn, g_sp, dA);
}

}  // namespace d2gs
