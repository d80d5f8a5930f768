// Fused photometric loss of the training step, forward and backward:
//   loss = (1 - l_dssim) * L1(image, gt) + l_dssim * (1 - SSIM(image, gt))
//        + l_normal * mean(1 - sum_c rend_normal_c * surf_normal_c) + l_dist * mean(rend_dist)
// Behavioural contract: utils/loss_utils.py:18-19 (l1_loss), :33-76 (11x11 Gaussian window, sigma 1.5, zero padding,
// C1 = 0.01^2, C2 = 0.03^2, mean over all elements) and train_gui.py:292-313 (normal-consistency and distortion terms,
// composition).  The reference runs 5 grouped conv2d + ~25 elementwise / reduction kernels forward and as many again
// through autograd; here it is one tiled kernel each way.  The 2-D window is the outer product of a 1-D Gaussian, so both
// directions are separable passes through shared memory.
//
// SSIM as a function of the five window means (mu1, mu2, E11 = w*x^2, E22 = w*y^2, E12 = w*xy):
//   A1 = 2 mu1 mu2 + C1, A2 = 2 (E12 - mu1 mu2) + C2, B1 = mu1^2 + mu2^2 + C1, B2 = (E11 - mu1^2) + (E22 - mu2^2) + C2
//   ssim = A1 A2 / (B1 B2)
// The forward stores d ssim / d(mu1, E11, E12) per pixel; the backward convolves those three maps with the same window:
//   dL/dx(p) = g * [ (w * d_mu1)(p) + 2 x(p) (w * d_E11)(p) + y(p) (w * d_E12)(p) ],   g = -l_dssim / (3 H W)
#include "raster_common.cuh"
#include "loss.cuh"

namespace d2gs {

constexpr int LW = 11, LR = 5;            // window size / radius
constexpr int TX = 32, TY = 16;           // output tile
constexpr int SX = TX + 2 * LR, SY = TY + 2 * LR;
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

__constant__ float c_win[LW];

static void upload_window() {
  static bool done = false;
  if (done) return;
  // loss_utils.py:33-35: exp() in double, stored as float32, normalised by the float32 sum
  float g[LW];
  float sum = 0.f;
  for (int i = 0; i < LW; i++) { g[i] = (float)exp(-(double)((i - LR) * (i - LR)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
  for (int i = 0; i < LW; i++) g[i] = g[i] / sum;
  cudaMemcpyToSymbol(c_win, g, sizeof(g));   // once per process (one process per GPU)
  done = true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums[0] = sum |x - y|, sums[1] = sum ssim_map, sums[2] = sum_pixels sum_c rn*sn, sums[3] = sum rend_dist
__global__ void __launch_bounds__(256) loss_fwd_kernel(LossArgs a) {
  __shared__ float s_x[SY][SX + 1], s_y[SY][SX + 1];
  __shared__ float s_h[5][SY][TX];
  __shared__ float s_part[4][8];
  const int W = a.W, H = a.H;
  const size_t HW = (size_t)W * H;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float acc_l1 = 0.f, acc_ssim = 0.f, acc_nd = 0.f, acc_dist = 0.f;

  for (int c = 0; c < 3; c++) {
    const float* X = a.image + c * HW;
    const float* Y = a.gt + c * HW;
    for (int e = tid; e < SX * SY; e += 256) {
      const int r = e / SX, q = e - r * SX;
      const int gx = x0 + q - LR, gy = y0 + r - LR;
      const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
      s_x[r][q] = in ? __ldg(X + (size_t)gy * W + gx) : 0.f;      // zero padding of F.conv2d
      s_y[r][q] = in ? __ldg(Y + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < SY * TX; e += 256) {                    // horizontal pass
      const int r = e >> 5, q = e & 31;
      float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, h4 = 0.f;
#pragma unroll
      for (int t = 0; t < LW; t++) {
        const float w = c_win[t], x = s_x[r][q + t], y = s_y[r][q + t];
        h0 = fmaf(w, x, h0); h1 = fmaf(w, y, h1);
        h2 = fmaf(w, x * x, h2); h3 = fmaf(w, y * y, h3); h4 = fmaf(w, x * y, h4);
      }
      s_h[0][r][q] = h0; s_h[1][r][q] = h1; s_h[2][r][q] = h2; s_h[3][r][q] = h3; s_h[4][r][q] = h4;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; half++) {                        // vertical pass, two rows per thread
      const int ly = ty + 8 * half;
      const int gx = x0 + tx, gy = y0 + ly;
      if (gx < W && gy < H) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int t = 0; t < LW; t++) {
          const float w = c_win[t];
          mu1 = fmaf(w, s_h[0][ly + t][tx], mu1); mu2 = fmaf(w, s_h[1][ly + t][tx], mu2);
          e11 = fmaf(w, s_h[2][ly + t][tx], e11); e22 = fmaf(w, s_h[3][ly + t][tx], e22);
          e12 = fmaf(w, s_h[4][ly + t][tx], e12);
        }
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
        const float A1 = 2.f * mu12 + SSIM_C1, A2 = 2.f * s12 + SSIM_C2;
        const float B1 = mu1_sq + mu2_sq + SSIM_C1, B2 = s1 + s2 + SSIM_C2;
        const float inv = 1.0f / (B1 * B2);
        const float ssim = A1 * A2 * inv;
        acc_ssim += ssim;
        const float x = s_x[ly + LR][tx + LR], y = s_y[ly + LR][tx + LR];
        acc_l1 += fabsf(x - y);
        if (a.d_mu1) {
          const size_t pix = c * HW + (size_t)gy * W + gx;
          // partials with (mu2, E22) fixed; sigma terms depend on mu1 through -mu1^2 and -mu1 mu2
          a.d_mu1[pix] = 2.f * mu2 * (A2 - A1) * inv - 2.f * mu1 * ssim * (B2 - B1) * (1.0f / B1) * (1.0f / B2);
          a.d_e11[pix] = -ssim / B2;
          a.d_e12[pix] = 2.f * A1 * inv;
        }
      }
    }
    __syncthreads();
  }
  // normal-consistency and distortion terms (train_gui.py:292-300)
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int gx = x0 + tx, gy = y0 + ty + 8 * half;
    if (gx < W && gy < H) {
      const size_t pix = (size_t)gy * W + gx;
      if (a.rend_normal && a.surf_normal)
        acc_nd += a.rend_normal[pix] * a.surf_normal[pix] + a.rend_normal[HW + pix] * a.surf_normal[HW + pix] +
                  a.rend_normal[2 * HW + pix] * a.surf_normal[2 * HW + pix];
      if (a.rend_dist) acc_dist += a.rend_dist[pix];
    }
  }
  acc_l1 = warp_sum(acc_l1); acc_ssim = warp_sum(acc_ssim); acc_nd = warp_sum(acc_nd); acc_dist = warp_sum(acc_dist);
  if (tx == 0) { s_part[0][ty] = acc_l1; s_part[1][ty] = acc_ssim; s_part[2][ty] = acc_nd; s_part[3][ty] = acc_dist; }
  __syncthreads();
  if (tid < 4) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) v += s_part[tid][w];
    atomicAdd(a.sums + tid, v);
  }
}

// out[0] = loss, out[1] = L1, out[2] = SSIM, out[3] = normal loss (weighted), out[4] = distortion loss (weighted)
__global__ void loss_finalize_kernel(LossArgs a) {
  const float n3 = 3.0f * (float)a.W * (float)a.H, n1 = (float)a.W * (float)a.H;
  const float l1 = a.sums[0] / n3, ssim = a.sums[1] / n3;
  const float nl = (a.rend_normal && a.surf_normal) ? a.l_normal * (1.0f - a.sums[2] / n1) : 0.f;
  const float dl = a.rend_dist ? a.l_dist * (a.sums[3] / n1) : 0.f;
  a.out[0] = (1.0f - a.l_dssim) * l1 + a.l_dssim * (1.0f - ssim) + nl + dl;
  a.out[1] = l1; a.out[2] = ssim; a.out[3] = nl; a.out[4] = dl;
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(LossArgs a) {
  __shared__ float s_d[3][SY][SX + 1];
  __shared__ float s_h[3][SY][TX];
  const int W = a.W, H = a.H;
  const size_t HW = (size_t)W * H;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const float up = a.upstream ? __ldg(a.upstream) : 1.0f;
  const float g_ssim = -a.l_dssim / (3.0f * (float)W * (float)H) * up;
  const float g_l1 = (1.0f - a.l_dssim) / (3.0f * (float)W * (float)H) * up;
  const float g_n = -a.l_normal / ((float)W * (float)H) * up;
  const float g_d = a.l_dist / ((float)W * (float)H) * up;

  for (int c = 0; c < 3; c++) {
    const float* D0 = a.d_mu1 + c * HW;
    const float* D1 = a.d_e11 + c * HW;
    const float* D2 = a.d_e12 + c * HW;
    for (int e = tid; e < SX * SY; e += 256) {
      const int r = e / SX, q = e - r * SX;
      const int gx = x0 + q - LR, gy = y0 + r - LR;
      const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
      const size_t p = (size_t)gy * W + gx;
      s_d[0][r][q] = in ? __ldg(D0 + p) : 0.f;
      s_d[1][r][q] = in ? __ldg(D1 + p) : 0.f;
      s_d[2][r][q] = in ? __ldg(D2 + p) : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < SY * TX; e += 256) {
      const int r = e >> 5, q = e & 31;
      float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
      for (int t = 0; t < LW; t++) {
        const float w = c_win[t];
        h0 = fmaf(w, s_d[0][r][q + t], h0); h1 = fmaf(w, s_d[1][r][q + t], h1); h2 = fmaf(w, s_d[2][r][q + t], h2);
      }
      s_h[0][r][q] = h0; s_h[1][r][q] = h1; s_h[2][r][q] = h2;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int ly = ty + 8 * half;
      const int gx = x0 + tx, gy = y0 + ly;
      if (gx < W && gy < H) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int t = 0; t < LW; t++) {
          const float w = c_win[t];
          c0 = fmaf(w, s_h[0][ly + t][tx], c0); c1 = fmaf(w, s_h[1][ly + t][tx], c1); c2 = fmaf(w, s_h[2][ly + t][tx], c2);
        }
        const size_t pix = c * HW + (size_t)gy * W + gx;
        const float x = __ldg(a.image + pix), y = __ldg(a.gt + pix);
        const float df = x - y;
        const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);          // torch.abs backward: sign(0) = 0
        a.g_image[pix] = g_ssim * (c0 + 2.f * x * c1 + y * c2) + g_l1 * sgn;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int gx = x0 + tx, gy = y0 + ty + 8 * half;
    if (gx < W && gy < H) {
      const size_t pix = (size_t)gy * W + gx;
      if (a.g_rend_normal && a.surf_normal) {
#pragma unroll
        for (int c = 0; c < 3; c++) a.g_rend_normal[c * HW + pix] = g_n * a.surf_normal[c * HW + pix];
      }
      if (a.g_surf_normal && a.rend_normal) {
#pragma unroll
        for (int c = 0; c < 3; c++) a.g_surf_normal[c * HW + pix] = g_n * a.rend_normal[c * HW + pix];
      }
      if (a.g_rend_dist) a.g_rend_dist[pix] = g_d;
    }
  }
}

void launch_loss_forward(const LossArgs& a, cudaStream_t s) {
  upload_window();
  cudaMemsetAsync(a.sums, 0, 4 * sizeof(float), s);
  dim3 grid((a.W + TX - 1) / TX, (a.H + TY - 1) / TY);
  loss_fwd_kernel<<<grid, 256, 0, s>>>(a);
  loss_finalize_kernel<<<1, 1, 0, s>>>(a);
}
void launch_loss_backward(const LossArgs& a, cudaStream_t s) {
  upload_window();
  dim3 grid((a.W + TX - 1) / TX, (a.H + TY - 1) / TY);
  loss_bwd_kernel<<<grid, 256, 0, s>>>(a);
}

}  // namespace d2gs
