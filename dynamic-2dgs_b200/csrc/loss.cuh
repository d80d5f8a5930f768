// Descriptor of the fused photometric loss (loss.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace d2gs {

struct LossArgs {
  int W, H;
  const float* image; const float* gt;                 // (3,H,W)
  const float* rend_normal; const float* surf_normal;  // (3,H,W) or NULL
  const float* rend_dist;                              // (1,H,W) or NULL
  float l_dssim, l_normal, l_dist;
  float* sums;                                         // (4) scratch, zeroed by the forward launch
  float* out;                                          // (5) loss, L1, SSIM, normal term, distortion term
  float* d_mu1; float* d_e11; float* d_e12;            // (3,H,W) each: d ssim / d(mu1, E11, E12); NULL = not saved
  const float* upstream;                               // device scalar or NULL (= 1)
  float* g_image; float* g_rend_normal; float* g_surf_normal; float* g_rend_dist;
};

void launch_loss_forward(const LossArgs& a, cudaStream_t s);
void launch_loss_backward(const LossArgs& a, cudaStream_t s);

}  // namespace d2gs
