// Optimiser step of the training loop, fused.
//
// The reference builds torch.optim.Adam(l, lr=0.0, eps=1e-15) over seven surfel parameter groups and a second one over the
// deformation network (scene/gaussian_model.py:181-203, scene/deform_model.py train_setting, train_gui.py:426-432): per
// step that is a foreach pipeline of ~10 kernels per dtype/device bucket with five passes over every tensor.  Here ONE
// launch updates every tensor of an optimiser: 4 reads + 3 writes per element, the arithmetic of torch's
// _single_tensor_adam (lerp for the first moment, mul/addcmul for the second, sqrt * 1/sqrt(bc2) + eps, addcdiv).
// Tensor descriptors travel in the kernel parameter space (no per-step H2D copy); a block finds its (tensor, chunk) by a
// binary search over the prefix sums of the chunk counts.
#include "raster_common.cuh"
#include "optim.cuh"

namespace d2gs {

constexpr int ADAM_CHUNK = 256 * 4 * 8;   // elements per block: 256 threads x 8 float4

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamBatch B) {
  int lo = 0, hi = B.count;               // first t with chunk_end[t] > blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (B.chunk_end[mid] > (int)blockIdx.x) hi = mid; else lo = mid + 1;
  }
  const AdamTensor& T = B.t[lo];
  const int chunk = (int)blockIdx.x - (lo ? B.chunk_end[lo - 1] : 0);
  const long long base = (long long)chunk * ADAM_CHUNK;
  const float w1 = T.w1, w2 = T.w2;
  auto update = [&](float& p, float g, float& m, float& v) {
    m = m + w1 * (g - m);                                  // exp_avg.lerp_(grad, 1 - beta1)
    v = v * T.beta2 + w2 * g * g;                          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) * T.inv_bc2_sqrt + T.eps; // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p + T.neg_step_size * (m / denom);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  const bool vec = T.vec4;
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const long long e = base + ((long long)it * 256 + threadIdx.x) * 4;
    if (e >= T.numel) break;
    if (vec && e + 3 < T.numel) {
      float4 p = *reinterpret_cast<float4*>(T.param + e);
      const float4 g = __ldg(reinterpret_cast<const float4*>(T.grad + e));
      float4 m = *reinterpret_cast<float4*>(T.exp_avg + e);
      float4 v = *reinterpret_cast<float4*>(T.exp_avg_sq + e);
      update(p.x, g.x, m.x, v.x); update(p.y, g.y, m.y, v.y); update(p.z, g.z, m.z, v.z); update(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(T.param + e) = p;
      *reinterpret_cast<float4*>(T.exp_avg + e) = m;
      *reinterpret_cast<float4*>(T.exp_avg_sq + e) = v;
    } else {
      for (int k = 0; k < 4 && e + k < T.numel; k++) {
        float p = T.param[e + k], m = T.exp_avg[e + k], v = T.exp_avg_sq[e + k];
        update(p, T.grad[e + k], m, v);
        T.param[e + k] = p; T.exp_avg[e + k] = m; T.exp_avg_sq[e + k] = v;
      }
    }
  }
}

int adam_chunks(long long numel) { return (int)((numel + ADAM_CHUNK - 1) / ADAM_CHUNK); }

void launch_adam(const AdamBatch& B, cudaStream_t s) {
  if (B.count <= 0) return;
  const int blocks = B.chunk_end[B.count - 1];
  if (blocks > 0) adam_step_kernel<<<blocks, 256, 0, s>>>(B);
}

// add_densification_stats (scene/gaussian_model.py:484-486) without the boolean-mask indexing (which synchronises):
//   accum[i] += |viewspace_grad[i, :2]|  and  denom[i] += 1  for the surfels of the update filter
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ vs_grad, int stride,
                                                            const unsigned char* __restrict__ filter,
                                                            float* __restrict__ accum, float* __restrict__ denom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || !filter[i]) return;
  const float gx = vs_grad[(size_t)i * stride], gy = vs_grad[(size_t)i * stride + 1];
  accum[i] += sqrtf(gx * gx + gy * gy);
  denom[i] += 1.0f;
}
void launch_densify_stats(int P, const float* vs_grad, int stride, const unsigned char* filter, float* accum, float* denom,
                          cudaStream_t s) {
  if (P > 0) densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, vs_grad, stride, filter, accum, denom);
}

}  // namespace d2gs
