// Descriptors of the fused optimiser step (optim.cu).
#pragma once
#include <cuda_runtime.h>

namespace d2gs {

constexpr int ADAM_MAX_TENSORS = 48;   // per launch: 48 x 72 B + 48 x 4 B stays inside the 4 KB kernel parameter space

struct AdamTensor {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  long long numel;
  float w1, beta2, w2, eps, neg_step_size, inv_bc2_sqrt;   // w1 = 1 - beta1, w2 = 1 - beta2 (rounded from double)
  int vec4;     // all four pointers 16-byte aligned
};
struct AdamBatch {
  AdamTensor t[ADAM_MAX_TENSORS];
  int chunk_end[ADAM_MAX_TENSORS];   // inclusive prefix sums of the per-tensor block counts
  int count;
};

int adam_chunks(long long numel);
void launch_adam(const AdamBatch& B, cudaStream_t s);
void launch_densify_stats(int P, const float* vs_grad, int stride, const unsigned char* filter, float* accum, float* denom,
                          cudaStream_t s);

}  // namespace d2gs
