// Host-side launch descriptors of the deformation kernels (deform.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace d2gs {

struct DeformFwdHost {
  int P, M, K, hyper;   // hyper = hyper_dim of the node table (stride 3+hyper); the query uses it only when feature != NULL
  const float* xyz; const float* feature; int fstride;
  const float* nodes; const float* radius_log; const float* weight_logit;
  const float* trans; const float* rot; const float* scale; const float* local_rot; const float* mask;
  int64_t* nn_idx; float* nn_dist; float* nn_weight;
  float* d_xyz; float* d_rot; float* d_scale;
  int attr_stride;      // > 0: trans/rot/scale/local_rot are columns of one (M, attr_stride) matrix
  const int* order;     // optional processing order (a permutation of 0..P-1): thread t handles surfel order[t]
};

struct DeformBwdHost {
  int P, M, K, hyper;
  const float* xyz; const float* feature; int fstride;
  const float* nodes; const float* radius_log; const float* weight_logit;
  const float* trans; const float* rot; const float* scale; const float* local_rot; const float* mask;
  const int64_t* nn_idx; const float* nn_dist; const float* nn_weight;
  const float* g_xyz; const float* g_rot; const float* g_scale;
  float* d_trans; float* d_rot; float* d_scale; float* d_local_rot; float* d_nodes; float* d_radius_log;
  float* d_weight_logit; float* d_feature; float* d_mask;
  int attr_stride;
  const int* order;
};

int deform_forward_launch(const DeformFwdHost& h, cudaStream_t s, const char** err);
int deform_backward_launch(const DeformBwdHost& h, cudaStream_t s, const char** err);

// Morton order of the surfel centres (see d2gs_deform_order)
void deform_order_keys_launch(int P, const float* xyz, unsigned int* bbox, unsigned int* keys, int* vals, cudaStream_t s);

}  // namespace d2gs
