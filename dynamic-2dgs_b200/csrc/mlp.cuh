// Descriptors of the fused deformation MLP (mlp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace d2gs {

struct MlpLayer {
  const float* W;   // (N, K) row-major, the nn.Linear layout
  const float* b;   // (N)
  int N, K, NP;     // NP = N rounded up to 4 (row length of the transposed copy)
  size_t wt_off;    // offset (floats) of W^T (K x NP) inside the transposed-weight scratch
};
struct MlpLayers {
  MlpLayer layer[12];   // [timenet.0, timenet.2,] linear.0..7, heads
  int count;
  float* wt;
};
struct MlpFwd {
  MlpLayers layers;
  int rows, Et, Tt, NH, has_timenet, t_stride;
  const float* x;       // (rows, 3)
  const float* t;       // (rows) with stride t_stride (0 = one time for every row)
  float* out;           // (rows, NH)
  float* save_inp;      // (rows, 96)  [x_emb | time feature]
  float* save_te;       // (rows, 32)  time embedding
  float* save_th;       // (rows, 256) timenet hidden
  float* save_h;        // (8, rows, 256) post-ReLU trunk activations
};
struct MlpBwd {
  MlpLayers layers;
  int rows, Et, Tt, NH, has_timenet;
  const float* g_out;   // (rows, NH)
  const float* save_th; // (rows, 256)
  const float* save_h;  // (8, rows, 256)
  float* G;             // (8, rows, 256) pre-activation gradients of the trunk
  float* G_t1;          // (rows, 256)
  float* g_tfeat;       // (rows, 32)
};
struct MlpWJob {
  const float* G; int ldG; int N;
  const float* A; int ldA; int Ka;
  const float* B; int ldB; int Kb;
  float* dW; float* db;
  int tiles;
};
struct MlpWJobs {
  MlpWJob job[12];
  int count;
};

void mlp_launch_transpose(const MlpLayers& L, cudaStream_t s);
void mlp_set_cluster(int which /*0 forward, 1 backward*/, int size /*1, 2, 4 or 8 CTAs*/);
void mlp_launch_forward(const MlpFwd& a, cudaStream_t s);
void mlp_launch_backward(const MlpBwd& a, const MlpWJobs& J, cudaStream_t s);

}  // namespace d2gs
