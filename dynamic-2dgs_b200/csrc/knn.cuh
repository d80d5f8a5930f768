// Internal launcher of knn.cu (mean squared distance to the 3 nearest neighbours).
#pragma once
#include <cuda_runtime.h>

namespace d2gs {
// order: Morton permutation of the points (deform_order_keys_launch + CUB sort); sp: 32*ceil(P/32) float4;
// leaf_box: 2*ceil(P/32) float4; group_box: 2*ceil(P/1024) float4
void knn_mean_dist2_launch(int P, const float* xyz, const int* order, float4* sp, float4* leaf_box, float4* group_box,
                           float* out, cudaStream_t s);
}  // namespace d2gs
