#pragma once
#include <cuda_runtime.h>
namespace d2gs {
void launch_epilogue_fwd(int W, int H, const float* allmap, const float* view, float fx, float fy, float* alpha,
                         float* rend_normal, float* rend_dist, float* depth, float* surf_normal, float* surf_point,
                         cudaStream_t s);
void launch_epilogue_bwd(int W, int H, const float* allmap, const float* view, float fx, float fy, const float* g_alpha,
                         const float* g_rn, const float* g_dist, const float* g_depth, const float* g_sn, const float* g_sp,
                         float* dA, cudaStream_t s);
}
