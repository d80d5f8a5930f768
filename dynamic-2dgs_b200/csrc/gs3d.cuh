// Internal interface of gs3d.cu (3-D Gaussian rasterizer with depth and alpha outputs).
#pragma once
#include "raster_common.cuh"

namespace d2gs {

// Projected Gaussian, 48 B:  q0 = (mean2D.x, mean2D.y, view depth, prefilter threshold on the exponent)
//                            q1 = (conic.x, conic.y, conic.z, opacity)      q2 = (r, g, b, 0)
struct __align__(16) G3Rec { float4 q0, q1, q2; };
static_assert(sizeof(G3Rec) == 48, "G3Rec must be 48 bytes");

struct G3Params {
  int P, D, M, W, H;
  const float* bg;
  const float* means3D;
  const float* shs;
  const float* colors_precomp;
  const float* opacities;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  const float* view;
  const float* proj;
  const float* campos;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int prefiltered;
  uint32_t gx, gy;
};

constexpr int G3_GRAD_FLOATS = 12;   // per-Gaussian gradient record of the blend backward

// tile_box (optional): packed tile rectangle + depth bits per Gaussian for the per-tile binning of tile_binning.cu
void g3_launch_preprocess_fwd(const G3Params& p, G3Rec* rec, float* cov3Ds, uint8_t* clamped, int* radii, uint32_t* tiles_touched,
                              uint4* tile_box, cudaStream_t s);
void g3_launch_duplicate(int P, const G3Rec* rec, const int* radii, const uint32_t* offsets, uint64_t* keys, uint32_t* vals,
                         uint32_t gx, uint32_t gy, cudaStream_t s);
// status (optional): {R, overflow, ...} of the deferred-count binning; overflow poisons the colour planes with NaN
void g3_launch_blend_fwd(const G3Params& p, const uint2* ranges, const uint32_t* point_list, const G3Rec* rec, float* out_color,
                         float* out_depth, float* out_alpha, uint32_t* n_contrib, const uint32_t* status, cudaStream_t s);
void g3_launch_blend_bwd(const G3Params& p, const uint2* ranges, const uint32_t* point_list, const G3Rec* rec, const float* alphas,
                         const uint32_t* n_contrib, const float* dL_dpix, const float* dL_ddepth, const float* dL_dalpha,
                         float* grad, cudaStream_t s);
void g3_launch_preprocess_bwd(const G3Params& p, const float* cov3Ds, const uint8_t* clamped, const int* radii, const float* grad,
                              float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D, float* dL_dcov3D,
                              float* dL_dsh, float* dL_dscales, float* dL_drot, cudaStream_t s);

}  // namespace d2gs
