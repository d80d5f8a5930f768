/*
 * d2gs.h — C ABI of the B200-native Dynamic-2DGS render hot path (libd2gs.so).
 *
 * Plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*, int status codes.  No torch
 * types.  Every entry point names the reference interface it replaces (paths relative to
 * /root/reference; DSR = submodules/diff-surfel-rasterization).
 *
 * Conventions
 *   - all tensors are contiguous float32 unless stated; "int" = int32
 *   - a NULL data pointer means "not provided" (DSR/cuda_rasterizer/forward.cu:215,247)
 *   - every call is asynchronous on `stream` except where noted
 *   - return 0 on success; on failure a negative D2GS_ERR_* code, text via d2gs_last_error()
 */
#ifndef D2GS_H_
#define D2GS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define D2GS_API __attribute__((visibility("default")))
#else
#define D2GS_API
#endif

#define D2GS_OK 0
#define D2GS_ERR_INVALID_ARG (-1)
#define D2GS_ERR_CUDA (-2)
#define D2GS_ERR_WORKSPACE (-3)      /* a workspace buffer is too small; *_required fields are filled */
#define D2GS_NEED_BINNING 1          /* forward stopped after the geometry stage: grow the binning buffer, call again with resume=1 */

/* Compile-time configuration baked into the kernels; mirrors DSR/cuda_rasterizer/config.h:15-17 and
 * auxiliary.h:20-37.  Exposed so parity tests can assert the two builds agree. */
typedef struct D2gsConfig {
  int num_channels;      /* 3 */
  int block_x, block_y;  /* 16, 16 */
  int tight_bbox;        /* 0 */
  int render_auxiliary;  /* 1 */
  int backface_cull;     /* 1 */
  int dual_visible;      /* 1 */
  int detach_weight;     /* 0 */
  double near_plane;     /* 0.2 */
  double far_plane;      /* 100.0 */
  double filter_size;    /* 0.7071067811865476 */
  int sm_arch;           /* 100 */
} D2gsConfig;

/* Per-stage device timing with CUDA events recorded on the launch stream (for bench.py's roofline line).
 * Stages: 0 preprocess_fwd, 1 scan, 2 duplicate, 3 sort, 4 ranges, 5 blend_fwd, 6 blend_bwd, 7 preprocess_bwd,
 *         8 deform_fwd, 9 deform_bwd, 10 epilogue_fwd, 11 epilogue_bwd, 12 mlp_fwd, 13 mlp_bwd, 14 loss_fwd, 15 loss_bwd.  d2gs_profile_collect synchronises the device, adds the elapsed times of all
 * recorded launches to total_ms[stage] / launches[stage] (arrays of D2GS_NUM_STAGES) and clears the record. */
#define D2GS_NUM_STAGES 16
/* Runtime switches (speed only; results are identical either way, the tests flip them to prove it).
 *   "cull"       (default 1): warp-level cull boxes in the blend kernels
 *   "knn_filter" (default 1): warp-level candidate filter of the K-nearest-node search in d2gs_deform_forward
 *   "tile_sort"  (default 1): binning of d2gs_raster_forward by per-tile buckets (count, scan, atomic scatter, one sort per
 *                tile on depth bits | surfel id) instead of one global radix sort on tile | depth bits (0); the per-tile
 *                lists, ranges and everything downstream are bit-identical
 *   "deform_bwd_smem" (default 1): per-CTA shared accumulators in the incoherent d2gs_deform_backward path
 *   "mlp_cluster_fwd" (default 4) / "mlp_cluster_bwd" (default 1): thread-block cluster size of the MLP kernels; the CTAs of a
 *                cluster receive the weight stream through ONE multicast bulk copy per chunk
 *   "tile_order" (default 1): the blend kernels visit tiles longest list first (the tile scan writes the order); set it
 *                before the forward of a frame
 *   "lane_walk"  (default 7): bit 0 / bit 1: the forward / backward blend kernel lets every lane walk its OWN list of
 *                prefilter hits (lanes of a warp work on different surfels at the same time) instead of visiting one surfel
 *                per warp iteration; bit 2: the forward stores its prefilter ballots — one 32-bit word per (instance, 8x4
 *                pixel patch of its tile), 32 B per instance at the end of the binning workspace — and the backward reads them
 *                instead of repeating the prefilter.  Images, n_contrib and final_T are bit-identical in every setting;
 *                gradients differ by the order of the floating-point sums */
D2GS_API int d2gs_set_option(const char* name, int value);
D2GS_API int d2gs_profile_enable(int on);
D2GS_API int d2gs_profile_collect(double* total_ms, int64_t* launches);

D2GS_API const char* d2gs_last_error(void);
D2GS_API const char* d2gs_version(void);
D2GS_API int d2gs_get_config(D2gsConfig* out);

/* ------------------------------------------------------------------------------------------------
 * Rasterizer.  Replaces CudaRasterizer::Rasterizer::{forward,backward,markVisible}
 * (DSR/cuda_rasterizer/rasterizer.h:24-87) and their torch wrappers RasterizeGaussiansCUDA /
 * RasterizeGaussiansBackwardCUDA / markVisible (DSR/rasterize_points.h:18-68).
 * ---------------------------------------------------------------------------------------------- */

/* Workspace sizes.  Replaces the required<GeometryState/ImageState/BinningState>() queries
 * (DSR/cuda_rasterizer/rasterizer_impl.h:64-72).  binning_bytes is for `num_rendered` instances. */
D2GS_API int d2gs_raster_workspace(int P, int width, int height, int64_t num_rendered,
                          size_t* geom_bytes, size_t* img_bytes, size_t* binning_bytes);

typedef struct D2gsRasterFwdArgs {
  /* sizes: P surfels, D = active SH degree, M = SH coefficients per surfel (0 when colours are given) */
  int P, D, M, width, height;
  const float* background;        /* (3) */
  const float* means3D;           /* (P,3) */
  const float* shs;               /* (P,M,3) or NULL; when sh_rest != NULL this is the DC block (P,1,3) */
  const float* sh_rest;           /* (P,M-1,3) or NULL: split layout, avoids the per-frame cat of gaussian_renderer/__init__.py:114,122 */
  const float* colors_precomp;    /* (P,3) or NULL */
  const float* opacities;         /* (P) */
  const float* scales;            /* (P,2) or NULL */
  const float* rotations;         /* (P,4) (w,x,y,z) or NULL */
  const float* transMat_precomp;  /* (P,9) or NULL */
  float scale_modifier;           /* accepted and ignored, like the reference (forward.cu:95) */
  const float* viewmatrix;        /* (4,4) transposed world->view */
  const float* projmatrix;        /* (4,4) transposed full projection */
  const float* campos;            /* (3) */
  float tan_fovx, tan_fovy;
  int prefiltered;
  int debug;                      /* 1: synchronise + check after every stage (auxiliary.h:271-278) */
  /* outputs */
  float* out_color;               /* (3,H,W) */
  float* out_others;              /* (8,H,W): depth, alpha, normal xyz, median depth, distortion, median weight */
  int* radii;                     /* (P) */
  /* workspaces (device). geometry+image buffers must be kept for backward; so must `binning`. */
  void* geom_buffer;   size_t geom_bytes;
  void* img_buffer;    size_t img_bytes;
  void* binning_buffer; size_t binning_bytes;
  int resume;                     /* 1: geometry stage already done by a call that returned D2GS_NEED_BINNING */
  /* results on the host */
  int64_t* num_rendered;          /* host: R, number of (surfel,tile) instances */
  size_t* binning_required;       /* host: bytes needed for R instances */
  /* Raw-parameter mode (raw_params = 1): the activations and deformation deltas of render()
   * (gaussian_renderer/__init__.py:83-99, scene/gaussian_model.py:67-75,101-124) are applied inside the per-surfel
   * kernel instead of by ~10 eager ops:  means3D += d_means3D;  scales = exp(scales) + d_scales;
   * rotations = normalize(rotations + d_rotations);  opacities = sigmoid(opacities).   Deltas may be NULL (= 0). */
  int raw_params;
  const float* d_means3D;         /* (P,3) or NULL */
  const float* d_scales;          /* (P,2) or NULL */
  const float* d_rotations;       /* (P,4) or NULL */
  /* Deferred-count mode (binning_capacity > 0): the instance count R is NOT read back, so the call never synchronises.
   * The binning stage works on exactly `binning_capacity` slots (binning_bytes must cover d2gs_raster_workspace(...,
   * binning_capacity)).  With the per-tile binning (option "tile_sort", default) the slots are only capacity: R, the
   * overflow flag and the tile ranges come from a device-side scan of per-tile counters.  With the global sort
   * ("tile_sort" = 0) slots past R carry all-ones keys, and the stable sort leaves the R real instances first in the
   * order a sort of R items produces.  Either way every result is identical to the synchronous mode.  *num_rendered is set to
   * binning_capacity: pass that value to d2gs_raster_backward / d2gs_raster_export_state (it fixes the workspace
   * layout).  If R > binning_capacity the frame renders nothing and out_color is filled with NaN (never silently
   * wrong).  num_rendered_async: optional PINNED HOST pair {R, overflow flag}, written by an asynchronous copy on
   * `stream` — read it after an event recorded behind this call. */
  int64_t binning_capacity;
  uint32_t* num_rendered_async;
} D2gsRasterFwdArgs;

/* Forward.  With binning_capacity == 0 it synchronises `stream` once (the reference's blocking readback of
 * num_rendered, rasterizer_impl.cu:281-282) to size the binning stage; if binning_bytes is too small it returns
 * D2GS_NEED_BINNING with *binning_required set.  With binning_capacity > 0 see "deferred-count mode" above. */
D2GS_API int d2gs_raster_forward(const D2gsRasterFwdArgs* args, void* stream);

typedef struct D2gsRasterBwdArgs {
  int P, D, M, width, height;
  int64_t num_rendered;
  const float* background;
  const float* means3D;
  const float* shs;
  const float* sh_rest;
  const float* colors_precomp;
  const float* scales;
  const float* rotations;
  const float* transMat_precomp;
  float scale_modifier;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float tan_fovx, tan_fovy;
  const int* radii;
  const void* geom_buffer;
  const void* binning_buffer;
  const void* img_buffer;
  const float* dL_dout_color;   /* (3,H,W) */
  const float* dL_dout_others;  /* (8,H,W) */
  int debug;
  /* scratch: (P,20) floats, zero on entry; left zero on exit (the per-surfel stage clears what it consumes) */
  float* grad_scratch;
  /* outputs; every element is written (no pre-zeroing needed), NULL = not wanted */
  float* dL_dmeans2D;     /* (P,3): xy = the densification "projected gradient" (backward.cu:645-648) */
  float* dL_dcolors;      /* (P,3) */
  float* dL_dopacity;     /* (P,1) */
  float* dL_dmeans3D;     /* (P,3) */
  float* dL_dtransMat;    /* (P,9) */
  float* dL_dsh;          /* (P,M,3), or the DC block (P,1,3) when dL_dsh_rest != NULL */
  float* dL_dsh_rest;     /* (P,M-1,3) or NULL */
  float* dL_dscales;      /* (P,2) */
  float* dL_drotations;   /* (P,4) */
  /* Raw-parameter mode (must match the forward call).  Then: opacities = the logits given to forward;
   * dL_dmeans3D is the gradient of xyz AND d_means3D; dL_dscales the gradient of d_scales (and of the activated
   * scale); dL_dscales_raw that of the log-scales; dL_drotations the gradient of rotations AND d_rotations (through
   * the normalisation); dL_dopacity the gradient of the opacity logits. */
  int raw_params;
  const float* opacities;
  const float* d_means3D;
  const float* d_scales;
  const float* d_rotations;
  float* dL_dscales_raw;  /* (P,2) */
} D2gsRasterBwdArgs;

D2GS_API int d2gs_raster_backward(const D2gsRasterBwdArgs* args, void* stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer_impl.cu:141-153). present: (P) bytes. */
D2GS_API int d2gs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                      uint8_t* present, void* stream);

/* Debug/parity access to the intermediates the reference keeps inside geomBuffer/binningBuffer/imgBuffer
 * (rasterizer_impl.cu:155-194).  Copies into caller-provided DEVICE arrays (NULL = skip). */
typedef struct D2gsRasterState {
  float* means2D;         /* (P,2) */
  float* depths;          /* (P) */
  float* transMat;        /* (P,9) */
  float* normal_opacity;  /* (P,4) */
  float* rgb;             /* (P,3) */
  uint8_t* clamped;       /* (P,3) */
  uint32_t* tiles_touched;   /* (P) */
  uint32_t* point_offsets;   /* (P) */
  uint64_t* keys_unsorted;   /* (R) reference format tile << 32 | depth bits; with per-tile binning: in bucket order */
  uint32_t* values_unsorted; /* (R) */
  uint64_t* keys_sorted;     /* (R) */
  uint32_t* point_list;      /* (R) */
  uint32_t* ranges;          /* (tiles,2) */
  float* final_T;            /* (3,H,W): T, dist1, dist2 */
  uint32_t* n_contrib;       /* (2,H,W): last contributor, median contributor */
} D2gsRasterState;

D2GS_API int d2gs_raster_export_state(int P, int width, int height, int64_t num_rendered, const void* geom_buffer,
                             const void* binning_buffer, const void* img_buffer, const D2gsRasterState* out,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * Image-space epilogue of render(): replaces gaussian_renderer/__init__.py:172-207 and depth_to_normal /
 * depths_to_points (utils/point_utils.py:9-38).  All planes are (c,H,W) float32 on the device.
 * ---------------------------------------------------------------------------------------------- */
typedef struct D2gsEpilogueArgs {
  int width, height;
  const float* allmap;        /* (8,H,W) rasterizer output */
  const float* viewmatrix;    /* (4,4) transposed world->view */
  float focal_x, focal_y;     /* W / (2 tan(FoVx/2)), H / (2 tan(FoVy/2)) */
  /* forward outputs */
  float* alpha;               /* (1,H,W) */
  float* rend_normal;         /* (3,H,W) world space */
  float* rend_dist;           /* (1,H,W) */
  float* depth;               /* (1,H,W) median depth, nan_to_num'd */
  float* surf_normal;         /* (3,H,W) normal of the depth map x alpha */
  float* surf_point;          /* (3,H,W) unprojected depth */
  /* backward: upstream gradients (NULL = zero) and the result */
  const float* g_alpha; const float* g_rend_normal; const float* g_rend_dist; const float* g_depth;
  const float* g_surf_normal; const float* g_surf_point;
  float* dL_dallmap;          /* (8,H,W), fully written */
} D2gsEpilogueArgs;

D2GS_API int d2gs_epilogue_forward(const D2gsEpilogueArgs* args, void* stream);
D2GS_API int d2gs_epilogue_backward(const D2gsEpilogueArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Photometric loss of the training step, fused.  Replaces l1_loss / ssim
 * (utils/loss_utils.py:18-19,33-76: 11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over
 * all elements) and the normal-consistency / distortion terms and their composition (train_gui.py:292-313):
 *   loss = (1 - lambda_dssim) * L1 + lambda_dssim * (1 - SSIM) + lambda_normal * mean(1 - <rend_normal, surf_normal>)
 *        + lambda_dist * mean(rend_dist)
 * ---------------------------------------------------------------------------------------------- */
typedef struct D2gsLossArgs {
  int width, height;
  const float* image;          /* (3,H,W) rendered */
  const float* gt;             /* (3,H,W) target */
  const float* rend_normal;    /* (3,H,W) or NULL: the normal term is dropped */
  const float* surf_normal;    /* (3,H,W) or NULL */
  const float* rend_dist;      /* (1,H,W) or NULL: the distortion term is dropped */
  float lambda_dssim, lambda_normal, lambda_dist;
  float* out;                  /* device (5): loss, L1, SSIM, weighted normal term, weighted distortion term */
  void* workspace;             /* d2gs_loss_workspace bytes: reduction scratch + the three SSIM partial-derivative maps */
  size_t workspace_bytes;
  int save_for_backward;       /* 0: forward only (the maps are not written) */
  /* backward (reads the workspace written by the forward call with the same inputs) */
  const float* upstream;       /* device scalar dL/dloss, or NULL (= 1) */
  float* g_image;              /* (3,H,W) fully written */
  float* g_rend_normal;        /* (3,H,W) or NULL */
  float* g_surf_normal;        /* (3,H,W) or NULL */
  float* g_rend_dist;          /* (1,H,W) or NULL */
} D2gsLossArgs;

D2GS_API int d2gs_loss_workspace(int width, int height, size_t* bytes);
D2GS_API int d2gs_loss_forward(const D2gsLossArgs* args, void* stream);
D2GS_API int d2gs_loss_backward(const D2gsLossArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step, fused.  Replaces the per-step work of the two torch.optim.Adam instances
 * the reference builds (scene/gaussian_model.py:181-203: seven surfel groups, lr=0.0, eps=1e-15; the deformation
 * optimiser; stepped in train_gui.py:426-432) with ONE launch per optimiser: the arithmetic of torch's
 * _single_tensor_adam (no weight decay, no amsgrad, maximize off).  The bias corrections are computed by the caller
 * from its step count, in double like torch:  step_size = lr / (1 - beta1^t),  bias_correction2_sqrt = sqrt(1 - beta2^t).
 * ---------------------------------------------------------------------------------------------- */
typedef struct D2gsAdamTensor {
  float* param;            /* updated in place */
  const float* grad;
  float* exp_avg;          /* first moment, updated in place */
  float* exp_avg_sq;       /* second moment, updated in place */
  int64_t numel;
  double beta1, beta2;          /* double: torch forms 1 - beta in double before rounding to the tensor's float */
  float eps;
  float step_size;              /* lr / bias_correction1 */
  float bias_correction2_sqrt;  /* sqrt(1 - beta2^t) */
} D2gsAdamTensor;

/* `tensors` is a HOST array; any count (split into launches of 48 descriptors). */
D2GS_API int d2gs_adam_step(const D2gsAdamTensor* tensors, int count, void* stream);

/* add_densification_stats (scene/gaussian_model.py:484-486): accum[i] += |viewspace_grad[i,:2]|, denom[i] += 1 where
 * update_filter[i] (bool, 1 byte) — one kernel, no boolean-mask indexing (which forces a device synchronisation). */
D2GS_API int d2gs_densification_stats(int P, const float* viewspace_grad, int grad_stride, const uint8_t* update_filter,
                                      float* xyz_gradient_accum, float* denom, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Deformation MLP, fused.  Replaces DeformNetwork.forward (utils/time_utils.py:410-453: embedders :208-256,
 * timenet :344-346, 8x256 trunk with the skip concat :348-352,416-420, heads :363-371,422-452) and its autograd
 * backward.  Fixed architecture of the reference: D=8, W=256, multires=10, skip after layer 4; is_blender selects the
 * 6-frequency time embedding + timenet (13->256->30) or the 10-frequency embedding fed directly (21).
 * Head rows are concatenated by the caller: [gaussian_warp 3 | gaussian_scaling 2 | gaussian_rotation 4 |
 * local_rotation 4 (optional) | gaussian_opacity 1 (optional)] -> num_out <= 16.
 * ---------------------------------------------------------------------------------------------- */
typedef struct D2gsMlpArgs {
  int rows;                 /* number of evaluated points (control nodes, or nodes x times) */
  int is_blender;
  int num_out;
  const float* x;           /* (rows,3) */
  const float* t;           /* time per row; element i at t[i * t_stride] (t_stride 0: one time for all rows) */
  int t_stride;
  const float* timenet0_w; const float* timenet0_b;   /* (256,13) (256)  — NULL unless is_blender */
  const float* timenet2_w; const float* timenet2_b;   /* (30,256) (30) */
  const float* linear_w[8]; const float* linear_b[8]; /* (256,in) (256) */
  const float* heads_w; const float* heads_b;         /* (num_out,256) (num_out) */
  void* workspace; size_t workspace_bytes;            /* from d2gs_mlp_workspace; forward fills it, backward reads it */
  /* forward output */
  float* out;               /* (rows,num_out) */
  /* backward */
  const float* g_out;       /* (rows,num_out) */
  float* g_timenet0_w; float* g_timenet0_b; float* g_timenet2_w; float* g_timenet2_b;
  float* g_linear_w[8]; float* g_linear_b[8];
  float* g_heads_w; float* g_heads_b;                 /* all fully written */
} D2gsMlpArgs;

D2GS_API int d2gs_mlp_workspace(int rows, int is_blender, int num_out, size_t* bytes);
D2GS_API int d2gs_mlp_forward(const D2gsMlpArgs* args, void* stream);
D2GS_API int d2gs_mlp_backward(const D2gsMlpArgs* args, void* stream);
/* hidden activations of the last trunk layer (rows,256) inside a workspace filled by d2gs_mlp_forward */
D2GS_API const float* d2gs_mlp_hidden(int rows, int is_blender, int num_out, const void* workspace);

/* ------------------------------------------------------------------------------------------------
 * Node-controlled deformation: KNN weights + blend.  Replaces ControlNodeWarp.cal_nn_weight
 * (utils/time_utils.py:934-967, incl. pytorch3d.ops.knn_points at :950) and the blend of
 * ControlNodeWarp.forward (utils/time_utils.py:1145-1157,1190-1194), fused with the activations of
 * render() (gaussian_renderer/__init__.py:83-99).
 * ---------------------------------------------------------------------------------------------- */
typedef struct D2gsDeformFwdArgs {
  int P, M, K, hyper_dim;        /* K <= 8 */
  const float* xyz;              /* (P,3) canonical surfel centres */
  const float* feature;          /* (P,feature_stride) first hyper_dim columns are the hyper coordinates; NULL: 3-D query */
  int feature_stride;
  const float* nodes;            /* (M,3+hyper_dim) */
  const float* node_radius_log;  /* (M) _node_radius */
  const float* node_weight_logit;/* (M) _node_weight or NULL */
  const float* node_trans;       /* (M,3) MLP output d_xyz */
  const float* node_rot;         /* (M,4) MLP output d_rotation */
  const float* node_scale;       /* (M,2) MLP output d_scaling */
  const float* node_local_rot;   /* (M,4) MLP output local_rotation or NULL (local_frame off) */
  const float* motion_mask;      /* (P) or NULL (= ones) */
  /* outputs */
  int64_t* nn_idx;               /* (P,K) ascending distance */
  float* nn_dist;                /* (P,K) squared distances */
  float* nn_weight;              /* (P,K) normalised */
  float* d_xyz;                  /* (P,3) */
  float* d_rotation;             /* (P,4) */
  float* d_scaling;              /* (P,2) */
  int node_attr_stride;          /* 0: node_trans/rot/scale/local_rot are separate packed tables; > 0: they are columns
                                    of ONE (M,node_attr_stride) row-major matrix (the MLP head output), no slicing copies */
  const int32_t* order;          /* optional (P): processing order from d2gs_deform_order; results do not depend on it */
} D2gsDeformFwdArgs;

D2GS_API int d2gs_deform_forward(const D2gsDeformFwdArgs* args, void* stream);

typedef struct D2gsDeformBwdArgs {
  int P, M, K, hyper_dim;
  const float* xyz;
  const float* feature; int feature_stride;
  const float* nodes;
  const float* node_radius_log;
  const float* node_weight_logit;
  const float* node_trans;
  const float* node_rot;
  const float* node_scale;
  const float* node_local_rot;
  const float* motion_mask;
  const int64_t* nn_idx;
  const float* nn_dist;
  const float* nn_weight;
  const float* dL_d_xyz;         /* (P,3) */
  const float* dL_d_rotation;    /* (P,4) */
  const float* dL_d_scaling;     /* (P,2) */
  /* outputs (accumulated with atomics: must be zero on entry) */
  float* dL_dnode_trans;         /* (M,3) */
  float* dL_dnode_rot;           /* (M,4) */
  float* dL_dnode_scale;         /* (M,2) */
  float* dL_dnode_local_rot;     /* (M,4) or NULL */
  float* dL_dnodes;              /* (M,3+hyper_dim): only the hyper columns receive gradient */
  float* dL_dnode_radius_log;    /* (M) */
  float* dL_dnode_weight_logit;  /* (M) or NULL */
  /* outputs (written) */
  float* dL_dfeature;            /* (P,feature_stride) or NULL */
  float* dL_dmotion_mask;        /* (P) or NULL */
  int node_attr_stride;          /* as in D2gsDeformFwdArgs; applies to node_* and dL_dnode_{trans,rot,scale,local_rot} */
  const int32_t* order;          /* optional (P): with an order the node gradients are reduced per warp over the distinct
                                    nodes of 32 neighbouring surfels instead of one atomic per (surfel, node, component) */
} D2gsDeformBwdArgs;

D2GS_API int d2gs_deform_backward(const D2gsDeformBwdArgs* args, void* stream);

/* Spatial processing order for the two calls above: a permutation of 0..P-1 that sorts the surfel centres along a
 * 30-bit Morton curve of their bounding box, so that the 32 surfels of a warp are neighbours and share their control
 * nodes (the reference has no counterpart: pytorch3d.ops.knn_points, utils/time_utils.py:950, is order-agnostic).
 * The order only affects speed; it may be reused for many frames while the centres move slowly. */
D2GS_API int d2gs_deform_order_workspace(int P, size_t* bytes);
D2GS_API int d2gs_deform_order(int P, const float* xyz, int32_t* order, void* workspace, size_t workspace_bytes, void* stream);

/* Mean squared distance of every point to its 3 nearest neighbours (the point itself excluded by index): replaces
 * simple_knn._C.distCUDA2 (submodules/simple-knn/spatial.cu:15-26 -> SimpleKNN::knn, simple_knn.cu:185-220), which
 * scene/gaussian_model.py:162 calls once to initialise the surfel scales.  points: (P,3) float32, mean_dist2: (P).
 * Exact 3-NN (two-level boxes over a Morton order, one warp per 32 neighbouring queries); all work on `stream`, no
 * allocation, no host synchronisation (the reference allocates with cudaMalloc/thrust and copies its bounding box to
 * the host twice). */
D2GS_API int d2gs_knn_mean_dist2_workspace(int P, size_t* bytes);
D2GS_API int d2gs_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3-D Gaussian rasterizer with colour, depth and alpha outputs.  Replaces
 * CudaRasterizer::Rasterizer::{forward,backward} of the reference's second rasterizer
 * (DGR = submodules/diff-gaussian-rasterization: DGR/cuda_rasterizer/rasterizer.h, rasterizer_impl.cu:196-446) and
 * its torch wrappers RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:35-217), which
 * render_flow (gaussian_renderer/__init__.py:222-337) calls.  d2gs_mark_visible serves both rasterizers.
 * Same protocol as d2gs_raster_forward: caller-owned workspaces, one stream synchronisation to read the instance
 * count (DGR rasterizer_impl.cu:270-271), D2GS_NEED_BINNING + resume when the binning workspace is too small.
 * Quaternions are used as given (the reference's variant does not normalise them, forward.cu:127).
 * ---------------------------------------------------------------------------------------------- */
D2GS_API int d2gs_gs3d_workspace(int P, int width, int height, int64_t num_rendered, size_t* geom_bytes, size_t* img_bytes,
                                 size_t* binning_bytes);

typedef struct D2gsGs3dFwdArgs {
  int P, D, M, width, height;     /* Gaussians, active SH degree, SH coefficients per Gaussian (0 with colours) */
  const float* background;        /* (3) */
  const float* means3D;           /* (P,3) */
  const float* shs;               /* (P,M,3) or NULL */
  const float* colors_precomp;    /* (P,3) or NULL */
  const float* opacities;         /* (P) */
  const float* scales;            /* (P,3) or NULL */
  const float* rotations;         /* (P,4) (r,x,y,z), 16-byte aligned, or NULL */
  const float* cov3D_precomp;     /* (P,6) upper triangle or NULL */
  float scale_modifier;
  const float* viewmatrix;        /* (4,4) transposed world->view */
  const float* projmatrix;        /* (4,4) transposed full projection */
  const float* campos;            /* (3) */
  float tan_fovx, tan_fovy;
  int prefiltered;
  int debug;
  float* out_color;               /* (3,H,W) */
  float* out_depth;               /* (1,H,W): sum_i w_i depth_i */
  float* out_alpha;               /* (1,H,W): sum_i w_i; kept by the caller for the backward */
  int* radii;                     /* (P) */
  void* geom_buffer;   size_t geom_bytes;
  void* img_buffer;    size_t img_bytes;
  void* binning_buffer; size_t binning_bytes;
  int resume;                     /* 1: per-Gaussian stage already done by a call that returned D2GS_NEED_BINNING */
  int64_t* num_rendered;          /* host */
  size_t* binning_required;       /* host */
  /* Deferred-count mode, as in D2gsRasterFwdArgs: binning_capacity > 0 = bin into exactly that many instance slots without
   * reading the count back (needs the per-tile binning, option "tile_sort"); *num_rendered = binning_capacity;
   * num_rendered_async (pinned host, optional) receives {R, R > capacity} through an asynchronous copy on the stream;
   * an overflowing frame renders nothing and its colour planes are NaN. */
  int64_t binning_capacity;
  int32_t* num_rendered_async;
} D2gsGs3dFwdArgs;

D2GS_API int d2gs_gs3d_forward(const D2gsGs3dFwdArgs* args, void* stream);

typedef struct D2gsGs3dBwdArgs {
  int P, D, M, width, height;
  int64_t num_rendered;
  const float* background;
  const float* means3D;
  const float* shs;
  const float* colors_precomp;
  const float* scales;
  const float* rotations;
  const float* cov3D_precomp;
  float scale_modifier;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float tan_fovx, tan_fovy;
  const int* radii;
  const float* out_alpha;        /* (1,H,W) from the forward */
  const void* geom_buffer;
  const void* binning_buffer;
  const void* img_buffer;
  const float* dL_dout_color;    /* (3,H,W) */
  const float* dL_dout_depth;    /* (1,H,W) */
  const float* dL_dout_alpha;    /* (1,H,W) */
  int debug;
  float* grad_scratch;           /* (P,12) floats; cleared by the call */
  /* outputs; every element is written, NULL = not wanted */
  float* dL_dmeans2D;            /* (P,3): xy in NDC units (the densification statistic), z = 0 */
  float* dL_dcolors;             /* (P,3) */
  float* dL_dopacity;            /* (P,1) */
  float* dL_dmeans3D;            /* (P,3) */
  float* dL_dcov3D;              /* (P,6) */
  float* dL_dsh;                 /* (P,M,3) */
  float* dL_dscales;             /* (P,3) */
  float* dL_drotations;          /* (P,4), 16-byte aligned */
} D2gsGs3dBwdArgs;

D2GS_API int d2gs_gs3d_backward(const D2gsGs3dBwdArgs* args, void* stream);

/* Parity access to the intermediates (device arrays, NULL = skip).  rec: (P,12) floats = mean2D.xy, depth, prefilter
 * threshold | conic.xyz, opacity | rgb, 0 */
typedef struct D2gsGs3dState {
  float* rec;
  float* cov3D;              /* (P,6) */
  uint8_t* clamped;          /* (P): bit c set = colour channel c clamped at 0 */
  uint32_t* tiles_touched;   /* (P) */
  uint64_t* keys_sorted;     /* (R) */
  uint32_t* point_list;      /* (R) */
  uint32_t* ranges;          /* (tiles,2) */
  uint32_t* n_contrib;       /* (H,W) */
} D2gsGs3dState;
D2GS_API int d2gs_gs3d_export_state(int P, int width, int height, int64_t num_rendered, const void* geom_buffer,
                                    const void* binning_buffer, const void* img_buffer, const D2gsGs3dState* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D2GS_H_ */
