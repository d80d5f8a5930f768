// C-ABI entry points of libd2gs.so (see include/d2gs.h).  Host-side orchestration only: workspace carving,
// stage sequencing on the caller's stream, CUB scan / radix sort for the binning stage.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/d2gs.h"
#include "raster_common.cuh"
#include "deform.cuh"
#include "epilogue.cuh"
#include "loss.cuh"
#include "optim.cuh"
#include "mlp.cuh"
#include "knn.cuh"
#include "gs3d.cuh"

namespace d2gs {

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define D2GS_CUDA_OK(expr)                                                                       \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return fail(D2GS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));           \
  } while (0)

// after a stage: launch errors always, execution errors too when debug is on (reference: auxiliary.h:271-278)
#define D2GS_STAGE(name, debug, stream)                                                          \
  do {                                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ == cudaSuccess && (debug)) e__ = cudaStreamSynchronize(stream);                      \
    if (e__ != cudaSuccess)                                                                      \
      return fail(D2GS_ERR_CUDA, std::string("stage ") + name + ": " + cudaGetErrorString(e__)); \
  } while (0)

// ---- optional per-stage timing (CUDA events on the launch stream) ------------------------------------------
enum Stage { ST_PRE = 0, ST_SCAN, ST_DUP, ST_SORT, ST_RANGES, ST_BLEND_F, ST_BLEND_B, ST_PRE_B, ST_DEF_F, ST_DEF_B, ST_EPI_F, ST_EPI_B, ST_MLP_F, ST_MLP_B, ST_LOSS_F, ST_LOSS_B };
struct StageRec { int stage; cudaEvent_t a, b; };
static bool g_profile = false;
int g_deform_bwd_smem = 1;  // node-gradient accumulation of deform_bwd: 1 = per-CTA shared accumulators, 0 = global reductions
int g_knn_filter = 1;       // warp-level candidate filter of the K-nearest-node search (0: every node is visited; same results)
static int g_tile_sort = 1;   // binning: 0 = global radix sort of (tile | depth) keys, 1 = per-tile buckets + segmented sort (same lists)
static int g_tile_order = 1;  // blend CTAs visit tiles longest list first (scheduling only; set before the forward of a frame)
static int g_lane_walk = 7;   // bit 0: forward blend, bit 1: backward blend walk per-lane hit lists; bit 2: the forward hands its
                              // prefilter ballots to the backward (A/B switches; same results)
static int g_cull = 1;   // warp-level cull boxes in the blend kernels (tests switch it off to prove it changes nothing)
static std::vector<StageRec> g_recs;
static std::vector<cudaEvent_t> g_free_events;

static cudaEvent_t get_event() {
  if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct StageTimer {
  bool on; cudaStream_t s; StageRec r;
  StageTimer(int stage, cudaStream_t stream) : on(g_profile), s(stream) {
    if (on) { r.stage = stage; r.a = get_event(); r.b = get_event(); cudaEventRecord(r.a, s); }
  }
  ~StageTimer() { if (on) { cudaEventRecord(r.b, s); g_recs.push_back(r); } }
};

GeomLayout geom_layout(int P) {
  GeomLayout L{};
  size_t o = 0;
  L.rec = o; o = align_up(o + sizeof(SurfelRec) * (size_t)P);
  L.clamped = o; o = align_up(o + (size_t)P);
  L.tiles_touched = o; o = align_up(o + 4 * (size_t)P);
  L.point_offsets = o; o = align_up(o + 4 * (size_t)P);
  size_t tmp = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, P > 0 ? P : 1);
  L.scan_temp_bytes = tmp;
  L.scan_temp = o; o = align_up(o + tmp);
  L.status = o; o = align_up(o + 4 * sizeof(uint32_t));   // {R, overflow flag, number of long tiles (per-tile binning)}
  L.tile_box = o; o = align_up(o + 16 * (size_t)P);       // packed tile rectangle + depth bits per surfel (per-tile binning)
  L.total = o + 256;
  return L;
}

ImgLayout img_layout(int W, int H) {
  ImgLayout L{};
  const size_t HW = (size_t)W * H;
  const size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
  size_t o = 0;
  L.ranges = o; o = align_up(o + 8 * tiles);
  L.final_T = o; o = align_up(o + 4 * 3 * HW);
  L.n_contrib = o; o = align_up(o + 4 * 2 * HW);
  L.tile_count = o; o = align_up(o + 4 * tiles);
  L.seg_begin = o; o = align_up(o + 4 * tiles);
  L.seg_end = o; o = align_up(o + 4 * tiles);
  L.tile_order = o; o = align_up(o + 4 * tiles);
  L.total = o + 256;
  return L;
}

BinLayout bin_layout(int64_t R) {
  BinLayout L{};
  const size_t n = R > 0 ? (size_t)R : 1;
  size_t o = 0;
  L.keys_unsorted = o; o = align_up(o + 8 * n);
  L.keys_sorted = o; o = align_up(o + 8 * n);
  L.vals_unsorted = o; o = align_up(o + 4 * n);
  L.point_list = o; o = align_up(o + 4 * n);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n);
  L.sort_temp_bytes = tmp;
  L.sort_temp = o; o = align_up(o + tmp);
  L.hit_mask = o; o = align_up(o + 32 * n);     // one 32-bit prefilter ballot per (instance, 8x4 patch of its tile)
  L.total = o + 256;
  return L;
}

static char* aligned_base(const void* p) {
  return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
}

// number of key bits above the 32 depth bits (reference: rasterizer_impl.cu:35-50)
static uint32_t higher_msb(uint32_t n) {
  uint32_t msb = 16, step = 16;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

}  // namespace d2gs

using namespace d2gs;

extern "C" {

const char* d2gs_last_error(void) { return g_last_error.c_str(); }

int d2gs_profile_enable(int on) { g_profile = on != 0; return D2GS_OK; }

int d2gs_set_option(const char* name, int value) {
  if (!name) return fail(D2GS_ERR_INVALID_ARG, "null option name");
  if (std::strcmp(name, "cull") == 0) { g_cull = value != 0; return D2GS_OK; }
  if (std::strcmp(name, "deform_bwd_smem") == 0) { g_deform_bwd_smem = value != 0; return D2GS_OK; }
  if (std::strcmp(name, "knn_filter") == 0) { g_knn_filter = value != 0; return D2GS_OK; }
  if (std::strcmp(name, "tile_sort") == 0) { g_tile_sort = value != 0; return D2GS_OK; }
  if (std::strcmp(name, "mlp_cluster_fwd") == 0) { mlp_set_cluster(0, value); return D2GS_OK; }
  if (std::strcmp(name, "mlp_cluster_bwd") == 0) { mlp_set_cluster(1, value); return D2GS_OK; }
  if (std::strcmp(name, "lane_walk") == 0) { g_lane_walk = value; return D2GS_OK; }
  if (std::strcmp(name, "tile_order") == 0) { g_tile_order = value != 0; return D2GS_OK; }
  return fail(D2GS_ERR_INVALID_ARG, std::string("unknown option ") + name);
}

int d2gs_profile_collect(double* total_ms, int64_t* launches) {
  if (!total_ms || !launches) return fail(D2GS_ERR_INVALID_ARG, "null output");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  for (const StageRec& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.stage >= 0 && r.stage < D2GS_NUM_STAGES) {
      total_ms[r.stage] += ms;
      launches[r.stage] += 1;
    }
    g_free_events.push_back(r.a);
    g_free_events.push_back(r.b);
  }
  g_recs.clear();
  return D2GS_OK;
}
const char* d2gs_version(void) { return "d2gs-b200 0.1 (sm_100a)"; }

int d2gs_get_config(D2gsConfig* c) {
  if (!c) return fail(D2GS_ERR_INVALID_ARG, "null config");
  c->num_channels = NUM_CH; c->block_x = TILE_X; c->block_y = TILE_Y;
  c->tight_bbox = 0; c->render_auxiliary = 1; c->backface_cull = 1; c->dual_visible = 1; c->detach_weight = 0;
  c->near_plane = D2GS_NEAR_PLANE; c->far_plane = D2GS_FAR_PLANE; c->filter_size = D2GS_FILTER_SIZE;
  c->sm_arch = 100;
  return D2GS_OK;
}

int d2gs_raster_workspace(int P, int width, int height, int64_t num_rendered, size_t* geom_bytes, size_t* img_bytes,
                          size_t* binning_bytes) {
  if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail(D2GS_ERR_INVALID_ARG, "negative size");
  if (geom_bytes) *geom_bytes = geom_layout(P).total;
  if (img_bytes) *img_bytes = img_layout(width, height).total;
  if (binning_bytes) *binning_bytes = bin_layout(num_rendered).total;
  return D2GS_OK;
}

int d2gs_raster_forward(const D2gsRasterFwdArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (!a->out_color || !a->out_others || !a->num_rendered) return fail(D2GS_ERR_INVALID_ARG, "missing outputs");
  const int P = a->P, W = a->width, H = a->height;
  const size_t HW = (size_t)W * H;
  if (P == 0) {   // reference returns zero images for an empty scene (rasterize_points.cu:92-94,106)
    D2GS_CUDA_OK(cudaMemsetAsync(a->out_color, 0, 4 * 3 * HW, stream));
    D2GS_CUDA_OK(cudaMemsetAsync(a->out_others, 0, 4 * 8 * HW, stream));
    *a->num_rendered = 0;
    return D2GS_OK;
  }
  if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->campos || !a->background || !a->radii)
    return fail(D2GS_ERR_INVALID_ARG, "missing inputs");
  {
    const void* v16[] = {a->shs, a->sh_rest, a->rotations, a->d_rotations};
    for (const void* q : v16)
      if (q && ((uintptr_t)q & 15)) return fail(D2GS_ERR_INVALID_ARG, "SH / rotation tables must be 16-byte aligned");
  }
  if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
    return fail(D2GS_ERR_INVALID_ARG, "provide exactly one of SHs or precomputed colours");
  if (((a->scales == nullptr) || (a->rotations == nullptr)) == (a->transMat_precomp == nullptr))
    return fail(D2GS_ERR_INVALID_ARG, "provide exactly one of scale/rotation pair or precomputed transMat");
  if (a->shs && (a->M < 1 || a->M > 16 || a->D < 0 || a->D > 3 || (a->D + 1) * (a->D + 1) > a->M))
    return fail(D2GS_ERR_INVALID_ARG, "SH degree / coefficient count out of range (deg<=3, M<=16)");

  const GeomLayout GL = geom_layout(P);
  const ImgLayout IL = img_layout(W, H);
  if (a->geom_bytes < GL.total || !a->geom_buffer) return fail(D2GS_ERR_WORKSPACE, "geometry workspace too small");
  if (a->img_bytes < IL.total || !a->img_buffer) return fail(D2GS_ERR_WORKSPACE, "image workspace too small");
  char* gb = aligned_base(a->geom_buffer);
  char* ib = aligned_base(a->img_buffer);
  SurfelRec* rec = (SurfelRec*)(gb + GL.rec);
  uint8_t* clamped = (uint8_t*)(gb + GL.clamped);
  uint32_t* tiles_touched = (uint32_t*)(gb + GL.tiles_touched);
  uint32_t* point_offsets = (uint32_t*)(gb + GL.point_offsets);
  uint2* ranges = (uint2*)(ib + IL.ranges);
  float* final_T = (float*)(ib + IL.final_T);
  uint32_t* n_contrib = (uint32_t*)(ib + IL.n_contrib);

  FwdParams p{};
  p.P = P; p.D = a->D; p.M = a->M; p.W = W; p.H = H;
  p.bg = a->background; p.means3D = a->means3D; p.shs = a->shs; p.sh_rest = a->sh_rest;
  p.colors_precomp = a->colors_precomp; p.opacities = a->opacities; p.scales = a->scales; p.rotations = a->rotations;
  p.transMat_precomp = a->transMat_precomp; p.view = a->viewmatrix; p.proj = a->projmatrix; p.campos = a->campos;
  p.tan_fovx = a->tan_fovx; p.tan_fovy = a->tan_fovy;
  p.focal_y = H / (2.0f * a->tan_fovy);
  p.focal_x = W / (2.0f * a->tan_fovx);
  p.prefiltered = a->prefiltered;
  p.gx = (W + TILE_X - 1) / TILE_X; p.gy = (H + TILE_Y - 1) / TILE_Y;
  p.raw = a->raw_params; p.d_means3D = a->d_means3D; p.d_scales = a->d_scales; p.d_rotations = a->d_rotations;
  if (p.raw && (a->transMat_precomp || !a->scales || !a->rotations))
    return fail(D2GS_ERR_INVALID_ARG, "raw-parameter mode needs scales and rotations (no transMat_precomp)");

  uint32_t* status = (uint32_t*)(gb + GL.status);
  const bool deferred = a->binning_capacity > 0;
  const uint32_t tiles = p.gx * p.gy;
  const bool tile_sort = g_tile_sort && tiles <= (uint32_t)TILE_SORT_MAX_TILES;
  uint32_t* tile_count = (uint32_t*)(ib + IL.tile_count);
  uint32_t* seg_begin = (uint32_t*)(ib + IL.seg_begin);
  uint32_t* big_list = (uint32_t*)(ib + IL.seg_end);
  uint32_t* tile_order = g_tile_order ? (uint32_t*)(ib + IL.tile_order) : nullptr;
  if (deferred && a->binning_capacity > 0xffffffffll) return fail(D2GS_ERR_INVALID_ARG, "binning_capacity exceeds 2^32-1 instances");
  uint4* tile_box = (uint4*)(gb + GL.tile_box);
  p.tile_box = tile_sort ? tile_box : nullptr;
  // hit-mask handshake with the backward: the per-surfel kernel stamps status[3] (every frame, so a frame rendered with
  // another setting never leaves a stale stamp behind)
  const bool write_masks = (g_lane_walk & 1) && (g_lane_walk & 4);
  p.frame_flag = status + 3;
  p.frame_flag_value = write_masks ? HIT_MASK_MAGIC : 0u;
  if (!a->resume) {
    { StageTimer t(ST_PRE, stream); launch_preprocess_fwd(p, rec, clamped, a->radii, tiles_touched, stream); }
    D2GS_STAGE("preprocess", a->debug, stream);
    if (tile_sort) {
      // per-tile counters -> ranges, instance total and overflow flag (no per-surfel scan)
      StageTimer t(ST_SCAN, stream);
      launch_tile_count(P, tile_box, p.gx, p.gy, tile_count, stream);
      launch_tile_scan(tiles, deferred ? (uint32_t)a->binning_capacity : 0xffffffffu, tile_count, seg_begin, ranges, big_list, status, tile_order, stream);
    } else {
      size_t tmp = GL.scan_temp_bytes;
      StageTimer t(ST_SCAN, stream);
      D2GS_CUDA_OK(cub::DeviceScan::InclusiveSum(gb + GL.scan_temp, tmp, tiles_touched, point_offsets, P, stream));
    }
    D2GS_STAGE("scan", a->debug, stream);
  }
  int64_t R = 0;       // instance slots the binning stage works on: the exact count, or the caller's capacity
  if (deferred) {
    // Deferred-count mode: nothing is read back, so the host never waits for the device.  The binning stage runs on
    // exactly `binning_capacity` slots; slots past the real count carry all-ones keys and sort to the end.
    R = a->binning_capacity;
    *a->num_rendered = R;
  } else {
    uint32_t R32 = 0;
    D2GS_CUDA_OK(cudaMemcpyAsync(&R32, tile_sort ? status : point_offsets + P - 1, 4, cudaMemcpyDeviceToHost, stream));
    D2GS_CUDA_OK(cudaStreamSynchronize(stream));
    R = R32;
    *a->num_rendered = R;
  }
  const BinLayout BL = bin_layout(R);
  if (a->binning_required) *a->binning_required = BL.total;
  if (a->binning_bytes < BL.total || !a->binning_buffer) {
    if (deferred) return fail(D2GS_ERR_WORKSPACE, "binning workspace smaller than binning_capacity instances need");
    return D2GS_NEED_BINNING;
  }

  char* bb = aligned_base(a->binning_buffer);
  uint64_t* keys_unsorted = (uint64_t*)(bb + BL.keys_unsorted);
  uint64_t* keys_sorted = (uint64_t*)(bb + BL.keys_sorted);
  uint32_t* vals_unsorted = (uint32_t*)(bb + BL.vals_unsorted);
  uint32_t* point_list = (uint32_t*)(bb + BL.point_list);

  if (tile_sort) {
    { StageTimer t(ST_DUP, stream);
      launch_tile_scatter(P, tile_box, p.gx, p.gy, seg_begin, tile_count, status, keys_unsorted, stream); }
    D2GS_STAGE("scatter", a->debug, stream);
    if (deferred && a->num_rendered_async)
      D2GS_CUDA_OK(cudaMemcpyAsync(a->num_rendered_async, status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    if (R > 0) {
      StageTimer t(ST_SORT, stream);
      launch_tile_sort(tiles, ranges, big_list, status, keys_unsorted, keys_sorted, point_list, stream);
      D2GS_STAGE("tile sort", a->debug, stream);
    }
  } else {
  { StageTimer t(ST_DUP, stream);
    launch_duplicate(P, rec, a->radii, point_offsets, keys_unsorted, vals_unsorted, p.gx, p.gy, (uint32_t)R, stream);
    if (deferred) launch_pad_keys((uint32_t)R, point_offsets + P - 1, keys_unsorted, status, stream); }
  D2GS_STAGE("duplicate", a->debug, stream);
  if (deferred && a->num_rendered_async)
    D2GS_CUDA_OK(cudaMemcpyAsync(a->num_rendered_async, status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  if (R > 0) {
    const int bit = (int)higher_msb(p.gx * p.gy);
    size_t tmp = BL.sort_temp_bytes;
    StageTimer t(ST_SORT, stream);
    D2GS_CUDA_OK(cub::DeviceRadixSort::SortPairs(bb + BL.sort_temp, tmp, keys_unsorted, keys_sorted, vals_unsorted,
                                                 point_list, (int)R, 0, 32 + bit, stream));
    D2GS_STAGE("sort", a->debug, stream);
  }
  { StageTimer t(ST_RANGES, stream);
    D2GS_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)p.gx * p.gy, stream));
    if (deferred) launch_ranges_deferred((uint32_t)R, status, keys_sorted, ranges, stream);
    else launch_ranges(R, keys_sorted, ranges, stream);
    if (tile_order) launch_tile_order(tiles, ranges, tile_order, stream); }
  D2GS_STAGE("ranges", a->debug, stream);
  }
  { StageTimer t(ST_BLEND_F, stream);
    launch_blend_fwd(p, ranges, point_list, rec, final_T, n_contrib, a->out_color, a->out_others, g_cull,
                     deferred ? status : nullptr, tile_order, g_lane_walk & 1,
                     write_masks ? (uint32_t*)(bb + BL.hit_mask) : nullptr, stream); }
  D2GS_STAGE("blend", a->debug, stream);
  return D2GS_OK;
}

int d2gs_raster_backward(const D2gsRasterBwdArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  const int P = a->P, W = a->width, H = a->height;
  if (P == 0) return D2GS_OK;
  if (P < 0 || W <= 0 || H <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (!a->geom_buffer || !a->img_buffer || !a->binning_buffer || !a->grad_scratch || !a->radii ||
      !a->dL_dout_color || !a->dL_dout_others)
    return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  {  // the per-surfel kernels use 16-byte vector loads/stores on these tables
    const void* v16[] = {a->shs, a->sh_rest, a->rotations, a->d_rotations, a->dL_dsh, a->dL_dsh_rest, a->dL_drotations,
                         a->grad_scratch};
    for (const void* q : v16)
      if (q && ((uintptr_t)q & 15)) return fail(D2GS_ERR_INVALID_ARG, "SH / rotation tables and their gradients must be 16-byte aligned");
  }
  const GeomLayout GL = geom_layout(P);
  const ImgLayout IL = img_layout(W, H);
  const BinLayout BL = bin_layout(a->num_rendered);
  char* gb = aligned_base(a->geom_buffer);
  char* ib = aligned_base(a->img_buffer);
  char* bb = aligned_base(a->binning_buffer);
  const SurfelRec* rec = (const SurfelRec*)(gb + GL.rec);
  const uint8_t* clamped = (const uint8_t*)(gb + GL.clamped);
  const uint2* ranges = (const uint2*)(ib + IL.ranges);
  const float* final_T = (const float*)(ib + IL.final_T);
  const uint32_t* n_contrib = (const uint32_t*)(ib + IL.n_contrib);
  const uint32_t* point_list = (const uint32_t*)(bb + BL.point_list);

  BwdParams p{};
  p.P = P; p.D = a->D; p.M = a->M; p.W = W; p.H = H;
  p.bg = a->background; p.means3D = a->means3D; p.shs = a->shs; p.sh_rest = a->sh_rest;
  p.colors_precomp = a->colors_precomp; p.scales = a->scales; p.rotations = a->rotations;
  p.transMat_precomp = a->transMat_precomp; p.view = a->viewmatrix; p.proj = a->projmatrix; p.campos = a->campos;
  p.tan_fovx = a->tan_fovx; p.tan_fovy = a->tan_fovy;
  p.focal_y = H / (2.0f * a->tan_fovy);
  p.focal_x = W / (2.0f * a->tan_fovx);
  p.gx = (W + TILE_X - 1) / TILE_X; p.gy = (H + TILE_Y - 1) / TILE_Y;
  p.raw = a->raw_params; p.opacities = a->opacities; p.d_means3D = a->d_means3D; p.d_scales = a->d_scales;
  p.d_rotations = a->d_rotations;
  if (p.raw && (!a->opacities || !a->scales || !a->rotations))
    return fail(D2GS_ERR_INVALID_ARG, "raw-parameter mode needs opacities, scales and rotations");

  if (a->num_rendered > 0) {
    { StageTimer t(ST_BLEND_B, stream);
      launch_blend_bwd(p, ranges, point_list, rec, final_T, n_contrib, a->dL_dout_color, a->dL_dout_others,
                       a->grad_scratch, g_cull, g_tile_order ? (const uint32_t*)(ib + IL.tile_order) : nullptr,
                       (g_lane_walk >> 1) & 1, (g_lane_walk & 4) ? (const uint32_t*)(bb + BL.hit_mask) : nullptr,
                       (const uint32_t*)(gb + GL.status), stream); }
    D2GS_STAGE("blend_bwd", a->debug, stream);
  }
  { StageTimer t(ST_PRE_B, stream);
    launch_preprocess_bwd(p, rec, clamped, a->radii, a->grad_scratch, a->dL_dmeans2D, a->dL_dcolors, a->dL_dopacity,
                          a->dL_dmeans3D, a->dL_dtransMat, a->dL_dsh, a->dL_dsh_rest, a->dL_dscales, a->dL_drotations,
                          p.raw ? a->dL_dscales_raw : nullptr, stream); }
  D2GS_STAGE("preprocess_bwd", a->debug, stream);
  return D2GS_OK;
}

int d2gs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                      void* stream) {
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(D2GS_ERR_INVALID_ARG, "bad args");
  launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

__global__ void export_geom_kernel(int P, const SurfelRec* rec, const uint8_t* clamped, D2gsRasterState o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const SurfelRec r = rec[i];
  if (o.means2D) { o.means2D[2 * i] = r.q2.y; o.means2D[2 * i + 1] = r.q2.z; }
  if (o.depths) o.depths[i] = r.q3.w;
  if (o.transMat) {
    float* t = o.transMat + 9 * (size_t)i;
    t[0] = r.q0.x; t[1] = r.q0.y; t[2] = r.q0.z; t[3] = r.q0.w; t[4] = r.q1.x; t[5] = r.q1.y; t[6] = r.q1.z;
    t[7] = r.q1.w; t[8] = r.q2.x;
  }
  if (o.normal_opacity) {
    float* t = o.normal_opacity + 4 * (size_t)i;
    t[0] = r.q3.x; t[1] = r.q3.y; t[2] = r.q3.z; t[3] = r.q4.w;
  }
  if (o.rgb) { o.rgb[3 * i] = r.q4.x; o.rgb[3 * i + 1] = r.q4.y; o.rgb[3 * i + 2] = r.q4.z; }
  if (o.clamped) {
    const uint8_t c = clamped[i];
    o.clamped[3 * i] = c & 1; o.clamped[3 * i + 1] = (c >> 1) & 1; o.clamped[3 * i + 2] = (c >> 2) & 1;
  }
}

int d2gs_raster_export_state(int P, int width, int height, int64_t R, const void* geom_buffer,
                             const void* binning_buffer, const void* img_buffer, const D2gsRasterState* out,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!out || P <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad args");
  const GeomLayout GL = geom_layout(P);
  const ImgLayout IL = img_layout(width, height);
  const BinLayout BL = bin_layout(R);
  const size_t HW = (size_t)width * height;
  const size_t tiles = (size_t)((width + TILE_X - 1) / TILE_X) * ((height + TILE_Y - 1) / TILE_Y);
  const bool tile_sort = g_tile_sort && tiles <= (size_t)TILE_SORT_MAX_TILES;
  if (geom_buffer) {
    char* gb = aligned_base(geom_buffer);
    export_geom_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, (const SurfelRec*)(gb + GL.rec),
                                                           (const uint8_t*)(gb + GL.clamped), *out);
    if (out->tiles_touched)
      D2GS_CUDA_OK(cudaMemcpyAsync(out->tiles_touched, gb + GL.tiles_touched, 4 * (size_t)P, cudaMemcpyDeviceToDevice, stream));
    if (out->point_offsets) {
      if (tile_sort) {   // the per-tile path never needs the per-surfel scan: produce it for the caller
        size_t tmp = GL.scan_temp_bytes;
        D2GS_CUDA_OK(cub::DeviceScan::InclusiveSum(gb + GL.scan_temp, tmp, (const uint32_t*)(gb + GL.tiles_touched), out->point_offsets, P, stream));
      } else {
        D2GS_CUDA_OK(cudaMemcpyAsync(out->point_offsets, gb + GL.point_offsets, 4 * (size_t)P, cudaMemcpyDeviceToDevice, stream));
      }
    }
  }
  if (binning_buffer && R > 0 && tile_sort && img_buffer) {
    // per-tile keys are (depth << 32 | id): hand out the reference's format (tile << 32 | depth) for the parity tests
    char* bb = aligned_base(binning_buffer);
    const uint2* rg = (const uint2*)(aligned_base(img_buffer) + IL.ranges);
    if (out->keys_unsorted || out->values_unsorted)
      launch_tile_export_keys((uint32_t)tiles, rg, (const uint64_t*)(bb + BL.keys_unsorted), out->keys_unsorted, out->values_unsorted, stream);
    if (out->keys_sorted) launch_tile_export_keys((uint32_t)tiles, rg, (const uint64_t*)(bb + BL.keys_sorted), out->keys_sorted, nullptr, stream);
    if (out->point_list) D2GS_CUDA_OK(cudaMemcpyAsync(out->point_list, bb + BL.point_list, 4 * (size_t)R, cudaMemcpyDeviceToDevice, stream));
  } else if (binning_buffer && R > 0) {
    char* bb = aligned_base(binning_buffer);
    if (out->keys_unsorted) D2GS_CUDA_OK(cudaMemcpyAsync(out->keys_unsorted, bb + BL.keys_unsorted, 8 * (size_t)R, cudaMemcpyDeviceToDevice, stream));
    if (out->keys_sorted) D2GS_CUDA_OK(cudaMemcpyAsync(out->keys_sorted, bb + BL.keys_sorted, 8 * (size_t)R, cudaMemcpyDeviceToDevice, stream));
    if (out->values_unsorted) D2GS_CUDA_OK(cudaMemcpyAsync(out->values_unsorted, bb + BL.vals_unsorted, 4 * (size_t)R, cudaMemcpyDeviceToDevice, stream));
    if (out->point_list) D2GS_CUDA_OK(cudaMemcpyAsync(out->point_list, bb + BL.point_list, 4 * (size_t)R, cudaMemcpyDeviceToDevice, stream));
  }
  if (img_buffer) {
    char* ib = aligned_base(img_buffer);
    if (out->ranges) D2GS_CUDA_OK(cudaMemcpyAsync(out->ranges, ib + IL.ranges, 8 * tiles, cudaMemcpyDeviceToDevice, stream));
    if (out->final_T) D2GS_CUDA_OK(cudaMemcpyAsync(out->final_T, ib + IL.final_T, 4 * 3 * HW, cudaMemcpyDeviceToDevice, stream));
    if (out->n_contrib) D2GS_CUDA_OK(cudaMemcpyAsync(out->n_contrib, ib + IL.n_contrib, 4 * 2 * HW, cudaMemcpyDeviceToDevice, stream));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

int d2gs_deform_forward(const D2gsDeformFwdArgs* a, void* stream) {
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  if (a->P < 0 || a->M <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (a->P > 0 && (!a->xyz || !a->nodes || !a->node_radius_log || !a->node_trans || !a->node_rot || !a->node_scale ||
                   !a->d_xyz || !a->d_rotation || !a->d_scaling))
    return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  DeformFwdHost h{};
  h.P = a->P; h.M = a->M; h.K = a->K; h.hyper = a->hyper_dim;
  h.xyz = a->xyz; h.feature = a->feature; h.fstride = a->feature_stride;
  h.nodes = a->nodes; h.radius_log = a->node_radius_log; h.weight_logit = a->node_weight_logit;
  h.trans = a->node_trans; h.rot = a->node_rot; h.scale = a->node_scale; h.local_rot = a->node_local_rot;
  h.mask = a->motion_mask; h.nn_idx = a->nn_idx; h.nn_dist = a->nn_dist; h.nn_weight = a->nn_weight;
  h.d_xyz = a->d_xyz; h.d_rot = a->d_rotation; h.d_scale = a->d_scaling;
  h.attr_stride = a->node_attr_stride; h.order = a->order;
  const char* err = nullptr;
  { StageTimer t(ST_DEF_F, (cudaStream_t)stream);
    if (deform_forward_launch(h, (cudaStream_t)stream, &err) != 0) return fail(D2GS_ERR_INVALID_ARG, err); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

namespace {
struct OrderLayout { size_t bbox, keys, keys_out, vals, temp, temp_bytes, total; };
OrderLayout order_layout(int P) {
  OrderLayout L{};
  const size_t n = (size_t)(P > 0 ? P : 1);
  size_t o = 0;
  L.bbox = o; o = align_up(o + 6 * sizeof(unsigned int));
  L.keys = o; o = align_up(o + 4 * n);
  L.keys_out = o; o = align_up(o + 4 * n);
  L.vals = o; o = align_up(o + 4 * n);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned int*)nullptr, (unsigned int*)nullptr, (int*)nullptr, (int*)nullptr, (int)n, 0, 30);
  L.temp = o; L.temp_bytes = tmp; o = align_up(o + tmp);
  L.total = o + 256;
  return L;
}
}  // namespace

int d2gs_deform_order_workspace(int P, size_t* bytes) {
  if (!bytes || P < 0) return fail(D2GS_ERR_INVALID_ARG, "bad arguments");
  *bytes = order_layout(P).total;
  return D2GS_OK;
}

int d2gs_deform_order(int P, const float* xyz, int32_t* order, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (P == 0) return D2GS_OK;
  if (!xyz || !order || !workspace) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  const OrderLayout L = order_layout(P);
  if (workspace_bytes < L.total) return fail(D2GS_ERR_WORKSPACE, "order workspace too small");
  char* ws = aligned_base(workspace);
  unsigned int* keys = (unsigned int*)(ws + L.keys);
  unsigned int* keys_out = (unsigned int*)(ws + L.keys_out);
  int* vals = (int*)(ws + L.vals);
  deform_order_keys_launch(P, xyz, (unsigned int*)(ws + L.bbox), keys, vals, stream);
  size_t tmp = L.temp_bytes;
  D2GS_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + L.temp, tmp, keys, keys_out, vals, (int*)order, P, 0, 30, stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

namespace {
struct KnnLayout { OrderLayout ord; size_t order, sp, leaf, group, total; };
KnnLayout knn_layout(int P) {
  KnnLayout L{};
  L.ord = order_layout(P);
  const size_t n = (size_t)(P > 0 ? P : 1);
  const size_t nleaf = (n + 31) / 32, ngroup = (nleaf + 31) / 32;
  size_t o = align_up(L.ord.total);
  L.order = o; o = align_up(o + 4 * n);
  L.sp = o; o = align_up(o + 16 * 32 * nleaf);
  L.leaf = o; o = align_up(o + 32 * nleaf);
  L.group = o; o = align_up(o + 32 * ngroup);
  L.total = o + 256;
  return L;
}
}  // namespace

int d2gs_knn_mean_dist2_workspace(int P, size_t* bytes) {
  if (!bytes || P < 0) return fail(D2GS_ERR_INVALID_ARG, "bad arguments");
  *bytes = knn_layout(P).total;
  return D2GS_OK;
}

int d2gs_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (P == 0) return D2GS_OK;
  if (!points || !mean_dist2 || !workspace) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  const KnnLayout L = knn_layout(P);
  if (workspace_bytes < L.total) return fail(D2GS_ERR_WORKSPACE, "knn workspace too small");
  char* ws = aligned_base(workspace);
  unsigned int* keys = (unsigned int*)(ws + L.ord.keys);
  unsigned int* keys_out = (unsigned int*)(ws + L.ord.keys_out);
  int* vals = (int*)(ws + L.ord.vals);
  int* order = (int*)(ws + L.order);
  deform_order_keys_launch(P, points, (unsigned int*)(ws + L.ord.bbox), keys, vals, stream);
  size_t tmp = L.ord.temp_bytes;
  D2GS_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + L.ord.temp, tmp, keys, keys_out, vals, order, P, 0, 30, stream));
  knn_mean_dist2_launch(P, points, order, (float4*)(ws + L.sp), (float4*)(ws + L.leaf), (float4*)(ws + L.group), mean_dist2, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

int d2gs_deform_backward(const D2gsDeformBwdArgs* a, void* stream) {
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  if (a->P < 0 || a->M <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (a->P > 0 && (!a->xyz || !a->nodes || !a->nn_idx || !a->nn_dist || !a->nn_weight || !a->dL_d_xyz ||
                   !a->dL_d_rotation || !a->dL_d_scaling || !a->dL_dnode_trans || !a->dL_dnode_rot ||
                   !a->dL_dnode_scale || !a->dL_dnodes || !a->dL_dnode_radius_log))
    return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  DeformBwdHost h{};
  h.P = a->P; h.M = a->M; h.K = a->K; h.hyper = a->hyper_dim;
  h.xyz = a->xyz; h.feature = a->feature; h.fstride = a->feature_stride;
  h.nodes = a->nodes; h.radius_log = a->node_radius_log; h.weight_logit = a->node_weight_logit;
  h.trans = a->node_trans; h.rot = a->node_rot; h.scale = a->node_scale; h.local_rot = a->node_local_rot;
  h.mask = a->motion_mask; h.nn_idx = a->nn_idx; h.nn_dist = a->nn_dist; h.nn_weight = a->nn_weight;
  h.g_xyz = a->dL_d_xyz; h.g_rot = a->dL_d_rotation; h.g_scale = a->dL_d_scaling;
  h.d_trans = a->dL_dnode_trans; h.d_rot = a->dL_dnode_rot; h.d_scale = a->dL_dnode_scale;
  h.d_local_rot = a->dL_dnode_local_rot; h.d_nodes = a->dL_dnodes; h.d_radius_log = a->dL_dnode_radius_log;
  h.d_weight_logit = a->dL_dnode_weight_logit; h.d_feature = a->dL_dfeature; h.d_mask = a->dL_dmotion_mask;
  h.attr_stride = a->node_attr_stride; h.order = a->order;
  const char* err = nullptr;
  { StageTimer t(ST_DEF_B, (cudaStream_t)stream);
    if (deform_backward_launch(h, (cudaStream_t)stream, &err) != 0) return fail(D2GS_ERR_INVALID_ARG, err); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

int d2gs_epilogue_forward(const D2gsEpilogueArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || a->width <= 0 || a->height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad args");
  if (!a->allmap || !a->viewmatrix || !a->alpha || !a->rend_normal || !a->rend_dist || !a->depth || !a->surf_normal || !a->surf_point)
    return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  { StageTimer t(ST_EPI_F, stream);
    launch_epilogue_fwd(a->width, a->height, a->allmap, a->viewmatrix, a->focal_x, a->focal_y, a->alpha, a->rend_normal,
                        a->rend_dist, a->depth, a->surf_normal, a->surf_point, stream); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

int d2gs_epilogue_backward(const D2gsEpilogueArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || a->width <= 0 || a->height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad args");
  if (!a->allmap || !a->viewmatrix || !a->dL_dallmap) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  { StageTimer t(ST_EPI_B, stream);
    launch_epilogue_bwd(a->width, a->height, a->allmap, a->viewmatrix, a->focal_x, a->focal_y, a->g_alpha, a->g_rend_normal,
                        a->g_rend_dist, a->g_depth, a->g_surf_normal, a->g_surf_point, a->dL_dallmap, stream); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

}  // extern "C"

extern "C" {
int d2gs_adam_step(const D2gsAdamTensor* tensors, int count, void* stream_) {
  if (count < 0 || (count > 0 && !tensors)) return fail(D2GS_ERR_INVALID_ARG, "bad arguments");
  d2gs::AdamBatch B{};
  int acc = 0;
  auto flush = [&]() { if (B.count) d2gs::launch_adam(B, (cudaStream_t)stream_); B.count = 0; acc = 0; };
  for (int i = 0; i < count; i++) {
    const D2gsAdamTensor& t = tensors[i];
    if (t.numel == 0) continue;
    if (t.numel < 0 || !t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq) return fail(D2GS_ERR_INVALID_ARG, "bad tensor descriptor");
    if (!(t.bias_correction2_sqrt > 0.f)) return fail(D2GS_ERR_INVALID_ARG, "bias_correction2_sqrt must be positive (step >= 1)");
    d2gs::AdamTensor& k = B.t[B.count];
    k.param = t.param; k.grad = t.grad; k.exp_avg = t.exp_avg; k.exp_avg_sq = t.exp_avg_sq; k.numel = t.numel;
    k.w1 = (float)(1.0 - t.beta1); k.beta2 = (float)t.beta2; k.w2 = (float)(1.0 - t.beta2); k.eps = t.eps;
    k.neg_step_size = -t.step_size;
    k.inv_bc2_sqrt = 1.0f / t.bias_correction2_sqrt;
    k.vec4 = ((((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0) ? 1 : 0;
    acc += d2gs::adam_chunks(t.numel);
    B.chunk_end[B.count] = acc;
    if (++B.count == d2gs::ADAM_MAX_TENSORS) flush();
  }
  flush();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}
int d2gs_densification_stats(int P, const float* viewspace_grad, int grad_stride, const uint8_t* update_filter,
                             float* xyz_gradient_accum, float* denom, void* stream_) {
  if (P < 0 || grad_stride < 2) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (P == 0) return D2GS_OK;
  if (!viewspace_grad || !update_filter || !xyz_gradient_accum || !denom) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  d2gs::launch_densify_stats(P, viewspace_grad, grad_stride, update_filter, xyz_gradient_accum, denom, (cudaStream_t)stream_);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}
}  // extern "C"

namespace {
struct LossLayout { size_t sums, maps, total; };
LossLayout loss_layout(int W, int H) {
  LossLayout L{};
  size_t o = 0;
  L.sums = o; o = align_up(o + 4 * sizeof(float));
  L.maps = o; o = align_up(o + 9 * sizeof(float) * (size_t)W * H);
  L.total = o + 256;
  return L;
}
bool loss_args(const D2gsLossArgs* a, d2gs::LossArgs& k, bool backward) {
  const LossLayout L = loss_layout(a->width, a->height);
  if (!a->workspace || a->workspace_bytes < L.total) return false;
  char* ws = aligned_base(a->workspace);
  const size_t HW3 = 3 * (size_t)a->width * a->height;
  k.W = a->width; k.H = a->height;
  k.image = a->image; k.gt = a->gt; k.rend_normal = a->rend_normal; k.surf_normal = a->surf_normal; k.rend_dist = a->rend_dist;
  k.l_dssim = a->lambda_dssim; k.l_normal = a->lambda_normal; k.l_dist = a->lambda_dist;
  k.sums = (float*)(ws + L.sums); k.out = a->out;
  float* maps = (float*)(ws + L.maps);
  const bool keep = backward || a->save_for_backward;
  k.d_mu1 = keep ? maps : nullptr; k.d_e11 = keep ? maps + HW3 : nullptr; k.d_e12 = keep ? maps + 2 * HW3 : nullptr;
  k.upstream = a->upstream;
  k.g_image = a->g_image; k.g_rend_normal = a->g_rend_normal; k.g_surf_normal = a->g_surf_normal; k.g_rend_dist = a->g_rend_dist;
  return true;
}
}  // namespace

extern "C" {
int d2gs_loss_workspace(int width, int height, size_t* bytes) {
  if (!bytes || width <= 0 || height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad arguments");
  *bytes = loss_layout(width, height).total;
  return D2GS_OK;
}
int d2gs_loss_forward(const D2gsLossArgs* a, void* stream_) {
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  if (a->width <= 0 || a->height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (!a->image || !a->gt || !a->out) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  d2gs::LossArgs k{};
  if (!loss_args(a, k, false)) return fail(D2GS_ERR_WORKSPACE, "loss workspace missing or too small");
  { StageTimer t(ST_LOSS_F, (cudaStream_t)stream_);
    d2gs::launch_loss_forward(k, (cudaStream_t)stream_); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}
int d2gs_loss_backward(const D2gsLossArgs* a, void* stream_) {
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  if (a->width <= 0 || a->height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (!a->image || !a->gt || !a->g_image) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  if ((a->g_rend_normal && !a->surf_normal) || (a->g_surf_normal && !a->rend_normal))
    return fail(D2GS_ERR_INVALID_ARG, "normal gradients need both normal maps");
  d2gs::LossArgs k{};
  if (!loss_args(a, k, true)) return fail(D2GS_ERR_WORKSPACE, "loss workspace missing or too small");
  { StageTimer t(ST_LOSS_B, (cudaStream_t)stream_);
    d2gs::launch_loss_backward(k, (cudaStream_t)stream_); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}
}  // extern "C"

namespace d2gs {
struct MlpPlan {
  MlpLayers L;
  int Et, Tt, in0, NH, has_timenet;
  size_t wt_floats, off_te, off_th, off_inp, off_h, off_G, off_Gt1, off_gt, total_bytes;
};
static bool mlp_plan(const D2gsMlpArgs* a, int rows, int is_blender, int NH, MlpPlan& P) {
  P.has_timenet = is_blender ? 1 : 0;
  P.Et = is_blender ? 13 : 21;
  P.Tt = is_blender ? 30 : P.Et;
  P.in0 = 63 + P.Tt;
  P.NH = NH;
  int c = 0;
  size_t off = 0;
  auto add = [&](const float* W, const float* b, int N, int K) {
    MlpLayer& l = P.L.layer[c++];
    l.W = W; l.b = b; l.N = N; l.K = K; l.NP = (N + 3) & ~3; l.wt_off = off;
    off += (size_t)K * l.NP;
  };
  if (is_blender) {
    add(a ? a->timenet0_w : nullptr, a ? a->timenet0_b : nullptr, 256, P.Et);
    add(a ? a->timenet2_w : nullptr, a ? a->timenet2_b : nullptr, 30, 256);
  }
  for (int l = 0; l < 8; l++) add(a ? a->linear_w[l] : nullptr, a ? a->linear_b[l] : nullptr, 256, l == 0 ? P.in0 : (l == 5 ? P.in0 + 256 : 256));
  add(a ? a->heads_w : nullptr, a ? a->heads_b : nullptr, NH, 256);
  P.L.count = c;
  P.wt_floats = off;
  size_t o = align_up(4 * off);
  const size_t R = (size_t)rows;
  P.off_te = o;  o = align_up(o + 4 * R * 32);
  P.off_th = o;  o = align_up(o + 4 * R * 256);
  P.off_inp = o; o = align_up(o + 4 * R * 96);
  P.off_h = o;   o = align_up(o + 4 * R * 256 * 8);
  P.off_G = o;   o = align_up(o + 4 * R * 256 * 8);
  P.off_Gt1 = o; o = align_up(o + 4 * R * 256);
  P.off_gt = o;  o = align_up(o + 4 * R * 32);
  P.total_bytes = o + 256;
  return NH >= 1 && NH <= 16 && rows >= 0;
}
}  // namespace d2gs

extern "C" {

int d2gs_mlp_workspace(int rows, int is_blender, int num_out, size_t* bytes) {
  MlpPlan P;
  if (!bytes || !mlp_plan(nullptr, rows, is_blender, num_out, P)) return fail(D2GS_ERR_INVALID_ARG, "bad mlp sizes");
  *bytes = P.total_bytes;
  return D2GS_OK;
}

const float* d2gs_mlp_hidden(int rows, int is_blender, int num_out, const void* workspace) {
  MlpPlan P;
  if (!workspace || !mlp_plan(nullptr, rows, is_blender, num_out, P)) return nullptr;
  return (const float*)(aligned_base(workspace) + P.off_h) + (size_t)7 * rows * 256;
}

int d2gs_mlp_forward(const D2gsMlpArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  MlpPlan P;
  if (!mlp_plan(a, a->rows, a->is_blender, a->num_out, P)) return fail(D2GS_ERR_INVALID_ARG, "bad mlp sizes");
  if (a->rows == 0) return D2GS_OK;
  if (!a->x || !a->t || !a->out || !a->heads_w || !a->workspace) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  for (int i = 0; i < P.L.count; i++)
    if (!P.L.layer[i].W || !P.L.layer[i].b) return fail(D2GS_ERR_INVALID_ARG, "missing weights");
  for (int i = 0; i < P.L.count; i++)   // the backward streams W with 16-byte-granular bulk copies
    if ((uintptr_t)P.L.layer[i].W & 15) return fail(D2GS_ERR_INVALID_ARG, "weight matrices must be 16-byte aligned");
  if (a->workspace_bytes < P.total_bytes) return fail(D2GS_ERR_WORKSPACE, "mlp workspace too small");
  char* ws = aligned_base(a->workspace);
  P.L.wt = (float*)ws;
  MlpFwd f{};
  f.layers = P.L; f.rows = a->rows; f.Et = P.Et; f.Tt = P.Tt; f.NH = P.NH; f.has_timenet = P.has_timenet;
  f.t_stride = a->t_stride; f.x = a->x; f.t = a->t; f.out = a->out;
  f.save_inp = (float*)(ws + P.off_inp); f.save_te = (float*)(ws + P.off_te); f.save_th = (float*)(ws + P.off_th);
  f.save_h = (float*)(ws + P.off_h);
  { StageTimer t(ST_MLP_F, stream);
    mlp_launch_transpose(P.L, stream);
    mlp_launch_forward(f, stream); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

int d2gs_mlp_backward(const D2gsMlpArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  MlpPlan P;
  if (!mlp_plan(a, a->rows, a->is_blender, a->num_out, P)) return fail(D2GS_ERR_INVALID_ARG, "bad mlp sizes");
  if (a->rows == 0) return D2GS_OK;
  if (!a->g_out || !a->workspace || !a->g_heads_w || !a->g_heads_b) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  if (a->workspace_bytes < P.total_bytes) return fail(D2GS_ERR_WORKSPACE, "mlp workspace too small");
  char* ws = aligned_base(a->workspace);
  P.L.wt = (float*)ws;
  const int rows = a->rows;
  float* save_inp = (float*)(ws + P.off_inp); float* save_te = (float*)(ws + P.off_te);
  float* save_th = (float*)(ws + P.off_th); float* save_h = (float*)(ws + P.off_h);
  float* G = (float*)(ws + P.off_G); float* Gt1 = (float*)(ws + P.off_Gt1); float* gt = (float*)(ws + P.off_gt);
  MlpBwd b{};
  b.layers = P.L; b.rows = rows; b.Et = P.Et; b.Tt = P.Tt; b.NH = P.NH; b.has_timenet = P.has_timenet;
  b.g_out = a->g_out; b.save_th = save_th; b.save_h = save_h; b.G = G; b.G_t1 = Gt1; b.g_tfeat = gt;
  MlpWJobs J{};
  int c = 0;
  auto job = [&](const float* Gp, int ldG, int N, const float* A, int ldA, int Ka, const float* B, int ldB, int Kb, float* dW,
                 float* db) {
    MlpWJob& j = J.job[c++];
    j.G = Gp; j.ldG = ldG; j.N = N; j.A = A; j.ldA = ldA; j.Ka = Ka; j.B = B; j.ldB = ldB; j.Kb = Kb; j.dW = dW; j.db = db;
    j.tiles = ((N + 63) / 64) * ((Ka + Kb + 63) / 64);
  };
  if (P.has_timenet) {
    if (!a->g_timenet0_w || !a->g_timenet2_w) return fail(D2GS_ERR_INVALID_ARG, "missing timenet gradient buffers");
    job(Gt1, 256, 256, save_te, 32, P.Et, nullptr, 0, 0, a->g_timenet0_w, a->g_timenet0_b);
    job(gt, 32, 30, save_th, 256, 256, nullptr, 0, 0, a->g_timenet2_w, a->g_timenet2_b);
  }
  for (int l = 0; l < 8; l++) {
    if (!a->g_linear_w[l]) return fail(D2GS_ERR_INVALID_ARG, "missing trunk gradient buffers");
    const float* Gl = G + (size_t)l * rows * 256;
    if (l == 0) job(Gl, 256, 256, save_inp, 96, P.in0, nullptr, 0, 0, a->g_linear_w[l], a->g_linear_b[l]);
    else if (l == 5) job(Gl, 256, 256, save_inp, 96, P.in0, save_h + (size_t)4 * rows * 256, 256, 256, a->g_linear_w[l], a->g_linear_b[l]);
    else job(Gl, 256, 256, save_h + (size_t)(l - 1) * rows * 256, 256, 256, nullptr, 0, 0, a->g_linear_w[l], a->g_linear_b[l]);
  }
  job(a->g_out, P.NH, P.NH, save_h + (size_t)7 * rows * 256, 256, 256, nullptr, 0, 0, a->g_heads_w, a->g_heads_b);
  J.count = c;
  { StageTimer t(ST_MLP_B, stream);
    mlp_launch_backward(b, J, stream); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(D2GS_ERR_CUDA, cudaGetErrorString(e));
  return D2GS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// 3-D Gaussian rasterizer with depth and alpha outputs (gs3d.cu)
// ------------------------------------------------------------------------------------------------------------
namespace {
struct G3GeomLayout { size_t rec, cov3D, clamped, tiles_touched, point_offsets, scan_temp, scan_temp_bytes, status, tile_box, total; };
G3GeomLayout g3_geom_layout(int P) {
  G3GeomLayout L{};
  const size_t n = (size_t)(P > 0 ? P : 1);
  size_t o = 0;
  L.rec = o; o = align_up(o + sizeof(G3Rec) * n);
  L.cov3D = o; o = align_up(o + 24 * n);
  L.clamped = o; o = align_up(o + n);
  L.tiles_touched = o; o = align_up(o + 4 * n);
  L.point_offsets = o; o = align_up(o + 4 * n);
  size_t tmp = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  L.scan_temp_bytes = tmp;
  L.scan_temp = o; o = align_up(o + tmp);
  L.status = o; o = align_up(o + 4 * sizeof(uint32_t));   // {R, overflow flag, number of long tiles} of the per-tile binning
  L.tile_box = o; o = align_up(o + 16 * n);
  L.total = o + 256;
  return L;
}
struct G3ImgLayout { size_t ranges, n_contrib, tile_count, seg_begin, big_list, total; };
G3ImgLayout g3_img_layout(int W, int H) {
  G3ImgLayout L{};
  const size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
  size_t o = 0;
  L.ranges = o; o = align_up(o + 8 * tiles);
  L.n_contrib = o; o = align_up(o + 4 * (size_t)W * H);
  L.tile_count = o; o = align_up(o + 4 * tiles);
  L.seg_begin = o; o = align_up(o + 4 * tiles);
  L.big_list = o; o = align_up(o + 4 * tiles);
  L.total = o + 256;
  return L;
}
G3Params g3_params(int P, int D, int M, int W, int H, const float* bg, const float* means3D, const float* shs,
                   const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                   const float* rotations, const float* cov3D_precomp, const float* view, const float* proj,
                   const float* campos, float tan_fovx, float tan_fovy, int prefiltered) {
  G3Params p{};
  p.P = P; p.D = D; p.M = M; p.W = W; p.H = H; p.bg = bg; p.means3D = means3D; p.shs = shs;
  p.colors_precomp = colors_precomp; p.opacities = opacities; p.scales = scales; p.scale_modifier = scale_modifier;
  p.rotations = rotations; p.cov3D_precomp = cov3D_precomp; p.view = view; p.proj = proj; p.campos = campos;
  p.tan_fovx = tan_fovx; p.tan_fovy = tan_fovy;
  p.focal_y = H / (2.0f * tan_fovy);
  p.focal_x = W / (2.0f * tan_fovx);
  p.prefiltered = prefiltered;
  p.gx = (W + TILE_X - 1) / TILE_X; p.gy = (H + TILE_Y - 1) / TILE_Y;
  return p;
}
}  // namespace

extern "C" {

int d2gs_gs3d_workspace(int P, int width, int height, int64_t num_rendered, size_t* geom_bytes, size_t* img_bytes,
                        size_t* binning_bytes) {
  if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (geom_bytes) *geom_bytes = g3_geom_layout(P).total;
  if (img_bytes) *img_bytes = g3_img_layout(width, height).total;
  if (binning_bytes) *binning_bytes = bin_layout(num_rendered).total;
  return D2GS_OK;
}

int d2gs_gs3d_forward(const D2gsGs3dFwdArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  if (a->P < 0 || a->width <= 0 || a->height <= 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (!a->out_color || !a->out_depth || !a->out_alpha || !a->num_rendered) return fail(D2GS_ERR_INVALID_ARG, "missing outputs");
  const int P = a->P, W = a->width, H = a->height;
  const size_t HW = (size_t)W * H;
  if (P == 0) {   // rasterize_points.cu:64-66,84: zero images for an empty scene
    D2GS_CUDA_OK(cudaMemsetAsync(a->out_color, 0, 4 * 3 * HW, stream));
    D2GS_CUDA_OK(cudaMemsetAsync(a->out_depth, 0, 4 * HW, stream));
    D2GS_CUDA_OK(cudaMemsetAsync(a->out_alpha, 0, 4 * HW, stream));
    *a->num_rendered = 0;
    return D2GS_OK;
  }
  if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->campos || !a->background || !a->radii)
    return fail(D2GS_ERR_INVALID_ARG, "missing inputs");
  if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
    return fail(D2GS_ERR_INVALID_ARG, "provide exactly one of SHs or precomputed colours");
  if (((a->scales == nullptr) || (a->rotations == nullptr)) == (a->cov3D_precomp == nullptr))
    return fail(D2GS_ERR_INVALID_ARG, "provide exactly one of scale/rotation pair or precomputed 3D covariance");
  if (a->shs && (a->M < 1 || a->M > 16 || a->D < 0 || a->D > 3 || (a->D + 1) * (a->D + 1) > a->M))
    return fail(D2GS_ERR_INVALID_ARG, "SH degree / coefficient count out of range (deg<=3, M<=16)");
  if (a->rotations && ((uintptr_t)a->rotations & 15)) return fail(D2GS_ERR_INVALID_ARG, "rotations must be 16-byte aligned");

  const G3GeomLayout GL = g3_geom_layout(P);
  const G3ImgLayout IL = g3_img_layout(W, H);
  if (a->geom_bytes < GL.total || !a->geom_buffer) return fail(D2GS_ERR_WORKSPACE, "geometry workspace too small");
  if (a->img_bytes < IL.total || !a->img_buffer) return fail(D2GS_ERR_WORKSPACE, "image workspace too small");
  char* gb = aligned_base(a->geom_buffer);
  char* ib = aligned_base(a->img_buffer);
  G3Rec* rec = (G3Rec*)(gb + GL.rec);
  float* cov3D = (float*)(gb + GL.cov3D);
  uint8_t* clamped = (uint8_t*)(gb + GL.clamped);
  uint32_t* tiles_touched = (uint32_t*)(gb + GL.tiles_touched);
  uint32_t* point_offsets = (uint32_t*)(gb + GL.point_offsets);
  uint2* ranges = (uint2*)(ib + IL.ranges);
  uint32_t* n_contrib = (uint32_t*)(ib + IL.n_contrib);
  const G3Params p = g3_params(P, a->D, a->M, W, H, a->background, a->means3D, a->shs, a->colors_precomp, a->opacities, a->scales,
                               a->scale_modifier, a->rotations, a->cov3D_precomp, a->viewmatrix, a->projmatrix, a->campos,
                               a->tan_fovx, a->tan_fovy, a->prefiltered);
  // Binning as in the surfel path (tile_binning.cu, option "tile_sort"): per-tile counters -> one-CTA scan (ranges, R,
  // overflow flag) -> atomic scatter of (depth | id) keys -> one sort per tile; the same per-tile lists as the reference's
  // global sort on (tile | depth) keys.  binning_capacity > 0 selects the deferred-count mode (no host readback).
  const uint32_t tiles = p.gx * p.gy;
  const bool tile_sort = g_tile_sort && tiles <= (uint32_t)TILE_SORT_MAX_TILES;
  const bool deferred = a->binning_capacity > 0;
  if (deferred && !tile_sort) return fail(D2GS_ERR_INVALID_ARG, "the deferred-count mode of gs3d needs the per-tile binning (tile_sort = 1)");
  if (deferred && a->binning_capacity > 0xffffffffll) return fail(D2GS_ERR_INVALID_ARG, "binning_capacity exceeds 2^32-1 instances");
  uint32_t* status = (uint32_t*)(gb + GL.status);
  uint4* tile_box = (uint4*)(gb + GL.tile_box);
  uint32_t* tile_count = (uint32_t*)(ib + IL.tile_count);
  uint32_t* seg_begin = (uint32_t*)(ib + IL.seg_begin);
  uint32_t* big_list = (uint32_t*)(ib + IL.big_list);
  if (!a->resume) {
    g3_launch_preprocess_fwd(p, rec, cov3D, clamped, a->radii, tiles_touched, tile_sort ? tile_box : nullptr, stream);
    D2GS_STAGE("gs3d preprocess", a->debug, stream);
    if (tile_sort) {
      launch_tile_count(P, tile_box, p.gx, p.gy, tile_count, stream);
      launch_tile_scan(tiles, deferred ? (uint32_t)a->binning_capacity : 0xffffffffu, tile_count, seg_begin, ranges, big_list, status,
                       nullptr, stream);
    } else {
      size_t tmp = GL.scan_temp_bytes;
      D2GS_CUDA_OK(cub::DeviceScan::InclusiveSum(gb + GL.scan_temp, tmp, tiles_touched, point_offsets, P, stream));
    }
    D2GS_STAGE("gs3d scan", a->debug, stream);
  }
  int64_t R = 0;
  if (deferred) {
    R = a->binning_capacity;
  } else {
    uint32_t R32 = 0;    // the reference's blocking readback (DGR/cuda_rasterizer/rasterizer_impl.cu:270-271)
    D2GS_CUDA_OK(cudaMemcpyAsync(&R32, tile_sort ? status : point_offsets + P - 1, 4, cudaMemcpyDeviceToHost, stream));
    D2GS_CUDA_OK(cudaStreamSynchronize(stream));
    R = R32;
  }
  *a->num_rendered = R;
  const BinLayout BL = bin_layout(R);
  if (a->binning_required) *a->binning_required = BL.total;
  if (a->binning_bytes < BL.total || !a->binning_buffer) {
    if (deferred) return fail(D2GS_ERR_WORKSPACE, "binning workspace smaller than binning_capacity instances need");
    return D2GS_NEED_BINNING;
  }
  char* bb = aligned_base(a->binning_buffer);
  uint64_t* keys_unsorted = (uint64_t*)(bb + BL.keys_unsorted);
  uint64_t* keys_sorted = (uint64_t*)(bb + BL.keys_sorted);
  uint32_t* vals_unsorted = (uint32_t*)(bb + BL.vals_unsorted);
  uint32_t* point_list = (uint32_t*)(bb + BL.point_list);
  if (tile_sort) {
    launch_tile_scatter(P, tile_box, p.gx, p.gy, seg_begin, tile_count, status, keys_unsorted, stream);
    D2GS_STAGE("gs3d scatter", a->debug, stream);
    if (deferred && a->num_rendered_async)
      D2GS_CUDA_OK(cudaMemcpyAsync(a->num_rendered_async, status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    if (R > 0) {
      launch_tile_sort(tiles, ranges, big_list, status, keys_unsorted, keys_sorted, point_list, stream);
      D2GS_STAGE("gs3d tile sort", a->debug, stream);
    }
  } else {
    g3_launch_duplicate(P, rec, a->radii, point_offsets, keys_unsorted, vals_unsorted, p.gx, p.gy, stream);
    D2GS_STAGE("gs3d duplicate", a->debug, stream);
    if (R > 0) {
      const int bit = (int)higher_msb(p.gx * p.gy);
      size_t tmp = BL.sort_temp_bytes;
      D2GS_CUDA_OK(cub::DeviceRadixSort::SortPairs(bb + BL.sort_temp, tmp, keys_unsorted, keys_sorted, vals_unsorted, point_list,
                                                   (int)R, 0, 32 + bit, stream));
      D2GS_STAGE("gs3d sort", a->debug, stream);
    }
    D2GS_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)p.gx * p.gy, stream));
    launch_ranges(R, keys_sorted, ranges, stream);
    D2GS_STAGE("gs3d ranges", a->debug, stream);
  }
  g3_launch_blend_fwd(p, ranges, point_list, rec, a->out_color, a->out_depth, a->out_alpha, n_contrib, deferred ? status : nullptr, stream);
  D2GS_STAGE("gs3d blend", a->debug, stream);
  return D2GS_OK;
}

int d2gs_gs3d_backward(const D2gsGs3dBwdArgs* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a) return fail(D2GS_ERR_INVALID_ARG, "null args");
  const int P = a->P, W = a->width, H = a->height;
  if (P == 0) return D2GS_OK;
  if (P < 0 || W <= 0 || H <= 0 || a->num_rendered < 0) return fail(D2GS_ERR_INVALID_ARG, "bad sizes");
  if (!a->geom_buffer || !a->img_buffer || !a->binning_buffer || !a->grad_scratch || !a->radii || !a->out_alpha ||
      !a->dL_dout_color || !a->dL_dout_depth || !a->dL_dout_alpha || !a->means3D || !a->viewmatrix || !a->projmatrix ||
      !a->campos || !a->background)
    return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  const G3GeomLayout GL = g3_geom_layout(P);
  const G3ImgLayout IL = g3_img_layout(W, H);
  const BinLayout BL = bin_layout(a->num_rendered);
  char* gb = aligned_base(a->geom_buffer);
  char* ib = aligned_base(a->img_buffer);
  char* bb = aligned_base(a->binning_buffer);
  const G3Rec* rec = (const G3Rec*)(gb + GL.rec);
  const G3Params p = g3_params(P, a->D, a->M, W, H, a->background, a->means3D, a->shs, a->colors_precomp, nullptr, a->scales,
                               a->scale_modifier, a->rotations, a->cov3D_precomp, a->viewmatrix, a->projmatrix, a->campos,
                               a->tan_fovx, a->tan_fovy, 0);
  D2GS_CUDA_OK(cudaMemsetAsync(a->grad_scratch, 0, sizeof(float) * G3_GRAD_FLOATS * (size_t)P, stream));
  g3_launch_blend_bwd(p, (const uint2*)(ib + IL.ranges), (const uint32_t*)(bb + BL.point_list), rec, a->out_alpha,
                      (const uint32_t*)(ib + IL.n_contrib), a->dL_dout_color, a->dL_dout_depth, a->dL_dout_alpha, a->grad_scratch,
                      stream);
  D2GS_STAGE("gs3d blend backward", a->debug, stream);
  g3_launch_preprocess_bwd(p, (const float*)(gb + GL.cov3D), (const uint8_t*)(gb + GL.clamped), a->radii, a->grad_scratch,
                           a->dL_dmeans2D, a->dL_dcolors, a->dL_dopacity, a->dL_dmeans3D, a->dL_dcov3D, a->dL_dsh, a->dL_dscales,
                           a->dL_drotations, stream);
  D2GS_STAGE("gs3d preprocess backward", a->debug, stream);
  return D2GS_OK;
}

/* intermediates for parity tests (what the reference keeps in its geometry / binning / image buffers) */
int d2gs_gs3d_export_state(int P, int width, int height, int64_t num_rendered, const void* geom_buffer, const void* binning_buffer,
                           const void* img_buffer, const D2gsGs3dState* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!out || !geom_buffer || !img_buffer) return fail(D2GS_ERR_INVALID_ARG, "missing buffers");
  if (P <= 0) return D2GS_OK;
  const G3GeomLayout GL = g3_geom_layout(P);
  const G3ImgLayout IL = g3_img_layout(width, height);
  char* gb = aligned_base(geom_buffer);
  char* ib = aligned_base(img_buffer);
  if (out->rec) D2GS_CUDA_OK(cudaMemcpyAsync(out->rec, gb + GL.rec, sizeof(G3Rec) * (size_t)P, cudaMemcpyDeviceToDevice, stream));
  if (out->cov3D) D2GS_CUDA_OK(cudaMemcpyAsync(out->cov3D, gb + GL.cov3D, 24 * (size_t)P, cudaMemcpyDeviceToDevice, stream));
  if (out->clamped) D2GS_CUDA_OK(cudaMemcpyAsync(out->clamped, gb + GL.clamped, (size_t)P, cudaMemcpyDeviceToDevice, stream));
  if (out->tiles_touched) D2GS_CUDA_OK(cudaMemcpyAsync(out->tiles_touched, gb + GL.tiles_touched, 4 * (size_t)P, cudaMemcpyDeviceToDevice, stream));
  if (out->n_contrib) D2GS_CUDA_OK(cudaMemcpyAsync(out->n_contrib, ib + IL.n_contrib, 4 * (size_t)width * height, cudaMemcpyDeviceToDevice, stream));
  const size_t tiles = (size_t)((width + TILE_X - 1) / TILE_X) * ((height + TILE_Y - 1) / TILE_Y);
  if (out->ranges) D2GS_CUDA_OK(cudaMemcpyAsync(out->ranges, ib + IL.ranges, 8 * tiles, cudaMemcpyDeviceToDevice, stream));
  if (binning_buffer && num_rendered > 0) {
    const BinLayout BL = bin_layout(num_rendered);
    char* bb = aligned_base(binning_buffer);
    const bool tile_sort = g_tile_sort && tiles <= (size_t)TILE_SORT_MAX_TILES;
    if (out->keys_sorted && tile_sort)     // per-tile keys are (depth << 32 | id): hand out the reference's (tile << 32 | depth)
      launch_tile_export_keys((uint32_t)tiles, (const uint2*)(ib + IL.ranges), (const uint64_t*)(bb + BL.keys_sorted), out->keys_sorted, nullptr, stream);
    else if (out->keys_sorted) D2GS_CUDA_OK(cudaMemcpyAsync(out->keys_sorted, bb + BL.keys_sorted, 8 * (size_t)num_rendered, cudaMemcpyDeviceToDevice, stream));
    if (out->point_list) D2GS_CUDA_OK(cudaMemcpyAsync(out->point_list, bb + BL.point_list, 4 * (size_t)num_rendered, cudaMemcpyDeviceToDevice, stream));
  }
  return D2GS_OK;
}

}  // extern "C"
