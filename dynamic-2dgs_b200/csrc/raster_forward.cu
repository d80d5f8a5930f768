// Forward kernels of the surfel rasterizer for sm_100a:
//   preprocess (per surfel) -> [scan] -> tile-instance emission -> [radix sort] -> tile ranges -> per-tile blend.
// Behavioural contract: DSR/cuda_rasterizer/forward.cu:166-260 (preprocess), :265-463 (blend),
// rasterizer_impl.cu:70-111 (instance emission), :116-138 (ranges).  Tile/sort indices must be bit-exact, so the
// float expressions that feed integer decisions are written in the same association order as the reference's.
#include "raster_common.cuh"
#include "sh_basis.cuh"

namespace d2gs {

// ------------------------------------------------------------------------------------------------------------
// preprocess
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_sh(const FwdParams& p, int idx, int ncoef, float* sh /*[48]*/) {
  if (p.sh_rest == nullptr) {
    const float* base = p.shs + (size_t)idx * p.M * 3;
    if (p.M == 16 && ((reinterpret_cast<uintptr_t>(p.shs) & 15) == 0)) {
      const float4* b4 = reinterpret_cast<const float4*>(base);   // 192 B per surfel, 16-B aligned
      const int n4 = (ncoef * 3 + 3) / 4;
#pragma unroll
      for (int i = 0; i < 12; i++) {
        if (i < n4) {
          float4 v = __ldg(b4 + i);
          sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 48; i++)
        if (i < ncoef * 3) sh[i] = __ldg(base + i);
    }
  } else {
    const float* dc = p.shs + (size_t)idx * 3;
    sh[0] = __ldg(dc); sh[1] = __ldg(dc + 1); sh[2] = __ldg(dc + 2);
    const float* rest = p.sh_rest + (size_t)idx * (p.M - 1) * 3;
#pragma unroll
    for (int i = 3; i < 48; i++)
      if (i < ncoef * 3) sh[i] = __ldg(rest + i - 3);
  }
}

constexpr int SHF_WARP_FLOATS = 32 * 45 + 32 * 3;   // per-warp staging of the split SH layout: rest block + DC block

__global__ void __launch_bounds__(256, 3) preprocess_fwd_kernel(FwdParams p, SurfelRec* __restrict__ rec,
                                                             uint8_t* __restrict__ clamped, int* __restrict__ radii,
                                                             uint32_t* __restrict__ tiles_touched, const bool staged) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  // Split SH layout (DC (P,1,3) + rest (P,15,3), the trainer's parameters): the 32 surfels of a warp own one contiguous
  // 5760-B block of `rest` and 384 B of DC.  The warp copies both to shared memory with 16-byte asynchronous copies
  // (LDGSTS: coalesced, no staging registers); each thread then reads its 48 coefficients at a 45-word stride
  // (conflict-free) instead of issuing 48 scalar global loads at a 180-B stride (lg_throttle 8.3 in profiles/ncu_r1i.md).
  // The wait sits right here, before any lane can leave the kernel: a lane that returned early would never publish its
  // copies.  (The projection code below is textually the round-1 kernel: its FMA contraction is part of the bit-exact
  // contract with the reference build, so it is not restructured around the copy.)
  extern __shared__ float s_shf[];
  const int lane = threadIdx.x & 31;
  float* wrest = s_shf + (threadIdx.x >> 5) * SHF_WARP_FLOATS;
  float* wdc = wrest + 32 * 45;
  const int wbase = idx - lane;
  const bool async_stage = staged && wbase + 32 <= p.P;
  if (async_stage) {
    const float4* src = reinterpret_cast<const float4*>(p.sh_rest + (size_t)wbase * 45);
    const uint32_t d0 = smem_addr(wrest);
    for (int v = lane; v < 360; v += 32) cp_async16(d0 + 16u * (uint32_t)v, src + v);
    if (lane < 24) cp_async16(smem_addr(wdc) + 16u * (uint32_t)lane, reinterpret_cast<const float4*>(p.shs + (size_t)wbase * 3) + lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
  }
  if (idx == 0 && p.frame_flag) *p.frame_flag = p.frame_flag_value;
  if (idx >= p.P) return;
  radii[idx] = 0;
  tiles_touched[idx] = 0;
  if (p.tile_box) p.tile_box[idx] = make_uint4(0u, 0u, 0u, 0u);

  const float* m = p.view;
  Activated act;
  if (p.raw) act = activate_surfel(idx, p.means3D, p.d_means3D, p.scales, p.d_scales, p.rotations, p.d_rotations, p.opacities);
  const v3 pw = p.raw ? act.pw : v3{p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
  // near-plane cull (the only frustum test the reference keeps active)
  const float depth = m[2] * pw.x + m[6] * pw.y + m[10] * pw.z + m[14];
  if (depth <= 0.2f) {
    if (p.prefiltered) {
      printf("Point is filtered although prefiltered is set. This shouldn't happen!");
      __trap();
    }
    return;
  }

  float T[9];
  v3 normal = {0.f, 0.f, 0.f};   // transMat_precomp path: the reference leaves this undefined; we define 0
  if (p.transMat_precomp != nullptr) {
#pragma unroll
    for (int i = 0; i < 9; i++) T[i] = p.transMat_precomp[9 * (size_t)idx + i];
  } else {
    const float cx = float(p.W) / 2.0f, cy = float(p.H) / 2.0f;
    const m3 Wm = view_rot(m);
    const v3 cam = {m[12], m[13], m[14]};
    const v3 p_view = Wm * pw + cam;
    const float2 sc = p.raw ? act.sc : reinterpret_cast<const float2*>(p.scales)[idx];
    const float4 q = p.raw ? act.q : reinterpret_cast<const float4*>(p.rotations)[idx];
    const m3 R = quat_to_rot(q);
    const v3 M0 = Wm * (R.c0 * sc.x);
    const v3 M1 = Wm * (R.c1 * sc.y);
    const v3 M2 = p_view;
    v3 tn = Wm * R.c2;
    const float cs = dot3(-tn, p_view);
    if (cs == 0.0f) return;   // exactly edge-on
    const float mult = cs > 0 ? 1.f : -1.f;
    tn = tn * mult;
    // row k of the 3x3 homography: (fx*M_k.x + cx*M_k.z, fy*M_k.y + cy*M_k.z, M_k.z); the association
    // order (product of focal first, then fused add of the principal-point term) is part of the contract.
    T[0] = __fmaf_rn(cx, M0.z, __fmul_rn(p.focal_x, M0.x));
    T[1] = __fmaf_rn(cx, M1.z, __fmul_rn(p.focal_x, M1.x));
    T[2] = __fmaf_rn(cx, M2.z, __fmul_rn(p.focal_x, M2.x));
    T[3] = __fmaf_rn(cy, M0.z, __fmul_rn(p.focal_y, M0.y));
    T[4] = __fmaf_rn(cy, M1.z, __fmul_rn(p.focal_y, M1.y));
    T[5] = __fmaf_rn(cy, M2.z, __fmul_rn(p.focal_y, M2.y));
    T[6] = M0.z; T[7] = M1.z; T[8] = M2.z;
    normal = tn;
  }

  // screen-space bounding box of the 1-sigma ellipse (forward.cu:133-163)
  const v3 Tu = {T[0], T[1], T[2]}, Tv = {T[3], T[4], T[5]}, Tw = {T[6], T[7], T[8]};
  const v3 sgn = {1.0f, 1.0f, -1.0f};
  const float d = dot3(sgn, Tw * Tw);
  if (d == 0.0f) return;
  const v3 f = sgn * (1.0f / d);
  const v3 pc = {dot3(f, Tu * Tw), dot3(f, Tv * Tw), dot3(f, Tw * Tw)};
  const v3 h0 = pc * pc - v3{dot3(f, Tu * Tu), dot3(f, Tv * Tv), dot3(f, Tw * Tw)};
  const float ex = sqrtf(fmaxf(0.0f, h0.x)), ey = sqrtf(fmaxf(0.0f, h0.y));
  // radius = ceil(3 * max(extent, FilterSize)) evaluated in double because FilterSize is a double literal
  const float radius = (float)ceil(3.f * fmax((double)fmaxf(ex, ey), D2GS_FILTER_SIZE));

  const RectU r = tile_rect(pc.x, pc.y, (int)radius, p.gx, p.gy);
  if ((r.x1 - r.x0) * (r.y1 - r.y0) == 0) return;

  v3 rgb = {0.f, 0.f, 0.f};
  if (p.colors_precomp == nullptr) {
    const v3 campos = {p.campos[0], p.campos[1], p.campos[2]};
    v3 dir = pw - campos;
    dir = dir / sqrtf(dot3(dir, dir));
    float sh[48];
    const int ncoef = (p.D + 1) * (p.D + 1);
    if (async_stage) {
#pragma unroll
      for (int i = 0; i < 48; i++)
        if (i < ncoef * 3) sh[i] = (i < 3) ? wdc[lane * 3 + i] : wrest[lane * 45 + (i - 3)];
    } else {
      load_sh(p, idx, ncoef, sh);
    }
    v3 res = eval_sh(p.D, dir, sh);
    res.x += 0.5f; res.y += 0.5f; res.z += 0.5f;
    clamped[idx] = (uint8_t)((res.x < 0 ? 1 : 0) | (res.y < 0 ? 2 : 0) | (res.z < 0 ? 4 : 0));
    rgb = {fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f)};
  } else {
    rgb = {p.colors_precomp[3 * (size_t)idx], p.colors_precomp[3 * (size_t)idx + 1], p.colors_precomp[3 * (size_t)idx + 2]};
  }

  radii[idx] = (int)radius;
  tiles_touched[idx] = (r.y1 - r.y0) * (r.x1 - r.x0);
  // packed tile rectangle + depth bits: all the per-tile binning needs of this surfel (16 B instead of two quads of the record)
  if (p.tile_box) p.tile_box[idx] = make_uint4(r.x0 | (r.x1 << 16), r.y0 | (r.y1 << 16), __float_as_uint(depth), 0u);
  SurfelRec o;
  o.q0 = make_float4(T[0], T[1], T[2], T[3]);
  o.q1 = make_float4(T[4], T[5], T[6], T[7]);
  const float opacity = p.raw ? act.opacity : p.opacities[idx];
  o.q2 = make_float4(T[8], pc.x, pc.y, 2.0f * logf(255.0f * opacity) + 1e-4f);
  o.q3 = make_float4(normal.x, normal.y, normal.z, depth);
  o.q4 = make_float4(rgb.x, rgb.y, rgb.z, opacity);
  o.q5 = cull_box(T, pc.x, pc.y, o.q2.w);
  rec[idx] = o;
}

void launch_preprocess_fwd(const FwdParams& p, SurfelRec* rec, uint8_t* clamped, int* radii, uint32_t* tiles_touched,
                           cudaStream_t s) {
  if (p.P == 0) return;
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool staged = p.colors_precomp == nullptr && p.shs && p.sh_rest && p.M == 16 && al16(p.sh_rest) && al16(p.shs);
  const size_t smem = staged ? sizeof(float) * SHF_WARP_FLOATS * 8 : 0;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(preprocess_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * SHF_WARP_FLOATS * 8));
    configured = true;
  }
  preprocess_fwd_kernel<<<(p.P + 255) / 256, 256, smem, s>>>(p, rec, clamped, radii, tiles_touched, staged);
}

// ------------------------------------------------------------------------------------------------------------
// (surfel, tile) instance emission: key = tile id << 32 | depth bits, value = surfel id, row-major over the rect
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) duplicate_kernel(int P, const SurfelRec* __restrict__ rec,
                                                        const int* __restrict__ radii,
                                                        const uint32_t* __restrict__ offsets,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        uint32_t gx, uint32_t gy, uint32_t capacity) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const int rad = radii[idx];
  if (rad <= 0) return;
  uint32_t off = (idx == 0) ? 0 : offsets[idx - 1];
  if (offsets[idx] > capacity) return;   // only possible in deferred-count mode; pad_keys flags the overflow
  const float4 q2 = __ldg(&rec[idx].q2);
  const float depth = __ldg(&rec[idx].q3.w);
  const RectU r = tile_rect(q2.y, q2.z, rad, gx, gy);
  const uint64_t dbits = (uint64_t)__float_as_uint(depth);
  for (uint32_t y = r.y0; y < r.y1; y++) {
    for (uint32_t x = r.x0; x < r.x1; x++) {
      keys[off] = ((uint64_t)(y * gx + x) << 32) | dbits;
      vals[off] = (uint32_t)idx;
      off++;
    }
  }
}

void launch_duplicate(int P, const SurfelRec* rec, const int* radii, const uint32_t* offsets, uint64_t* keys,
                      uint32_t* vals, uint32_t gx, uint32_t gy, uint32_t capacity, cudaStream_t s) {
  if (P == 0) return;
  duplicate_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, rec, radii, offsets, keys, vals, gx, gy, capacity);
}

__global__ void __launch_bounds__(256) pad_keys_kernel(uint32_t capacity, const uint32_t* __restrict__ total,
                                                       uint64_t* __restrict__ keys, uint32_t* __restrict__ status) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t R = *total;
  if (idx == 0) { status[0] = R; status[1] = R > capacity ? 1u : 0u; }
  // (on overflow the frame renders nothing and is poisoned — ranges_deferred_kernel / blend_fwd_kernel — so the
  // stale keys the sort then sees are never used)
  if (idx < capacity && idx >= R) keys[idx] = ~0ull;
}

void launch_pad_keys(uint32_t capacity, const uint32_t* total, uint64_t* keys, uint32_t* status, cudaStream_t s) {
  pad_keys_kernel<<<(capacity + 255) / 256 + (capacity == 0), 256, 0, s>>>(capacity, total, keys, status);
}

__global__ void __launch_bounds__(256) ranges_kernel(int64_t L, const uint64_t* __restrict__ keys,
                                                     uint2* __restrict__ ranges) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> 32);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == L - 1) ranges[cur].y = (uint32_t)L;
}

void launch_ranges(int64_t R, const uint64_t* keys_sorted, uint2* ranges, cudaStream_t s) {
  if (R == 0) return;
  ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(R, keys_sorted, ranges);
}

__global__ void __launch_bounds__(256) ranges_deferred_kernel(uint32_t capacity, const uint32_t* __restrict__ status,
                                                              const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t L = status[1] ? 0u : min(status[0], capacity);   // an overflowed frame renders nothing (and is poisoned)
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> 32);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = idx;
      ranges[cur].x = idx;
    }
  }
  if (idx == L - 1) ranges[cur].y = L;
}

void launch_ranges_deferred(uint32_t capacity, const uint32_t* status, const uint64_t* keys_sorted, uint2* ranges, cudaStream_t s) {
  if (capacity == 0) return;
  ranges_deferred_kernel<<<(capacity + 255) / 256, 256, 0, s>>>(capacity, status, keys_sorted, ranges);
}

// ------------------------------------------------------------------------------------------------------------
// per-tile front-to-back blend.  One CTA per 16x16 tile, one thread per pixel; each warp owns a compact
// 8x4 pixel patch.  Instances are staged 256 at a time into shared memory as five float4 planes
// (broadcast LDS.128 in the inner loop).
// ------------------------------------------------------------------------------------------------------------
#ifndef D2GS_FWD_BATCH
#define D2GS_FWD_BATCH 128
#endif
constexpr int BLEND_BATCH = D2GS_FWD_BATCH;   // instances per staged batch (double buffered)
// Two CTAs of 4 warps per 16x16 tile (rows 0-7 / 8-15, blockIdx.z): barriers wait for 4 patches instead of 8, a half
// tile whose pixels are all saturated stops on its own, and eight small CTAs per SM interleave.
#ifndef D2GS_FWD_WARPS
#define D2GS_FWD_WARPS 4
#endif
constexpr int FWD_NWARP = D2GS_FWD_WARPS;        // warps (8x4 pixel patches) per CTA: 4, 2 or 1
constexpr int FWD_THREADS = 32 * FWD_NWARP;
constexpr int FWD_Z = 8 / FWD_NWARP;             // CTAs per 16x16 tile (blockIdx.z)
constexpr int FWD_SPT = (D2GS_FWD_BATCH + FWD_THREADS - 1) / FWD_THREADS;   // staging slots per thread

__device__ __forceinline__ void pixel_of_thread(int tid, int& lx, int& ly) {
  const int w = tid >> 5, l = tid & 31;
  lx = ((w & 1) << 3) | (l & 7);
  ly = ((w >> 1) << 2) | (l >> 3);
}

// One CTA per 16x16 tile, one thread per pixel, each warp a compact 8x4 patch.  Per staged instance a warp first tests
// the instance's cull box against its patch (uniform branch, one broadcast LDS.128), then each lane runs the exact
// prefilter (q0..q2) and only survivors pay for the divisions / exp / accumulation.
#ifndef D2GS_FWD_MINBLOCKS
#define D2GS_FWD_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(FWD_THREADS, D2GS_FWD_MINBLOCKS) blend_fwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const SurfelRec* __restrict__ rec, const float* __restrict__ bg, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_others, int cull,
    const uint32_t* __restrict__ status, const uint32_t* __restrict__ tile_order, uint32_t gx) {
  __shared__ float4 s_q[2][REC_QUADS][BLEND_BATCH];

  const int tid = threadIdx.x;
  // 1-D grid: CTA b works on sub-tile b % FWD_Z of the (b / FWD_Z)-th LONGEST tile list, so the in-order CTA dispatch
  // starts the long lists first and the empty tiles fill the tail of the kernel
  const uint32_t sub = blockIdx.x % FWD_Z;
  const uint32_t tile = tile_order ? __ldg(tile_order + blockIdx.x / FWD_Z) : blockIdx.x / FWD_Z;
  const uint32_t tile_x = tile % gx, tile_y = tile / gx;
  int lx, ly;
  pixel_of_thread(tid + FWD_THREADS * (int)sub, lx, ly);
  const uint32_t pix_x = tile_x * TILE_X + lx, pix_y = tile_y * TILE_Y + ly;
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const uint32_t pix_id = W * pix_y + pix_x;
  const float2 pixf = {(float)pix_x + 0.5f, (float)pix_y + 0.5f};
  bool done = !inside;
  // pixel-centre bounds of this warp's patch
  const int wq = (tid >> 5) + FWD_NWARP * (int)sub;
  const float pcx0 = (float)(tile_x * TILE_X + ((wq & 1) << 3)) + 0.5f, pcx1 = pcx0 + 7.0f;
  const float pcy0 = (float)(tile_y * TILE_Y + ((wq >> 1) << 2)) + 0.5f, pcy1 = pcy0 + 3.0f;

  const uint2 range = ranges[tile];
  const int rounds = (range.y - range.x + BLEND_BATCH - 1) / BLEND_BATCH;
  int toDo = range.y - range.x;
  const uint32_t sq_base = smem_addr(&s_q[0][0][0]);
  constexpr uint32_t QS = 16u * BLEND_BATCH;              // bytes per staged quad plane
  constexpr uint32_t BUF = QS * REC_QUADS;                // bytes per staging buffer

  float T = 1.0f;
  uint32_t last_contributor = 0, median_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float Dacc = 0.f, N[3] = {0.f, 0.f, 0.f};
  float dist1 = 0.f, dist2 = 0.f, distortion = 0.f;
  float median_depth = 0.f, median_weight = 0.f;

  // instance ids of this thread's slots in batch bi; fetched one batch ahead of the cp.async that needs them
  struct Ids { uint32_t v[FWD_SPT]; };
  auto slot_id = [&](int bi) -> Ids {
    Ids r;
#pragma unroll
    for (int k = 0; k < FWD_SPT; k++) {
      const int t = tid + k * FWD_THREADS;
      const uint32_t pos = range.x + (uint32_t)(bi * BLEND_BATCH + t);
      r.v[k] = (bi < rounds && t < BLEND_BATCH && pos < range.y) ? __ldg(&point_list[pos]) : 0xffffffffu;
    }
    return r;
  };
  auto stage = [&](int buf, const Ids& ids) {
#pragma unroll
    for (int k = 0; k < FWD_SPT; k++) {
      const uint32_t id = ids.v[k];
      if (id != 0xffffffffu) {
        const float4* r4 = reinterpret_cast<const float4*>(rec + id);
        const uint32_t dst = sq_base + (uint32_t)buf * BUF + ((uint32_t)(tid + k * FWD_THREADS) << 4);
#pragma unroll
        for (int q = 0; q < REC_QUADS; q++) cp_async16(dst + q * QS, r4 + q);
      }
    }
    cp_async_commit();
  };
  stage(0, slot_id(0));
  Ids pre_id = slot_id(1);
  for (int i = 0; i < rounds; i++, toDo -= BLEND_BATCH) {
    // every warp has left batch i-1 (so its buffer may be refilled) — and if all pixels are finished the tile is done
    if (__syncthreads_count(done) == FWD_THREADS) break;
    const int buf = i & 1;
    if (i + 1 < rounds) {
      stage(buf ^ 1, pre_id);
      pre_id = slot_id(i + 2);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sb0 = sq_base + (uint32_t)buf * BUF;
    const uint32_t sb1 = sb0 + QS, sb2 = sb0 + 2 * QS, sb3 = sb0 + 3 * QS, sb4 = sb0 + 4 * QS, sb5 = sb0 + 5 * QS;

    const int n = min(BLEND_BATCH, toDo);
    const uint32_t cbase = (uint32_t)(i * BLEND_BATCH) + 1u;
    // Warp-level compaction: each lane tests 8 staged instances against this warp's 8x4 patch (one conflict-free
    // LDS.128 each); the ballots form a 256-bit survivor mask and the warp then visits survivors only.  At C3 ~75 % of
    // (patch, instance) pairs die here for ~1/30 of the instructions a per-iteration test costs.
    uint32_t keepmask[BLEND_BATCH / 32];
#pragma unroll
    for (int w = 0; w < BLEND_BATCH / 32; w++) {
      const int jj = w * 32 + (tid & 31);
      bool keep = jj < n;
      if (keep && cull) {
        const float4 bb = lds128(sb5 + ((uint32_t)jj << 4));
        keep = !(bb.z < pcx0 || bb.x > pcx1 || bb.w < pcy0 || bb.y > pcy1);
      }
      keepmask[w] = __ballot_sync(0xffffffffu, keep);
    }
#pragma unroll
    for (int w = 0; w < BLEND_BATCH / 32; w++) {
      uint32_t m = keepmask[w];
      while (m != 0u && !done) {
      const int j = w * 32 + (__ffs(m) - 1);
      m &= m - 1u;
      const uint32_t off = (uint32_t)j << 4;
      const float4 a = lds128(sb0 + off), b = lds128(sb1 + off), c = lds128(sb2 + off);
      const float3 Tu = {a.x, a.y, a.z}, Tv = {a.w, b.x, b.y}, Tw = {b.z, b.w, c.x};
      // ray / splat intersection: two planes through the pixel, their cross product is the homogeneous hit point.
      // Roundings pinned (intrinsics) to the instruction sequence of the reference build: every alpha — and with it
      // n_contrib, the saturation test and the median — must come out bit-identical (forward.cu:357-398).
      const float3 k = {__fmaf_rn(pixf.x, Tw.x, -Tu.x), __fmaf_rn(pixf.x, Tw.y, -Tu.y), __fmaf_rn(pixf.x, Tw.z, -Tu.z)};
      const float3 l = {__fmaf_rn(pixf.y, Tw.x, -Tv.x), __fmaf_rn(pixf.y, Tw.y, -Tv.y), __fmaf_rn(pixf.y, Tw.z, -Tv.z)};
      const float3 p = {__fmaf_rn(k.y, l.z, -__fmul_rn(k.z, l.y)), __fmaf_rn(k.z, l.x, -__fmul_rn(k.x, l.z)),
                        __fmaf_rn(k.x, l.y, -__fmul_rn(k.y, l.x))};
      const float2 dd = {__fsub_rn(c.y, pixf.x), __fsub_rn(c.z, pixf.y)};
      // 1/FilterSize^2 * r^2 evaluated in double and rounded equals 2*r^2 in float exactly
      const float rho2d = 2.0f * __fmaf_rn(dd.x, dd.x, __fmul_rn(dd.y, dd.y));
      if (pair_rejected(p.x, p.y, p.z, rho2d, c.w)) continue;   // alpha < 1/255 for certain
      if (p.z == 0.0f) continue;
      const float2 s = {__fdiv_rn(p.x, p.z), __fdiv_rn(p.y, p.z)};
      const float rho3d = __fmaf_rn(s.x, s.x, __fmul_rn(s.y, s.y));
      const float rho = fminf(rho3d, rho2d);
      const float depth = (rho3d <= rho2d) ? __fadd_rn(Tw.z, __fmaf_rn(Tw.x, s.x, __fmul_rn(Tw.y, s.y))) : Tw.z;
      if (depth < 0.2f) continue;   // (double)depth < 0.2 <=> depth < 0.2f
      const float power = -0.5f * rho;
      if (power > 0.0f) continue;
      const float4 col = lds128(sb4 + off);   // rgb + opacity
      const float alpha = fminf(0.99f, __fmul_rn(col.w, expf(power)));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const uint32_t contributor = cbase + (uint32_t)j;
      const float4 nrm = lds128(sb3 + off);
      // distortion bookkeeping (forward.cu:399-421): every accumulator is fma(T, value*alpha, acc) like the reference
      const float A = __fsub_rn(1.f, T);
      const float md = mapped_depth(depth);
      const float md2 = __fmul_rn(md, md);
      const float error = __fmaf_rn(-dist1, __fadd_rn(md, md), __fmaf_rn(A, md2, dist2));
      distortion = __fmaf_rn(T, __fmul_rn(alpha, error), distortion);
      if (T > 0.5f) {
        median_depth = depth;
        median_weight = __fmul_rn(T, alpha);
        median_contributor = contributor;
      }
      N[0] = __fmaf_rn(T, __fmul_rn(nrm.x, alpha), N[0]);
      N[1] = __fmaf_rn(T, __fmul_rn(nrm.y, alpha), N[1]);
      N[2] = __fmaf_rn(T, __fmul_rn(nrm.z, alpha), N[2]);
      Dacc = __fmaf_rn(T, __fmul_rn(depth, alpha), Dacc);
      dist1 = __fmaf_rn(T, __fmul_rn(alpha, md), dist1);
      dist2 = __fmaf_rn(T, __fmul_rn(alpha, md2), dist2);
      C[0] = __fmaf_rn(T, __fmul_rn(col.x, alpha), C[0]);
      C[1] = __fmaf_rn(T, __fmul_rn(col.y, alpha), C[1]);
      C[2] = __fmaf_rn(T, __fmul_rn(col.z, alpha), C[2]);
      T = test_T;
      last_contributor = contributor;
      }
    }
  }
  cp_async_wait<0>();   // an early exit (all pixels saturated) may leave the prefetch of the next batch in flight

  if (inside) {
    const size_t HW = (size_t)H * W;
    final_T[pix_id] = T;
    final_T[pix_id + HW] = dist1;
    final_T[pix_id + 2 * HW] = dist2;
    n_contrib[pix_id] = last_contributor;
    n_contrib[pix_id + HW] = median_contributor;   // 0 when nothing contributed (the reference converts -1.0f: undefined)
    // deferred-count mode: a frame whose instance count exceeded the binning capacity is never silently wrong
    const bool poisoned = status != nullptr && status[1] != 0u;
    const float qnan = __int_as_float(0x7fc00000);
    out_color[0 * HW + pix_id] = poisoned ? qnan : __fmaf_rn(T, bg[0], C[0]);
    out_color[1 * HW + pix_id] = poisoned ? qnan : __fmaf_rn(T, bg[1], C[1]);
    out_color[2 * HW + pix_id] = poisoned ? qnan : __fmaf_rn(T, bg[2], C[2]);
    out_others[0 * HW + pix_id] = Dacc;
    out_others[1 * HW + pix_id] = 1 - T;
    out_others[2 * HW + pix_id] = N[0];
    out_others[3 * HW + pix_id] = N[1];
    out_others[4 * HW + pix_id] = N[2];
    out_others[5 * HW + pix_id] = median_depth;
    out_others[6 * HW + pix_id] = distortion;
    out_others[7 * HW + pix_id] = median_weight;
  }
}


// ------------------------------------------------------------------------------------------------------------
// Lane-walk variant of the forward blend (default).  Same staging, same arithmetic, different SIMD mapping:
//   phase 1  all lanes run the exact prefilter for every cull-box survivor of the warp's 8x4 patch (broadcast LDS.128)
//            and each lane records ITS OWN hits as bits of a 64-bit mask (bit = ordinal of the survivor in the chunk;
//            the ordinal -> staged slot map is a per-warp byte array in shared memory);
//   phase 2  every lane walks its own hit list front to back: intersection (recomputed, same roundings), divisions,
//            exp, saturation test, accumulation.  Lanes work on DIFFERENT instances at the same time, so the expensive
//            part runs at ~20 of 32 active lanes instead of ~9 (a splat covers ~9 pixels of a patch; measured on the
//            C3 scene: 0.59 M warp iterations instead of 1.31 M).
// Per pixel the instances are still consumed strictly in list order with the same float operations, so every output
// is bit-identical to blend_fwd_kernel (and to the reference where that one is).
// ------------------------------------------------------------------------------------------------------------
#ifndef D2GS_FWD_LW_CHUNK
#define D2GS_FWD_LW_CHUNK 32
#endif
constexpr int LW_CHUNK = D2GS_FWD_LW_CHUNK;
static_assert(D2GS_FWD_LW_CHUNK == 32, "the hit-mask transpose works on one 32-bit mask per lane");
__global__ void __launch_bounds__(FWD_THREADS, D2GS_FWD_MINBLOCKS) blend_fwd_lw_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const SurfelRec* __restrict__ rec, const float* __restrict__ bg, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_others, int cull,
    const uint32_t* __restrict__ status, const uint32_t* __restrict__ tile_order, uint32_t gx,
    uint32_t* __restrict__ hit_mask) {
  __shared__ float4 s_q[2][REC_QUADS][BLEND_BATCH];
  __shared__ uint8_t s_slot[FWD_NWARP][BLEND_BATCH];

  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t sub = blockIdx.x % FWD_Z;
  const uint32_t tile = tile_order ? __ldg(tile_order + blockIdx.x / FWD_Z) : blockIdx.x / FWD_Z;
  const uint32_t tile_x = tile % gx, tile_y = tile / gx;
  int lx, ly;
  pixel_of_thread(tid + FWD_THREADS * (int)sub, lx, ly);
  const uint32_t pix_x = tile_x * TILE_X + lx, pix_y = tile_y * TILE_Y + ly;
  const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
  const uint32_t pix_id = W * pix_y + pix_x;
  const float2 pixf = {(float)pix_x + 0.5f, (float)pix_y + 0.5f};
  bool done = !inside;
  const int wq = (tid >> 5) + FWD_NWARP * (int)sub;
  const float pcx0 = (float)(tile_x * TILE_X + ((wq & 1) << 3)) + 0.5f, pcx1 = pcx0 + 7.0f;
  const float pcy0 = (float)(tile_y * TILE_Y + ((wq >> 1) << 2)) + 0.5f, pcy1 = pcy0 + 3.0f;

  const uint2 range = ranges[tile];
  const int rounds = (range.y - range.x + BLEND_BATCH - 1) / BLEND_BATCH;
  int toDo = range.y - range.x;
  const uint32_t sq_base = smem_addr(&s_q[0][0][0]);
  const uint32_t slot_base = smem_addr(&s_slot[tid >> 5][0]);
  constexpr uint32_t QS = 16u * BLEND_BATCH;
  constexpr uint32_t BUF = QS * REC_QUADS;
  // this patch's row of the tile's hit-mask segment (see raster_common.cuh)
  uint32_t* const hm_row = hit_mask ? hit_mask + (size_t)range.x * 8u + (size_t)wq * (range.y - range.x) : nullptr;

  float T = 1.0f;
  uint32_t last_contributor = 0, median_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};
  float Dacc = 0.f, N[3] = {0.f, 0.f, 0.f};
  float dist1 = 0.f, dist2 = 0.f, distortion = 0.f;
  float median_depth = 0.f, median_weight = 0.f;

  struct Ids { uint32_t v[FWD_SPT]; };
  auto slot_id = [&](int bi) -> Ids {
    Ids r;
#pragma unroll
    for (int k = 0; k < FWD_SPT; k++) {
      const int t = tid + k * FWD_THREADS;
      const uint32_t pos = range.x + (uint32_t)(bi * BLEND_BATCH + t);
      r.v[k] = (bi < rounds && t < BLEND_BATCH && pos < range.y) ? __ldg(&point_list[pos]) : 0xffffffffu;
    }
    return r;
  };
  auto stage = [&](int buf, const Ids& ids) {
#pragma unroll
    for (int k = 0; k < FWD_SPT; k++) {
      const uint32_t id = ids.v[k];
      if (id != 0xffffffffu) {
        const float4* r4 = reinterpret_cast<const float4*>(rec + id);
        const uint32_t dst = sq_base + (uint32_t)buf * BUF + ((uint32_t)(tid + k * FWD_THREADS) << 4);
#pragma unroll
        for (int q = 0; q < REC_QUADS; q++) cp_async16(dst + q * QS, r4 + q);
      }
    }
    cp_async_commit();
  };
  stage(0, slot_id(0));
  Ids pre_id = slot_id(1);
  for (int i = 0; i < rounds; i++, toDo -= BLEND_BATCH) {
    if (__syncthreads_count(done) == FWD_THREADS) break;
    const int buf = i & 1;
    if (i + 1 < rounds) {
      stage(buf ^ 1, pre_id);
      pre_id = slot_id(i + 2);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sb0 = sq_base + (uint32_t)buf * BUF;
    const uint32_t sb1 = sb0 + QS, sb2 = sb0 + 2 * QS, sb3 = sb0 + 3 * QS, sb4 = sb0 + 4 * QS, sb5 = sb0 + 5 * QS;
    const int n = min(BLEND_BATCH, toDo);
    const uint32_t cbase = (uint32_t)(i * BLEND_BATCH) + 1u;

    uint32_t keepmask[BLEND_BATCH / 32];
#pragma unroll
    for (int w = 0; w < BLEND_BATCH / 32; w++) {
      const int jj = w * 32 + lane;
      bool keep = jj < n;
      if (keep && cull) {
        const float4 bb = lds128(sb5 + ((uint32_t)jj << 4));
        keep = !(bb.z < pcx0 || bb.x > pcx1 || bb.w < pcy0 || bb.y > pcy1);
      }
      keepmask[w] = __ballot_sync(0xffffffffu, keep);
    }
    if (__all_sync(0xffffffffu, done)) continue;      // this patch is saturated: it only keeps the barriers company
    // compact list of the surviving slots (ballot compaction, one byte per survivor): the loops below walk it with a
    // plain counter instead of bit-scanning four mask words
    int ns = 0;
#pragma unroll
    for (int w = 0; w < BLEND_BATCH / 32; w++) {
      const uint32_t km = keepmask[w];
      if ((km >> lane) & 1u)
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(slot_base + (uint32_t)(ns + __popc(km & ((1u << lane) - 1u)))), "r"(w * 32 + lane) : "memory");
      ns += __popc(km);
      // hit masks for the backward: zeros first (slots the cull box dropped stay 0), survivors' ballots overwrite them below
      if (hm_row && w * 32 + lane < n) hm_row[i * BLEND_BATCH + w * 32 + lane] = 0u;
    }
    __syncwarp();
    for (int o0 = 0; o0 < ns; o0 += LW_CHUNK) {
      // ---- phase 1: up to LW_CHUNK survivors, exact prefilter on all lanes, per-lane hit masks
      const int nchunk = min(LW_CHUNK, ns - o0);
      const uint32_t sbase = slot_base + (uint32_t)o0;
      uint32_t mine = 0u;
      for (int ord = 0; ord < nchunk; ord++) {
        uint32_t j;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(j) : "r"(sbase + (uint32_t)ord));
        const uint32_t off = j << 4;
        const float4 a = lds128(sb0 + off), b = lds128(sb1 + off), c = lds128(sb2 + off);
        const float3 k = {__fmaf_rn(pixf.x, b.z, -a.x), __fmaf_rn(pixf.x, b.w, -a.y), __fmaf_rn(pixf.x, c.x, -a.z)};
        const float3 l = {__fmaf_rn(pixf.y, b.z, -a.w), __fmaf_rn(pixf.y, b.w, -b.x), __fmaf_rn(pixf.y, c.x, -b.y)};
        const float3 p = {__fmaf_rn(k.y, l.z, -__fmul_rn(k.z, l.y)), __fmaf_rn(k.z, l.x, -__fmul_rn(k.x, l.z)),
                          __fmaf_rn(k.x, l.y, -__fmul_rn(k.y, l.x))};
        const float2 dd = {__fsub_rn(c.y, pixf.x), __fsub_rn(c.z, pixf.y)};
        const float rho2d = 2.0f * __fmaf_rn(dd.x, dd.x, __fmul_rn(dd.y, dd.y));
        const bool pass = !pair_rejected(p.x, p.y, p.z, rho2d, c.w) && p.z != 0.0f;
        mine |= (pass ? 1u : 0u) << ord;
      }
      if (hm_row) {
        // hit masks for the backward: lane p holds the chunk's hit bits of pixel p; the transposed bit matrix gives lane o the
        // ballot of survivor o over the 32 pixels, stored at the survivor's list position (30 instructions per chunk instead
        // of a ballot + select per survivor)
        const uint32_t hm = transpose_bits32(mine, lane);
        if (lane < nchunk && hm != 0u) {
          uint32_t j;
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(j) : "r"(sbase + (uint32_t)lane));
          hm_row[i * BLEND_BATCH + j] = hm;
        }
      }
      // ---- phase 2: every lane consumes its own hits in list order
      {
        while (mine != 0u && !done) {
          const uint32_t o = (uint32_t)(__ffs(mine) - 1);
          mine &= mine - 1u;
          uint32_t j;
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(j) : "r"(sbase + o));
          const uint32_t off = j << 4;
          const float4 a = lds128(sb0 + off), b = lds128(sb1 + off), c = lds128(sb2 + off);
          const float3 Tw = {b.z, b.w, c.x};
          const float3 k = {__fmaf_rn(pixf.x, b.z, -a.x), __fmaf_rn(pixf.x, b.w, -a.y), __fmaf_rn(pixf.x, c.x, -a.z)};
          const float3 l = {__fmaf_rn(pixf.y, b.z, -a.w), __fmaf_rn(pixf.y, b.w, -b.x), __fmaf_rn(pixf.y, c.x, -b.y)};
          const float3 p = {__fmaf_rn(k.y, l.z, -__fmul_rn(k.z, l.y)), __fmaf_rn(k.z, l.x, -__fmul_rn(k.x, l.z)),
                            __fmaf_rn(k.x, l.y, -__fmul_rn(k.y, l.x))};
          const float2 dd = {__fsub_rn(c.y, pixf.x), __fsub_rn(c.z, pixf.y)};
          const float rho2d = 2.0f * __fmaf_rn(dd.x, dd.x, __fmul_rn(dd.y, dd.y));
          const float2 s = {__fdiv_rn(p.x, p.z), __fdiv_rn(p.y, p.z)};
          const float rho3d = __fmaf_rn(s.x, s.x, __fmul_rn(s.y, s.y));
          const float rho = fminf(rho3d, rho2d);
          const float depth = (rho3d <= rho2d) ? __fadd_rn(Tw.z, __fmaf_rn(Tw.x, s.x, __fmul_rn(Tw.y, s.y))) : Tw.z;
          if (depth < 0.2f) continue;
          const float power = -0.5f * rho;
          if (power > 0.0f) continue;
          const float4 col = lds128(sb4 + off);
          const float alpha = fminf(0.99f, __fmul_rn(col.w, expf(power)));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
          if (test_T < 0.0001f) {
            done = true;
            continue;
          }
          const uint32_t contributor = cbase + j;
          const float4 nrm = lds128(sb3 + off);
          const float A = __fsub_rn(1.f, T);
          const float md = mapped_depth(depth);
          const float md2 = __fmul_rn(md, md);
          const float error = __fmaf_rn(-dist1, __fadd_rn(md, md), __fmaf_rn(A, md2, dist2));
          distortion = __fmaf_rn(T, __fmul_rn(alpha, error), distortion);
          if (T > 0.5f) {
            median_depth = depth;
            median_weight = __fmul_rn(T, alpha);
            median_contributor = contributor;
          }
          N[0] = __fmaf_rn(T, __fmul_rn(nrm.x, alpha), N[0]);
          N[1] = __fmaf_rn(T, __fmul_rn(nrm.y, alpha), N[1]);
          N[2] = __fmaf_rn(T, __fmul_rn(nrm.z, alpha), N[2]);
          Dacc = __fmaf_rn(T, __fmul_rn(depth, alpha), Dacc);
          dist1 = __fmaf_rn(T, __fmul_rn(alpha, md), dist1);
          dist2 = __fmaf_rn(T, __fmul_rn(alpha, md2), dist2);
          C[0] = __fmaf_rn(T, __fmul_rn(col.x, alpha), C[0]);
          C[1] = __fmaf_rn(T, __fmul_rn(col.y, alpha), C[1]);
          C[2] = __fmaf_rn(T, __fmul_rn(col.z, alpha), C[2]);
          T = test_T;
          last_contributor = contributor;
        }
      }
    }
    __syncwarp();     // the slot list is rewritten by the next batch
  }
  cp_async_wait<0>();

  if (inside) {
    const size_t HW = (size_t)H * W;
    final_T[pix_id] = T;
    final_T[pix_id + HW] = dist1;
    final_T[pix_id + 2 * HW] = dist2;
    n_contrib[pix_id] = last_contributor;
    n_contrib[pix_id + HW] = median_contributor;
    const bool poisoned = status != nullptr && status[1] != 0u;
    const float qnan = __int_as_float(0x7fc00000);
    out_color[0 * HW + pix_id] = poisoned ? qnan : __fmaf_rn(T, bg[0], C[0]);
    out_color[1 * HW + pix_id] = poisoned ? qnan : __fmaf_rn(T, bg[1], C[1]);
    out_color[2 * HW + pix_id] = poisoned ? qnan : __fmaf_rn(T, bg[2], C[2]);
    out_others[0 * HW + pix_id] = Dacc;
    out_others[1 * HW + pix_id] = 1 - T;
    out_others[2 * HW + pix_id] = N[0];
    out_others[3 * HW + pix_id] = N[1];
    out_others[4 * HW + pix_id] = N[2];
    out_others[5 * HW + pix_id] = median_depth;
    out_others[6 * HW + pix_id] = distortion;
    out_others[7 * HW + pix_id] = median_weight;
  }
}

void launch_blend_fwd(const FwdParams& p, const uint2* ranges, const uint32_t* point_list, const SurfelRec* rec,
                      float* final_T, uint32_t* n_contrib, float* out_color, float* out_others, int cull,
                      const uint32_t* status, const uint32_t* tile_order, int lane_walk, uint32_t* hit_mask, cudaStream_t s) {
  const uint32_t grid = p.gx * p.gy * FWD_Z;
  if (lane_walk) {
    blend_fwd_lw_kernel<<<grid, FWD_THREADS, 0, s>>>(ranges, point_list, p.W, p.H, rec, p.bg, final_T, n_contrib, out_color,
                                                     out_others, cull, status, tile_order, p.gx, hit_mask);
    return;
  }
  blend_fwd_kernel<<<grid, FWD_THREADS, 0, s>>>(ranges, point_list, p.W, p.H, rec, p.bg, final_T, n_contrib, out_color,
                                            out_others, cull, status, tile_order, p.gx);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float z = view[2] * means3D[3 * idx] + view[6] * means3D[3 * idx + 1] + view[10] * means3D[3 * idx + 2] + view[14];
  present[idx] = !(z <= 0.2f);
}

void launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s) {
  if (P == 0) return;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
}

}  // namespace d2gs
