// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement of the reference 2-D-Gaussian ("surfel") rasterizer, forward and backward,
// written from the reference's algorithm (not its code) so the CUDA path can be checked against it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Parity pin: the reference ships no golden vectors.  This oracle is pinned
// against outputs of the UNMODIFIED reference CUDA extension (oracle/_ref, built by
// oracle/build_ref.sh) recorded on a B200 into tests/golden/ by tests/golden/make_golden.py.
//
// Reference citations (relative to /root/reference/submodules/diff-surfel-rasterization/):
//   preprocess fwd      cuda_rasterizer/forward.cu:166-260  (+:20-71 SH, :75-128 transMat, :133-163 AABB)
//   frustum / rect      cuda_rasterizer/auxiliary.h:160-185, :64-74
//   quaternion          cuda_rasterizer/auxiliary.h:188-210 (fwd), :213-257 (vjp)
//   binning             cuda_rasterizer/rasterizer_impl.cu:35-50, :70-111, :116-138, :278-318
//   blend fwd           cuda_rasterizer/forward.cu:265-463
//   blend bwd           cuda_rasterizer/backward.cu:143-449
//   AABB bwd            cuda_rasterizer/backward.cu:599-649
//   preprocess bwd      cuda_rasterizer/backward.cu:451-597 (+:20-139 SH bwd)
//
// Two instantiations: real = float follows the reference's fp32 arithmetic including the places
// where its macros are double literals (FilterSize, NEAR/FAR_PLANE, "+0.5"); real = double is
// the high-precision yardstick used to judge gradient error of both implementations.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int TILE = 16;                       // config.h:16-17
constexpr double FILTER_SIZE = 0.7071067811865476;  // auxiliary.h:20
constexpr double NEAR_PLANE_D = 0.2;           // auxiliary.h:35
constexpr double FAR_PLANE_D = 100.0;          // auxiliary.h:36

const float C0 = 0.28209479177387814f;
const float C1 = 0.4886025119029199f;
const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                     -1.0925484305920792f, 0.5462742152960396f};
const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                     -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

template <typename R> struct V3 { R x, y, z; };
template <typename R> inline V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename R> inline V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename R> inline V3<R> operator*(V3<R> a, V3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <typename R> inline V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename R> inline V3<R> operator*(R s, V3<R> a) { return {a.x * s, a.y * s, a.z * s}; }
template <typename R> inline R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R> inline V3<R> cross(V3<R> a, V3<R> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// float -> int conversion with the GPU's saturating semantics (NaN -> 0), so degenerate surfels
// take the same path as on the device instead of hitting C++ UB.
template <typename R> inline int to_int_sat(R v) {
  if (v != v) return 0;
  if (v >= (R)2147483647.0) return std::numeric_limits<int>::max();
  if (v <= (R)-2147483648.0) return std::numeric_limits<int>::min();
  return (int)v;
}
template <typename R> inline uint32_t to_u32_sat(R v) {
  if (v != v || v <= (R)0) return 0u;
  if (v >= (R)4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)v;
}

// view-matrix helpers: matrices arrive transposed (row-vector convention) and are indexed column-major.
template <typename R> inline V3<R> mul_W(const R* v, V3<R> p) {   // 3x3 block, no translation
  return {v[0] * p.x + v[4] * p.y + v[8] * p.z, v[1] * p.x + v[5] * p.y + v[9] * p.z,
          v[2] * p.x + v[6] * p.y + v[10] * p.z};
}
template <typename R> inline V3<R> mul_Wt(const R* v, V3<R> p) {  // transpose of the 3x3 block
  return {v[0] * p.x + v[1] * p.y + v[2] * p.z, v[4] * p.x + v[5] * p.y + v[6] * p.z,
          v[8] * p.x + v[9] * p.y + v[10] * p.z};
}

template <typename R> struct Rot { V3<R> c0, c1, c2; };   // columns
template <typename R> inline Rot<R> quat_to_rot(const R* q) {
  R s = (R)1 / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  R w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
  Rot<R> r;
  r.c0 = {(R)1 - (R)2 * (y * y + z * z), (R)2 * (x * y + w * z), (R)2 * (x * z - w * y)};
  r.c1 = {(R)2 * (x * y - w * z), (R)1 - (R)2 * (x * x + z * z), (R)2 * (y * z + w * x)};
  r.c2 = {(R)2 * (x * z + w * y), (R)2 * (y * z - w * x), (R)1 - (R)2 * (x * x + y * y)};
  return r;
}

struct Rect { uint32_t x0, y0, x1, y1; };
template <typename R> inline Rect get_rect(R px, R py, int radius, int gx, int gy) {
  auto clampi = [](int v, int hi) { return (uint32_t)std::min(hi, std::max(0, v)); };
  Rect r;
  R rad = (R)radius;
  r.x0 = clampi(to_int_sat<R>((px - rad) / (R)TILE), gx);
  r.y0 = clampi(to_int_sat<R>((py - rad) / (R)TILE), gy);
  r.x1 = clampi(to_int_sat<R>((((px + rad) + (R)TILE) - (R)1) / (R)TILE), gx);
  r.y1 = clampi(to_int_sat<R>((((py + rad) + (R)TILE) - (R)1) / (R)TILE), gy);
  return r;
}

inline uint32_t higher_msb(uint32_t n) {   // rasterizer_impl.cu:35-50, restated as a loop
  uint32_t msb = 16, step = 16;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

template <typename R> inline uint32_t depth_bits(R d) {
  float f = (float)d;
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

template <typename R>
struct Ctx {
  int P, D, M, W, H;
  const R *bg, *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *transMat_precomp;
  const R *view, *proj, *campos;
  R tanfovx, tanfovy;
};

template <typename R>
void sh_forward(int idx, int deg, int M, const R* means, const R* campos, const R* shs, uint8_t* clamped, R* out) {
  V3<R> pos{means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]};
  V3<R> dir = pos - V3<R>{campos[0], campos[1], campos[2]};
  R len = std::sqrt(dot(dir, dir));
  dir = {dir.x / len, dir.y / len, dir.z / len};
  const R* sh = shs + (size_t)idx * M * 3;
  auto S = [&](int k) { return V3<R>{sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]}; };
  V3<R> res = (R)C0 * S(0);
  if (deg > 0) {
    R x = dir.x, y = dir.y, z = dir.z;
    res = res - ((R)C1 * y) * S(1) + ((R)C1 * z) * S(2) - ((R)C1 * x) * S(3);
    if (deg > 1) {
      R xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      res = res + ((R)C2[0] * xy) * S(4) + ((R)C2[1] * yz) * S(5) + ((R)C2[2] * ((R)2 * zz - xx - yy)) * S(6) +
            ((R)C2[3] * xz) * S(7) + ((R)C2[4] * (xx - yy)) * S(8);
      if (deg > 2) {
        res = res + ((R)C3[0] * y * ((R)3 * xx - yy)) * S(9) + ((R)C3[1] * xy * z) * S(10) +
              ((R)C3[2] * y * ((R)4 * zz - xx - yy)) * S(11) +
              ((R)C3[3] * z * ((R)2 * zz - (R)3 * xx - (R)3 * yy)) * S(12) +
              ((R)C3[4] * x * ((R)4 * zz - xx - yy)) * S(13) + ((R)C3[5] * z * (xx - yy)) * S(14) +
              ((R)C3[6] * x * (xx - (R)3 * yy)) * S(15);
      }
    }
  }
  res = res + V3<R>{(R)0.5, (R)0.5, (R)0.5};
  clamped[3 * idx + 0] = res.x < 0;
  clamped[3 * idx + 1] = res.y < 0;
  clamped[3 * idx + 2] = res.z < 0;
  out[0] = std::max(res.x, (R)0);
  out[1] = std::max(res.y, (R)0);
  out[2] = std::max(res.z, (R)0);
}

// forward.cu:75-128.  Returns false for a surfel seen exactly edge-on.
template <typename R>
bool trans_mat(const R* p, const R* quat, const R* scale, const R* view, R fx, R fy, R cx, R cy, R* T, R* normal) {
  V3<R> pw{p[0], p[1], p[2]};
  V3<R> cam{view[12], view[13], view[14]};
  V3<R> pv = mul_W(view, pw) + cam;
  Rot<R> Rm = quat_to_rot(quat);
  V3<R> M0 = mul_W(view, Rm.c0 * scale[0]);
  V3<R> M1 = mul_W(view, Rm.c1 * scale[1]);
  V3<R> M2 = pv;
  V3<R> tn = mul_W(view, Rm.c2);
  R c = dot(V3<R>{-tn.x, -tn.y, -tn.z}, pv);
  if (c == (R)0) return false;
  R mult = c > 0 ? (R)1 : (R)-1;
  tn = tn * mult;
  T[0] = fx * M0.x + cx * M0.z; T[1] = fx * M1.x + cx * M1.z; T[2] = fx * M2.x + cx * M2.z;
  T[3] = fy * M0.y + cy * M0.z; T[4] = fy * M1.y + cy * M1.z; T[5] = fy * M2.y + cy * M2.z;
  T[6] = M0.z; T[7] = M1.z; T[8] = M2.z;
  normal[0] = tn.x; normal[1] = tn.y; normal[2] = tn.z;
  return true;
}

template <typename R> bool aabb(const R* T, R* center, R* extent) {   // forward.cu:133-163
  V3<R> Tu{T[0], T[1], T[2]}, Tv{T[3], T[4], T[5]}, Tw{T[6], T[7], T[8]};
  V3<R> sgn{(R)1, (R)1, (R)-1};
  R d = dot(sgn, Tw * Tw);
  if (d == (R)0) return false;
  V3<R> f = sgn * ((R)1 / d);
  R px = dot(f, Tu * Tw), py = dot(f, Tv * Tw);
  R hx = px * px - dot(f, Tu * Tu), hy = py * py - dot(f, Tv * Tv);
  center[0] = px; center[1] = py;
  extent[0] = std::sqrt(std::max((R)0, hx));
  extent[1] = std::sqrt(std::max((R)0, hy));
  return true;
}

}  // namespace

template <typename R>
int64_t forward_impl(const Ctx<R>& c,
                     int* radii, R* means2D, R* depths, R* transMat, R* normal_opacity, R* rgb, uint8_t* clamped,
                     uint32_t* tiles_touched, uint32_t* point_offsets,
                     R* out_color, R* out_others, R* final_T, uint32_t* n_contrib, uint32_t* ranges,
                     int64_t cap, uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys_sorted,
                     uint32_t* point_list) {
  const int P = c.P, W = c.W, H = c.H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const R focal_y = (R)H / ((R)2 * c.tanfovy), focal_x = (R)W / ((R)2 * c.tanfovx);
  const R cx = (R)W / (R)2, cy = (R)H / (R)2;

#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const R* p = c.means3D + 3 * i;
    // auxiliary.h:160-185: only the near plane culls
    R pvz = c.view[2] * p[0] + c.view[6] * p[1] + c.view[10] * p[2] + c.view[14];
    if (pvz <= (R)0.2f) continue;
    const R* T;
    R nrm[3] = {0, 0, 0};
    if (c.transMat_precomp) {
      T = c.transMat_precomp + 9 * i;   // normal is undefined in the reference here; we define it as 0
    } else {
      if (!trans_mat(p, c.rotations + 4 * i, c.scales + 2 * i, c.view, focal_x, focal_y, cx, cy, transMat + 9 * i, nrm))
        continue;
      T = transMat + 9 * i;
    }
    R ctr[2], ext[2];
    if (!aabb(T, ctr, ext)) continue;
    // forward.cu:239: ceil(3.f * max(max(ex,ey), FilterSize)) — the macro is a double literal
    double m = std::max((double)std::max(ext[0], ext[1]), FILTER_SIZE);
    R radius = (R)std::ceil(3.0 * m);
    int irad = to_int_sat<R>(radius);
    Rect r = get_rect(ctr[0], ctr[1], irad, gx, gy);
    if ((r.x1 - r.x0) * (r.y1 - r.y0) == 0) continue;
    if (!c.colors_precomp) sh_forward(i, c.D, c.M, c.means3D, c.campos, c.shs, clamped, rgb + 3 * i);
    depths[i] = pvz;
    radii[i] = irad;
    means2D[2 * i] = ctr[0]; means2D[2 * i + 1] = ctr[1];
    normal_opacity[4 * i] = nrm[0]; normal_opacity[4 * i + 1] = nrm[1]; normal_opacity[4 * i + 2] = nrm[2];
    normal_opacity[4 * i + 3] = c.opacities[i];
    tiles_touched[i] = (r.y1 - r.y0) * (r.x1 - r.x0);
  }

  uint32_t run = 0;
  for (int i = 0; i < P; i++) { run += tiles_touched[i]; point_offsets[i] = run; }
  const int64_t Rn = P ? (int64_t)point_offsets[P - 1] : 0;
  if (Rn > cap) return Rn;   // caller re-invokes with enough room

#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (radii[i] <= 0) continue;
    uint32_t off = i == 0 ? 0 : point_offsets[i - 1];
    Rect r = get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy);
    uint32_t db = depth_bits(depths[i]);
    for (uint32_t y = r.y0; y < r.y1; y++)
      for (uint32_t x = r.x0; x < r.x1; x++) {
        keys_unsorted[off] = ((uint64_t)(y * gx + x) << 32) | db;
        vals_unsorted[off] = (uint32_t)i;
        off++;
      }
  }

  // stable sort on the low 32+bit bits (rasterizer_impl.cu:301-309)
  const uint32_t bit = higher_msb((uint32_t)(gx * gy));
  const uint64_t mask = (32 + bit) >= 64 ? ~0ull : ((1ull << (32 + bit)) - 1);
  std::vector<uint32_t> order((size_t)Rn);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    return (keys_unsorted[a] & mask) < (keys_unsorted[b] & mask);
  });
  for (int64_t k = 0; k < Rn; k++) { keys_sorted[k] = keys_unsorted[order[k]]; point_list[k] = vals_unsorted[order[k]]; }

  std::memset(ranges, 0, sizeof(uint32_t) * 2 * gx * gy);
  for (int64_t k = 0; k < Rn; k++) {
    uint32_t cur = (uint32_t)(keys_sorted[k] >> 32);
    if (k == 0) ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(keys_sorted[k - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)k; ranges[2 * cur] = (uint32_t)k; }
    }
    if (k == Rn - 1) ranges[2 * cur + 1] = (uint32_t)Rn;
  }

  const R* feat = c.colors_precomp ? c.colors_precomp : rgb;
  const R* Tm = c.transMat_precomp ? c.transMat_precomp : transMat;
  const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
        if (pxi >= W || pyi >= H) continue;
        const size_t pix = (size_t)W * pyi + pxi;
        const R pfx = (R)((double)pxi + 0.5), pfy = (R)((double)pyi + 0.5);
        R T = 1, C[3] = {0, 0, 0}, Dd = 0, N[3] = {0, 0, 0}, dist1 = 0, dist2 = 0, distortion = 0;
        R median_depth = 0, median_weight = 0, median_contrib = -1;
        uint32_t contributor = 0, last_contributor = 0;
        for (uint32_t k = r0; k < r1; k++) {
          contributor++;
          const uint32_t id = point_list[k];
          const R* t = Tm + 9 * (size_t)id;
          V3<R> Tu{t[0], t[1], t[2]}, Tv{t[3], t[4], t[5]}, Tw{t[6], t[7], t[8]};
          V3<R> kk{-Tu.x + pfx * Tw.x, -Tu.y + pfx * Tw.y, -Tu.z + pfx * Tw.z};
          V3<R> ll{-Tv.x + pfy * Tw.x, -Tv.y + pfy * Tw.y, -Tv.z + pfy * Tw.z};
          V3<R> p = cross(kk, ll);
          if (p.z == (R)0) continue;
          R sx = p.x / p.z, sy = p.y / p.z;
          R rho3d = sx * sx + sy * sy;
          R dx = means2D[2 * id] - pfx, dy = means2D[2 * id + 1] - pfy;
          R rho2d = (R)(1 / (FILTER_SIZE * FILTER_SIZE) * (double)(dx * dx + dy * dy));
          R rho = std::min(rho3d, rho2d);
          R depth = (rho3d <= rho2d) ? (sx * Tw.x + sy * Tw.y) + Tw.z : Tw.z;
          if ((double)depth < NEAR_PLANE_D) continue;
          const R* no = normal_opacity + 4 * (size_t)id;
          R power = (R)-0.5 * rho;
          if (power > 0) continue;
          R alpha = std::min((R)0.99f, no[3] * std::exp(power));
          if (alpha < (R)(1.0f / 255.0f)) continue;
          R test_T = T * (1 - alpha);
          if (test_T < (R)0.0001f) break;   // "done": pixel stops consuming the list
          R A = 1 - T;
          R md = (R)((FAR_PLANE_D * (double)depth - FAR_PLANE_D * NEAR_PLANE_D) / ((FAR_PLANE_D - NEAR_PLANE_D) * (double)depth));
          R err = md * md * A + dist2 - 2 * md * dist1;
          distortion += err * alpha * T;
          if ((double)T > 0.5) { median_depth = depth; median_weight = alpha * T; median_contrib = (R)contributor; }
          for (int ch = 0; ch < 3; ch++) N[ch] += no[ch] * alpha * T;
          Dd += depth * alpha * T;
          dist1 += md * alpha * T;
          dist2 += md * md * alpha * T;
          for (int ch = 0; ch < 3; ch++) C[ch] += feat[3 * (size_t)id + ch] * alpha * T;
          T = test_T;
          last_contributor = contributor;
        }
        final_T[pix] = T;
        final_T[pix + HW] = dist1;
        final_T[pix + 2 * HW] = dist2;
        n_contrib[pix] = last_contributor;
        n_contrib[pix + HW] = to_u32_sat<R>(median_contrib);
        for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = C[ch] + T * c.bg[ch];
        out_others[0 * HW + pix] = Dd;
        out_others[1 * HW + pix] = 1 - T;
        for (int ch = 0; ch < 3; ch++) out_others[(2 + ch) * HW + pix] = N[ch];
        out_others[5 * HW + pix] = median_depth;
        out_others[6 * HW + pix] = distortion;
        out_others[7 * HW + pix] = median_weight;
      }
  }
  return Rn;
}

template <typename R> inline void atomic_add(R* p, R v) {
#pragma omp atomic
  *p += v;
}

template <typename R>
void backward_impl(const Ctx<R>& c, int64_t Rn, const int* radii, const R* means2D, const R* transMat,
                   const R* normal_opacity, const R* rgb, const uint8_t* clamped, const R* final_T,
                   const uint32_t* n_contrib, const uint32_t* ranges, const uint32_t* point_list,
                   const R* dL_dpix, const R* dL_dothers,
                   R* dL_dmean2D, R* dL_dnormal, R* dL_dopacity, R* dL_dcolor, R* dL_dmean3D, R* dL_dtransMat,
                   R* dL_dsh, R* dL_dscale, R* dL_drot) {
  (void)Rn;
  const int P = c.P, W = c.W, H = c.H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const R focal_y = (R)H / ((R)2 * c.tanfovy), focal_x = (R)W / ((R)2 * c.tanfovx);
  const size_t HW = (size_t)H * W;
  const R* feat = c.colors_precomp ? c.colors_precomp : rgb;
  const R* Tm = c.transMat_precomp ? c.transMat_precomp : transMat;

  // ---- blend backward (backward.cu:143-449): back-to-front per pixel -------------------------
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
        if (pxi >= W || pyi >= H) continue;
        const size_t pix = (size_t)W * pyi + pxi;
        const R pfx = (R)((double)pxi + 0.5), pfy = (R)((double)pyi + 0.5);
        const R T_final = final_T[pix];
        R T = T_final;
        const uint32_t last_contributor = n_contrib[pix];
        const int median_contributor = (int)n_contrib[pix + HW];
        R dpix[3];
        for (int ch = 0; ch < 3; ch++) dpix[ch] = dL_dpix[ch * HW + pix];
        const R dL_ddepth = dL_dothers[0 * HW + pix], dL_daccum = dL_dothers[1 * HW + pix];
        const R dL_dreg = dL_dothers[6 * HW + pix];
        R dL_dn2[3];
        for (int ch = 0; ch < 3; ch++) dL_dn2[ch] = dL_dothers[(2 + ch) * HW + pix];
        const R dL_dmedian_depth = dL_dothers[5 * HW + pix], dL_dmax_dweight = dL_dothers[7 * HW + pix];
        const R final_D = final_T[pix + HW], final_D2 = final_T[pix + 2 * HW], final_A = 1 - T_final;
        R accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
        R last_depth = 0, last_normal[3] = {0, 0, 0}, accum_depth_rec = 0, accum_alpha_rec = 0;
        R accum_normal_rec[3] = {0, 0, 0}, last_dL_dT = 0;
        R bg_dot = 0;
        for (int ch = 0; ch < 3; ch++) bg_dot += c.bg[ch] * dpix[ch];
        uint32_t contributor = r1 - r0;
        for (uint32_t k = r1; k-- > r0;) {
          contributor--;
          if (contributor >= last_contributor) continue;
          const uint32_t id = point_list[k];
          const R* t = Tm + 9 * (size_t)id;
          V3<R> Tu{t[0], t[1], t[2]}, Tv{t[3], t[4], t[5]}, Tw{t[6], t[7], t[8]};
          V3<R> kk{-Tu.x + pfx * Tw.x, -Tu.y + pfx * Tw.y, -Tu.z + pfx * Tw.z};
          V3<R> ll{-Tv.x + pfy * Tw.x, -Tv.y + pfy * Tw.y, -Tv.z + pfy * Tw.z};
          V3<R> p = cross(kk, ll);
          if (p.z == (R)0) continue;
          R sx = p.x / p.z, sy = p.y / p.z;
          R rho3d = sx * sx + sy * sy;
          R dx = means2D[2 * id] - pfx, dy = means2D[2 * id + 1] - pfy;
          R rho2d = (R)(1 / (FILTER_SIZE * FILTER_SIZE) * (double)(dx * dx + dy * dy));
          R rho = std::min(rho3d, rho2d);
          R c_d = (rho3d <= rho2d) ? (sx * Tw.x + sy * Tw.y) + Tw.z : Tw.z;
          if ((double)c_d < NEAR_PLANE_D) continue;
          const R* no = normal_opacity + 4 * (size_t)id;
          R power = (R)-0.5 * rho;
          if (power > 0) continue;
          const R G = std::exp(power);
          const R alpha = std::min((R)0.99f, no[3] * G);
          if (alpha < (R)(1.0f / 255.0f)) continue;
          T = T / ((R)1 - alpha);
          const R w = alpha * T;
          R dL_dalpha = 0;
          for (int ch = 0; ch < 3; ch++) {
            const R col = feat[3 * (size_t)id + ch];
            accum_rec[ch] = last_alpha * last_color[ch] + ((R)1 - last_alpha) * accum_rec[ch];
            last_color[ch] = col;
            dL_dalpha += (col - accum_rec[ch]) * dpix[ch];
            atomic_add(&dL_dcolor[3 * (size_t)id + ch], w * dpix[ch]);
          }
          R dL_dz = 0, dL_dweight = 0;
          const double cd = (double)c_d;
          const R m_d = (R)((FAR_PLANE_D * cd - FAR_PLANE_D * NEAR_PLANE_D) / ((FAR_PLANE_D - NEAR_PLANE_D) * cd));
          const R dmd_dd = (R)((FAR_PLANE_D * NEAR_PLANE_D) / ((FAR_PLANE_D - NEAR_PLANE_D) * cd * cd));
          if (contributor == (uint32_t)(median_contributor - 1)) { dL_dz += dL_dmedian_depth; dL_dweight += dL_dmax_dweight; }
          dL_dweight += (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
          dL_dalpha += dL_dweight - last_dL_dT;
          last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
          const R dL_dmd = (R)2 * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
          dL_dz += dL_dmd * dmd_dd;
          accum_depth_rec = last_alpha * last_depth + ((R)1 - last_alpha) * accum_depth_rec;
          last_depth = c_d;
          dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
          accum_alpha_rec = (R)((double)last_alpha * 1.0 + (double)(((R)1 - last_alpha) * accum_alpha_rec));
          dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
          for (int ch = 0; ch < 3; ch++) {
            accum_normal_rec[ch] = last_alpha * last_normal[ch] + ((R)1 - last_alpha) * accum_normal_rec[ch];
            last_normal[ch] = no[ch];
            dL_dalpha += (no[ch] - accum_normal_rec[ch]) * dL_dn2[ch];
            atomic_add(&dL_dnormal[3 * (size_t)id + ch], alpha * T * dL_dn2[ch]);
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / ((R)1 - alpha)) * bg_dot;
          const R dL_dG = no[3] * dL_dalpha;
          dL_dz += alpha * T * dL_ddepth;
          if (rho3d <= rho2d) {
            const R dsx = dL_dG * -G * sx + dL_dz * Tw.x, dsy = dL_dG * -G * sy + dL_dz * Tw.y;
            const R dsx_pz = dsx / p.z, dsy_pz = dsy / p.z;
            V3<R> dL_dp{dsx_pz, dsy_pz, -(dsx_pz * sx + dsy_pz * sy)};
            V3<R> dL_dk = cross(ll, dL_dp), dL_dl = cross(dL_dp, kk);
            R g[9] = {-dL_dk.x, -dL_dk.y, -dL_dk.z, -dL_dl.x, -dL_dl.y, -dL_dl.z,
                      pfx * dL_dk.x + pfy * dL_dl.x + dL_dz * sx, pfx * dL_dk.y + pfy * dL_dl.y + dL_dz * sy,
                      pfx * dL_dk.z + pfy * dL_dl.z + dL_dz};
            for (int q = 0; q < 9; q++) atomic_add(&dL_dtransMat[9 * (size_t)id + q], g[q]);
          } else {
            // the macro expands textually: "-G * 1/(F*F) * d" is evaluated left to right in double
            const R dG_ddelx = (R)(((double)(-G) * 1) / (FILTER_SIZE * FILTER_SIZE) * (double)dx);
            const R dG_ddely = (R)(((double)(-G) * 1) / (FILTER_SIZE * FILTER_SIZE) * (double)dy);
            atomic_add(&dL_dmean2D[3 * (size_t)id + 0], dL_dG * dG_ddelx);
            atomic_add(&dL_dmean2D[3 * (size_t)id + 1], dL_dG * dG_ddely);
            atomic_add(&dL_dtransMat[9 * (size_t)id + 8], dL_dz);
          }
          atomic_add(&dL_dopacity[id], G * dL_dalpha);
        }
      }
  }

  // ---- per-surfel backward (backward.cu:599-649 then :533-597) --------------------------------
  const R Wn = focal_x * c.tanfovx, Hn = focal_y * c.tanfovy;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(radii[i] > 0)) continue;
    const R* t = Tm + 9 * (size_t)i;
    R* gT = dL_dtransMat + 9 * (size_t)i;
    {
      V3<R> Tu{t[0], t[1], t[2]}, Tv{t[3], t[4], t[5]}, Tw{t[6], t[7], t[8]};
      V3<R> sgn{(R)1, (R)1, (R)-1};
      const R gx2 = dL_dmean2D[3 * (size_t)i], gy2 = dL_dmean2D[3 * (size_t)i + 1];
      R d = dot(sgn, Tw * Tw);
      V3<R> f = sgn * ((R)1 / d);
      V3<R> dT0 = gx2 * (f * Tw), dT1 = gy2 * (f * Tw);
      V3<R> dT3 = gx2 * (f * Tu) + gy2 * (f * Tv);
      V3<R> dL_df = gx2 * (Tu * Tw) + gy2 * (Tv * Tw);
      R dL_dd = (R)((double)dot(dL_df, f) * (-1.0 / (double)d));
      V3<R> dd_dT3 = (sgn * Tw) * (R)2;
      dT3 = dT3 + dL_dd * dd_dT3;
      gT[0] += dT0.x; gT[1] += dT0.y; gT[2] += dT0.z;
      gT[3] += dT1.x; gT[4] += dT1.y; gT[5] += dT1.z;
      gT[6] += dT3.x; gT[7] += dT3.y; gT[8] += dT3.z;
      const R z = t[8];
      dL_dmean2D[3 * (size_t)i + 0] = gT[2] * z * Wn;   // the densification "projected gradient" (:645-648)
      dL_dmean2D[3 * (size_t)i + 1] = gT[5] * z * Hn;
    }
    const R* view = c.view;
    const R fx = focal_x, fy = focal_y, cxx = focal_x * c.tanfovx, cyy = focal_y * c.tanfovy;
    const R* q = c.rotations + 4 * (size_t)i;
    const R* sc = c.scales + 2 * (size_t)i;
    V3<R> pw{c.means3D[3 * i], c.means3D[3 * i + 1], c.means3D[3 * i + 2]};
    Rot<R> Rm = quat_to_rot(q);
    V3<R> pv = mul_W(view, pw) + V3<R>{view[12], view[13], view[14]};
    V3<R> dM[3];
    for (int k = 0; k < 3; k++) dM[k] = {fx * gT[k], fy * gT[3 + k], cxx * gT[k] + cyy * gT[3 + k] + gT[6 + k] + (R)0};
    V3<R> dRS0 = mul_Wt(view, dM[0]), dRS1 = mul_Wt(view, dM[1]), dpw = mul_Wt(view, dM[2]);
    V3<R> dtn = mul_Wt(view, V3<R>{dL_dnormal[3 * (size_t)i], dL_dnormal[3 * (size_t)i + 1], dL_dnormal[3 * (size_t)i + 2]});
    V3<R> tn = mul_W(view, Rm.c2);
    R cs = dot(V3<R>{-tn.x, -tn.y, -tn.z}, pv);
    dtn = dtn * (cs > 0 ? (R)1 : (R)-1);
    V3<R> vR0 = dRS0 * sc[0], vR1 = dRS1 * sc[1], vR2 = dtn;   // columns of dL/dR
    {
      R s = (R)1 / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      R w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
      // v[c][r]: column c, row r  (auxiliary.h:213-257)
      const R v00 = vR0.x, v01 = vR0.y, v02 = vR0.z, v10 = vR1.x, v11 = vR1.y, v12 = vR1.z, v20 = vR2.x, v21 = vR2.y, v22 = vR2.z;
      R* o = dL_drot + 4 * (size_t)i;
      o[0] = (R)2 * (x * (v12 - v21) + y * (v20 - v02) + z * (v01 - v10));
      o[1] = (R)2 * ((R)-2 * x * (v11 + v22) + y * (v01 + v10) + z * (v02 + v20) + w * (v12 - v21));
      o[2] = (R)2 * (x * (v01 + v10) - (R)2 * y * (v00 + v22) + z * (v12 + v21) + w * (v20 - v02));
      o[3] = (R)2 * (x * (v02 + v20) + y * (v12 + v21) - (R)2 * z * (v00 + v11) + w * (v01 - v10));
    }
    dL_dscale[2 * (size_t)i] = dot(dRS0, Rm.c0);
    dL_dscale[2 * (size_t)i + 1] = dot(dRS1, Rm.c1);
    dL_dmean3D[3 * (size_t)i] = dpw.x; dL_dmean3D[3 * (size_t)i + 1] = dpw.y; dL_dmean3D[3 * (size_t)i + 2] = dpw.z;

    if (c.shs) {   // backward.cu:20-139
      const int M = c.M, deg = c.D;
      V3<R> dir_o = pw - V3<R>{c.campos[0], c.campos[1], c.campos[2]};
      R len = std::sqrt(dot(dir_o, dir_o));
      V3<R> dir{dir_o.x / len, dir_o.y / len, dir_o.z / len};
      const R* sh = c.shs + (size_t)i * M * 3;
      auto S = [&](int k) { return V3<R>{sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]}; };
      V3<R> g{dL_dcolor[3 * (size_t)i], dL_dcolor[3 * (size_t)i + 1], dL_dcolor[3 * (size_t)i + 2]};
      g.x *= clamped[3 * i] ? 0 : 1; g.y *= clamped[3 * i + 1] ? 0 : 1; g.z *= clamped[3 * i + 2] ? 0 : 1;
      R* out = dL_dsh + (size_t)i * M * 3;
      auto put = [&](int k, R b) { out[3 * k] = b * g.x; out[3 * k + 1] = b * g.y; out[3 * k + 2] = b * g.z; };
      V3<R> dx{0, 0, 0}, dy{0, 0, 0}, dz{0, 0, 0};
      R x = dir.x, y = dir.y, z = dir.z;
      put(0, (R)C0);
      if (deg > 0) {
        put(1, -(R)C1 * y); put(2, (R)C1 * z); put(3, -(R)C1 * x);
        dx = -(R)C1 * S(3); dy = -(R)C1 * S(1); dz = (R)C1 * S(2);
        if (deg > 1) {
          R xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          put(4, (R)C2[0] * xy); put(5, (R)C2[1] * yz); put(6, (R)C2[2] * ((R)2 * zz - xx - yy));
          put(7, (R)C2[3] * xz); put(8, (R)C2[4] * (xx - yy));
          dx = dx + ((R)C2[0] * y) * S(4) + ((R)C2[2] * (R)2 * -x) * S(6) + ((R)C2[3] * z) * S(7) + ((R)C2[4] * (R)2 * x) * S(8);
          dy = dy + ((R)C2[0] * x) * S(4) + ((R)C2[1] * z) * S(5) + ((R)C2[2] * (R)2 * -y) * S(6) + ((R)C2[4] * (R)2 * -y) * S(8);
          dz = dz + ((R)C2[1] * y) * S(5) + ((R)C2[2] * (R)2 * (R)2 * z) * S(6) + ((R)C2[3] * x) * S(7);
          if (deg > 2) {
            put(9, (R)C3[0] * y * ((R)3 * xx - yy)); put(10, (R)C3[1] * xy * z);
            put(11, (R)C3[2] * y * ((R)4 * zz - xx - yy)); put(12, (R)C3[3] * z * ((R)2 * zz - (R)3 * xx - (R)3 * yy));
            put(13, (R)C3[4] * x * ((R)4 * zz - xx - yy)); put(14, (R)C3[5] * z * (xx - yy));
            put(15, (R)C3[6] * x * (xx - (R)3 * yy));
            dx = dx + ((R)C3[0] * (R)3 * (R)2 * xy) * S(9) + ((R)C3[1] * yz) * S(10) + ((R)C3[2] * (R)-2 * xy) * S(11) +
                 ((R)C3[3] * (R)-3 * (R)2 * xz) * S(12) + ((R)C3[4] * ((R)-3 * xx + (R)4 * zz - yy)) * S(13) +
                 ((R)C3[5] * (R)2 * xz) * S(14) + ((R)C3[6] * (R)3 * (xx - yy)) * S(15);
            dy = dy + ((R)C3[0] * (R)3 * (xx - yy)) * S(9) + ((R)C3[1] * xz) * S(10) +
                 ((R)C3[2] * ((R)-3 * yy + (R)4 * zz - xx)) * S(11) + ((R)C3[3] * (R)-3 * (R)2 * yz) * S(12) +
                 ((R)C3[4] * (R)-2 * xy) * S(13) + ((R)C3[5] * (R)-2 * yz) * S(14) + ((R)C3[6] * (R)-3 * (R)2 * xy) * S(15);
            dz = dz + ((R)C3[1] * xy) * S(10) + ((R)C3[2] * (R)4 * (R)2 * yz) * S(11) +
                 ((R)C3[3] * (R)3 * ((R)2 * zz - xx - yy)) * S(12) + ((R)C3[4] * (R)4 * (R)2 * xz) * S(13) +
                 ((R)C3[5] * (xx - yy)) * S(14);
          }
        }
      }
      V3<R> ddir{dot(dx, g), dot(dy, g), dot(dz, g)};
      // gradient through dir = v/|v|  (auxiliary.h:125-135)
      R sum2 = dot(dir_o, dir_o);
      R inv32 = (R)1 / std::sqrt(sum2 * sum2 * sum2);
      V3<R> v = dir_o;
      V3<R> dm{((sum2 - v.x * v.x) * ddir.x - v.y * v.x * ddir.y - v.z * v.x * ddir.z) * inv32,
               (-v.x * v.y * ddir.x + (sum2 - v.y * v.y) * ddir.y - v.z * v.y * ddir.z) * inv32,
               (-v.x * v.z * ddir.x - v.y * v.z * ddir.y + (sum2 - v.z * v.z) * ddir.z) * inv32};
      dL_dmean3D[3 * (size_t)i] += dm.x; dL_dmean3D[3 * (size_t)i + 1] += dm.y; dL_dmean3D[3 * (size_t)i + 2] += dm.z;
    }
  }
}

#define ORC_API extern "C" __attribute__((visibility("default")))

#define DEFINE_API(SUFFIX, REAL)                                                                                   \
  ORC_API int64_t orc_forward_##SUFFIX(                                                                            \
      int P, int D, int M, int W, int H, const REAL* bg, const REAL* means3D, const REAL* shs,                     \
      const REAL* colors_precomp, const REAL* opacities, const REAL* scales, const REAL* rotations,                \
      const REAL* transMat_precomp, const REAL* view, const REAL* proj, const REAL* campos, REAL tanfovx,          \
      REAL tanfovy, int* radii, REAL* means2D, REAL* depths, REAL* transMat, REAL* normal_opacity, REAL* rgb,      \
      uint8_t* clamped, uint32_t* tiles_touched, uint32_t* point_offsets, REAL* out_color, REAL* out_others,       \
      REAL* final_T, uint32_t* n_contrib, uint32_t* ranges, int64_t cap, uint64_t* keys_unsorted,                  \
      uint32_t* vals_unsorted, uint64_t* keys_sorted, uint32_t* point_list) {                                      \
    Ctx<REAL> c{P, D, M, W, H, bg, means3D, shs, colors_precomp, opacities, scales, rotations, transMat_precomp,   \
                view, proj, campos, tanfovx, tanfovy};                                                             \
    return forward_impl<REAL>(c, radii, means2D, depths, transMat, normal_opacity, rgb, clamped, tiles_touched,    \
                              point_offsets, out_color, out_others, final_T, n_contrib, ranges, cap,               \
                              keys_unsorted, vals_unsorted, keys_sorted, point_list);                              \
  }                                                                                                                \
  ORC_API void orc_backward_##SUFFIX(                                                                              \
      int P, int D, int M, int W, int H, int64_t Rn, const REAL* bg, const REAL* means3D, const REAL* shs,         \
      const REAL* colors_precomp, const REAL* scales, const REAL* rotations, const REAL* transMat_precomp,         \
      const REAL* view, const REAL* proj, const REAL* campos, REAL tanfovx, REAL tanfovy, const int* radii,        \
      const REAL* means2D, const REAL* transMat, const REAL* normal_opacity, const REAL* rgb,                      \
      const uint8_t* clamped, const REAL* final_T, const uint32_t* n_contrib, const uint32_t* ranges,              \
      const uint32_t* point_list, const REAL* dL_dpix, const REAL* dL_dothers, REAL* dL_dmean2D,                   \
      REAL* dL_dnormal, REAL* dL_dopacity, REAL* dL_dcolor, REAL* dL_dmean3D, REAL* dL_dtransMat, REAL* dL_dsh,    \
      REAL* dL_dscale, REAL* dL_drot) {                                                                            \
    Ctx<REAL> c{P, D, M, W, H, bg, means3D, shs, colors_precomp, nullptr, scales, rotations, transMat_precomp,     \
                view, proj, campos, tanfovx, tanfovy};                                                             \
    backward_impl<REAL>(c, Rn, radii, means2D, transMat, normal_opacity, rgb, clamped, final_T, n_contrib,         \
                        ranges, point_list, dL_dpix, dL_dothers, dL_dmean2D, dL_dnormal, dL_dopacity, dL_dcolor,   \
                        dL_dmean3D, dL_dtransMat, dL_dsh, dL_dscale, dL_drot);                                     \
  }

DEFINE_API(f32, float)
DEFINE_API(f64, double)

ORC_API void orc_mark_visible_f32(int P, const float* means3D, const float* view, uint8_t* present) {
  // rasterizer_impl.cu:54-66 + auxiliary.h:160-185
  for (int i = 0; i < P; i++) {
    const float* p = means3D + 3 * i;
    float z = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
    present[i] = !(z <= 0.2f);
  }
}

ORC_API int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
