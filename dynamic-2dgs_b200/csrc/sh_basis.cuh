// Real spherical-harmonics basis up to degree 3 (shared by the surfel and the 3-D Gaussian rasterizers).
#pragma once
#include "raster_common.cuh"

namespace d2gs {

// SH basis evaluation, degree <= 3 (reference: forward.cu:20-71).  Returns the colour before "+0.5 / clamp".
__device__ __forceinline__ v3 eval_sh(int deg, v3 dir, const float* sh) {
  auto S = [&](int k) { return v3{sh[3 * k], sh[3 * k + 1], sh[3 * k + 2]}; };
  v3 result = kSH_C0 * S(0);
  if (deg > 0) {
    float x = dir.x, y = dir.y, z = dir.z;
    result = result - kSH_C1 * y * S(1) + kSH_C1 * z * S(2) - kSH_C1 * x * S(3);
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      result = result + kSH_C2[0] * xy * S(4) + kSH_C2[1] * yz * S(5) + kSH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
               kSH_C2[3] * xz * S(7) + kSH_C2[4] * (xx - yy) * S(8);
      if (deg > 2) {
        result = result + kSH_C3[0] * y * (3.0f * xx - yy) * S(9) + kSH_C3[1] * xy * z * S(10) +
                 kSH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                 kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                 kSH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + kSH_C3[5] * z * (xx - yy) * S(14) +
                 kSH_C3[6] * x * (xx - 3.0f * yy) * S(15);
      }
    }
  }
  return result;
}

// Basis function k (0..15) at the unit direction (x, y, z) and its gradient w.r.t. the direction:
//   colour = sum_k b_k(dir) sh_k;   dL/dsh_k = b_k gc;   dL/ddir += grad b_k (sh_k . gc)
// (the analytic derivatives of forward.cu:20-71; backward.cu:20-140 of both reference rasterizers expand the same sums).
__device__ __forceinline__ void sh_basis_grad(int k, float x, float y, float z, float& bk, float& bx, float& by, float& bz) {
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  switch (k) {
    case 0: bk = kSH_C0; bx = 0.f; by = 0.f; bz = 0.f; break;
    case 1: bk = -kSH_C1 * y; bx = 0.f; by = -kSH_C1; bz = 0.f; break;
    case 2: bk = kSH_C1 * z; bx = 0.f; by = 0.f; bz = kSH_C1; break;
    case 3: bk = -kSH_C1 * x; bx = -kSH_C1; by = 0.f; bz = 0.f; break;
    case 4: bk = kSH_C2[0] * xy; bx = kSH_C2[0] * y; by = kSH_C2[0] * x; bz = 0.f; break;
    case 5: bk = kSH_C2[1] * yz; bx = 0.f; by = kSH_C2[1] * z; bz = kSH_C2[1] * y; break;
    case 6: bk = kSH_C2[2] * (2.f * zz - xx - yy); bx = kSH_C2[2] * -2.f * x; by = kSH_C2[2] * -2.f * y; bz = kSH_C2[2] * 4.f * z; break;
    case 7: bk = kSH_C2[3] * xz; bx = kSH_C2[3] * z; by = 0.f; bz = kSH_C2[3] * x; break;
    case 8: bk = kSH_C2[4] * (xx - yy); bx = kSH_C2[4] * 2.f * x; by = kSH_C2[4] * -2.f * y; bz = 0.f; break;
    case 9: bk = kSH_C3[0] * y * (3.f * xx - yy); bx = kSH_C3[0] * 6.f * xy; by = kSH_C3[0] * 3.f * (xx - yy); bz = 0.f; break;
    case 10: bk = kSH_C3[1] * xy * z; bx = kSH_C3[1] * yz; by = kSH_C3[1] * xz; bz = kSH_C3[1] * xy; break;
    case 11: bk = kSH_C3[2] * y * (4.f * zz - xx - yy); bx = kSH_C3[2] * -2.f * xy; by = kSH_C3[2] * (-3.f * yy + 4.f * zz - xx); bz = kSH_C3[2] * 8.f * yz; break;
    case 12: bk = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); bx = kSH_C3[3] * -6.f * xz; by = kSH_C3[3] * -6.f * yz; bz = kSH_C3[3] * 3.f * (2.f * zz - xx - yy); break;
    case 13: bk = kSH_C3[4] * x * (4.f * zz - xx - yy); bx = kSH_C3[4] * (-3.f * xx + 4.f * zz - yy); by = kSH_C3[4] * -2.f * xy; bz = kSH_C3[4] * 8.f * xz; break;
    case 14: bk = kSH_C3[5] * z * (xx - yy); bx = kSH_C3[5] * 2.f * xz; by = kSH_C3[5] * -2.f * yz; bz = kSH_C3[5] * (xx - yy); break;
    default: bk = kSH_C3[6] * x * (xx - 3.f * yy); bx = kSH_C3[6] * 3.f * (xx - yy); by = kSH_C3[6] * -6.f * xy; bz = 0.f; break;
  }
}

}  // namespace d2gs
