// Probe (not product code): can the deformation MLP's trunk run on tensor cores within the 1e-4 parity bar, and how fast?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mlp_tc_probe tools/mlp_tc_probe.cu && /tmp/mlp_tc_probe
// Workload = the trunk of DeformNetwork at C3: 512 rows (control nodes), 8 dense layers 256 -> 256 with bias + ReLU
// (utils/time_utils.py:410-453).  Kernel: 32 CTAs x 8 warps, one m16 row tile per CTA kept in shared memory across the
// layers, warp w owns output columns [32w, 32w+32), mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate).  Weights are read straight from L2 (each CTA streams all 2 MB,
// as the product kernel's TMA ring does).  Prints time per pass and the error against an fp64 evaluation, next to the
// error of a plain fp32 evaluation (the yardstick of tests/test_deform_gpu.py::test_fused_mlp_matches_eager_layers).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int ROWS = 512, W = 256, L = 8, LDA = W + 8;   // +8 floats: conflict-free fragment loads

__device__ __forceinline__ uint32_t tf32_hi(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int SPLIT>   // 1: plain TF32, 3: 3xTF32
__global__ void __launch_bounds__(256) trunk_tc(const float* __restrict__ x, const float* __restrict__ Wt /*[L][n][k]*/,
                                                const float* __restrict__ bias, float* __restrict__ y) {
  __shared__ float act[2][16][LDA];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row0 = blockIdx.x * 16;
  for (int i = tid; i < 16 * W; i += 256) act[0][i / W][i % W] = x[(size_t)(row0 + i / W) * W + i % W];
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  int cur = 0;
  for (int l = 0; l < L; l++) {
    const float* Wl = Wt + (size_t)l * W * W;
    float c[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f; }
#pragma unroll 4
    for (int k0 = 0; k0 < W; k0 += 8) {
      // k permutation inside the 8-step: fragment slots (t, t+4) hold memory columns (2t, 2t+1) for A and B alike
      const float2 a_lo_row = *reinterpret_cast<const float2*>(&act[cur][g][k0 + 2 * t]);
      const float2 a_hi_row = *reinterpret_cast<const float2*>(&act[cur][g + 8][k0 + 2 * t]);
      const float av[4] = {a_lo_row.x, a_hi_row.x, a_lo_row.y, a_hi_row.y};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { ah[i] = tf32_hi(av[i]); al[i] = tf32_hi(av[i] - __uint_as_float(ah[i])); }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int n = warp * 32 + nt * 8 + g;
        const float2 bv = __ldg(reinterpret_cast<const float2*>(Wl + (size_t)n * W + k0 + 2 * t));
        uint32_t bh[2] = {tf32_hi(bv.x), tf32_hi(bv.y)};
        mma_tf32(c[nt], ah, bh);
        if (SPLIT == 3) {
          uint32_t bl[2] = {tf32_hi(bv.x - __uint_as_float(bh[0])), tf32_hi(bv.y - __uint_as_float(bh[1]))};
          mma_tf32(c[nt], al, bh);
          mma_tf32(c[nt], ah, bl);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int n = warp * 32 + nt * 8 + 2 * t;
      const float b0 = bias[l * W + n], b1 = bias[l * W + n + 1];
      act[cur ^ 1][g][n] = fmaxf(c[nt][0] + b0, 0.f); act[cur ^ 1][g][n + 1] = fmaxf(c[nt][1] + b1, 0.f);
      act[cur ^ 1][g + 8][n] = fmaxf(c[nt][2] + b0, 0.f); act[cur ^ 1][g + 8][n + 1] = fmaxf(c[nt][3] + b1, 0.f);
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int i = tid; i < 16 * W; i += 256) y[(size_t)(row0 + i / W) * W + i % W] = act[cur][i / W][i % W];
}

int main() {
  std::vector<float> x(ROWS * W), Wt((size_t)L * W * W), b(L * W);
  srand(3);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& v : x) v = rnd();
  for (auto& v : Wt) v = rnd() * 0.108f;          // ~ kaiming-uniform bound sqrt(6/256)/sqrt(2): activations stay O(1)
  for (auto& v : b) v = rnd() * 0.06f;
  std::vector<double> ref(x.begin(), x.end()), tmp(ROWS * W);
  std::vector<float> r32(x), t32(ROWS * W);
  for (int l = 0; l < L; l++) {
    for (int r = 0; r < ROWS; r++)
      for (int n = 0; n < W; n++) {
        double s = b[l * W + n]; float s32 = 0.f;
        for (int k = 0; k < W; k++) { s += ref[r * W + k] * (double)Wt[((size_t)l * W + n) * W + k]; s32 = fmaf(r32[r * W + k], Wt[((size_t)l * W + n) * W + k], s32); }
        tmp[r * W + n] = s > 0 ? s : 0; t32[r * W + n] = fmaxf(s32 + b[l * W + n], 0.f);
      }
    ref = tmp; r32 = t32;
  }
  float *dx, *dW, *db, *dy;
  cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dW, Wt.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dy, x.size() * 4);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, Wt.data(), Wt.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  auto err = [&](const std::vector<float>& y) {
    double num = 0, den = 0;
    for (size_t i = 0; i < y.size(); i++) { num += (y[i] - ref[i]) * (y[i] - ref[i]); den += ref[i] * ref[i]; }
    return sqrt(num / den);
  };
  printf("fp32 fma chain (CPU) vs fp64: rel. L2 error %.3e\n", err(r32));
  for (int split : {1, 3}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 5; it++) { if (split == 1) trunk_tc<1><<<ROWS / 16, 256>>>(dx, dW, db, dy); else trunk_tc<3><<<ROWS / 16, 256>>>(dx, dW, db, dy); }
    cudaEventRecord(e0);
    const int reps = 50;
    for (int it = 0; it < reps; it++) { if (split == 1) trunk_tc<1><<<ROWS / 16, 256>>>(dx, dW, db, dy); else trunk_tc<3><<<ROWS / 16, 256>>>(dx, dW, db, dy); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> y(x.size());
    cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost);
    printf("%dxTF32 mma.sync trunk (8 layers, 512 rows, 32 CTAs): %.2f us per pass, rel. L2 error vs fp64 %.3e  [%s]\n", split, ms * 1e3 / reps, err(y),
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
