// Micro-benchmark: issue rate of scalar FADD/FMUL vs packed add/mul.f32x2 on
// sm_100a (decides whether pairing fused iterations pays).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) spin(float* out, int iters, float seed) {
  float2 a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = make_float2(seed + k + threadIdx.x, seed - k);
  const float2 c = make_float2(seed * 0.5f, seed * 0.25f);
  const float2 m = make_float2(0.999f, 1.001f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) {          // scalar: 2 FADD + 2 FMUL per pair
        a[k].x = (a[k].x + c.x) * m.x;
        a[k].y = (a[k].y + c.y) * m.y;
      } else {                  // packed: 1 FADD2 + 1 FMUL2 per pair
        a[k] = __fmul2_rn(__fadd2_rn(a[k], c), m);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int blocks_per_sm = 1; blocks_per_sm <= 8; blocks_per_sm *= 2) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) spin<0><<<148 * blocks_per_sm, 256>>>(out, iters, 1.5f);
        else spin<1><<<148 * blocks_per_sm, 256>>>(out, iters, 1.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double flops = 148.0 * blocks_per_sm * 256 * iters * 8 * 4;
      printf("%s blocks/SM %d: %.3f ms  %.2f TFLOP/s (add+mul counted 1 each) "
             "= %.1f flop/clk/SM at 1.9 GHz\n", mode ? "packed" : "scalar",
             blocks_per_sm, best, flops / best / 1e9,
             flops / best / 1e-3 / 148 / 1.9e9);
    }
  }
  return 0;
}
