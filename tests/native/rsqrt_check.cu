// Test helper (not product code): runs soda::RecipSqrtF32 over EVERY positive
// normal float x for a given numerator a and compares the float-arithmetic
// decision with the FP64 evaluation it stands for.
//   rsqrt_check(a, &mismatches, &undecided) -> cudaError_t as int
#include "soda_cuda_device.cuh"

namespace {

__global__ void sweep(float a, unsigned first, unsigned long long count,
                      unsigned long long* mismatches,
                      unsigned long long* undecided) {
  unsigned long long bad = 0, open = 0;
  for (unsigned long long n = blockIdx.x * 1ull * blockDim.x + threadIdx.x;
       n < count; n += 1ull * gridDim.x * blockDim.x) {
    const soda::RecipSqrtF32 v{a, __uint_as_float(first + unsigned(n))};
    float fast;
    const bool decided = v.decide(&fast);
    const float want = static_cast<float>(v.exact());
    if (!decided) ++open;
    else if (__float_as_uint(fast) != __float_as_uint(want)) ++bad;
    // the product entry point must agree always
    if (__float_as_uint(v.to_float()) != __float_as_uint(want)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
  if (open) atomicAdd(undecided, open);
}

}  // namespace

extern "C" int rsqrt_check(float a, unsigned first_bits,
                           unsigned long long count,
                           unsigned long long* mismatches,
                           unsigned long long* undecided) {
  unsigned long long* dev = nullptr;
  cudaError_t rc = cudaMalloc(&dev, 16);
  if (rc != cudaSuccess) return rc;
  cudaMemset(dev, 0, 16);
  sweep<<<148 * 8, 256>>>(a, first_bits, count, dev, dev + 1);
  rc = cudaDeviceSynchronize();
  unsigned long long host[2] = {0, 0};
  if (rc == cudaSuccess) rc = cudaMemcpy(host, dev, 16, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  *mismatches = host[0];
  *undecided = host[1];
  return rc;
}
