// Test helper (not product code): soda::div_try(float, float, rare) against
// __fdiv_rn.  Two sweeps per call:
//   consecutive  b = bits(first_b) + n for n < count, numerator a fixed;
//   random       (a, b) from a counter hash, exponents spread over
//                [2^-45, 2^45] (so both sides of the quick sequence's range
//                test occur), every mantissa bit random.
// Wherever div_try does not raise `rare` its result must be __fdiv_rn's bit
// for bit; `rare` must be raised for every operand outside [2^-40, 2^41).
//   div_check(a, first_b_bits, count, random_count, &mismatches, &rare, &missed)
#include "soda_cuda_device.cuh"

namespace {

__device__ unsigned mix(unsigned long long n) {
  n ^= n >> 33; n *= 0xff51afd7ed558ccdull;
  n ^= n >> 33; n *= 0xc4ceb9fe1a85ec53ull;
  n ^= n >> 33;
  return static_cast<unsigned>(n);
}

__device__ float random_operand(unsigned long long n) {
  const unsigned h = mix(n);
  const unsigned exponent = 127u - 45u + (mix(n ^ 0x9e3779b97f4a7c15ull) % 91u);
  return __uint_as_float((h & 0x807fffffu) | (exponent << 23));
}

__device__ void one(float a, float b, unsigned long long* bad,
                    unsigned long long* rare_count,
                    unsigned long long* missed) {
  bool rare = false;
  const float got = soda::div_try(a, b, rare);
  const float want = __fdiv_rn(a, b);
  const float lo = 9.094947017729282e-13f, hi = 2.199023255552e12f;  // 2^-40, 2^41
  const bool inside = fabsf(a) >= lo && fabsf(a) < hi && fabsf(b) >= lo &&
                      fabsf(b) < hi;
  if (rare) ++*rare_count;
  else if (__float_as_uint(got) != __float_as_uint(want)) ++*bad;
  if (!inside && !rare) ++*missed;
}

__global__ void sweep(float a, unsigned first, unsigned long long count,
                      unsigned long long random_count,
                      unsigned long long* out) {
  unsigned long long bad = 0, rare = 0, missed = 0;
  const unsigned long long tid = blockIdx.x * 1ull * blockDim.x + threadIdx.x;
  const unsigned long long step = 1ull * gridDim.x * blockDim.x;
  for (unsigned long long n = tid; n < count; n += step) {
    one(a, __uint_as_float(first + unsigned(n)), &bad, &rare, &missed);
    one(__uint_as_float(first + unsigned(n)), a, &bad, &rare, &missed);
  }
  for (unsigned long long n = tid; n < random_count; n += step)
    one(random_operand(2 * n), random_operand(2 * n + 1), &bad, &rare, &missed);
  if (bad) atomicAdd(out, bad);
  if (rare) atomicAdd(out + 1, rare);
  if (missed) atomicAdd(out + 2, missed);
}

}  // namespace

extern "C" int div_check(float a, unsigned first_bits, unsigned long long count,
                         unsigned long long random_count,
                         unsigned long long* mismatches,
                         unsigned long long* rare,
                         unsigned long long* missed) {
  unsigned long long* dev = nullptr;
  cudaError_t rc = cudaMalloc(&dev, 24);
  if (rc != cudaSuccess) return rc;
  cudaMemset(dev, 0, 24);
  sweep<<<148 * 8, 256>>>(a, first_bits, count, random_count, dev);
  rc = cudaDeviceSynchronize();
  unsigned long long host[3] = {0, 0, 0};
  if (rc == cudaSuccess) rc = cudaMemcpy(host, dev, 24, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  *mismatches = host[0];
  *rare = host[1];
  *missed = host[2];
  return rc;
}
