/* Test helper (not product code): C model of soda::RecipSqrtF32::decide
 * (csrc/soda_cuda_device.cuh) with the hardware's MUFU.RSQ replaced by the
 * correctly rounded 1/sqrt(x) pushed `skew` ulps off — the algorithm may only
 * rely on the approximation's error bound, not on its bits.
 *   rsqrt_model(a, first_bits, count, skew, &mismatches, &undecided)
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (fmaf from libm). */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static int decide(float a, float x, int skew, float* result) {
  float y0 = (float)(1.0 / sqrt((double)x));
  y0 = from_bits(to_bits(y0) + skew);
  const float t = x * y0;
  const float t_err = fmaf(x, y0, -t);
  float rho = fmaf(-t, y0, 1.0f);
  rho = fmaf(-t_err, y0, rho);
  const float c = rho * fmaf(rho, 0.375f, 0.5f);
  const float q = a * y0;
  const float q_err = fmaf(a, y0, -q);
  const float d = fmaf(q, c, q_err);
  const float r = q + d;
  const float e = d + -(r + -q);
  const uint32_t bits = to_bits(r);
  const float edge = from_bits((bits & 0x7f800000u) - (24u << 23)) * 0.99999f;
  const float ax = fabsf(a);
  *result = r;
  return fabsf(e) < edge && (bits & 0x007fffffu) != 0u && x > 1e-30f &&
         x < 1e30f && ax > 1e-15f && ax < 1e15f;
}

int rsqrt_model(float a, uint32_t first_bits, uint64_t count, int skew,
                uint64_t* mismatches, uint64_t* undecided) {
  uint64_t bad = 0, open = 0;
#pragma omp parallel for reduction(+ : bad, open)
  for (uint64_t n = 0; n < count; ++n) {
    const float x = from_bits(first_bits + (uint32_t)n);
    float fast;
    const double exact = (double)a / sqrt((double)x);
    if (!decide(a, x, skew, &fast)) ++open;
    else if (to_bits(fast) != to_bits((float)exact)) ++bad;
  }
  *mismatches = bad;
  *undecided = open;
  return 0;
}
