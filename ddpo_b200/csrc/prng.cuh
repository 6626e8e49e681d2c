// JAX-compatible threefry2x32 PRNG pieces shared by the scheduler (ddim.cu) and RWR (rwr.cu) kernels.
// 3P jax==0.4.8 jax/_src/prng.py (threefry_2x32, random_bits) and jax/_src/random.py (normal, randint);
// reference call sites: pipeline_flax_stable_diffusion.py:196-197,232,252; scheduling_ddim_flax.py:347;
// ddpo/training/diffusion.py:14-29.
#pragma once
#include "common.cuh"

namespace ddpo {

// ------------------------------------------------------------------ threefry ----
struct u32x2 {
  uint32_t a, b;
};
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ u32x2 threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, rot[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + static_cast<uint32_t>(i + 1);
  }
  return {x0, x1};
}

// bits of element i of random_bits(key, n): counters are iota(n) padded to even and split in halves
__device__ __forceinline__ uint32_t random_bits_at(uint32_t k0, uint32_t k1, uint32_t i, uint32_t half,
                                                   uint32_t n) {
  if (i < half) {
    uint32_t hi = i + half;
    u32x2 r = threefry2x32(k0, k1, i, hi < n ? hi : 0u);  // odd n: padded counter is 0
    return r.a;
  }
  u32x2 r = threefry2x32(k0, k1, i - half, i);
  return r.b;
}

__device__ __forceinline__ float erfinv_xla(float x) {
  // XLA ErfInv (float32): Giles' polynomial, evaluated without FMA contraction to match the
  // CPU oracle / XLA:CPU op order
  float w = -log1pf(-__fmul_rn(x, x));
  const bool lt = w < 5.0f;
  w = lt ? __fadd_rn(w, -2.5f) : __fadd_rn(sqrtf(w), -3.0f);
  float p = lt ? 2.81022636e-08f : -0.000200214257f;
  p = __fadd_rn(lt ? 3.43273939e-07f : 0.000100950558f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -3.5233877e-06f : 0.00134934322f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -4.39150654e-06f : -0.00367342844f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? 0.00021858087f : 0.00573950773f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -0.00125372503f : -0.0076224613f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -0.00417768164f : 0.00943887047f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? 0.246640727f : 1.00167406f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? 1.50140941f : 2.83297682f, __fmul_rn(p, w));
  return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : __fmul_rn(p, x);
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float scale = 1.0f - lo;  // rounds to 2.0f exactly as in float32 numpy / XLA
  float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, scale), lo));
  return __fmul_rn(1.41421356237309504880f, erfinv_xla(u));
}

}  // namespace ddpo
