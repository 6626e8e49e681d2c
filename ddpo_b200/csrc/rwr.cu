// Reward-weighted-regression (RWR) training step pieces around the U-Net:
//   * posterior sample of the stored VAE moments + forward diffusion (add_noise) in one pass
//   * MSE on the classifier-free-guidance-combined prediction, optional per-sample weights, and its gradient
//
// Reference semantics:
//   ddpo/training/diffusion.py:16-43  (FlaxDiagonalGaussianDistribution.sample, *0.18215, noise, add_noise)
//   ddpo/training/diffusion.py:62-90  (CFG combine, per-sample MSE, mean / weighted sum)
//   3P diffusers==0.12.1 vae_flax.py FlaxDiagonalGaussianDistribution (logvar clipped to [-30, 20]),
//   scheduling_utils_flax.py add_noise_common (sqrt(a_t) x + sqrt(1 - a_t) n), 3P jax.random.normal
//
// HBM/latency bound (16 K elements per sample): one pass, 128-bit accesses where the layout allows, noise
// generated in registers, fixed-order reductions (no float atomics -> bit-reproducible loss).
#include "common.cuh"
#include "prng.cuh"

namespace ddpo {

constexpr int RWR_THREADS = 256;

// One thread per (b, y, x): reads the 2C moments of the pixel (NHWC), draws the C posterior normals (counter =
// NHWC flat index, as jax.random.normal(key, mean.shape) on the channels-last tensor) and the C noise normals
// (counter = NCHW flat index, jax.random.normal(noise_rng, latents.shape)), writes noise / noisy latents NCHW.
template <int C>
__global__ void __launch_bounds__(RWR_THREADS) rwr_noisy_latents_kernel(
    const float* __restrict__ moments, const uint32_t* __restrict__ key_sample, const uint32_t* __restrict__ key_noise,
    const int32_t* __restrict__ timesteps, const float* __restrict__ alphas_cumprod, float scaling, int B, int HW,
    float* __restrict__ noise_out, float* __restrict__ noisy_out, float* __restrict__ latents_out) {
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * RWR_THREADS + threadIdx.x;
  if (pix >= static_cast<int64_t>(B) * HW) return;
  const int b = static_cast<int>(pix / HW), hw = static_cast<int>(pix % HW);
  const uint32_t ntot = static_cast<uint32_t>(B) * C * HW;
  const uint32_t half = (ntot + 1) / 2;
  const uint32_t ks0 = key_sample[0], ks1 = key_sample[1], kn0 = key_noise[0], kn1 = key_noise[1];
  const float a_t = alphas_cumprod[timesteps[b]];
  const float sa = sqrtf(a_t), sb = sqrtf(1.0f - a_t);
  float mom[2 * C];
  const float4* mp = reinterpret_cast<const float4*>(moments + pix * 2 * C);
#pragma unroll
  for (int i = 0; i < 2 * C / 4; ++i) {
    const float4 v = mp[i];
    mom[4 * i] = v.x, mom[4 * i + 1] = v.y, mom[4 * i + 2] = v.z, mom[4 * i + 3] = v.w;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float logvar = fminf(fmaxf(mom[C + c], -30.0f), 20.0f);
    const float std = expf(0.5f * logvar);
    const uint32_t i_nhwc = static_cast<uint32_t>(pix) * C + c;
    const uint32_t i_nchw = (static_cast<uint32_t>(b) * C + c) * HW + hw;
    const float z = bits_to_normal(random_bits_at(ks0, ks1, i_nhwc, half, ntot));
    const float nz = bits_to_normal(random_bits_at(kn0, kn1, i_nchw, half, ntot));
    const float lat = __fmul_rn(__fadd_rn(mom[c], __fmul_rn(std, z)), scaling);
    noise_out[i_nchw] = nz;
    noisy_out[i_nchw] = __fadd_rn(__fmul_rn(sa, lat), __fmul_rn(sb, nz));
    if (latents_out != nullptr) latents_out[i_nchw] = lat;
  }
}

// grid (DDPO_DDIM_CHUNKS, B).  pred = e_u + g (e_c - e_u); per-sample mse = mean (noise - pred)^2;
// loss = mean_b mse_b (weights == NULL) or sum_b w_b mse_b; d pred = 2 (pred - noise) / n * (w_b or 1/B).
__global__ void __launch_bounds__(RWR_THREADS) rwr_mse_loss_kernel(
    const float* __restrict__ eps_u, const float* __restrict__ eps_c, const float* __restrict__ noise,
    const float* __restrict__ weights, float g, int B, int n, float* __restrict__ loss_out,
    float* __restrict__ per_sample, float* __restrict__ d_eu, float* __restrict__ d_ec, float* __restrict__ ws) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int per_chunk = ((n / 4 + DDPO_DDIM_CHUNKS - 1) / DDPO_DDIM_CHUNKS) * 4;
  const int begin = chunk * per_chunk;
  const int end = min(n, begin + per_chunk);
  const size_t base = static_cast<size_t>(b) * n;
  const float wb = weights != nullptr ? weights[b] : 1.0f / static_cast<float>(B);
  const float gscale = 2.0f * wb / static_cast<float>(n);
  float acc = 0.0f;
  for (int i = begin + threadIdx.x * 4; i < end; i += RWR_THREADS * 4) {
    const float4 eu = *reinterpret_cast<const float4*>(eps_u + base + i);
    const float4 ec = *reinterpret_cast<const float4*>(eps_c + base + i);
    const float4 nz = *reinterpret_cast<const float4*>(noise + base + i);
    const float e_u[4] = {eu.x, eu.y, eu.z, eu.w}, e_c[4] = {ec.x, ec.y, ec.z, ec.w}, nn[4] = {nz.x, nz.y, nz.z, nz.w};
    float du[4], dc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float pred = e_u[j] + g * (e_c[j] - e_u[j]);
      const float d = pred - nn[j];
      acc += d * d;
      const float de = d * gscale;
      dc[j] = g * de;
      du[j] = (1.0f - g) * de;
    }
    if (d_ec != nullptr) *reinterpret_cast<float4*>(d_ec + base + i) = make_float4(dc[0], dc[1], dc[2], dc[3]);
    if (d_eu != nullptr) *reinterpret_cast<float4*>(d_eu + base + i) = make_float4(du[0], du[1], du[2], du[3]);
  }
  __shared__ float warp_part[RWR_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  float* partials = ws;                                                       // [B][CHUNKS]
  float* sample_mse = ws + B * DDPO_DDIM_CHUNKS;                              // [B]
  unsigned int* counters = reinterpret_cast<unsigned int*>(sample_mse + B);   // [B] + [1]
  if (threadIdx.x == 0) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < RWR_THREADS / 32; ++w) s += warp_part[w];
    partials[b * DDPO_DDIM_CHUNKS + chunk] = s;
    __threadfence();
    const unsigned int ticket = atomicAdd(&counters[b], 1u);
    if (ticket == DDPO_DDIM_CHUNKS - 1) {
      __threadfence();
      float tot = 0.0f;
      for (int q = 0; q < DDPO_DDIM_CHUNKS; ++q) tot += *(volatile float*)&partials[b * DDPO_DDIM_CHUNKS + q];
      const float mse = tot / static_cast<float>(n);
      sample_mse[b] = mse;
      if (per_sample != nullptr) per_sample[b] = mse;
      counters[b] = 0;
      __threadfence();
      const unsigned int t2 = atomicAdd(&counters[B], 1u);
      if (t2 == static_cast<unsigned int>(B) - 1) {
        __threadfence();
        float L = 0.0f;
        for (int q = 0; q < B; ++q) {
          const float m = *(volatile float*)&sample_mse[q];
          L += weights != nullptr ? m * weights[q] : m;
        }
        loss_out[0] = weights != nullptr ? L : L / static_cast<float>(B);
        counters[B] = 0;
      }
    }
  }
}

}  // namespace ddpo

using namespace ddpo;

// jax.random.randint(key, (n,), minval, maxval) for int32 on the host (n is a batch size): see oracle/threefry.py
// and 3P jax/_src/random.py:_randint.  random_bits(k, (n,)) pairs counter i with i + ceil(n/2) (odd n: padded 0).
static uint32_t host_random_bits_at(const uint32_t k[2], uint32_t i, uint32_t n) {
  const uint32_t half = (n + 1) / 2;
  if (i < half) {
    const uint32_t hi = i + half;
    return threefry2x32(k[0], k[1], i, hi < n ? hi : 0u).a;
  }
  return threefry2x32(k[0], k[1], i - half, i).b;
}

extern "C" int ddpo_threefry_randint_host(const uint32_t key[2], int n, int32_t minval, int32_t maxval, int32_t* out) {
  DDPO_REQUIRE(key && out && n > 0, "threefry_randint_host: bad arguments");
  const u32x2 a = threefry2x32(key[0], key[1], 0u, 2u), b = threefry2x32(key[0], key[1], 1u, 3u);  // split(key)
  const uint32_t k1[2] = {a.a, b.a}, k2[2] = {a.b, b.b};
  const uint32_t span = maxval > minval ? static_cast<uint32_t>(maxval - minval) : 1u;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  for (int i = 0; i < n; ++i) {
    const uint32_t hi = host_random_bits_at(k1, static_cast<uint32_t>(i), static_cast<uint32_t>(n));
    const uint32_t lo = host_random_bits_at(k2, static_cast<uint32_t>(i), static_cast<uint32_t>(n));
    const uint32_t off = ((hi % span) * mult + (lo % span)) % span;
    out[i] = minval + static_cast<int32_t>(off);
  }
  return DDPO_OK;
}

extern "C" int64_t ddpo_rwr_workspace_floats(int batch) {
  return static_cast<int64_t>(batch) * DDPO_DDIM_CHUNKS + 2 * static_cast<int64_t>(batch) + 1;
}

extern "C" int ddpo_rwr_noisy_latents(const float* moments_nhwc, const uint32_t* key_sample_dev,
                                      const uint32_t* key_noise_dev, const int32_t* timesteps,
                                      const float* alphas_cumprod, float scaling, int batch, int channels, int h, int w,
                                      float* noise_out, float* noisy_out, float* latents_out, void* stream) {
  DDPO_REQUIRE(moments_nhwc && key_sample_dev && key_noise_dev && timesteps && alphas_cumprod && noise_out && noisy_out,
               "rwr_noisy_latents: null pointer");
  DDPO_REQUIRE(channels == 4, "rwr_noisy_latents: %d latent channels (only 4 is built)", channels);
  DDPO_REQUIRE(batch > 0 && h > 0 && w > 0 && static_cast<int64_t>(batch) * channels * h * w < (int64_t(1) << 32),
               "rwr_noisy_latents: bad shape");
  const int64_t pix = static_cast<int64_t>(batch) * h * w;
  rwr_noisy_latents_kernel<4><<<static_cast<int>((pix + RWR_THREADS - 1) / RWR_THREADS), RWR_THREADS, 0,
                                static_cast<cudaStream_t>(stream)>>>(moments_nhwc, key_sample_dev, key_noise_dev,
                                                                     timesteps, alphas_cumprod, scaling, batch, h * w,
                                                                     noise_out, noisy_out, latents_out);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}

extern "C" int ddpo_rwr_mse_loss(const float* eps_uncond, const float* eps_cond, const float* noise,
                                 const float* weights, float guidance_scale, int batch, int n, float* loss_out,
                                 float* per_sample_out, float* d_eps_uncond, float* d_eps_cond, float* workspace,
                                 void* stream) {
  DDPO_REQUIRE(eps_uncond && eps_cond && noise && loss_out && workspace, "rwr_mse_loss: null pointer");
  DDPO_REQUIRE(batch > 0 && n > 0 && n % 4 == 0, "rwr_mse_loss: batch=%d n=%d (n must be a multiple of 4)", batch, n);
  dim3 grid(DDPO_DDIM_CHUNKS, batch);
  rwr_mse_loss_kernel<<<grid, RWR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      eps_uncond, eps_cond, noise, weights, guidance_scale, batch, n, loss_out, per_sample_out, d_eps_uncond,
      d_eps_cond, workspace);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}
