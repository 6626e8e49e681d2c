// VAE-decoder pieces that are not GEMM shaped (the convolutions / linears / GroupNorms of the decoder reuse
// igemm*.cu and norm.cu):
//   * latent scaling + post_quant_conv (1x1, 4 -> 4) on the NCHW latents
//   * row softmax of the single-head mid-block attention (scores materialised per sample: 4096 x 4096 fp32)
//   * conv_out (3x3, C -> 3) fused with the image post-processing (x/2 + 0.5).clip(0, 1), NHWC output
//
// Reference semantics: pipeline/policy_gradient.py:174-182 and ddpo/training/diffusion.py:105-112 (vae_decode);
// 3P diffusers==0.12.1 vae_flax.py: FlaxAutoencoderKL.decode (post_quant_conv -> FlaxDecoder), FlaxAttentionBlock
// (softmax((q s)(k s)^T), s = (C / heads)^-1/4), FlaxDecoder.conv_out.
#include "common.cuh"

namespace ddpo {

// y[b, co, p] = bias[co] + sum_ci w[ci, co] * (z[b, ci, p] * inv_scaling);  NCHW -> NCHW, C = 4
__global__ void __launch_bounds__(256) vae_post_quant_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float inv_scaling, int B,
                                                             int HW, float* __restrict__ y) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * HW) return;
  const int b = static_cast<int>(i / HW), p = static_cast<int>(i % HW);
  const float* zp = z + static_cast<size_t>(b) * 4 * HW + p;
  float in[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) in[c] = zp[static_cast<size_t>(c) * HW] * inv_scaling;
  float* yp = y + static_cast<size_t>(b) * 4 * HW + p;
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    float acc = bias[co];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) acc = fmaf(in[ci], w[ci * 4 + co], acc);
    yp[static_cast<size_t>(co) * HW] = acc;
  }
}

// P[r, :] = softmax(scale * S[r, :]) as bf16; one CTA per row, n a multiple of 4, any n (loops).
constexpr int SM_THREADS = 256;
__global__ void __launch_bounds__(SM_THREADS) softmax_rows_kernel(const float* __restrict__ S, int64_t ld_s, float scale,
                                                                  __nv_bfloat16* __restrict__ P, int64_t ld_p, int n) {
  const int64_t r = blockIdx.x;
  const float* s = S + r * ld_s;
  __nv_bfloat16* p = P + r * ld_p;
  __shared__ float red[SM_THREADS / 32];
  __shared__ float bcast;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int i = threadIdx.x * 4; i < n; i += SM_THREADS * 4) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < SM_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x * 4; i < n; i += SM_THREADS * 4) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    sum += (expf((v.x - mx) * scale) + expf((v.y - mx) * scale)) + (expf((v.z - mx) * scale) + expf((v.w - mx) * scale));
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < SM_THREADS / 32; ++w) t += red[w];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int i = threadIdx.x * 4; i < n; i += SM_THREADS * 4) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    const float e0 = expf((v.x - mx) * scale) * inv, e1 = expf((v.y - mx) * scale) * inv;
    const float e2 = expf((v.z - mx) * scale) * inv, e3 = expf((v.w - mx) * scale) * inv;
    *reinterpret_cast<uint2*>(p + i) = make_uint2(pack_bf16(e0, e1), pack_bf16(e2, e3));
  }
}

// x fp32 NHWC [B,H,W,Cin] (already GroupNorm+SiLU'ed), w fp32 HWIO [3,3,Cin,3]; one warp per output pixel.
// raw_nchw (optional): the decoder's .sample, NCHW [B,3,H,W]; img_nhwc: (raw/2 + 0.5).clip(0,1), NHWC [B,H,W,3].
__global__ void __launch_bounds__(256) vae_conv_out_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ raw_nchw,
                                                           float* __restrict__ img_nhwc, int B, int H, int W, int Cin) {
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t HW = static_cast<int64_t>(H) * W;
  if (pix >= B * HW) return;
  const int b = static_cast<int>(pix / HW);
  const int hw = static_cast<int>(pix % HW), h = hw / W, ww = hw % W;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const int C4 = Cin >> 2;
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = h + tap / 3 - 1, xx = ww + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float4* xr = reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * H + yy) * W + xx) * Cin);
    const float4* wr = reinterpret_cast<const float4*>(w + static_cast<size_t>(tap) * Cin * 3);
    for (int c4 = lane; c4 < C4; c4 += 32) {  // a lane owns 4 channels: one 16-byte activation load, 3 weight loads
      const float4 v = xr[c4];
      const float4 w0 = __ldg(wr + c4 * 3), w1 = __ldg(wr + c4 * 3 + 1), w2 = __ldg(wr + c4 * 3 + 2);
      // [c][o] for c = 0..3, o = 0..2 laid out as 12 consecutive floats
      a0 = fmaf(v.x, w0.x, a0), a1 = fmaf(v.x, w0.y, a1), a2 = fmaf(v.x, w0.z, a2);
      a0 = fmaf(v.y, w0.w, a0), a1 = fmaf(v.y, w1.x, a1), a2 = fmaf(v.y, w1.y, a2);
      a0 = fmaf(v.z, w1.z, a0), a1 = fmaf(v.z, w1.w, a1), a2 = fmaf(v.z, w2.x, a2);
      a0 = fmaf(v.w, w2.y, a0), a1 = fmaf(v.w, w2.z, a1), a2 = fmaf(v.w, w2.w, a2);
    }
  }
  a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2);
  if (lane == 0) {
    const float r0 = a0 + bias[0], r1 = a1 + bias[1], r2 = a2 + bias[2];
    if (raw_nchw != nullptr) {
      float* ro = raw_nchw + static_cast<size_t>(b) * 3 * HW + hw;
      ro[0] = r0, ro[HW] = r1, ro[2 * HW] = r2;
    }
    if (img_nhwc != nullptr) {
      float* io = img_nhwc + static_cast<size_t>(pix) * 3;
      io[0] = fminf(fmaxf(r0 * 0.5f + 0.5f, 0.f), 1.f);
      io[1] = fminf(fmaxf(r1 * 0.5f + 0.5f, 0.f), 1.f);
      io[2] = fminf(fmaxf(r2 * 0.5f + 0.5f, 0.f), 1.f);
    }
  }
}

// ------------------------------------------------------------------------------------ encoder pieces ----
// images NHWC [B,H,W,3] in [0,1] -> NCHW (x - 0.5) / 0.5  (reference ddpo/training/callbacks.py:43-44)
__global__ void __launch_bounds__(256) vae_image_to_nchw_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                                int64_t pixels, int64_t HW) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= pixels) return;
  const int64_t b = i / HW, p = i % HW;
  const float* s = img + i * 3;
  float* o = out + b * 3 * HW + p;
  o[0] = (s[0] - 0.5f) / 0.5f, o[HW] = (s[1] - 0.5f) / 0.5f, o[2 * HW] = (s[2] - 0.5f) / 0.5f;
}

// Encoder head in fp32: conv_out 3x3 (pad 1) Cin -> 8, quant_conv 1x1 8 -> 8, logvar clip.  One warp per pixel; a lane
// owns 4 input channels per step (one 16-byte activation load, eight 16-byte weight loads = [4 c][8 o]).
__global__ void __launch_bounds__(256) vae_encoder_head_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, const float* __restrict__ wq,
                                                               const float* __restrict__ bq, float* __restrict__ moments,
                                                               int B, int H, int W, int Cin) {
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t HW = static_cast<int64_t>(H) * W;
  if (pix >= B * HW) return;
  const int b = static_cast<int>(pix / HW);
  const int hw = static_cast<int>(pix % HW), h = hw / W, ww = hw % W;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int C4 = Cin >> 2;
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = h + tap / 3 - 1, xx = ww + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float4* xr = reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * H + yy) * W + xx) * Cin);
    const float4* wr = reinterpret_cast<const float4*>(w + static_cast<size_t>(tap) * Cin * 8);
    for (int c4 = lane; c4 < C4; c4 += 32) {
      const float4 v = xr[c4];
      const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 w0 = __ldg(wr + (c4 * 4 + c) * 2), w1 = __ldg(wr + (c4 * 4 + c) * 2 + 1);
        acc[0] = fmaf(xv[c], w0.x, acc[0]), acc[1] = fmaf(xv[c], w0.y, acc[1]);
        acc[2] = fmaf(xv[c], w0.z, acc[2]), acc[3] = fmaf(xv[c], w0.w, acc[3]);
        acc[4] = fmaf(xv[c], w1.x, acc[4]), acc[5] = fmaf(xv[c], w1.y, acc[5]);
        acc[6] = fmaf(xv[c], w1.z, acc[6]), acc[7] = fmaf(xv[c], w1.w, acc[7]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = warp_sum(acc[o]) + bias[o];
  if (lane < 8) {
    float m = bq[lane];
#pragma unroll
    for (int i = 0; i < 8; ++i) m = fmaf(acc[i], wq[i * 8 + lane], m);   // quant_conv kernel [in, out]
    if (lane >= 4) m = fminf(fmaxf(m, -30.0f), 20.0f);                    // logvar half of the moments
    moments[pix * 8 + lane] = m;
  }
}

}  // namespace ddpo

using namespace ddpo;

// uint8 hand-off of decoded images to the host-side rewards: u8 = (uint8)(x * 255) -- the truncating cast the reference
// applies on the host (ddpo/training/callbacks.py:181, utils/hdf5.py:31), here before the device -> host copy (4x fewer bytes)
__global__ void __launch_bounds__(256) image_to_uint8_kernel(const float4* __restrict__ img, uchar4* __restrict__ out, int64_t n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = img[i];
  out[i] = make_uchar4(static_cast<unsigned char>(v.x * 255.f), static_cast<unsigned char>(v.y * 255.f),
                       static_cast<unsigned char>(v.z * 255.f), static_cast<unsigned char>(v.w * 255.f));
}

extern "C" int ddpo_image_to_uint8(const float* img, unsigned char* out, long long n, void* stream) {
  DDPO_REQUIRE(img && out && n > 0 && n % 4 == 0, "image_to_uint8: n must be a positive multiple of 4");
  DDPO_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0, "image_to_uint8: alignment");
  const int64_t n4 = n / 4;
  image_to_uint8_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(img), reinterpret_cast<uchar4*>(out), n4);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}

extern "C" int ddpo_vae_image_to_nchw(const float* img_nhwc, float* out_nchw, int batch, int h, int w, void* stream) {
  DDPO_REQUIRE(img_nhwc && out_nchw && batch > 0 && h > 0 && w > 0, "vae_image_to_nchw: bad arguments");
  const int64_t px = static_cast<int64_t>(batch) * h * w;
  vae_image_to_nchw_kernel<<<static_cast<unsigned>((px + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img_nhwc, out_nchw, px, static_cast<int64_t>(h) * w);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}

extern "C" int ddpo_vae_encoder_head(const float* x_nhwc, const float* w_hwio, const float* bias, const float* wq_in_out,
                                     const float* bq, float* moments_nhwc, int batch, int h, int w, int cin, void* stream) {
  DDPO_REQUIRE(x_nhwc && w_hwio && bias && wq_in_out && bq && moments_nhwc && batch > 0 && h > 0 && w > 0 && cin > 0 &&
                   cin % 4 == 0,
               "vae_encoder_head: bad arguments (cin=%d must be a multiple of 4)", cin);
  const int64_t pix = static_cast<int64_t>(batch) * h * w;
  vae_encoder_head_kernel<<<static_cast<unsigned>((pix + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nhwc, w_hwio, bias, wq_in_out, bq, moments_nhwc, batch, h, w, cin);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}

extern "C" int ddpo_vae_post_quant(const float* latents_nchw, const float* w_in_out, const float* bias, float scaling,
                                   int batch, int channels, int h, int w, float* out_nchw, void* stream) {
  DDPO_REQUIRE(latents_nchw && w_in_out && bias && out_nchw, "vae_post_quant: null pointer");
  DDPO_REQUIRE(channels == 4 && batch > 0 && h > 0 && w > 0 && scaling != 0.f,
               "vae_post_quant: only 4 latent channels are built (got %d)", channels);
  const int64_t n = static_cast<int64_t>(batch) * h * w;
  vae_post_quant_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      latents_nchw, w_in_out, bias, 1.0f / scaling, batch, h * w, out_nchw);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}

extern "C" int ddpo_softmax_rows(const float* scores, int64_t ld_scores, float scale, void* probs_bf16, int64_t ld_probs,
                                 int rows, int n, void* stream) {
  DDPO_REQUIRE(scores && probs_bf16 && rows > 0 && n > 0 && n % 4 == 0 && ld_scores % 4 == 0 && ld_probs % 4 == 0,
               "softmax_rows: bad arguments (rows=%d n=%d)", rows, n);
  softmax_rows_kernel<<<rows, SM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, ld_scores, scale, static_cast<__nv_bfloat16*>(probs_bf16), ld_probs, n);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}

extern "C" int ddpo_vae_conv_out(const float* x_nhwc, const float* w_hwio, const float* bias, float* raw_nchw,
                                 float* img_nhwc, int batch, int h, int w, int cin, void* stream) {
  DDPO_REQUIRE(x_nhwc && w_hwio && bias && (raw_nchw || img_nhwc) && batch > 0 && h > 0 && w > 0 && cin > 0 && cin % 4 == 0,
               "vae_conv_out: bad arguments (cin=%d must be a multiple of 4)", cin);
  const int64_t pix = static_cast<int64_t>(batch) * h * w;
  DDPO_REQUIRE((pix + 7) / 8 < (int64_t(1) << 31), "vae_conv_out: too many pixels");
  vae_conv_out_kernel<<<static_cast<unsigned>((pix + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nhwc, w_hwio, bias, raw_nchw, img_nhwc, batch, h, w, cin);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}
