// Weight-gradient GEMM on tcgen05:  dW[(tap, c_in), c_out] += sum_pixels X[pixel + tap, c_in] * dY[pixel, c_out]
//
// Backward of nn.Conv / nn.Dense of the 3P Flax U-Net w.r.t. their kernels -- the part of
// `jax.grad(compute_loss)` (reference ddpo/training/policy_gradient.py:138) that XLA emits as
// conv_general_dilated / dot_general transposes.  Output is written straight in the Flax
// parameter layout (HWIO == [(tap, c_in), c_out]; Dense [in, out]) and ACCUMULATED into the flat
// fp32 gradient buffer, so `grad_acc += g` (AccumulatingTrainState, :44-47) costs no extra pass.
//
// The reduction (UMMA K) dimension is the pixel index, which is the slow dimension of both NHWC
// operands -> both operands are "MN-major": TMA boxes of [64 pixels x 64 channels] land as
// 128-byte rows (SWIZZLE_128B) and are consumed through MN-major UMMA descriptors
// (LBO = 8 KB between 64-channel groups, SBO = 1 KB between 8-pixel groups).  As in the forward
// kernel the 3x3 taps are shifted boxes with TMA zero fill as padding.
// Pixels are split into `splits` fixed ranges to fill the machine; partial tiles go to a workspace
// and are summed in split order by a second kernel (deterministic, no atomics).
#include "common.cuh"

namespace ddpo {

constexpr int WG_THREADS = 384;  // warps 4-11: epilogue (alternating 32-column chunks)
constexpr int WG_BKP = 64;                         // pixels per pipeline stage
constexpr int WG_BOX_BYTES = WG_BKP * 64 * 2;      // [64 pixels x 64 channels] bf16 = 8 KB
constexpr int WG_A_BYTES = 2 * WG_BOX_BYTES;       // 128 input channels
constexpr int WG_SMEM_BUDGET = 227 * 1024 - 1024 - 256;

struct WgradKArgs {
  CUtensorMap tmX0, tmX1, tmDY;
  int c0, c1, n;              // input channels per source, output channels
  int taps, is_conv, W, H, conv_stride, pad;
  int pix_total;              // rows of dY
  int splits, pblocks_per_split, pblocks;
  int mt_per_tap0, mt_per_tap1;  // 128-channel blocks per source
  int stages;
  float* dst;                 // splits == 1: dW (+=) ; else workspace [splits][rows][n]
  int accumulate;
};

__device__ __forceinline__ int wg_bn(int n, int nt) {  // N tiles of 256 + remainder (multiple of 64)
  const int rem = n - nt * 256;
  return rem >= 256 ? 256 : rem;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradKArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  const int stage_bytes = WG_A_BYTES + 4 * WG_BOX_BYTES;  // A (16 KB) + up to 256 dY channels (32 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int mt_per_tap = p.mt_per_tap0 + p.mt_per_tap1;
  const int MT = p.taps * mt_per_tap;
  const int NT = (p.n + 255) / 256;
  const int num_tiles = p.splits * MT * NT;
  const int cin = p.c0 + p.c1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmX0);
    prefetch_tmap(&p.tmX1);
    prefetch_tmap(&p.tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (split, tap, source, channel block, n tile)
  auto decode = [&](int tile, int& s, int& tap, int& src, int& blk, int& nt) {
    nt = tile % NT;
    int r = tile / NT;
    const int mt = r % MT;
    s = r / MT;
    tap = mt / mt_per_tap;
    const int mb = mt % mt_per_tap;
    src = mb >= p.mt_per_tap0;
    blk = src ? mb - p.mt_per_tap0 : mb;
  };

  // single-role warps run their loops converged and elect a lane for the issue instructions only (operands stay in
  // uniform registers; see igemm2.cu)
  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      const int HW = p.W * p.H;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int s, tap, src, blk, nt;
        decode(tile, s, tap, src, blk, nt);
        const int bn = wg_bn(p.n, nt);
        const int pb0 = s * p.pblocks_per_split;
        const int pb1 = min(p.pblocks, pb0 + p.pblocks_per_split);
        int dy = 0, dx = 0;
        if (p.taps == 9) dy = tap / 3, dx = tap - dy * 3;
        const CUtensorMap* tmX = src ? &p.tmX1 : &p.tmX0;
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * stage_bytes;
          uint8_t* sB = sA + WG_A_BYTES;
          const int row0 = pb * WG_BKP;
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], WG_A_BYTES + (bn / 64) * WG_BOX_BYTES);
            if (p.is_conv) {
              const int b0 = row0 / HW, h0 = (row0 % HW) / p.W;
              const int cx = dx - p.pad, cy = h0 * p.conv_stride + dy - p.pad;
              tma_load_4d(sA, tmX, &full_bar[stage], blk * 128, cx, cy, b0);
              tma_load_4d(sA + WG_BOX_BYTES, tmX, &full_bar[stage], blk * 128 + 64, cx, cy, b0);
            } else {
              tma_load_2d(sA, tmX, &full_bar[stage], blk * 128, row0);
              tma_load_2d(sA + WG_BOX_BYTES, tmX, &full_bar[stage], blk * 128 + 64, row0);
            }
            for (int j = 0; j < bn / 64; ++j)
              tma_load_2d(sB + j * WG_BOX_BYTES, &p.tmDY, &full_bar[stage], nt * 256 + j * 64, row0);
          }
          __syncwarp();
          if (++stage == stages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        int s, tap, src, blk, nt;
        decode(tile, s, tap, src, blk, nt);
        const int bn = wg_bn(p.n, nt);
        const uint32_t idesc = umma_idesc_bf16(128, bn, 1, 1);  // both operands MN-major
        const int pb0 = s * p.pblocks_per_split;
        const int pb1 = min(p.pblocks, pb0 + p.pblocks_per_split);
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_base = a_base + WG_A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < WG_BKP / 16; ++k) {
              // 16 pixels = 16 rows of 128 B; LBO = stride between 64-channel groups, SBO = 8-row groups
              umma_bf16(d_tmem, umma_desc(a_base + k * 2048, WG_BOX_BYTES, 1024),
                        umma_desc(b_base + k * 2048, WG_BOX_BYTES, 1024), idesc, (pb > pb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == stages) stage = 0, phase ^= 1;
        }
        if (elect_one()) umma_commit(&tmem_full[buf]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int cgrp = (warp - 4) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int s, tap, src, blk, nt;
      decode(tile, s, tap, src, blk, nt);
      const int bn = wg_bn(p.n, nt);
      const int buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const int ch = blk * 128 + q * 32 + lane;          // input channel within the source
      const int csrc = src ? p.c1 : p.c0;
      const bool ok = ch < csrc;
      const size_t row = static_cast<size_t>(tap) * cin + (src ? p.c0 : 0) + ch;
      const size_t rows_total = static_cast<size_t>(p.taps) * cin;
      float* dst = p.dst + (p.splits > 1 ? static_cast<size_t>(s) * rows_total * p.n : 0) + row * p.n + nt * 256;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
      for (int c0 = cgrp * 32; c0 < bn; c0 += 64) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            float4* d = reinterpret_cast<float4*>(dst + c0 + j);
            if (p.splits == 1 && p.accumulate) {
              const float4 old = *d;
              o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
            }
            *d = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int64_t n4, int splits,
                                    int accumulate) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 acc = accumulate ? reinterpret_cast<const float4*>(dw)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(part)[static_cast<int64_t>(s) * n4 + i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    reinterpret_cast<float4*>(dw)[i] = acc;
  }
}

}  // namespace ddpo

using namespace ddpo;

int ddpo_wgrad2_tiles(int taps, int cin, int n);  // wgrad2.cu
int ddpo_wgrad2_launch(const CUtensorMap& tmX0, const CUtensorMap& tmX1, const CUtensorMap& tmDY, int c0, int c1, int n,
                       int taps, int is_conv, int W, int H, int conv_stride, int splits, int pblocks, float* dst,
                       cudaStream_t stream);

static bool wgrad_use_pair(const ddpo_wgrad_args* a) { return a->kernel_override != 2; }

static int wgrad_plan(const ddpo_wgrad_args* a, int* splits, int* pblocks) {
  const int m = a->is_conv ? a->batch * a->h * a->w : a->m;
  const int pb = (m + WG_BKP - 1) / WG_BKP;
  const int cin = a->c0 + a->c1;
  int s;
  if (wgrad_use_pair(a)) {
    // CTA-pair kernel, 74 pairs.  Number of pixel splits from a small cost model (microseconds):
    //   main loop  = rounds x pixel blocks per split x 0.42 us per 256-wide 64-pixel stage (measured; 0.6 of that for
    //                the 128-wide padded remainder tile)
    //   reduction  = (splits + 2) x |dW| x 4 B at ~4 TB/s (partials written, read back, dW read-modify-written)
    const int tiles = ddpo_wgrad2_tiles(a->taps, cin, a->n);
    const int workers = num_sms() / 2;
    const int nt = (a->n + 255) / 256;
    const double wide = (a->n - (nt - 1) * 256) > 128 ? 1.0 : 0.6;
    const double stage_us = 0.42 * ((nt - 1) + wide) / nt;
    const double dw_mb = static_cast<double>(a->taps) * cin * a->n * 4.0 / 1e6;
    int hi = pb / 4;
    if (hi < 1) hi = 1;
    if (hi > 64) hi = 64;
    s = 1;
    double best = 1e30;
    for (int sk = 1; sk <= hi; ++sk) {
      const int units = sk * tiles;
      const int rounds = (units + workers - 1) / workers;
      const double t = rounds * ((pb + sk - 1) / sk) * stage_us + (sk > 1 ? (sk + 2) * dw_mb / 4.0 : 0.0);
      if (t < best - 1e-9) best = t, s = sk;
    }
  } else {
    const int mt = a->taps * ((a->c0 + 127) / 128 + (a->c1 + 127) / 128);
    const int nt = (a->n + 255) / 256;
    const int tiles = mt * nt;
    // fill ~2 waves of 148 CTAs, keep >= 8 pixel blocks per split, bound the workspace
    s = (2 * 148 + tiles - 1) / tiles;
    if (s > pb / 8) s = pb / 8;
    if (s < 1) s = 1;
  }
  const int64_t wsz = static_cast<int64_t>(a->taps) * cin * a->n;
  while (s > 1 && wsz * s > (int64_t(48) << 20)) --s;  // <= 48M floats of partials
  *splits = s;
  *pblocks = pb;
  return 0;
}

extern "C" int64_t ddpo_wgrad_workspace_floats(const ddpo_wgrad_args* a) {
  int s, pb;
  wgrad_plan(a, &s, &pb);
  return s > 1 ? static_cast<int64_t>(s) * a->taps * (a->c0 + a->c1) * a->n : 0;
}

extern "C" int ddpo_wgrad(const ddpo_wgrad_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPO_REQUIRE(a && a->dy && a->x0 && a->dw, "wgrad: null pointer");
  DDPO_REQUIRE(a->c0 % 64 == 0 && a->c1 % 64 == 0 && a->n % 64 == 0, "wgrad: channels must be multiples of 64");
  DDPO_REQUIRE(a->taps == 1 || a->taps == 9, "wgrad: taps must be 1 or 9");
  WgradKArgs p;
  memset(&p, 0, sizeof(p));
  int splits, pblocks;
  wgrad_plan(a, &splits, &pblocks);
  const int64_t need = ddpo_wgrad_workspace_floats(a);
  DDPO_REQUIRE(need == 0 || (a->workspace != nullptr && a->workspace_floats >= need),
               "wgrad: workspace too small (%lld < %lld floats)", (long long)a->workspace_floats, (long long)need);
  int M;
  if (a->is_conv) {
    const int W = a->w, H = a->h, B = a->batch, s = a->conv_stride;
    DDPO_REQUIRE((W & (W - 1)) == 0 && (H & (H - 1)) == 0 && W <= 64, "wgrad: W,H powers of two, W <= 64");
    M = B * W * H;
    int bw = W, bh = (WG_BKP / W < H) ? WG_BKP / W : H;
    int bb = WG_BKP / (bw * bh);
    const int Wi = W * s, Hi = H * s;
    for (int src = 0; src < (a->c1 > 0 ? 2 : 1); ++src) {
      const int C = src ? a->c1 : a->c0;
      const int ld = src ? (a->ldx1 > 0 ? a->ldx1 : a->c1) : (a->ldx0 > 0 ? a->ldx0 : a->c0);
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
      uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * Wi, (uint64_t)ld * 2 * Wi * Hi};
      uint32_t box[4] = {64, (uint32_t)(bw * s), (uint32_t)(bh * s), (uint32_t)bb};
      uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
      int rc = make_tensor_map(src ? &p.tmX1 : &p.tmX0, src ? a->x1 : a->x0, 2, 4, dims, strides, box, es, 1);
      if (rc) return rc;
    }
    if (a->c1 == 0) p.tmX1 = p.tmX0;
    p.W = W, p.H = H, p.conv_stride = s, p.pad = a->taps == 9 ? 1 : 0;
  } else {
    DDPO_REQUIRE(a->taps == 1 && a->c1 == 0, "wgrad: linear mode takes one source, one tap");
    M = a->m;
    uint64_t dims[2] = {(uint64_t)a->c0, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)(a->ldx0 > 0 ? a->ldx0 : a->c0) * 2};
    uint32_t box[2] = {64, (uint32_t)WG_BKP};
    uint32_t es[2] = {1, 1};
    int rc = make_tensor_map(&p.tmX0, a->x0, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
    p.tmX1 = p.tmX0;
    p.W = 1, p.H = 1, p.conv_stride = 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->n, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)(a->ldy > 0 ? a->ldy : a->n) * 2};
    uint32_t box[2] = {64, (uint32_t)WG_BKP};
    uint32_t es[2] = {1, 1};
    int rc = make_tensor_map(&p.tmDY, a->dy, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
  }
  p.c0 = a->c0, p.c1 = a->c1, p.n = a->n, p.taps = a->taps, p.is_conv = a->is_conv;
  p.pix_total = M, p.pblocks = pblocks, p.splits = splits;
  p.pblocks_per_split = (pblocks + splits - 1) / splits;
  p.mt_per_tap0 = (a->c0 + 127) / 128, p.mt_per_tap1 = (a->c1 + 127) / 128;
  p.dst = splits > 1 ? a->workspace : a->dw;
  p.accumulate = 1;
  if (wgrad_use_pair(a)) {
    int rc = ddpo_wgrad2_launch(p.tmX0, p.tmX1, p.tmDY, a->c0, a->c1, a->n, a->taps, a->is_conv, p.W, p.H, p.conv_stride,
                                splits, pblocks, p.dst, stream);
    if (rc) return rc;
  } else {
    const int stage_bytes = WG_A_BYTES + 4 * WG_BOX_BYTES;
    p.stages = WG_SMEM_BUDGET / stage_bytes;
    const size_t smem = (size_t)p.stages * stage_bytes + 256 + 1024;
    static bool attr = false;
    if (!attr) {
      DDPO_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr = true;
    }
    const int tiles = splits * p.taps * (p.mt_per_tap0 + p.mt_per_tap1) * ((a->n + 255) / 256);
    int grid = num_sms();
    if (grid > tiles) grid = tiles;
    wgrad_kernel<<<grid, WG_THREADS, smem, stream>>>(p);
    DDPO_LAUNCH_OK();
  }
  if (splits > 1) {
    const int64_t n4 = static_cast<int64_t>(a->taps) * (a->c0 + a->c1) * a->n / 4;
    int blocks = static_cast<int>((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(a->workspace, a->dw, n4, splits, 1);
    DDPO_LAUNCH_OK();
  }
  return DDPO_OK;
}
