// CTA-pair (tcgen05 cta_group::2) weight-gradient GEMM:
//     dW[(tap, c_in), c_out] += sum_pixels X[pixel + tap, c_in] * dY[pixel, c_out]
// Same mathematics, operand layouts (both MN-major: the reduction runs over pixels, the slow dimension of both
// NHWC operands), pixel-range splitting and fixed-order reduction as wgrad.cu; what changes is the tiling:
//
//  * rows of dW are handled as a flat list of 64-channel GROUPS g = (tap, 64-channel block of the concatenated
//    input).  A pair tile is 4 consecutive groups (2 per CTA = 128 TMEM lanes each) -- groups of one tile may belong
//    to different taps, because every group is its own [64 pixels x 64 channels] TMA box with its own tap shift.
//    320 input channels x 9 taps = 45 groups = 11.25 tiles instead of 9 x 3 blocks of 128 (one third of them half
//    empty) in the 1-CTA kernel.
//  * each CTA stages its 2 groups of X and HALF of the dY tile; the leader's MMA thread issues 256 x BN x 16 UMMAs
//    over both CTAs' shared memory.  Per 64-pixel stage a CTA moves 16 KB + BN/2 x 128 B through its 128 B/clk
//    shared-memory port (once written by TMA, once read by the tensor core) against 2*BN/... MMA clocks: balanced
//    at BN = 256, where the 1-CTA kernel (16 KB + BN x 128 B per 128 x BN tile) is port-bound at 2/3 of the MMA rate.
//  * column tiles are 256 wide; a remainder of 64 / 192 columns is run as a 128 / 256 tile whose missing channels
//    are TMA out-of-bounds zero fill (never stored).
#include "common.cuh"

namespace ddpo {

constexpr int WG2_THREADS = 384;                   // warps 0-3: TMA (X) / MMA / TMEM alloc / TMA (dY); warps 4-11: epilogue
constexpr int WG2_BKP = 64;                        // pixels per pipeline stage
constexpr int WG2_BOX = WG2_BKP * 64 * 2;          // [64 pixels x 64 channels] bf16 = 8 KB
constexpr int WG2_STAGE = 4 * WG2_BOX;             // per CTA: 2 X groups + up to 2 dY boxes = 32 KB
constexpr int WG2_SMEM_BUDGET = 227 * 1024 - 1024 - 256;

struct Wgrad2Args {
  CUtensorMap tmX0, tmX1, tmDY;
  int c0, c1, n;  // input channels per source, output channels
  int taps, is_conv, W, H, conv_stride, pad;
  int splits, pblocks_per_split, pblocks;
  int groups_per_tap, groups, MT, NT;
  int stages;
  float* dst;  // splits == 1: dW (+=) ; else workspace [splits][rows][n]
  int accumulate;
};

__device__ __forceinline__ int wg2_bn(int n, int nt) {  // width of column tile nt (256, or the padded remainder)
  const int rem = n - nt * 256;
  return rem > 128 ? 256 : 128;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG2_THREADS, 1)
    wgrad2_kernel(const __grid_constant__ Wgrad2Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * WG2_STAGE);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.splits * p.MT * p.NT;
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
      mbar_init(&tmem_empty[i], 16);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (split, row tile, column tile).  Column tiles differ in cost (256 wide vs a padded remainder); when the
  // number of pairs is a multiple of NT a pair would see the same column tile in every round (even pairs all the
  // 256-wide tiles of N = 320, odd pairs all the narrow ones), so the column index is rotated by the round number.
  const bool skew = (num_pairs % p.NT) == 0;
  auto decode = [&](int tile, int& s, int& mt, int& nt) {
    nt = tile % p.NT;
    if (skew) nt = (nt + tile / num_pairs) % p.NT;
    const int r = tile / p.NT;
    mt = r % p.MT;
    s = r / p.MT;
  };

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producers (both CTAs)
    // Two producer warps share a stage: warp 0 arms the barrier and fetches the X groups, warp 3 fetches the dY boxes.
    // One warp issuing all four 8 KB boxes of a 64-pixel stage needed about as long as the tensor core needs for the
    // stage (ncu: 64 % of the samples in the MMA warp's wait for `full`); transaction bytes may reach the barrier before
    // the arming arrive -- the phase cannot complete without it.
    // (loops run warp-converged, one elected lane issues: operands stay in uniform registers; see igemm2.cu)
    const bool x_role = warp == 0;
    {
      int stage = 0;
      uint32_t phase = 0;
      const int HW = p.W * p.H;
      const int g0_per_tap = p.c0 >> 6;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int s, mt, nt;
        decode(tile, s, mt, nt);
        const int bn = wg2_bn(p.n, nt);
        const int b_boxes = bn >> 7;  // dY boxes per CTA (BN/2 channels)
        const uint32_t tx = 2u * (2 * WG2_BOX + b_boxes * WG2_BOX);
        const int pb0 = s * p.pblocks_per_split;
        const int pb1 = min(p.pblocks, pb0 + p.pblocks_per_split);
        // this CTA's two row groups
        const CUtensorMap* tmx[2];
        int gch[2], gdx[2], gdy[2];
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          const int g = 4 * mt + 2 * static_cast<int>(rank) + gi;
          if (g < p.groups) {
            const int tap = g / p.groups_per_tap, cg = g - tap * p.groups_per_tap;
            const bool src1 = cg >= g0_per_tap;
            tmx[gi] = src1 ? &p.tmX1 : &p.tmX0;
            gch[gi] = (src1 ? cg - g0_per_tap : cg) * 64;
            gdy[gi] = p.taps == 9 ? tap / 3 : 0;
            gdx[gi] = p.taps == 9 ? tap - gdy[gi] * 3 : 0;
          } else {  // past the last group: a box entirely outside the tensor -> zero fill, rows never stored
            tmx[gi] = &p.tmX0;
            gch[gi] = p.c0;
            gdy[gi] = gdx[gi] = 0;
          }
        }
        const int nch = nt * 256 + static_cast<int>(rank) * (bn >> 1);
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * WG2_STAGE;
          uint8_t* sB = sA + 2 * WG2_BOX;
          const int row0 = pb * WG2_BKP;
          if (elect_one()) {
            if (x_role) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
              if (p.is_conv) {
                const int b0 = row0 / HW, h0 = (row0 % HW) / p.W;
#pragma unroll
                for (int gi = 0; gi < 2; ++gi)
                  tma_load_4d_2sm(sA + gi * WG2_BOX, tmx[gi], &full_bar[stage], gch[gi], gdx[gi] - p.pad,
                                  h0 * p.conv_stride + gdy[gi] - p.pad, b0);
              } else {
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) tma_load_2d_2sm(sA + gi * WG2_BOX, tmx[gi], &full_bar[stage], gch[gi], row0);
              }
            } else {
              for (int j = 0; j < b_boxes; ++j)
                tma_load_2d_2sm(sB + j * WG2_BOX, &p.tmDY, &full_bar[stage], nch + j * 64, row0);
            }
          }
          __syncwarp();
          if (++stage == stages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        int s, mt, nt;
        decode(tile, s, mt, nt);
        const int bn = wg2_bn(p.n, nt);
        const uint32_t idesc = umma_idesc_bf16(256, bn, 1, 1);  // both operands MN-major
        const int pb0 = s * p.pblocks_per_split;
        const int pb1 = min(p.pblocks, pb0 + p.pblocks_per_split);
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * WG2_STAGE);
          const uint32_t b_base = a_base + 2 * WG2_BOX;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < WG2_BKP / 16; ++k) {
              // 16 pixels = 16 rows of 128 B; LBO = stride between 64-channel boxes, SBO = 8-row groups
              umma_bf16_2sm(d_tmem, umma_desc(a_base + k * 2048, WG2_BOX, 1024), umma_desc(b_base + k * 2048, WG2_BOX, 1024),
                            idesc, (pb > pb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);
          }
          __syncwarp();
          if (++stage == stages) stage = 0, phase ^= 1;
        }
        if (elect_one()) umma_commit_2sm(&tmem_full[buf], 0x3);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int cgrp = (warp - 4) >> 2;
    const size_t rows_total = static_cast<size_t>(p.taps) * cin;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      int s, mt, nt;
      decode(tile, s, mt, nt);
      const int bn = wg2_bn(p.n, nt);
      const int buf = it & 1;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const int r = q * 32 + lane;  // TMEM lane = row of this CTA's 128
      const int g = 4 * mt + 2 * static_cast<int>(rank) + (r >> 6);
      const bool ok = g < p.groups;
      const int tap = g / p.groups_per_tap, cg = g - tap * p.groups_per_tap;
      const size_t row = static_cast<size_t>(tap) * cin + cg * 64 + (r & 63);
      float* dst = p.dst + (p.splits > 1 ? static_cast<size_t>(s) * rows_total * p.n : 0) + row * p.n + nt * 256;
      const int ncols = min(bn, p.n - nt * 256);  // stored columns (the rest is out-of-bounds padding)
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
      for (int c0 = cgrp * 32; c0 < ncols; c0 += 64) {
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
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[buf], 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace ddpo

using namespace ddpo;

// number of pair tiles of one pixel split (wgrad.cu's planner uses it to choose the number of splits)
int ddpo_wgrad2_tiles(int taps, int cin, int n) {
  const int groups = taps * (cin / 64);
  return ((groups + 3) / 4) * ((n + 255) / 256);
}

// called by ddpo_wgrad (wgrad.cu) with the tensor maps encoded (boxes of [64 channels x 64 pixels])
int ddpo_wgrad2_launch(const CUtensorMap& tmX0, const CUtensorMap& tmX1, const CUtensorMap& tmDY, int c0, int c1, int n,
                       int taps, int is_conv, int W, int H, int conv_stride, int splits, int pblocks, float* dst,
                       cudaStream_t stream) {
  Wgrad2Args p;
  memset(&p, 0, sizeof(p));
  p.tmX0 = tmX0, p.tmX1 = tmX1, p.tmDY = tmDY;
  p.c0 = c0, p.c1 = c1, p.n = n, p.taps = taps, p.is_conv = is_conv;
  p.W = W, p.H = H, p.conv_stride = conv_stride, p.pad = taps == 9 ? 1 : 0;
  p.splits = splits, p.pblocks = pblocks, p.pblocks_per_split = (pblocks + splits - 1) / splits;
  p.groups_per_tap = (c0 + c1) / 64;
  p.groups = taps * p.groups_per_tap;
  p.MT = (p.groups + 3) / 4;
  p.NT = (n + 255) / 256;
  p.dst = dst;
  p.accumulate = 1;
  p.stages = WG2_SMEM_BUDGET / WG2_STAGE;
  const size_t smem = (size_t)p.stages * WG2_STAGE + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    DDPO_CUDA_OK(cudaFuncSetAttribute(wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  const int tiles = splits * p.MT * p.NT;
  int pairs = num_sms() / 2;
  if (pairs > tiles) pairs = tiles;
  wgrad2_kernel<<<2 * pairs, WG2_THREADS, smem, stream>>>(p);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}
