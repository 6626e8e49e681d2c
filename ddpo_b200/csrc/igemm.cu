// Implicit-GEMM convolution / linear layer on tcgen05 (sm_100a).
//
//   out[m, n] = epilogue( sum_{tap, c} A[(b, y*s+dy-p, x*s+dx-p), c] * Wt[n, (tap, c)] )
//
// replaces every dense contraction of the Flax U-Net the reference runs through
// `unet.apply` (reference call sites: pipeline_flax_stable_diffusion.py:219-224,
// training/policy_gradient.py:87-102): nn.Conv 3x3 / 3x3 stride-2 / 1x1 and nn.Dense.
//
// Design (B200-first, not a translation of anything):
//  * persistent kernel, one CTA per SM, static round-robin tile scheduler (n fastest so
//    the 128-row activation tile is re-used out of L2 by consecutive CTAs);
//  * warp 0 = TMA producer.  The activation operand is fetched straight from the NHWC
//    bf16 tensor with a 4-D tiled tensor map whose box is (64 ch, bw, bh, bb) pixels;
//    the 3x3 taps are the same box shifted by (dx-1, dy-1) and TMA's out-of-bounds
//    zero fill IS the conv padding -> no im2col buffer ever exists.  Channel-concatenated
//    inputs (U-Net skip connections) are two tensor maps walked back to back;
//  * warp 1 = single-thread tcgen05.mma issuer, 128 x BN x 16 bf16 UMMAs, fp32
//    accumulators in TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i
//    overlaps the main loop of tile i+1;
//  * warps 4..11 = epilogue (two warps per TMEM lane quarter, alternating 32-column chunks):
//    tcgen05.ld 32x32b, fused bias / per-sample time-embedding
//    row vector / fp32 residual / GEGLU, 128-bit stores;
//  * K order is fixed and independent of the batch size and of the position of a row in
//    its tile -> results are batch-invariant and bit-reproducible (the PPO ratio of an
//    unchanged policy must be exactly 1; reference config/base.py:88 clip 1e-4).
#include "igemm_common.cuh"

namespace ddpo {

__global__ void __launch_bounds__(IGEMM_THREADS, 1) igemm_kernel(const __grid_constant__ IGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.BN;
  const int stages = p.stages;
  const int b_tile_bytes = BN * BK * 2;
  const int MT = p.MT;
  const int a_bytes = MT * A_TILE_BYTES;
  const int stage_bytes = a_bytes + b_tile_bytes;
  const int nbuf = MT == 2 ? 1 : 2;  // MT = 2 uses all 512 TMEM columns for one tile (no accumulator double buffering)
  uint8_t* epi_base = smem;  // [EPI_STAGE_BYTES] of the coalescing register epilogue
  if (p.epi_stage) smem += EPI_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int TM = BM * MT;
  const int tiles_m = (p.M_total + TM - 1) / TM;
  const int tiles_n = p.N_total / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kcs = p.kc0 + p.kc1;
  const int kiters = p.taps * kcs;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA0);
    prefetch_tmap(&p.tmA1);
    prefetch_tmap(&p.tmB);
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

  // single-role warps: loops run converged, one elected lane issues (operands in uniform registers; see igemm2.cu)
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const int HW = p.W * p.H;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / tiles_n, tn = tile % tiles_n;
      const int m0 = tm * TM, n0 = tn * BN;
      int kit = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int dy = p.taps == 9 ? tap / 3 : 0, dx = p.taps == 9 ? tap - dy * 3 : 0;
        for (int ch = 0; ch < kcs; ++ch, ++kit) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * stage_bytes;
          uint8_t* sB = sA + a_bytes;
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            for (int sub = 0; sub < MT; ++sub) {
              const int ms = m0 + sub * BM;
              if (p.is_conv) {
                const int b0 = ms / HW, h0 = (ms % HW) / p.W;
                const int w0 = ms % p.W;  // non-zero only when a pixel row is wider than a tile (W > 128)
                const int cx = w0 * p.conv_stride + dx - p.pad;
                const int cy = h0 * p.conv_stride + dy - p.pad;
                if (ch < p.kc0)
                  tma_load_4d(sA + sub * A_TILE_BYTES, &p.tmA0, &full_bar[stage], ch * BK, cx, cy, b0);
                else
                  tma_load_4d(sA + sub * A_TILE_BYTES, &p.tmA1, &full_bar[stage], (ch - p.kc0) * BK, cx, cy, b0);
              } else {
                tma_load_2d(sA + sub * A_TILE_BYTES, &p.tmA0, &full_bar[stage], kit * BK, ms);
              }
            }
            tma_load_2d(sB, &p.tmB, &full_bar[stage], kit * BK, n0);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    const uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it % nbuf;
      mbar_wait(&tmem_empty[buf], ((it / nbuf) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256;
      for (int kit = 0; kit < kiters; ++kit) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
        const uint32_t b_base = a_base + a_bytes;
        if (elect_one()) {
          for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16(d_tmem + sub * 256, umma_desc(a_base + sub * A_TILE_BYTES + k * 32, 16, 1024),
                        umma_desc(b_base + k * 32, 16, 1024), idesc, (kit | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full[buf]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;
    const int wgrp = (warp - 4) >> 2;
    // MT = 1: the two warps of a lane quarter alternate 32-column chunks; MT = 2: one warp group per 128-row sub-tile
    const int cgrp = MT == 2 ? 0 : wgrp;
    const int cstep = MT == 2 ? 32 : 64;
    const int sub = MT == 2 ? wgrp : 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / tiles_n, tn = tile % tiles_n;
      const int m0 = tm * TM + sub * BM, n0 = tn * BN;
      const int buf = it % nbuf;
      mbar_wait(&tmem_full[buf], (it / nbuf) & 1);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M_total;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256 + sub * 256;
      if (p.epi_stage) {
        igemm_epilogue_staged(p, epi_base + (warp - 4) * EPI_STAGE_WARP_BYTES, t_row, m0 + q * 32, lane, n0, BN, cgrp, cstep, [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        });
      } else {
        igemm_epilogue(p, t_row, row, row_ok, n0, tn, BN, cgrp, cstep);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Width of the CTA(-pair) tile: BN = UMMA N (multiple of 32, <= 256, divides N; GEGLU needs BN/2 % 32 == 0), or -- CTA
// pairs only -- 320 = two 160-column sub-tiles that share ONE fetch of the activation tile (igemm2.cu, NS = 2).  Tile
// shape is a pure scheduling choice (the K order of every output element is fixed), so it is picked by a cost model:
//   time ~ waves x (k-blocks x clks per k-block + fixed tile overhead)
// with, per 64-deep k-block and per SM,
//   * the tensor pipe: 2 * width clks (4096 MAC/clk/SM);
//   * operand delivery: every tap and every column tile re-reads its operands out of L2, and what a CTA's producer warps
//     get into shared memory per clock is bounded in practice (TMA issue + L2 round trips; 32-40 B/clk/SM observed in
//     every long-K launch, 64-76 B/clk/SM in a pure pull loop: tests/microbench/l2_tma_bw.cu, profiles/r2_igemm.md).  The
//     model charges 16 KB of activations + 64 B x width of weights per CTA of a pair (128 B x width for a single CTA)
//     at 35 B/clk: an empirical constant that ranks the tilings the way the measurements do -- wider tiles = fewer
//     operand bytes per FLOP;
//   * the shared-memory port (128 B/clk: TMA writes + MMA reads; the activation tile is read once per sub-tile).
// Narrower tiles win only when the widest tiling leaves most SMs idle in the last wave (e.g. 80 tiles on 74 pairs).
static int pick_width(int N, int geglu, int M_total, int kiters, bool pair, bool allow_wide) {
  const int step = geglu ? 64 : 32;
  const int workers = pair ? num_sms() / 2 : num_sms();
  const int rows = pair ? 2 * BM : BM;
  const long tiles_m = (M_total + rows - 1) / rows;
  int best = 0;
  long best_cost = 0;
  for (int w = (pair && allow_wide && !geglu) ? 320 : 256; w >= step; w -= step) {
    if (w > 256 && w != 320) continue;
    if (N % w != 0) continue;
    const int ns = w > 256 ? 2 : 1;
    const long tiles = tiles_m * (N / w);
    const long waves = (tiles + workers - 1) / workers;
    const long mma = 2L * w;
    const long bytes = 16384L + (pair ? 64L : 128L) * w;
    const long l2 = bytes / 35;
    const long port = 128L + 128L * ns + (pair ? 1L : 2L) * w;
    long kb = mma > l2 ? mma : l2;
    if (port > kb) kb = port;
    // fixed cost per tile: pipeline hand-over; the wide tile cannot double-buffer its accumulators completely (3 slots of
    // 160 TMEM columns): the drain of its first sub-tile is exposed
    const long cost = waves * (kb * kiters + (ns == 2 ? 4000 : 1000));
    if (best == 0 || cost < best_cost) best = w, best_cost = cost;
  }
  return best;
}

}  // namespace ddpo

using namespace ddpo;

int ddpo_igemm2_launch(IGemmArgs& p, cudaStream_t stream);  // igemm2.cu

extern "C" int ddpo_igemm(const ddpo_igemm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPO_REQUIRE(a != nullptr, "ddpo_igemm: null args");
  const int cin0 = a->c0, cin1 = a->c1;
  DDPO_REQUIRE(cin0 > 0 && cin0 % BK == 0 && cin1 % BK == 0, "ddpo_igemm: channel counts must be multiples of 64 (c0=%d c1=%d)", cin0, cin1);
  DDPO_REQUIRE(a->taps == 1 || a->taps == 9, "ddpo_igemm: taps must be 1 or 9");
  DDPO_REQUIRE(a->n % 32 == 0, "ddpo_igemm: N=%d must be a multiple of 32", a->n);
  DDPO_REQUIRE(a->out_f32 != nullptr || a->out_bf16 != nullptr, "ddpo_igemm: no output");
  const int m_rows = a->is_conv ? a->batch * a->w * a->h : a->m;
  // CTA pairs (cta_group::2, 256 x BN tiles) whenever the problem has enough rows; see igemm2.cu
  const bool use_pair = a->pair_override == 1 || (a->pair_override == 0 && a->mt_override == 0 && m_rows >= 1024);
  // DDPO_IGEMM_WIDE=0 keeps every tile at one UMMA width (A/B switch for the 2 x 160 tiles)
  static const bool allow_wide = []() {
    const char* e = getenv("DDPO_IGEMM_WIDE");
    return e == nullptr || e[0] != '0';
  }();
  int BN = a->bn_override > 0 ? a->bn_override
                              : pick_width(a->n, a->geglu, m_rows, a->taps * (cin0 + cin1) / BK, use_pair,
                                           // short-K layers run the TMA-staged epilogue, whose buffers leave too few stages
                                           allow_wide && a->epi_override != 1 &&
                                               (a->epi_override == 2 || a->taps * (cin0 + cin1) / BK > 24));
  int NS = 1;
  if (BN > 256) {  // bn_override = 320 (or the cost model): two sub-tiles of 160 columns per CTA-pair tile
    DDPO_REQUIRE(BN == 320 && use_pair && !a->geglu, "ddpo_igemm: 320-wide tiles need the CTA-pair kernel and no GEGLU");
    DDPO_REQUIRE(a->n % 320 == 0, "ddpo_igemm: N=%d is not a multiple of 320", a->n);
    NS = 2, BN = 160;
  }
  DDPO_REQUIRE(BN > 0 && a->n % BN == 0 && BN % 32 == 0 && BN <= 256, "ddpo_igemm: no valid BN for N=%d", a->n);
  if (a->geglu) DDPO_REQUIRE(BN % 64 == 0 && a->out_bf16 != nullptr && a->bias != nullptr, "ddpo_igemm: bad GEGLU config");

  IGemmArgs p;
  memset(&p, 0, sizeof(p));
  const int ktot = a->taps * (cin0 + cin1);
  int M_total;
  if (a->is_conv) {
    // output grid W x H per sample (input grid is stride x larger)
    const int W = a->w, H = a->h, B = a->batch, s = a->conv_stride;
    DDPO_REQUIRE(s == 1 || s == 2, "ddpo_igemm: stride must be 1 or 2");
    DDPO_REQUIRE(W > 0 && H > 0 && (W & (W - 1)) == 0 && (H & (H - 1)) == 0 && W <= 1024,
                 "ddpo_igemm: W,H must be powers of two, W<=1024 (W=%d H=%d)", W, H);
    // rows wider than a tile are cut into W/128 tiles; with stride 2 the 128-pixel box spans 256 input elements (the TMA limit)
    DDPO_REQUIRE(W <= BM || a->mt_override == 0, "ddpo_igemm: rows wider than %d pixels cannot use 256-row CTA tiles (W=%d)", BM, W);
    // a 128-row tile is bb samples x bh rows x bw pixels; rows wider than the tile are cut into W/128 tiles
    int bw = W < BM ? W : BM, bh = (BM / bw < H) ? BM / bw : H;
    int bb = BM / (bw * bh);
    M_total = B * W * H;
    const int Wi = W * s, Hi = H * s;
    for (int src = 0; src < (cin1 > 0 ? 2 : 1); ++src) {
      const int C = src == 0 ? cin0 : cin1;
      const void* base = src == 0 ? a->a0 : a->a1;
      const int ld = src == 0 ? a->lda0 : a->lda1;  // channel pitch (elements) of a pixel row
      DDPO_REQUIRE(base != nullptr && ld >= C, "ddpo_igemm: bad A source %d", src);
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
      uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * Wi, (uint64_t)ld * 2 * Wi * Hi};
      uint32_t box[4] = {(uint32_t)BK, (uint32_t)(bw * s), (uint32_t)(bh * s), (uint32_t)bb};
      uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
      int rc = make_tensor_map(src == 0 ? &p.tmA0 : &p.tmA1, base, 2, 4, dims, strides, box, es, 1);
      if (rc) return rc;
    }
    if (cin1 == 0) p.tmA1 = p.tmA0;
    p.W = W, p.H = H, p.conv_stride = s, p.pad = (a->taps == 9 && a->conv_pad == 0) ? 1 : 0;
  } else {
    DDPO_REQUIRE(a->taps == 1 && cin1 == 0, "ddpo_igemm: linear mode takes one source, one tap");
    M_total = a->m;
    uint64_t dims[2] = {(uint64_t)cin0, (uint64_t)M_total};
    uint64_t strides[1] = {(uint64_t)a->lda0 * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    uint32_t es[2] = {1, 1};
    int rc = make_tensor_map(&p.tmA0, a->a0, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
    p.tmA1 = p.tmA0;
    p.W = 1, p.H = 1, p.conv_stride = 1, p.pad = 0;
  }
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)a->n};
    uint64_t strides[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_pair ? BN / 2 : BN)};
    uint32_t es[2] = {1, 1};
    int rc = make_tensor_map(&p.tmB, a->wt, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
  }
  p.M_total = M_total, p.N_total = a->n, p.BN = BN, p.NS = NS;
  p.taps = a->taps, p.kc0 = cin0 / BK, p.kc1 = cin1 / BK, p.is_conv = a->is_conv;
  p.bias = a->bias, p.rowvec = a->rowvec, p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : 1;
  p.rowvec_ld = a->rowvec_ld;
  p.residual = a->residual, p.ld_res = a->ld_res > 0 ? a->ld_res : a->n;
  p.out_f32 = a->out_f32, p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16);
  p.ld_out = a->ld_out > 0 ? a->ld_out : (a->geglu ? a->n / 2 : a->n);
  p.geglu = a->geglu, p.accumulate_out = a->accumulate_out;
  p.aux_bf16 = static_cast<__nv_bfloat16*>(a->aux_bf16);
  p.gn_stats = a->gn_stats;
  if (a->gn_stats != nullptr)
    DDPO_REQUIRE(a->out_f32 != nullptr && !a->geglu && !a->accumulate_out,
                 "ddpo_igemm: gn_stats describes a plain fp32 output (no GEGLU, no accumulate)");
  {
    // DDPO_IGEMM_STAGED=0: register epilogue with one row per lane (A/B switch)
    static const bool staged = []() {
      const char* e = getenv("DDPO_IGEMM_STAGED");
      return e == nullptr || e[0] != '0';
    }();
    p.epi_stage = staged && !a->geglu && p.ld_out % 8 == 0 && p.ld_res % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0;
  }
  if (use_pair) {
    p.MT = 1;
    {
      const char* e = getenv("DDPO_IGEMM_DEBUG");
      p.dbg = e != nullptr ? atoi(e) : 0;
      e = getenv("DDPO_IGEMM_PREFETCH");
      p.res_prefetch = e == nullptr || e[0] != '0';
    }
    // TMA epilogue where the epilogue, not the main loop, bounds the tile: short K (the 1x1 / linear layers).
    // Long-K convolutions hide the register epilogue behind the next tile's MMAs and keep every stage for operands.
    const int kiters = a->taps * (cin0 + cin1) / BK;
    const bool aligned = (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->aux_bf16) & 15) == 0 && p.ld_out % 8 == 0 && p.ld_res % 4 == 0;
    const bool can_tma = aligned && !(a->residual != nullptr && a->accumulate_out) &&
                         (!a->geglu || (a->residual == nullptr && !a->accumulate_out && a->out_f32 == nullptr));
    const bool want_tma = a->epi_override == 1 || (a->epi_override == 0 && kiters <= 24);
    if (can_tma && want_tma) {
      p.epi_tma = 1, p.epi_stage = 0;
      uint32_t box[2] = {32, 32}, es[2] = {1, 1};
      int rc;
      if (a->geglu) {
        uint64_t dims_o[2] = {(uint64_t)a->n / 2, (uint64_t)M_total}, st_o[1] = {(uint64_t)p.ld_out * 2};
        if ((rc = make_tensor_map(&p.tmOutB, a->out_bf16, 2, 2, dims_o, st_o, box, es, 2))) return rc;
        if (a->aux_bf16 != nullptr) {
          uint64_t dims_a[2] = {(uint64_t)a->n, (uint64_t)M_total}, st_a[1] = {(uint64_t)a->n * 2};
          if ((rc = make_tensor_map(&p.tmIn, a->aux_bf16, 2, 2, dims_a, st_a, box, es, 2))) return rc;
        }
      } else {
        const float* in = a->residual != nullptr ? a->residual : (a->accumulate_out ? a->out_f32 : nullptr);
        const int ld_in = a->residual != nullptr ? p.ld_res : p.ld_out;
        p.epi_in = in != nullptr;
        uint64_t dims[2] = {(uint64_t)a->n, (uint64_t)M_total};
        if (in != nullptr) {
          uint64_t st[1] = {(uint64_t)ld_in * 4};
          if ((rc = make_tensor_map(&p.tmIn, in, 4, 2, dims, st, box, es, 1))) return rc;
        }
        if (a->out_f32 != nullptr) {
          uint64_t st[1] = {(uint64_t)p.ld_out * 4};
          if ((rc = make_tensor_map(&p.tmOutF, a->out_f32, 4, 2, dims, st, box, es, 1))) return rc;
        }
        if (a->out_bf16 != nullptr) {
          uint64_t st[1] = {(uint64_t)p.ld_out * 2};
          if ((rc = make_tensor_map(&p.tmOutB, a->out_bf16, 2, 2, dims, st, box, es, 2))) return rc;
        }
      }
    }
    return ddpo_igemm2_launch(p, stream);
  }
  // fat tiles (BM = 256) when there are still >= 2 waves of them and the main loop is long enough to amortise the
  // non-overlapped epilogue: operand bytes per MMA cycle drop from 8192(1/128+1/BN) to 8192(1/256+1/BN)
  int MT = (a->mt_override > 0) ? a->mt_override : 1;
  p.MT = MT;
  const int stage_bytes = MT * A_TILE_BYTES + BN * BK * 2;
  int stages = (SMEM_BUDGET - (p.epi_stage ? EPI_STAGE_BYTES : 0)) / stage_bytes;
  if (stages > 8) stages = 8;
  DDPO_REQUIRE(stages >= 2, "ddpo_igemm: not enough shared memory for BN=%d MT=%d", BN, MT);
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 256 + 1024 + (p.epi_stage ? EPI_STAGE_BYTES : 0);
  static bool attr_set = false;
  if (!attr_set) {
    DDPO_CUDA_OK(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int tiles = ((M_total + BM * MT - 1) / (BM * MT)) * (a->n / BN);
  int grid = num_sms();
  if (grid > tiles) grid = tiles;
  igemm_kernel<<<grid, IGEMM_THREADS, smem, stream>>>(p);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}
