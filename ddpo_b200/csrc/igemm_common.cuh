// Shared pieces of the implicit-GEMM kernels (1-CTA `igemm_kernel` and 2-CTA-pair `igemm2_kernel`):
// argument block and the fused epilogue (TMEM -> registers -> bias / time-embedding row vector / residual / GEGLU
// -> fp32 and/or bf16 global stores).
#pragma once
#include "common.cuh"

namespace ddpo {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int IGEMM_THREADS = 384;  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int SMEM_BUDGET = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/;

struct IGemmArgs {
  CUtensorMap tmA0, tmA1, tmB;
  int M_total, N_total, BN, stages;
  int epi_stage;  // register epilogue: 1 = global accesses go through a per-warp 4 KB shared-memory tile so that every
                  // LDG / STG covers whole 128-byte row segments (igemm_epilogue_staged); 0 = one row per lane
  int dbg;  // profiling only (DDPO_IGEMM_DEBUG, CTA-pair kernel): bit 0 = no operand loads, bit 1 = no MMAs, bit 2 = no epilogue
            // body -- results are garbage, the barrier protocol is unchanged: what each role costs on its own
  int res_prefetch;  // CTA-pair register epilogue: prefetch the fp32 residual lines into L2 during the main loop
  int NS;  // CTA-pair kernel: BN-wide sub-tiles per tile (1 or 2), see igemm2.cu
  int MT;  // 128-row sub-tiles per CTA tile (1 or 2).  MT = 2: BM = 256 sharing one B tile -> 33% less operand traffic
  int taps, kc0, kc1;
  int is_conv, W, H, conv_stride, pad;
  const float* bias;      // [N] or null
  const float* rowvec;    // [B, rowvec_ld] or null: added per sample (time-embedding projection)
  int rows_per_sample, rowvec_ld;
  const float* residual;  // [M, ld_res] fp32 or null
  int ld_res;
  float* out_f32;         // [M, ld_out] or null
  __nv_bfloat16* out_bf16;  // [M, ld_out] or null  (GEGLU: [M, ld_out] with N/2 useful columns)
  int ld_out;
  int geglu;
  int accumulate_out;     // out_f32 += result (used by backward passes that sum two branches)
  float* gn_stats;        // [ceil(M/32), N, 2] or null: per 32-row slab, per column (sum, sum of squares) of the fp32 OUTPUT
                          // values -- the GroupNorm that consumes out_f32 takes its statistics from these instead of
                          // re-reading the tensor (norm.cu gn_finalize_kernel)
  __nv_bfloat16* aux_bf16;  // GEGLU only: pre-activation [M, N] (tile-interleaved, bias included) kept for backward
  // TMA epilogue (CTA-pair kernel): residual / previous output fetched and results written as 32 x 32 boxes through
  // per-warp shared-memory staging, so that every global access is a full 128-byte (fp32) / 64-byte (bf16) row
  // segment issued by the TMA engine instead of 16-byte pieces of 32 different rows per LSU instruction.
  int epi_tma;               // 0 = register epilogue (igemm_epilogue), 1 = igemm_epilogue_tma
  int epi_in;                // epi_tma: an fp32 input tile (residual, or the old output when accumulate_out) is added
  CUtensorMap tmIn, tmOutF, tmOutB;  // GEGLU: tmIn = bf16 map of the pre-activation buffer (aux), tmOutB = [M, N/2]
};

constexpr int EPI_F32_TILE = 32 * 32 * 4;   // 4 KB, SWIZZLE_128B (rows of 128 B)
constexpr int EPI_BF16_TILE = 32 * 32 * 2;  // 2 KB, SWIZZLE_64B  (rows of 64 B)
constexpr int EPI_RING = 3;                 // fp32 tiles per warp: input fetched up to two chunks ahead, in-place output
constexpr int EPI_WARP_BYTES = EPI_RING * EPI_F32_TILE + 2 * EPI_BF16_TILE;
constexpr int EPI_BYTES = 8 * EPI_WARP_BYTES;  // 128 KB
constexpr int EPI_BAR_BYTES = 256;             // 8 warps x EPI_RING barriers

// Column sums over the 32 rows a warp holds (lane = row, v[j] = column j): a transposing butterfly -- at step h every lane
// hands the half of its values that belongs to its xor-h partner over and adds what it receives, so 31 shuffles (not
// 32 x 5) leave lane j with the sum of column j.  The pairing order is fixed: the result depends on the 32 values only.
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

// slab statistics of one 32 x 32 chunk of output values f (rows outside the matrix must already be zero)
__device__ __forceinline__ void gn_slab_stats(float* __restrict__ stats, int N_total, int m_slab, int n, int lane,
                                              const float (&f)[32]) {
  float t[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = f[j];
  const float s = colsum32(t, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = f[j] * f[j];
  const float ss = colsum32(t, lane);
  *reinterpret_cast<float2*>(stats + (static_cast<size_t>(m_slab >> 5) * N_total + n + lane) * 2) = make_float2(s, ss);
}

struct EpiWarp {
  uint8_t* buf;   // this warp's EPI_WARP_BYTES
  uint64_t* bar;  // [EPI_RING] input-tile barriers
  uint32_t g;     // chunks processed so far (chunk g uses fp32 buffer g % EPI_RING, bf16 buffer g & 1)
  uint32_t req;   // input tiles requested so far
};

// lane 0: fetch the fp32 input tiles of this tile's chunks [req - g0, nch) as far as the ring allows: chunk r may be
// requested once the store of chunk r - EPI_RING has left its buffer, i.e. (bulk_wait_read<1> after every commit)
// once chunk r - EPI_RING + 1 has been committed: r <= done + EPI_RING - 2 where done = chunks committed so far.
__device__ __forceinline__ void epi_request(const IGemmArgs& p, EpiWarp& e, uint32_t g0, int nch, uint32_t done, int n_first,
                                            int cstep, int m) {
  while (e.req < g0 + nch && e.req + 2 <= done + EPI_RING) {
    const uint32_t b = e.req % EPI_RING;
    mbar_expect_tx(&e.bar[b], EPI_F32_TILE);
    tma_load_2d(e.buf + b * EPI_F32_TILE, &p.tmIn, &e.bar[b], n_first + static_cast<int>(e.req - g0) * cstep, m);
    ++e.req;
  }
}

// One epilogue warp, one tile: 32 rows (TMEM lane quarter; global rows m_slab..m_slab+31) x the 32-column chunks
// c0 = cgrp*32, += cstep.  The caller has already called epi_request for this tile (before waiting for the
// accumulator) when p.epi_in.  Arithmetic order is the register epilogue's: acc + bias + rowvec + input.
__device__ __forceinline__ void igemm_epilogue_tma(const IGemmArgs& p, EpiWarp& e, uint32_t t_row, int m_slab, int lane,
                                                   int n0, int BN, int cgrp, int cstep) {
  const int row = m_slab + lane;
  const float* rv = nullptr;
  if (p.rowvec != nullptr && row < p.M_total) rv = p.rowvec + static_cast<size_t>(row / p.rows_per_sample) * p.rowvec_ld;
  const uint32_t sw128 = static_cast<uint32_t>(lane & 7), sw64 = static_cast<uint32_t>((lane >> 1) & 3);
  if (p.geglu) {
    // GEGLU: this N tile is [BN/2 linear | BN/2 gate] columns of the same output channels.  Per 32-channel chunk three
    // bf16 tiles leave through TMA: out = lin * gelu(gate) -> out_bf16[:, tn*BN/2 + c0], and (training) the
    // pre-activations lin / gate -> aux[:, n0 + c0] / aux[:, n0 + BN/2 + c0].  Same arithmetic as igemm_epilogue.
    const int half = BN >> 1;
    const int tn = n0 / BN;
    for (int c0 = cgrp * 32; c0 < half; c0 += cstep) {
      const uint32_t sl = e.g & 1;
      uint8_t* ob = e.buf + EPI_RING * EPI_F32_TILE + sl * EPI_BF16_TILE;
      uint8_t* lb = e.buf + sl * EPI_F32_TILE;
      uint8_t* gb = lb + EPI_BF16_TILE;
      const int n = n0 + c0;
      uint32_t a[32], g[32];
      tmem_ld_32x32(t_row + c0, a);
      tmem_ld_32x32(t_row + half + c0, g);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bl = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + n + half + j));
        const float blv[4] = {bl.x, bl.y, bl.z, bl.w}, bgv[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float lin = __uint_as_float(a[j + q]) + blv[q];
          float gate = __uint_as_float(g[j + q]) + bgv[q];
          f[j + q] = lin * gelu_tanh_f(gate);
          a[j + q] = __float_as_uint(lin), g[j + q] = __float_as_uint(gate);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = lane * 64 + ((j ^ sw64) << 4);
        uint4 u;
        u.x = pack_bf16(f[8 * j], f[8 * j + 1]), u.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
        u.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]), u.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
        *reinterpret_cast<uint4*>(ob + off) = u;
        if (p.aux_bf16 != nullptr) {
          uint4 v, w;
          v.x = pack_bf16(__uint_as_float(a[8 * j]), __uint_as_float(a[8 * j + 1]));
          v.y = pack_bf16(__uint_as_float(a[8 * j + 2]), __uint_as_float(a[8 * j + 3]));
          v.z = pack_bf16(__uint_as_float(a[8 * j + 4]), __uint_as_float(a[8 * j + 5]));
          v.w = pack_bf16(__uint_as_float(a[8 * j + 6]), __uint_as_float(a[8 * j + 7]));
          w.x = pack_bf16(__uint_as_float(g[8 * j]), __uint_as_float(g[8 * j + 1]));
          w.y = pack_bf16(__uint_as_float(g[8 * j + 2]), __uint_as_float(g[8 * j + 3]));
          w.z = pack_bf16(__uint_as_float(g[8 * j + 4]), __uint_as_float(g[8 * j + 5]));
          w.w = pack_bf16(__uint_as_float(g[8 * j + 6]), __uint_as_float(g[8 * j + 7]));
          *reinterpret_cast<uint4*>(lb + off) = v;
          *reinterpret_cast<uint4*>(gb + off) = w;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      ++e.g;
      if (lane == 0) {
        tma_store_2d(&p.tmOutB, ob, tn * half + c0, m_slab);
        if (p.aux_bf16 != nullptr) {
          tma_store_2d(&p.tmIn, lb, n, m_slab);
          tma_store_2d(&p.tmIn, gb, n + half, m_slab);
        }
        bulk_commit();
        bulk_wait_read<1>();
      }
      __syncwarp();
    }
    e.req = e.g;
    return;
  }
  const uint32_t g0 = e.g;
  const int nch = (BN - cgrp * 32 + cstep - 1) / cstep;
  for (int c0 = cgrp * 32; c0 < BN; c0 += cstep) {
    const uint32_t b = e.g % EPI_RING;
    uint8_t* fb = e.buf + b * EPI_F32_TILE;
    uint8_t* bb = e.buf + EPI_RING * EPI_F32_TILE + (e.g & 1) * EPI_BF16_TILE;
    const int n = n0 + c0;
    uint32_t v[32];
    tmem_ld_32x32(t_row + c0, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
      }
    }
    if (rv != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + n + j));
        f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
      }
    }
    if (p.epi_in) {
      mbar_wait(&e.bar[b], (e.g / EPI_RING) & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(fb + lane * 128 + ((j ^ sw128) << 4));
        f[4 * j] += b4.x, f[4 * j + 1] += b4.y, f[4 * j + 2] += b4.z, f[4 * j + 3] += b4.w;
      }
    }
    if (p.out_f32 != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(fb + lane * 128 + ((j ^ sw128) << 4)) =
            make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    }
    if (p.out_bf16 != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16(f[8 * j], f[8 * j + 1]);
        u.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
        u.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
        u.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
        *reinterpret_cast<uint4*>(bb + lane * 64 + ((j ^ sw64) << 4)) = u;
      }
    }
    if (p.gn_stats != nullptr) {
      if (row >= p.M_total) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f;
      }
      gn_slab_stats(p.gn_stats, p.N_total, m_slab, n, lane, f);
    }
    fence_proxy_async_smem();
    __syncwarp();
    ++e.g;
    if (lane == 0) {
      if (p.out_f32 != nullptr) tma_store_2d(&p.tmOutF, fb, n, m_slab);
      if (p.out_bf16 != nullptr) tma_store_2d(&p.tmOutB, bb, n, m_slab);
      bulk_commit();
      bulk_wait_read<1>();  // every store but the one just issued has left its buffers
      if (p.epi_in) epi_request(p, e, g0, nch, e.g, n0 + cgrp * 32, cstep, m_slab);
    }
    __syncwarp();
  }
  e.req = e.g > e.req ? e.g : e.req;  // keeps lanes other than 0 (which never request) and the no-input mode consistent
}

constexpr int EPI_STAGE_WARP_BYTES = 32 * 32 * 4;           // one fp32 32 x 32 tile per epilogue warp
constexpr int EPI_STAGE_BYTES = 8 * EPI_STAGE_WARP_BYTES;   // 32 KB

// Register epilogue with coalesced global accesses.  tcgen05.ld hands every lane one ROW of the 32 x 32 chunk; a lane
// that loads / stores its own row touches 16 bytes of 32 different 128-byte lines per instruction: 32 LSU wavefronts for
// what fits in 4 -- 512 wavefronts per chunk with an fp32 residual and an fp32 output, ~10k LSU clocks per 128 x 160
// sub-tile, and that drain is what the wide (2 x 160) tiles expose (tests/prof_igemm_roles.py: 82 us without the epilogue
// body, 142 us with it).  Here rows change hands through a per-warp 4 KB tile (16-byte chunks XOR-swizzled by the row,
// conflict-free for both access patterns): global instruction i of lane l covers chunk (l & 7) of row 4 i + (l >> 3), i.e.
// four whole 128-byte row segments.  Arithmetic and its order are igemm_epilogue's: acc + bias + rowvec + residual
// (+ old output) -> bit-identical results.  GEGLU layers keep the plain path (they run the TMA-staged epilogue anyway).
// `release()` hands the accumulator slot back (arrive on its tmem_empty barrier): it is called as soon as the LAST chunk of
// this warp has left TMEM, before that chunk's arithmetic, stores and statistics -- the next tile's MMAs wait for exactly
// this hand-over when the accumulators cannot be double-buffered (320-wide tiles).
template <typename Release>
__device__ __forceinline__ void igemm_epilogue_staged(const IGemmArgs& p, uint8_t* __restrict__ stg, uint32_t t_row,
                                                      int m_slab, int lane, int n0, int BN, int cgrp, int cstep, Release release) {
  if (m_slab >= p.M_total) {  // warp-uniform: a slab entirely below the matrix has nothing to do
    release();
    return;
  }
  const int row = m_slab + lane;
  const bool row_ok = row < p.M_total;
  const float* rv = nullptr;
  if (p.rowvec != nullptr && row_ok) rv = p.rowvec + static_cast<size_t>(row / p.rows_per_sample) * p.rowvec_ld;
  const uint32_t swr = static_cast<uint32_t>(lane & 7);
  const int crow = lane >> 3, cch = lane & 7;  // fp32 tiles: instruction i <-> row 4 i + crow, logical chunk cch
  const int brow = lane >> 2, bch = lane & 3;  // bf16 tiles (64-byte rows): instruction i <-> row 8 i + brow, chunk bch
  uint8_t* own = stg + lane * 128;
  for (int c0 = cgrp * 32; c0 < BN; c0 += cstep) {
    const int n = n0 + c0;
    uint32_t v[32];
    tmem_ld_32x32(t_row + c0, v);
    // the residual's (coalesced) loads are in flight while the accumulator is fetched
    float4 in[8];
    if (p.residual != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = m_slab + 4 * i + crow;
        in[i] = r < p.M_total ? *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(r) * p.ld_res + n + 4 * cch)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    tmem_ld_wait();
    if (c0 + cstep >= BN) release();   // (a warp whose chunk list is empty releases after the loop)
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
      }
    }
    if (rv != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + n + j));
        f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
      }
    }
    if (p.residual != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + crow;
        *reinterpret_cast<float4*>(stg + rr * 128 + ((cch ^ (rr & 7)) << 4)) = in[i];
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(own + ((j ^ swr) << 4));
        f[4 * j] += b4.x, f[4 * j + 1] += b4.y, f[4 * j + 2] += b4.z, f[4 * j + 3] += b4.w;
      }
      __syncwarp();
    }
    if (p.out_f32 != nullptr && p.accumulate_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + crow, r = m_slab + rr;
        const float4 o4 = r < p.M_total ? *reinterpret_cast<const float4*>(p.out_f32 + static_cast<size_t>(r) * p.ld_out + n + 4 * cch)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(stg + rr * 128 + ((cch ^ (rr & 7)) << 4)) = o4;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(own + ((j ^ swr) << 4));
        f[4 * j] += b4.x, f[4 * j + 1] += b4.y, f[4 * j + 2] += b4.z, f[4 * j + 3] += b4.w;
      }
      __syncwarp();
    }
    if (p.out_f32 != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(own + ((j ^ swr) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + crow, r = m_slab + rr;
        const float4 o4 = *reinterpret_cast<const float4*>(stg + rr * 128 + ((cch ^ (rr & 7)) << 4));
        if (r < p.M_total) *reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(r) * p.ld_out + n + 4 * cch) = o4;
      }
      __syncwarp();
    }
    if (p.out_bf16 != nullptr) {
      const uint32_t swb = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16(f[8 * j], f[8 * j + 1]);
        u.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
        u.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
        u.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
        *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ swb) << 4)) = u;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + brow, r = m_slab + rr;
        const uint4 u = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((bch ^ ((rr >> 1) & 3)) << 4));
        if (r < p.M_total) *reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(r) * p.ld_out + n + 8 * bch) = u;
      }
      __syncwarp();
    }
    if (p.gn_stats != nullptr) {
      if (!row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f;
      }
      gn_slab_stats(p.gn_stats, p.N_total, m_slab, n, lane, f);
    }
  }
  if (cgrp * 32 >= BN) release();
}

// One epilogue warp: rows (row .. ) of TMEM lane quarter `q`, 32-column chunks c0 = cgrp*32, += cstep.
// `t_row` = TMEM address of (lane quarter, accumulator buffer / sub-tile); row = global output row of this thread.
__device__ __forceinline__ void igemm_epilogue(const IGemmArgs& p, uint32_t t_row, int row, bool row_ok, int n0, int tn,
                                               int BN, int cgrp, int cstep) {
  const int lane = threadIdx.x & 31;
  const float* rv = nullptr;
      if (p.rowvec != nullptr && row_ok) rv = p.rowvec + static_cast<size_t>(row / p.rows_per_sample) * p.rowvec_ld;
      if (!p.geglu) {
        for (int c0 = cgrp * 32; c0 < BN; c0 += cstep) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c0, v);
          tmem_ld_wait();
          const int n = n0 + c0;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = row_ok ? __uint_as_float(v[j]) : 0.f;
          if (row_ok) {
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
              }
            }
            if (rv != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + n + j));
                f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
              }
            }
            if (p.residual != nullptr) {
              const float* r = p.residual + static_cast<size_t>(row) * p.ld_res + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b4 = *reinterpret_cast<const float4*>(r + j);
                f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
              }
            }
            if (p.out_f32 != nullptr) {
              float* o = p.out_f32 + static_cast<size_t>(row) * p.ld_out + n;
              if (p.accumulate_out) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 b4 = *reinterpret_cast<const float4*>(o + j);
                  f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            }
            if (p.out_bf16 != nullptr) {
              __nv_bfloat16* o = p.out_bf16 + static_cast<size_t>(row) * p.ld_out + n;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16(f[j], f[j + 1]);
                u.y = pack_bf16(f[j + 2], f[j + 3]);
                u.z = pack_bf16(f[j + 4], f[j + 5]);
                u.w = pack_bf16(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = u;
              }
            }
          }
          // statistics for the consuming GroupNorm (warp-uniform branch: all 32 lanes shuffle; rows outside M are zero)
          if (p.gn_stats != nullptr && row - lane < p.M_total) gn_slab_stats(p.gn_stats, p.N_total, row - lane, n, lane, f);
        }
      } else {
        // GEGLU: the weight rows of this N tile are [BN/2 linear | BN/2 gate] for the same
        // output channels (host-side row permutation), out = lin * gelu_tanh(gate)
        const int half = BN >> 1;
        for (int c0 = cgrp * 32; c0 < half; c0 += cstep) {
          uint32_t a[32], g[32];
          tmem_ld_32x32(t_row + c0, a);
          tmem_ld_32x32(t_row + half + c0, g);
          tmem_ld_wait();
          if (row_ok) {
            const int n = n0 + c0;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bl = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + n + half + j));
              const float blv[4] = {bl.x, bl.y, bl.z, bl.w}, bgv[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float lin = __uint_as_float(a[j + q]) + blv[q];
                float gate = __uint_as_float(g[j + q]) + bgv[q];
                f[j + q] = lin * gelu_tanh_f(gate);
                a[j + q] = __float_as_uint(lin), g[j + q] = __float_as_uint(gate);
              }
            }
            if (p.aux_bf16 != nullptr) {
              __nv_bfloat16* ax = p.aux_bf16 + static_cast<size_t>(row) * p.N_total + n;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u, w;
                u.x = pack_bf16(__uint_as_float(a[j]), __uint_as_float(a[j + 1]));
                u.y = pack_bf16(__uint_as_float(a[j + 2]), __uint_as_float(a[j + 3]));
                u.z = pack_bf16(__uint_as_float(a[j + 4]), __uint_as_float(a[j + 5]));
                u.w = pack_bf16(__uint_as_float(a[j + 6]), __uint_as_float(a[j + 7]));
                w.x = pack_bf16(__uint_as_float(g[j]), __uint_as_float(g[j + 1]));
                w.y = pack_bf16(__uint_as_float(g[j + 2]), __uint_as_float(g[j + 3]));
                w.z = pack_bf16(__uint_as_float(g[j + 4]), __uint_as_float(g[j + 5]));
                w.w = pack_bf16(__uint_as_float(g[j + 6]), __uint_as_float(g[j + 7]));
                *reinterpret_cast<uint4*>(ax + j) = u;
                *reinterpret_cast<uint4*>(ax + half + j) = w;
              }
            }
            __nv_bfloat16* o = p.out_bf16 + static_cast<size_t>(row) * p.ld_out + tn * half + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16(f[j], f[j + 1]);
              u.y = pack_bf16(f[j + 2], f[j + 3]);
              u.z = pack_bf16(f[j + 4], f[j + 5]);
              u.w = pack_bf16(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = u;
            }
          }
        }
      }
}

}  // namespace ddpo
