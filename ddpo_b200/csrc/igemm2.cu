// CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM kernel: two CTAs of a cluster (one TPC) compute one
// 256 x BN tile.  Each CTA stages ITS 128 activation rows and HALF of the weight tile; the leader CTA's single MMA
// thread issues 256 x BN x 16 UMMAs that read both CTAs' shared memory and write both CTAs' TMEM.
//
// Why: with one CTA per tile the tensor core re-reads (128 + BN) x 32 B of shared memory per UMMA while TMA writes
// the same amount -> ~230 B/clk/SM of shared-memory traffic against a 128 B/clk port: the 1-CTA kernel tops out
// near 50 % of the tensor peak (measured, profiles/).  In a pair each CTA reads 128 x 32 B of A and only BN/2 x 32 B
// of B per UMMA -> the shared-memory port is no longer the limiter.
//
// Protocol (all barriers live at identical offsets in both CTAs):
//   full[s]  (leader's is used) : leader producer arrive.expect_tx(bytes of BOTH CTAs); both CTAs' TMA loads
//                                 complete_tx on the leader's barrier (cp.async.bulk.tensor ... cta_group::2)
//   empty[s] (one per CTA)      : tcgen05.commit.cta_group::2 multicast from the leader's MMA thread
//   tmem_full[b] (one per CTA)  : multicast commit after the last k-block of a tile
//   tmem_empty[b] (leader's)    : 8 epilogue warps of EACH CTA arrive (peer: mapa + remote arrive)
// Everything else (TMA im2col-free operand fetch, fused epilogue, fixed K order => bit-identical results to the
// 1-CTA kernel) is shared with igemm.cu.
#include "igemm_common.cuh"

namespace ddpo {

template <int NS>  // BN-wide sub-tiles per tile (1 or 2): NS = 2 fetches the activation tile ONCE for 2 * BN columns
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(IGEMM_THREADS, 1)
    igemm2_kernel(const __grid_constant__ IGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.BN;
  const int HB = BN >> 1;  // weight rows staged by each CTA (per sub-tile)
  const int stages = p.stages;
  const int b_sub_bytes = HB * BK * 2;
  const int stage_bytes = A_TILE_BYTES + NS * b_sub_bytes;
  // accumulator ring in TMEM: sub-tile number jg (counted per CTA pair) lives in slot jg % nslot.  NS = 1: two slots of
  // 256 columns (double buffering).  NS = 2: three slots of 160 columns -- the epilogue drains the first sub-tile of a
  // tile first, which is exactly the slot the NEXT tile needs besides the free one.
  constexpr int nslot = NS == 2 ? 3 : 2;
  constexpr int slot_cols = NS == 2 ? 160 : 256;
  uint8_t* epi_base = smem;  // [EPI_BYTES] staging of the TMA epilogue / [EPI_STAGE_BYTES] of the coalescing register epilogue
  if (p.epi_tma) smem += EPI_BYTES;
  else if (p.epi_stage) smem += EPI_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
  uint64_t* epi_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes + 256);  // [8 warps][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int tiles_m = (p.M_total + 2 * BM - 1) / (2 * BM);
  const int tiles_n = p.N_total / (NS * BN);
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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);
    }
    if (p.epi_tma)
      for (int i = 0; i < 8 * EPI_RING; ++i) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The two single-role warps run their loops CONVERGED (all 32 lanes wait on the barriers and advance the counters) and
  // elect one lane only for the issue instructions: loop state, coordinates and UMMA descriptors then live in uniform
  // registers.  Inside an `if (lane == 0)` region the compiler keeps them in vector registers and wraps every UTMALDG /
  // UTCHMMA in an R2UR + ELECT "waterfall": ~165 (producer) / ~120 (MMA) dependent instructions per 64-deep k-block,
  // i.e. more clocks than the 2 * BN the tensor pipe needs for it -- the issue threads, not the tensor pipe, L2 or the
  // shared-memory port, were what bounded the main loop (tests/prof_igemm_roles.py).
  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producers (both CTAs)
    // warp 0 arms the stage barrier and fetches the activation tile, warp 3 (otherwise idle) the weight sub-tiles: with one
    // issuing warp a 160-wide stage (320 tensor clocks) was still bounded by the two-op issue loop (role isolation: 104 us
    // without the epilogue vs 87 us MMA-only).  Transaction bytes may reach the barrier before the arming arrive.
    const bool a_role = warp == 0;
    int stage = 0;
    uint32_t phase = 0;
    const int HW = p.W * p.H;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int tm = tile / tiles_n, tn = tile % tiles_n;
      const int m0 = tm * 2 * BM + static_cast<int>(rank) * BM;    // this CTA's 128 rows
      const int n0 = tn * NS * BN + static_cast<int>(rank) * HB;  // this CTA's half of (each sub-tile of) the weight tile
      int b0 = 0, h0 = 0, w0 = 0;
      if (p.is_conv) {
        b0 = m0 / HW;
        h0 = (m0 % HW) / p.W;
        w0 = m0 % p.W;  // non-zero only for rows wider than a tile (W > 128: the VAE's 256 / 512 px levels)
      }
      int kit = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        int cx = 0, cy = 0;
        if (p.is_conv) {
          const int dy = p.taps == 9 ? tap / 3 : 0, dx = p.taps == 9 ? tap - dy * 3 : 0;
          cx = w0 * p.conv_stride + dx - p.pad;
          cy = h0 * p.conv_stride + dy - p.pad;
        }
        for (int ch = 0; ch < kcs; ++ch, ++kit) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * stage_bytes;
          uint8_t* sB = sA + A_TILE_BYTES;
          if (elect_one()) {
            if (p.dbg & 1) {
              if (rank == 0 && a_role) mbar_arrive(&full_bar[stage]);
            } else if (a_role) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);
              if (p.is_conv) {
                if (ch < p.kc0)
                  tma_load_4d_2sm(sA, &p.tmA0, &full_bar[stage], ch * BK, cx, cy, b0);
                else
                  tma_load_4d_2sm(sA, &p.tmA1, &full_bar[stage], (ch - p.kc0) * BK, cx, cy, b0);
              } else {
                tma_load_2d_2sm(sA, &p.tmA0, &full_bar[stage], kit * BK, m0);
              }
            } else {
#pragma unroll
              for (int j = 0; j < NS; ++j)
                tma_load_2d_2sm(sB + j * b_sub_bytes, &p.tmB, &full_bar[stage], kit * BK, n0 + j * BN);
            }
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
    // -------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int jg0 = it * NS, slot0 = jg0 % nslot, slot1 = (jg0 + 1) % nslot;
        mbar_wait(&tmem_empty[slot0], ((jg0 / nslot) & 1) ^ 1);
        if (NS == 2) mbar_wait(&tmem_empty[slot1], (((jg0 + 1) / nslot) & 1) ^ 1);
        const uint32_t d_tmem0 = tmem_base + slot0 * slot_cols, d_tmem1 = tmem_base + slot1 * slot_cols;
        tc_fence_after();
        for (int kit = 0; kit < kiters; ++kit) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_base = a_base + A_TILE_BYTES;
          if (elect_one()) {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t a_desc = umma_desc(a_base + k * 32, 16, 1024);
                umma_bf16_2sm(d_tmem0, a_desc, umma_desc(b_base + k * 32, 16, 1024), idesc, (kit | k) != 0);
                if (NS == 2)
                  umma_bf16_2sm(d_tmem1, a_desc, umma_desc(b_base + b_sub_bytes + k * 32, 16, 1024), idesc, (kit | k) != 0);
              }
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          umma_commit_2sm(&tmem_full[slot0], 0x3);
          if (NS == 2) umma_commit_2sm(&tmem_full[slot1], 0x3);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int cgrp = (warp - 4) >> 2;
    int it = 0;
    if (p.epi_tma) {
      if (lane == 0) {
        prefetch_tmap(&p.tmIn);
        prefetch_tmap(&p.tmOutF);
        prefetch_tmap(&p.tmOutB);
      }
      EpiWarp e{epi_base + (warp - 4) * EPI_WARP_BYTES, epi_bar + (warp - 4) * EPI_RING, 0u, 0u};
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int tm = tile / tiles_n, tn = tile % tiles_n;
        const int m_slab = tm * 2 * BM + static_cast<int>(rank) * BM + q * 32;
        const bool slab_ok = m_slab < p.M_total;  // warp-uniform; a slab entirely below the matrix has nothing to do
        for (int j = 0; j < NS; ++j) {
          const int jg = it * NS + j, sl = jg % nslot;
          const int n0 = (tn * NS + j) * BN;
          if (slab_ok && p.epi_in && lane == 0 && !(p.dbg & 4))
            epi_request(p, e, e.g, (BN - cgrp * 32 + 63) / 64, e.g, n0 + cgrp * 32, 64, m_slab);
          mbar_wait(&tmem_full[sl], (jg / nslot) & 1);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sl * slot_cols;
          if (slab_ok && !(p.dbg & 4)) igemm_epilogue_tma(p, e, t_row, m_slab, lane, n0, BN, cgrp, 64);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[sl], 0);
        }
      }
      if (lane == 0) bulk_wait_all();
    } else
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int tm = tile / tiles_n, tn = tile % tiles_n;
      const int m0 = tm * 2 * BM + static_cast<int>(rank) * BM;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M_total;
      // while the main loop of this tile runs: the fp32 residual segments this thread will add (one 128-byte line per
      // 32-column chunk) are pulled into L2, so the drain of the accumulators does not wait on DRAM
      if (p.res_prefetch && p.residual != nullptr && row_ok) {
        const float* r = p.residual + static_cast<size_t>(row) * p.ld_res + tn * NS * BN;
        for (int j = 0; j < NS; ++j)
          for (int c0 = cgrp * 32; c0 < BN; c0 += 64) prefetch_l2(r + j * BN + c0);
      }
      for (int j = 0; j < NS; ++j) {
        const int jg = it * NS + j, sl = jg % nslot;
        mbar_wait(&tmem_full[sl], (jg / nslot) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sl * slot_cols;
        bool released = false;
        if (!(p.dbg & 4)) {
          if (p.epi_stage) {
            igemm_epilogue_staged(p, epi_base + (warp - 4) * EPI_STAGE_WARP_BYTES, t_row, m0 + q * 32, lane, (tn * NS + j) * BN, BN,
                                  cgrp, 64, [&]() {
                                    tc_fence_before();
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive_cluster(&tmem_empty[sl], 0);
                                  });
            released = true;
          } else {
            igemm_epilogue(p, t_row, row, row_ok, (tn * NS + j) * BN, tn * NS + j, BN, cgrp, 64);
          }
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[sl], 0);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair still reads its shared memory / barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace ddpo

using namespace ddpo;

// called by ddpo_igemm (igemm.cu) once the argument block is filled; tmB must have been encoded with box rows BN/2
int ddpo_igemm2_launch(IGemmArgs& p, cudaStream_t stream) {
  const int stage_bytes = A_TILE_BYTES + p.NS * (p.BN / 2) * BK * 2;
  const int epi = p.epi_tma ? EPI_BYTES + EPI_BAR_BYTES : (p.epi_stage ? EPI_STAGE_BYTES : 0);
  int stages = (SMEM_BUDGET - epi) / stage_bytes;
  if (stages > 10) stages = 10;
  DDPO_REQUIRE(stages >= 2, "ddpo_igemm: not enough shared memory for BN=%d NS=%d with the TMA epilogue", p.BN, p.NS);
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 256 + 1024 + epi;
  static bool attr_set = false;
  if (!attr_set) {
    DDPO_CUDA_OK(cudaFuncSetAttribute(igemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DDPO_CUDA_OK(cudaFuncSetAttribute(igemm2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int tiles = ((p.M_total + 2 * BM - 1) / (2 * BM)) * (p.N_total / (p.NS * p.BN));
  int pairs = num_sms() / 2;
  if (pairs > tiles) pairs = tiles;
  if (p.NS == 2)
    igemm2_kernel<2><<<2 * pairs, IGEMM_THREADS, smem, stream>>>(p);
  else
    igemm2_kernel<1><<<2 * pairs, IGEMM_THREADS, smem, stream>>>(p);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}
