// Attention backward on tcgen05 (head dim 64): the hand-written counterpart of what
// `jax.grad(compute_loss)` (reference ddpo/training/policy_gradient.py:138) derives for
// FlaxAttention (3P diffusers attention_flax.py):  P = softmax(Q K^T / sqrt(d)), O = P V.
//
//   delta_i = rowsum(dO_i * O_i)
//   dP = dO V^T ; dS = P * (dP - delta) / sqrt(d)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
//
// Two deterministic kernels (no atomics): `dkdv` owns a 128-key block and streams the query
// blocks, accumulating dK/dV in TMEM; `dq` owns a 128-query block and streams the key blocks,
// accumulating dQ in TMEM.  P is recomputed from the saved log-sum-exp (never stored).
// P / dS tiles are written once to shared memory as [query rows][64 keys] 128-byte swizzled rows and
// consumed both as a K-major operand (dS K) and as an MN-major operand (P^T dO, dS^T Q).
#include "common.cuh"

namespace ddpo {

constexpr int AB_THREADS = 320;  // warp0 TMA + TMEM alloc, warp1 MMA, warps 2..9 softmax/epilogue (2 per TMEM lane quarter)
constexpr int AB_T = 128 * 64 * 2;  // one [128 x 64] bf16 tile

struct AttnBwdArgs {
  CUtensorMap tmQ, tmK, tmV, tmDO, tmK64, tmV64;  // *64: 64-row boxes for the dQ kernel
  const float* lse;    // [B, heads, nq]
  const float* delta;  // [B, heads, nq]
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int nq, nk, heads, lddq, lddk, lddv;
  float scale, scale_log2e;
};

// one query row (thread r) x 128 keys: read S and dP from TMEM, write P (optional) and dS as bf16
// into [128 rows][64 keys] SW128 tiles.  `row_ok` false -> zeros.
template <bool WRITE_P, int NCOLS = 64>
__device__ __forceinline__ void softmax_bwd_row(uint32_t t_s, uint32_t t_dp, uint8_t* sP, uint8_t* sDS, int r, int valid_keys,
                                                bool row_ok, float lse_l2, float delta, float c, float scale, int col_begin, bool full) {
  const float delta_s = delta * scale;  // dS = P (dP - delta) scale = P (dP scale - delta scale): one FFMA + one FMUL
#pragma unroll 1
  for (int c0 = col_begin; c0 < col_begin + NCOLS; c0 += 32) {
    uint32_t s[32], d[32];
    tmem_ld_32x32(t_s + c0, s);
    tmem_ld_32x32(t_dp + c0, d);
    tmem_ld_wait();
    float p[32], ds[32];
    if (row_ok && full) {  // predicate-free fast path (the softmax warps are instruction bound)
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float xe = __uint_as_float(s[i]) * c - lse_l2;
        const float pv = ex2_sel<DDPO_EXP_POLY_BWD>(i, xe);
        p[i] = pv;
        ds[i] = pv * fmaf(__uint_as_float(d[i]), scale, -delta_s);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool ok = row_ok && (c0 + i < valid_keys);
        const float xe = __uint_as_float(s[i]) * c - lse_l2;
        const float pv = ok ? ex2_sel<DDPO_EXP_POLY_BWD>(i, xe) : 0.f;
        p[i] = pv;
        ds[i] = pv * fmaf(__uint_as_float(d[i]), scale, -delta_s);
      }
    }
    const int tile_off = (c0 >> 6) * AB_T + r * 128;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int chunk = ((c0 & 63) >> 3) + g;
      const int off = tile_off + ((chunk ^ (r & 7)) << 4);
      if (WRITE_P) {
        uint4 u;
        u.x = pack_bf16(p[g * 8 + 0], p[g * 8 + 1]), u.y = pack_bf16(p[g * 8 + 2], p[g * 8 + 3]);
        u.z = pack_bf16(p[g * 8 + 4], p[g * 8 + 5]), u.w = pack_bf16(p[g * 8 + 6], p[g * 8 + 7]);
        *reinterpret_cast<uint4*>(sP + off) = u;
      }
      uint4 w;
      w.x = pack_bf16(ds[g * 8 + 0], ds[g * 8 + 1]), w.y = pack_bf16(ds[g * 8 + 2], ds[g * 8 + 3]);
      w.z = pack_bf16(ds[g * 8 + 4], ds[g * 8 + 5]), w.w = pack_bf16(ds[g * 8 + 6], ds[g * 8 + 7]);
      *reinterpret_cast<uint4*>(sDS + off) = w;
    }
  }
}

// ------------------------------------------------------------------ dK, dV ----
// smem: K | V | 2 x (Q | dO) | P(2 tiles) | dS(2 tiles) | barriers
constexpr int KV_SMEM_K = 0, KV_SMEM_V = AB_T, KV_SMEM_RING = 2 * AB_T, KV_SMEM_P = 6 * AB_T, KV_SMEM_DS = 8 * AB_T,
              KV_SMEM_BAR = 10 * AB_T, KV_SMEM_TOTAL = KV_SMEM_BAR + 256;

__global__ void __launch_bounds__(AB_THREADS, 1) attention_bwd_dkdv_kernel(const __grid_constant__ AttnBwdArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KV_SMEM_BAR);
  uint64_t* kv_full = bars;          // 1
  uint64_t* qdo_full = bars + 1;     // [2]
  uint64_t* qdo_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;     // 1
  uint64_t* sdp_empty = bars + 6;    // 1 (count 4)
  uint64_t* pds_full = bars + 7;     // 1 (count 4)
  uint64_t* pds_empty = bars + 8;    // 1
  uint64_t* acc_full = bars + 9;     // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int nqb = (p.nq + 127) / 128;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&p.tmQ), prefetch_tmap(&p.tmK), prefetch_tmap(&p.tmV), prefetch_tmap(&p.tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&qdo_full[i], 1), mbar_init(&qdo_empty[i], 1);
    mbar_init(sdp_full, 1), mbar_init(sdp_empty, 8), mbar_init(pds_full, 8), mbar_init(pds_empty, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DV = tmem_base + 256, T_DK = tmem_base + 320;

  // warps 0 / 1: loops run converged, one elected lane issues TMA / MMA (operands stay in uniform registers; inside an
  // `if (lane == 0)` region each of the 24 UMMAs per query block carried an R2UR + ELECT loop of ~20 instructions)
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * AB_T);
      tma_load_4d(smem + KV_SMEM_K, &p.tmK, kv_full, 0, head, k0, b);
      tma_load_4d(smem + KV_SMEM_V, &p.tmV, kv_full, 0, head, k0, b);
    }
    __syncwarp();
    for (int i = 0; i < nqb; ++i) {
      const int st = i & 1;
      mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
      uint8_t* sQ = smem + KV_SMEM_RING + st * 2 * AB_T;
      if (elect_one()) {
        mbar_expect_tx(&qdo_full[st], 2 * AB_T);
        tma_load_4d(sQ, &p.tmQ, &qdo_full[st], 0, head, i * 128, b);
        tma_load_4d(sQ + AB_T, &p.tmDO, &qdo_full[st], 0, head, i * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    {
      const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);  // S = Q K^T, dP = dO V^T
      const uint32_t id_g = umma_idesc_bf16(128, 64, 1, 1);   // dV = P^T dO, dK = dS^T Q (both MN-major)
      const uint32_t k_addr = smem_u32(smem + KV_SMEM_K), v_addr = smem_u32(smem + KV_SMEM_V);
      const uint32_t p_addr = smem_u32(smem + KV_SMEM_P), ds_addr = smem_u32(smem + KV_SMEM_DS);
      mbar_wait(kv_full, 0);
      // Software-pipelined issue order: S/dP of query block i+1 go to the tensor pipe BEFORE dV/dK of block i.  The
      // softmax warps wait only for S/dP; queueing them behind the two gradient GEMMs of the previous block (the
      // tensor pipe executes in order) left those warps idle for ~30 % of their time (profiles/r1_attention.md).
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        const uint32_t q_addr = smem_u32(smem + KV_SMEM_RING + st * 2 * AB_T), do_addr = q_addr + AB_T;
        mbar_wait(&qdo_full[st], (i >> 1) & 1);
        mbar_wait(sdp_empty, (i & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(T_S, umma_desc(q_addr + k * 32, 16, 1024), umma_desc(k_addr + k * 32, 16, 1024), id_s, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(T_DP, umma_desc(do_addr + k * 32, 16, 1024), umma_desc(v_addr + k * 32, 16, 1024), id_s, k != 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };
      issue_sdp(0);
      for (int i = 0; i < nqb; ++i) {
        const int st = i & 1;
        const uint32_t q_addr = smem_u32(smem + KV_SMEM_RING + st * 2 * AB_T), do_addr = q_addr + AB_T;
        mbar_wait(pds_full, i & 1);
        if (i + 1 < nqb) issue_sdp(i + 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // K = 128 query rows, 16 per instruction
            umma_bf16(T_DV, umma_desc(p_addr + k * 2048, AB_T, 1024), umma_desc(do_addr + k * 2048, AB_T, 1024), id_g,
                      (i | k) != 0);
            umma_bf16(T_DK, umma_desc(ds_addr + k * 2048, AB_T, 1024), umma_desc(q_addr + k * 2048, AB_T, 1024), id_g,
                      (i | k) != 0);
          }
          umma_commit(&qdo_empty[st]);
          umma_commit(pds_empty);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;  // which 64 key columns of the 128-key block this warp handles
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int valid_keys = min(128, p.nk - k0);
    for (int i = 0; i < nqb; ++i) {
      const int row = i * 128 + r;
      const bool row_ok = row < p.nq;
      float lse_l2 = 0.f, delta = 0.f;
      if (row_ok) {
        const size_t o = (static_cast<size_t>(b) * p.heads + head) * p.nq + row;
        lse_l2 = p.lse[o] * 1.4426950408889634f;
        delta = p.delta[o];
      }
      mbar_wait(sdp_full, i & 1);
      mbar_wait(pds_empty, (i & 1) ^ 1);
      tc_fence_after();
      softmax_bwd_row<true>(T_S + lane_off, T_DP + lane_off, smem + KV_SMEM_P, smem + KV_SMEM_DS, r, valid_keys, row_ok,
                            lse_l2, delta, p.scale_log2e, p.scale, chalf * 64, valid_keys == 128);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(sdp_empty);
        mbar_arrive(pds_full);
      }
    }
    // epilogue: thread r <-> key row k0 + r
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int key = k0 + r;
#pragma unroll 1
    for (int which = chalf; which < chalf + 1; ++which) {
      __nv_bfloat16* dst = which == 0 ? p.dv : p.dk;
      const int ld = which == 0 ? p.lddv : p.lddk;
      const uint32_t t = (which == 0 ? T_DV : T_DK) + lane_off;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t + c0, v);
        tmem_ld_wait();
        if (key < p.nk) {
          __nv_bfloat16* o = dst + (static_cast<size_t>(b) * p.nk + key) * ld + head * 64 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            u.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            u.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            u.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
            *reinterpret_cast<uint4*>(o + j) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------- dQ ----
// 64-key blocks, 192 TMEM columns and 97 KB of shared memory per CTA -> TWO CTAs per SM, so that one CTA's
// softmax/dS phase overlaps the other's MMAs (a single CTA alternates strictly between the two).
// smem: Q | dO | 3 x (K | V) [64 keys each] | dS (1 tile) | barriers
constexpr int DQ_BKV = 64;
constexpr int DQ_KVT = DQ_BKV * 64 * 2;  // 8 KB
constexpr int DQ_STAGES = 3;
constexpr int DQ_THREADS = 320;          // warp0 TMA + TMEM alloc, warp1 MMA, warps 2..9 softmax/epilogue:
                                         // 4 TMEM lane quarters x 2 halves of 32 key columns -> with 2 CTAs/SM four
                                         // latency-bound softmax warps per scheduler instead of two
constexpr int DQ_SMEM_Q = 0, DQ_SMEM_DO = AB_T, DQ_SMEM_RING = 2 * AB_T, DQ_SMEM_DS = DQ_SMEM_RING + DQ_STAGES * 2 * DQ_KVT,
              DQ_SMEM_BAR = DQ_SMEM_DS + AB_T, DQ_SMEM_TOTAL = DQ_SMEM_BAR + 256;

__global__ void __launch_bounds__(DQ_THREADS, 2) attention_bwd_dq_kernel(const __grid_constant__ AttnBwdArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_SMEM_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [3]
  uint64_t* kv_empty = bars + 4;   // [3]
  uint64_t* sdp_full = bars + 7;
  uint64_t* sdp_empty = bars + 8;  // count 8
  uint64_t* ds_full = bars + 9;    // count 8
  uint64_t* ds_empty = bars + 10;
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.nk + DQ_BKV - 1) / DQ_BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    prefetch_tmap(&p.tmQ), prefetch_tmap(&p.tmK64), prefetch_tmap(&p.tmV64), prefetch_tmap(&p.tmDO);
    mbar_init(q_full, 1);
    for (int i = 0; i < DQ_STAGES; ++i) mbar_init(&kv_full[i], 1), mbar_init(&kv_empty[i], 1);
    mbar_init(sdp_full, 1), mbar_init(sdp_empty, 8), mbar_init(ds_full, 8), mbar_init(ds_empty, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 64, T_DQ = tmem_base + 128;

  if (warp == 0) {   // converged loops, elected issue (see the dK/dV kernel)
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * AB_T);
      tma_load_4d(smem + DQ_SMEM_Q, &p.tmQ, q_full, 0, head, q0, b);
      tma_load_4d(smem + DQ_SMEM_DO, &p.tmDO, q_full, 0, head, q0, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(&kv_empty[st], ph ^ 1);
      uint8_t* sK = smem + DQ_SMEM_RING + st * 2 * DQ_KVT;
      if (elect_one()) {
        mbar_expect_tx(&kv_full[st], 2 * DQ_KVT);
        tma_load_4d(sK, &p.tmK64, &kv_full[st], 0, head, j * DQ_BKV, b);
        tma_load_4d(sK + DQ_KVT, &p.tmV64, &kv_full[st], 0, head, j * DQ_BKV, b);
      }
      __syncwarp();
      if (++st == DQ_STAGES) st = 0, ph ^= 1;
    }
  } else if (warp == 1) {
    {
      const uint32_t id_s = umma_idesc_bf16(128, DQ_BKV, 0, 0);
      const uint32_t id_q = umma_idesc_bf16(128, 64, 0, 1);  // dQ = dS K : A K-major, B (K) MN-major
      const uint32_t q_addr = smem_u32(smem + DQ_SMEM_Q), do_addr = smem_u32(smem + DQ_SMEM_DO);
      const uint32_t ds_addr = smem_u32(smem + DQ_SMEM_DS);
      mbar_wait(q_full, 0);
      // same issue order as the dK/dV kernel: S/dP of key block j+1 before dQ of block j
      auto issue_sdp = [&](int j) {
        const int st = j % DQ_STAGES;
        const uint32_t k_addr = smem_u32(smem + DQ_SMEM_RING + st * 2 * DQ_KVT), v_addr = k_addr + DQ_KVT;
        mbar_wait(&kv_full[st], (j / DQ_STAGES) & 1);
        mbar_wait(sdp_empty, (j & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(T_S, umma_desc(q_addr + k * 32, 16, 1024), umma_desc(k_addr + k * 32, 16, 1024), id_s, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(T_DP, umma_desc(do_addr + k * 32, 16, 1024), umma_desc(v_addr + k * 32, 16, 1024), id_s, k != 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };
      issue_sdp(0);
      for (int j = 0; j < nkb; ++j) {
        const int st = j % DQ_STAGES;
        const uint32_t k_addr = smem_u32(smem + DQ_SMEM_RING + st * 2 * DQ_KVT);
        mbar_wait(ds_full, j & 1);
        if (j + 1 < nkb) issue_sdp(j + 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DQ_BKV / 16; ++k)  // A = dS [128 q x 64 keys] K-major; B = K_j rows of 16 keys (MN-major)
            umma_bf16(T_DQ, umma_desc(ds_addr + k * 32, 16, 1024), umma_desc(k_addr + k * 2048, DQ_KVT, 1024), id_q,
                      (j | k) != 0);
          umma_commit(&kv_empty[st]);
          umma_commit(ds_empty);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;  // which 32 of the 64 key columns (and of the 64 dQ columns in the epilogue)
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int row = q0 + r;
    const bool row_ok = row < p.nq;
    float lse_l2 = 0.f, delta = 0.f;
    if (row_ok) {
      const size_t o = (static_cast<size_t>(b) * p.heads + head) * p.nq + row;
      lse_l2 = p.lse[o] * 1.4426950408889634f;
      delta = p.delta[o];
    }
    for (int j = 0; j < nkb; ++j) {
      const int valid_keys = min(DQ_BKV, p.nk - j * DQ_BKV);
      mbar_wait(sdp_full, j & 1);
      mbar_wait(ds_empty, (j & 1) ^ 1);
      tc_fence_after();
      softmax_bwd_row<false, 32>(T_S + lane_off, T_DP + lane_off, nullptr, smem + DQ_SMEM_DS, r, valid_keys, row_ok,
                                 lse_l2, delta, p.scale_log2e, p.scale, chalf * 32, valid_keys == DQ_BKV);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(sdp_empty);
        mbar_arrive(ds_full);
      }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(T_DQ + lane_off + c0, v);
      tmem_ld_wait();
      if (row_ok) {
        __nv_bfloat16* o = p.dq + (static_cast<size_t>(b) * p.nq + row) * p.lddq + head * 64 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
          u.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          u.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
          u.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
          *reinterpret_cast<uint4*>(o + j) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// delta[b,h,q] = sum_d dO * O ; 8 threads per (row, head), 16-byte loads
__global__ void attention_delta_kernel(const __nv_bfloat16* __restrict__ o, int ldo, const __nv_bfloat16* __restrict__ d_o,
                                       int lddo, float* __restrict__ delta, int B, int nq, int heads) {
  const int64_t gid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const int64_t total = static_cast<int64_t>(B) * nq * heads;
  float acc = 0.f;
  int64_t row = 0;
  int h = 0;
  const bool ok = gid < total;
  if (ok) {
    row = gid / heads;
    h = gid % heads;
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * ldo + h * 64 + sub * 8);
    const uint4 b4 = *reinterpret_cast<const uint4*>(d_o + row * lddo + h * 64 + sub * 8);
    acc = bf16_lo(a.x) * bf16_lo(b4.x) + bf16_hi(a.x) * bf16_hi(b4.x) + bf16_lo(a.y) * bf16_lo(b4.y) +
          bf16_hi(a.y) * bf16_hi(b4.y) + bf16_lo(a.z) * bf16_lo(b4.z) + bf16_hi(a.z) * bf16_hi(b4.z) +
          bf16_lo(a.w) * bf16_lo(b4.w) + bf16_hi(a.w) * bf16_hi(b4.w);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && sub == 0) {
    const int b = row / nq, qi = row % nq;
    delta[(static_cast<size_t>(b) * heads + h) * nq + qi] = acc;
  }
}

static int make_map(CUtensorMap* m, const void* base, int heads, int n, int batch, int ld, int box_rows = 128) {
  uint64_t dims[4] = {64, (uint64_t)heads, (uint64_t)n, (uint64_t)batch};
  uint64_t strides[3] = {128, (uint64_t)ld * 2, (uint64_t)ld * 2 * n};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  uint32_t es[4] = {1, 1, 1, 1};
  return make_tensor_map(m, base, 2, 4, dims, strides, box, es, 1);
}

}  // namespace ddpo

using namespace ddpo;

extern "C" int ddpo_attention_bwd(const ddpo_attention_bwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPO_REQUIRE(a && a->q && a->k && a->v && a->out && a->dout && a->lse && a->delta && a->dq && a->dk && a->dv,
               "attention_bwd: null pointer");
  DDPO_REQUIRE(a->head_dim == 64, "attention_bwd: head_dim must be 64");
  AttnBwdArgs p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_map(&p.tmQ, a->q, a->heads, a->nq, a->batch, a->ldq))) return rc;
  if ((rc = make_map(&p.tmK, a->k, a->heads, a->nk, a->batch, a->ldk))) return rc;
  if ((rc = make_map(&p.tmV, a->v, a->heads, a->nk, a->batch, a->ldv))) return rc;
  if ((rc = make_map(&p.tmDO, a->dout, a->heads, a->nq, a->batch, a->lddo))) return rc;
  if ((rc = make_map(&p.tmK64, a->k, a->heads, a->nk, a->batch, a->ldk, DQ_BKV))) return rc;
  if ((rc = make_map(&p.tmV64, a->v, a->heads, a->nk, a->batch, a->ldv, DQ_BKV))) return rc;
  p.lse = a->lse, p.delta = a->delta;
  p.dq = static_cast<__nv_bfloat16*>(a->dq), p.dk = static_cast<__nv_bfloat16*>(a->dk);
  p.dv = static_cast<__nv_bfloat16*>(a->dv);
  p.nq = a->nq, p.nk = a->nk, p.heads = a->heads, p.lddq = a->lddq, p.lddk = a->lddk, p.lddv = a->lddv;
  p.scale = 0.125f;
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    DDPO_CUDA_OK(cudaFuncSetAttribute(attention_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM_TOTAL));
    DDPO_CUDA_OK(cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM_TOTAL));
    attr = true;
  }
  const int64_t items = static_cast<int64_t>(a->batch) * a->nq * a->heads;
  attention_delta_kernel<<<static_cast<int>((items * 8 + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(a->out), a->ldo, static_cast<const __nv_bfloat16*>(a->dout), a->lddo, a->delta,
      a->batch, a->nq, a->heads);
  DDPO_LAUNCH_OK();
  dim3 g1((a->nk + 127) / 128, a->heads, a->batch);
  attention_bwd_dkdv_kernel<<<g1, AB_THREADS, KV_SMEM_TOTAL, stream>>>(p);
  DDPO_LAUNCH_OK();
  dim3 g2((a->nq + 127) / 128, a->heads, a->batch);
  attention_bwd_dq_kernel<<<g2, DQ_THREADS, DQ_SMEM_TOTAL, stream>>>(p);
  DDPO_LAUNCH_OK();
  return DDPO_OK;
}
