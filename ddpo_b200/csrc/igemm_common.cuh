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
constexpr int SMEM_BUDGET = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - 8 * 32 * 33 * 4 /*epilogue staging*/;

struct IGemmArgs {
  CUtensorMap tmA0, tmA1, tmB;
  int M_total, N_total, BN, stages;
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
  __nv_bfloat16* aux_bf16;  // GEGLU only: pre-activation [M, N] (tile-interleaved, bias included) kept for backward
};


constexpr int EPI_PITCH = 33;                                  // floats per staged row (conflict-free transposition)
constexpr int EPI_STAGE_BYTES = 32 * EPI_PITCH * 4;            // per epilogue warp
constexpr int EPI_SMEM_BYTES = 8 * EPI_STAGE_BYTES;            // 8 epilogue warps

// Fused epilogue of one warp for its 32 accumulator rows (TMEM lane quarter) and its 32-column chunks.
// tcgen05.ld hands each THREAD one ROW (32 consecutive columns): stored like that, a warp-wide store touches 32
// different rows with 16 bytes each.  The chunk is therefore transposed through a per-warp shared-memory tile
// ([32][33] floats, bank-conflict free both ways) so that 8 lanes cover 128 contiguous bytes of one output row:
// bias / time-embedding row vector / residual loads and the fp32 / bf16 stores are fully coalesced.
// t_row: TMEM address (lane quarter, buffer); row0: global output row of lane 0 of this warp.
__device__ __forceinline__ void igemm_epilogue(const IGemmArgs& p, uint32_t t_row, int row0, int n0, int tn, int BN, int cgrp,
                                               int cstep, float* stage, int lane) {
  const int rsub = lane >> 3;        // row within a group of 4
  const int c4 = (lane & 7) * 4;     // column quad within the 32-column chunk
  if (!p.geglu) {
    for (int c0 = cgrp * 32; c0 < BN; c0 += cstep) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * EPI_PITCH + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int n = n0 + c0 + c4;
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + rsub;
        const int grow = row0 + r;
        if (grow < p.M_total) {
          const float* sp = stage + r * EPI_PITCH + c4;
          float4 f = make_float4(sp[0] + bias4.x, sp[1] + bias4.y, sp[2] + bias4.z, sp[3] + bias4.w);
          if (p.rowvec != nullptr) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(
                p.rowvec + static_cast<size_t>(grow / p.rows_per_sample) * p.rowvec_ld + n));
            f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
          }
          if (p.residual != nullptr) {
            const float4 t = *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(grow) * p.ld_res + n);
            f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
          }
          if (p.out_f32 != nullptr) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(grow) * p.ld_out + n);
            if (p.accumulate_out) {
              const float4 t = *o;
              f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
            }
            *o = f;
          }
          if (p.out_bf16 != nullptr)
            *reinterpret_cast<uint2*>(p.out_bf16 + static_cast<size_t>(grow) * p.ld_out + n) =
                make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
        }
      }
      __syncwarp();
    }
  } else {
    // GEGLU: the weight rows of this N tile are [BN/2 linear | BN/2 gate] for the same output channels
    // (host-side row permutation), out = lin * gelu_tanh(gate)
    const int half = BN >> 1;
    for (int c0 = cgrp * 32; c0 < half; c0 += cstep) {
      uint32_t v[32];
      float4 lin[8];
      tmem_ld_32x32(t_row + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * EPI_PITCH + j] = __uint_as_float(v[j]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* sp = stage + (i * 4 + rsub) * EPI_PITCH + c4;
        lin[i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
      }
      __syncwarp();
      tmem_ld_32x32(t_row + half + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * EPI_PITCH + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int n = n0 + c0 + c4;
      const float4 bl = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + n + half));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + rsub;
        const int grow = row0 + r;
        if (grow < p.M_total) {
          const float* sp = stage + r * EPI_PITCH + c4;
          const float4 l = make_float4(lin[i].x + bl.x, lin[i].y + bl.y, lin[i].z + bl.z, lin[i].w + bl.w);
          const float4 g = make_float4(sp[0] + bg.x, sp[1] + bg.y, sp[2] + bg.z, sp[3] + bg.w);
          if (p.aux_bf16 != nullptr) {
            __nv_bfloat16* ax = p.aux_bf16 + static_cast<size_t>(grow) * p.N_total + n;
            *reinterpret_cast<uint2*>(ax) = make_uint2(pack_bf16(l.x, l.y), pack_bf16(l.z, l.w));
            *reinterpret_cast<uint2*>(ax + half) = make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
          }
          *reinterpret_cast<uint2*>(p.out_bf16 + static_cast<size_t>(grow) * p.ld_out + tn * half + c0 + c4) =
              make_uint2(pack_bf16(l.x * gelu_tanh_f(g.x), l.y * gelu_tanh_f(g.y)),
                         pack_bf16(l.z * gelu_tanh_f(g.z), l.w * gelu_tanh_f(g.w)));
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace ddpo
