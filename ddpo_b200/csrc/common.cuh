// Shared device/host helpers for libddpo_b200 (sm_100a only).
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor) and tcgen05 (alloc / mma /
// commit / ld), UMMA shared-memory + instruction descriptors, status handling.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ddpo_b200.h"

namespace ddpo {

// ---------------------------------------------------------------- status ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define DDPO_CUDA_OK(expr)                                        \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::ddpo::cuda_fail(_e, #expr);   \
  } while (0)
#define DDPO_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      ::ddpo::set_error(__VA_ARGS__);                             \
      return DDPO_ERR_INVALID;                                    \
    }                                                             \
  } while (0)
#define DDPO_LAUNCH_OK() DDPO_CUDA_OK(cudaGetLastError())

int num_sms();

// Encode a (<=5-D) bf16/f32 tiled tensor map through the driver entry point
// (resolved lazily with cudaGetDriverEntryPoint so the library has no link-time
// dependency on libcuda and loads on a CPU-only box).
int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes /* rank-1 entries, dims 1.. */, const uint32_t* box,
                    const uint32_t* elem_strides, int swizzle_128b);

#ifdef __CUDACC__
// ------------------------------------------------------------- small utils ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
// packed fp32 pairs (sm_100: FFMA2 / FMUL2 execute two lanes' worth of fp32 in one issue slot; each half is an ordinary
// round-to-nearest fma / mul, so results are bit-identical to the scalar instructions) -- the softmax warps of the
// attention kernels are issue-bound
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// x[i] = x[i] * b + c for 32 values held as 32-bit registers (pairs are consecutive registers)
__device__ __forceinline__ void scale_add32(uint32_t (&v)[32], float b, float c, float (&out)[32]) {
  const uint64_t b2 = pack2(b, b), c2 = pack2(c, c);
#pragma unroll
  for (int i = 0; i < 32; i += 2)
    unpack2(fma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), b2, c2), out[i], out[i + 1]);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid / SiLU on the two MUFU ops ex2 + rcp (relative error ~2e-7); the IEEE division of x / (1 + __expf(-x)) cost
// ~10 more instructions per element in the GroupNorm apply kernels, which are issue- as much as bandwidth-limited.
__device__ __forceinline__ float sigmoid_f(float x) { return rcpf_approx(1.0f + exp2f_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// SiLU on ONE MUFU op: x sigmoid(x) = z + z tanh(z), z = x / 2 (tanh.approx: max relative error 2^-11, an order below
// bf16 resolution).  Used where the result is rounded to bf16 anyway (the GroupNorm -> GEMM-operand path, which is
// MUFU-/issue-bound with the two-MUFU form); fp32 consumers keep silu_f.
__device__ __forceinline__ float silu_bf16_f(float x) {
  const float z = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z));
  return fmaf(z, t, z);
}
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  float u = k * (x + 0.044715f * x * x * x);
  // 0.5 (1 + tanh u) == 1 / (1 + e^{-2u}): two MUFU ops (ex2, rcp) instead of libm tanhf's ~25 instructions;
  // relative error ~2e-7, exact limits at +-inf
  const float e = exp2f_approx(-2.885390081777927f * u);  // e^{-2u}
  return x * rcpf_approx(1.0f + e);
}
__device__ __forceinline__ float gelu_tanh_grad_f(float x) {
  const float k = 0.7978845608028654f;
  float u = k * (x + 0.044715f * x * x * x);
  const float e = exp2f_approx(fminf(-2.885390081777927f * u, 80.0f));  // e^{-2u}, clamped so that e * sg stays finite
  const float sg = rcpf_approx(1.0f + e);  // (1 + tanh u) / 2
  float du = k * (1.0f + 3.0f * 0.044715f * x * x);
  // d/dx [x sg] with 1 - tanh^2 u = 4 e sg^2 (no cancellation near |tanh| = 1)
  return sg + 2.0f * x * du * (e * sg) * sg;
}

// 2^x for x <= 0 two ways: the MUFU unit (16/clk/SM) and a Cody-Waite + degree-3 polynomial on the FMA/ALU pipes
// (max relative error 8.8e-5, far below bf16 resolution).  The attention kernels alternate them by COLUMN
// PARITY (a fixed function of the element position -> deterministic, batch-invariant) because the softmax
// exponentials, not the tensor cores, bound the kernel when every element goes through MUFU.
__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float r = __fadd_rd(x, 12582912.0f);  // 1.5 * 2^23: low mantissa bits now hold floor(x)
  const float f = x - (r - 12582912.0f);      // in [0, 1)
  const float p = fmaf(fmaf(fmaf(0.0771190897f, f, 0.2275643945f), f, 0.6951461434f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((__float_as_int(r) - 0x4B400000) << 23));
}

// Which columns of a 32-wide chunk take the polynomial: every DDPO_EXP_POLY_{FWD,BWD}-th one (0 = none).  The split
// is a fixed function of the element's position, so results stay deterministic and batch invariant.
// Measured on B200 (64x64 self-attention, batch 40, us fwd / bwd): period 2: 1664 / 5699, 4: 1501 / 5541,
// 8: 1485 / 5457, none: 1539 / 5392 -- the polynomial costs ~10 issue slots against 1 for MUFU.EX2 and the softmax
// warps are issue bound, so only the forward kernel (fewer other instructions per score) keeps a small share.
#ifndef DDPO_EXP_POLY_FWD
#define DDPO_EXP_POLY_FWD 8
#endif
#ifndef DDPO_EXP_POLY_BWD
#define DDPO_EXP_POLY_BWD 0
#endif
template <int PERIOD>
__device__ __forceinline__ float ex2_sel(int i, float x) {
  if (PERIOD > 0 && (i % (PERIOD > 0 ? PERIOD : 1)) == PERIOD - 1) return ex2_poly(x);
  return ex2_mufu(x);
}

// --------------------------------------------------------------- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
// try_wait already suspends the warp for a hardware-chosen interval; the clock is read once every 256 retries so the
// retry loop stays at 3 instructions.  (A long explicit suspend-time hint was measured: no gain for the attention
// kernels, -3 % on the large GEMMs whose MMA/TMA threads then wake up late.)
#ifndef DDPO_MBAR_TIMEOUT_CYCLES
#define DDPO_MBAR_TIMEOUT_CYCLES 8000000000ll  // ~4 s at 2 GHz
#endif
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > DDPO_MBAR_TIMEOUT_CYCLES) __trap();
    }
  }
}

// -------------------------------------------------------------------- TMA ----
// pull the 128-byte line holding `ptr` into L2 (no register, no scoreboard entry)
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<uint64_t>(ptr)));
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): `bytes` contiguous bytes, 16-byte aligned on both sides, completion
// credited to an mbarrier like a tensor load
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// TMA store (shared -> global, bulk async-group completion); rows/columns outside the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // all but the N most recent groups have finished READING smem
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, cta_group::1
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp gets lane (base_lane+i), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------- CTA-pair (cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// In a CTA pair the shared::cluster address of the SAME offset in the even (leader) CTA is obtained by clearing
// bit 24 of the local shared address (CUTLASS Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// TMA loads of a CTA pair: data lands in the issuing CTA's shared memory, completion bytes are credited to the
// LEADER CTA's mbarrier (only the leader issues the MMAs).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs, 128 rows each] * B[smem of both CTAs, N/2 rows each]; leader only
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs completed) on the mbarrier at the same offset in the CTAs of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster.  RELAXED: the only data the waiter depends on
// are TMEM reads, ordered by tcgen05.fence::before_thread_sync; a .release.cluster arrive compiles to
// MEMBAR.ALL.GPU + ERRBAR and stalls the epilogue warp until all of its global stores are acknowledged
// (16 % of the stall samples of the GEGLU linear, profiles/r1_linear_epilogue.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

// UMMA shared-memory matrix descriptor (sm_100: version=1 at bit 46, layout type at [61,64)).
// SWIZZLE_128B = 2.  Addresses/offsets are encoded >>4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 with BF16 A/B, FP32 D.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                                  // D format F32
         | (1u << 7)                                // A format BF16
         | (1u << 10)                               // B format BF16
         | (static_cast<uint32_t>(a_mn_major) << 15)  // A major (0 = K)
         | (static_cast<uint32_t>(b_mn_major) << 16)  // B major (0 = K)
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace ddpo
