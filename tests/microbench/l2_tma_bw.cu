// Microbenchmark (not product code): how many bytes per clock can the SMs pull out of L2 / DRAM with bulk async copies?
// The implicit-GEMM main loop re-reads its operands from L2 (9 taps x column tiles); this measures the ceiling of that path
// on the box at hand.  One CTA per SM, one thread drives a STAGES-deep ring of CHUNK-byte cp.async.bulk loads.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ddpo_b200/csrc -o tests/microbench/l2_tma_bw tests/microbench/l2_tma_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace ddpo;

// span is a power of two: the wrap is a mask (a 64-bit modulo in this loop costs ~600 clocks per copy and hides everything)
__global__ void __launch_bounds__(128, 1) pull_kernel(const uint8_t* src, size_t span, size_t cta_stride, int share, int chunk,
                                                      int stages, int iters, unsigned long long* clocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t mask = span - 1;
    size_t off = ((blockIdx.x / share) * cta_stride) & mask;
    int s = 0;
    uint32_t phase = 0;
    unsigned long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      if (i >= stages) mbar_wait(&bar[s], phase ^ 1);
      if (i < iters) {
        mbar_expect_tx(&bar[s], chunk);
        bulk_load_1d(smem + (size_t)s * chunk, src + off, chunk, &bar[s]);
        off = (off + chunk) & mask;
      }
      if (++s == stages) s = 0, phase ^= 1;
    }
    clocks[blockIdx.x] = clock64() - t0;
  }
}

static void run(const char* name, const uint8_t* buf, size_t span, int share, int grid, int chunk, int stages, unsigned long long* clocks) {
  const size_t per_cta = (size_t)32 << 20;  // bytes pulled by each CTA per launch
  const int iters = (int)(per_cta / chunk);
  const size_t stride = share == 0 ? 0 : span / 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    pull_kernel<<<grid, 128, (size_t)stages * chunk + 2048>>>(buf, span, stride, share == 0 ? 1 : share, chunk, stages, iters, clocks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  unsigned long long h[1024];
  cudaMemcpy(h, clocks, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
  double mean_clk = 0;
  for (int i = 0; i < grid; ++i) mean_clk += (double)h[i] / grid;
  printf("%-34s grid %3d chunk %6d stages %2d: %8.1f us %7.2f TB/s %7.1f B/clk/SM\n", name, grid, chunk, stages, best * 1e3,
         (double)per_cta * grid / (best * 1e-3) / 1e12, (double)per_cta / mean_clk);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); exit(1); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t cap = (size_t)4 << 30;
  uint8_t* buf;
  cudaMalloc(&buf, cap);
  cudaMemset(buf, 1, cap);
  unsigned long long* clocks;
  cudaMalloc(&clocks, sizeof(unsigned long long) * 1024);
  cudaFuncSetAttribute(pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const size_t l2 = (size_t)64 << 20;
  for (int grid : {1, 8, 37, 74, 111, 148}) run("64 MB in L2, own addresses", buf, l2, 1, grid, 16384, 8, clocks);
  for (int chunk : {4096, 8192, 16384, 32768}) run("64 MB in L2, own addresses", buf, l2, 1, sms, chunk, 192 * 1024 / chunk > 12 ? 12 : 192 * 1024 / chunk, clocks);
  for (int stages : {1, 2, 4, 8, 12}) run("64 MB in L2, own addresses", buf, l2, 1, sms, 16384, stages, clocks);
  run("64 MB in L2, CTA pairs share", buf, l2, 2, sms, 16384, 8, clocks);
  run("64 MB in L2, 4 CTAs share", buf, l2, 4, sms, 16384, 8, clocks);
  run("8 MB in L2, all CTAs same", buf, (size_t)8 << 20, 0, sms, 16384, 8, clocks);
  for (int chunk : {8192, 16384, 32768}) run("DRAM 4 GB, own addresses", buf, cap, 1, sms, chunk, 192 * 1024 / chunk > 12 ? 12 : 192 * 1024 / chunk, clocks);
  return 0;
}
