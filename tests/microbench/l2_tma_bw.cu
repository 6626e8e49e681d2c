// Microbenchmark (not product code): how many bytes per clock can the 148 SMs pull out of L2 with bulk async copies?
// The implicit-GEMM main loop re-reads its operands from L2 (9 taps x column tiles); this measures the ceiling of
// that path on the box at hand.  One CTA per SM, one thread drives a STAGES-deep ring of CHUNK-byte cp.async.bulk loads.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ddpo_b200/csrc -o tests/microbench/l2_tma_bw tests/microbench/l2_tma_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace ddpo;

__global__ void __launch_bounds__(128, 1) pull_kernel(const uint8_t* src, size_t span, size_t cta_stride, int chunk, int stages,
                                                      int iters, unsigned long long* clocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t base = (blockIdx.x * cta_stride) % span;
    unsigned long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&bar[s], ((i / stages) - 1) & 1);
      if (i < iters) {
        const size_t off = (base + (size_t)i * chunk) % span;
        mbar_expect_tx(&bar[s], chunk);
        bulk_load_1d(smem + (size_t)s * chunk, src + off, chunk, &bar[s]);
      }
    }
    clocks[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  int dev_clock_khz = 0, sms = 0;
  cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t cap = (size_t)4 << 30;
  uint8_t* buf;
  cudaMalloc(&buf, cap);
  cudaMemset(buf, 1, cap);
  unsigned long long* clocks;
  cudaMalloc(&clocks, sizeof(unsigned long long) * 1024);
  cudaFuncSetAttribute(pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("SMs %d, nominal %d MHz\n", sms, dev_clock_khz / 1000);
  printf("%-34s %8s %6s %10s %10s %12s\n", "case", "chunk", "stages", "us", "TB/s", "B/clk/SM");
  struct Case { const char* name; size_t span; int shared; };
  // span = distinct bytes touched by the whole grid per pass; shared = every CTA reads the same addresses (weight-tile pattern)
  Case cases[] = {{"L2-resident 48 MB, private", (size_t)48 << 20, 0}, {"L2-resident 8 MB, all CTAs same", (size_t)8 << 20, 1},
                  {"L2-resident 48 MB, pairs share", (size_t)48 << 20, 2}, {"DRAM 4 GB, private", cap, 0}};
  // per-SM ingest vs chip-wide L2 output: the same private L2-resident pull with fewer CTAs
  for (int grid : {1, 8, 37, 74, 111, 148}) {
    const int chunk = 16384, stages = 8;
    const size_t per_cta = (size_t)24 << 20;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      pull_kernel<<<grid, 128, (size_t)stages * chunk + 2048>>>(buf, (size_t)48 << 20, ((size_t)48 << 20) / 148 / 1024 * 1024, chunk, stages,
                                                                 (int)(per_cta / chunk), clocks);
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
    printf("grid %3d CTAs, private 48 MB in L2: %8.1f us  %6.2f TB/s  %6.1f B/clk/SM\n", grid, best * 1e3,
           (double)per_cta * grid / (best * 1e-3) / 1e12, (double)per_cta / mean_clk);
  }
  for (const Case& c : cases)
    for (int chunk : {8192, 16384, 32768}) {
      const int stages = (200 * 1024) / chunk > 8 ? 8 : (200 * 1024) / chunk;
      const size_t per_cta = (size_t)24 << 20;  // bytes pulled by each CTA per launch
      const int iters = (int)(per_cta / chunk);
      size_t stride = c.shared == 1 ? 0 : (c.shared == 2 ? 0 : c.span / sms / 1024 * 1024);
      const size_t smem = (size_t)stages * chunk + 2048;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0), cudaEventCreate(&e1);
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (c.shared == 2) {
          // neighbouring CTAs (2k, 2k+1) read the same region: stride applied per pair through span / (sms/2)
          pull_kernel<<<sms, 128, smem>>>(buf, c.span / 2, c.span / sms / 1024 * 1024, chunk, stages, iters, clocks);
        } else {
          pull_kernel<<<sms, 128, smem>>>(buf, c.span, stride, chunk, stages, iters, clocks);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      unsigned long long h[1024];
      cudaMemcpy(h, clocks, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
      double mean_clk = 0;
      for (int i = 0; i < sms; ++i) mean_clk += (double)h[i] / sms;
      const double bytes = (double)per_cta * sms;
      printf("%-34s %8d %6d %10.1f %10.2f %12.1f\n", c.name, chunk, stages, best * 1e3, bytes / (best * 1e-3) / 1e12,
             (double)per_cta / mean_clk);
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    }
  return 0;
}
