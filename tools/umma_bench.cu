// Hardware probe (not part of the product library): cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of
// N and of the operand major-ness, with all operands resident in shared memory (no loads in the timed region).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I joligen_b200/csrc -o tools/umma_bench.bin tools/umma_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"
using namespace jg;

struct Cfg {
  int N;        // 64 / 128 / 256
  int mn_major; // 0: K-major A and B; 1: MN-major A and B
  int lbo_a;    // LBO of A in bytes (MN-major only): 8192 = disjoint 64-wide blocks, 128 = overlapping windows
  int sbo_a;    // SBO of A in bytes
  int iters;
};

__global__ void __launch_bounds__(128, 1) bench_kernel(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.mn_major, c.mn_major);
    const uint32_t a_addr = smem_u32(smem);
    const uint32_t b_addr = smem_u32(smem + 32768);
    uint64_t a_desc, b_desc;
    uint32_t a_step, b_step;
    if (c.mn_major) {
      a_desc = make_smem_desc_sw128(a_addr, c.lbo_a, c.sbo_a);
      b_desc = make_smem_desc_sw128(b_addr, 8192, 1024);
      a_step = 2 * c.sbo_a;
      b_step = 2048;
    } else {
      a_desc = make_smem_desc_sw128(a_addr, 16, c.sbo_a);
      b_desc = make_smem_desc_sw128(b_addr, 16, 1024);
      a_step = 32;
      b_step = 32;
    }
    const long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + (it & 1) * 256, desc_advance(a_desc, k * a_step), desc_advance(b_desc, k * b_step), idesc,
                    1u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Cfg cfgs[] = {
      {64, 0, 0, 1024, 2000},  {128, 0, 0, 1024, 2000}, {256, 0, 0, 1024, 2000}, {64, 0, 0, 1280, 2000},
      {256, 0, 0, 1280, 2000}, {64, 1, 8192, 1024, 2000}, {128, 1, 8192, 1024, 2000}, {256, 1, 8192, 1024, 2000},
      {64, 1, 128, 1280, 2000}, {64, 1, 1024, 1280, 2000}, {128, 1, 128, 1280, 2000},
  };
  for (const Cfg& c : cfgs) {
    for (int grid : {1, 148}) {
      bench_kernel<<<grid, 128, smem>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("error %s\n", cudaGetErrorString(e));
        return 1;
      }
      long long cyc;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)cyc / (c.iters * 4.0);
      printf("UMMA M128 N%-3d %s lboA %-5d sboA %-5d grid %-3d : %.1f cycles/MMA  (ideal %d) -> %.0f%% of peak\n", c.N,
             c.mn_major ? "MN-major" : "K-major ", c.lbo_a, c.sbo_a, grid, per, c.N / 2, 100.0 * (c.N / 2) / per);
    }
  }
  return 0;
}
