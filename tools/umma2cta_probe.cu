// Hardware probe (not part of the product library): tcgen05.mma.cta_group::2 on a CTA pair.
//   D[256 x N] = A[256 x 64] * B[N x 64]^T, bf16 -> fp32, K-major SWIZZLE_128B operands.
// CTA r of the pair stages rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B in ITS OWN shared
// memory by TMA (.cta_group::2 loads that signal the LEADER's mbarrier), the leader's elected thread issues the MMAs,
// the completion is multicast to both CTAs, each CTA reads its 128 accumulator lanes from its own TMEM.
// What this establishes before conv_halo.cu relies on it: the B-operand split, the descriptor semantics (same
// shared-memory offsets in both CTAs), the barrier / commit / TMA signalling protocol, TMEM allocation for pairs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I joligen_b200/csrc -o tools/umma2cta_probe.bin tools/umma2cta_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "ptx.cuh"
using namespace jg;

struct Params {
  int N;       // 64, 128 or 256
  float* out;  // [256][N]
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // 128 rows x 128 B = 16 KB
  uint8_t* sB = smem + 16384;   // N/2 rows x 128 B <= 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);  // [0] full (leader's is used), [1] mma done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t ncols = p.N < 32 ? 32 : p.N;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc_2cta(tmem_ptr, ncols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t a_bytes = 16384, b_bytes = (p.N / 2) * 128;
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(&bars[0], 2 * (a_bytes + b_bytes));  // both CTAs' loads land on the leader's barrier
    const uint32_t lead_bar = mapa_u32(smem_u32(&bars[0]), 0);
    tma_load_2d_2cta(sA, &tmA, lead_bar, 0, rank * 128);
    tma_load_2d_2cta(sB, &tmB, lead_bar, 0, rank * (p.N / 2));
    if (rank == 0) {
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(256, p.N, 0, 0);
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_smem_desc_sw128(smem_u32(sA) + k * 32, 16, 1024);
        const uint64_t bd = make_smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024);
        umma_bf16_2cta(tmem, ad, bd, idesc, k != 0);
      }
      umma_commit_2cta(&bars[1], 0x3);
    }
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);  // both CTAs: the multicast commit arrives on each CTA's own barrier
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < p.N; c += 32) {
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[(size_t)(rank * 128 + warp * 32 + lane) * p.N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) tmem_dealloc_2cta(tmem, ncols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  srand(3);
  std::vector<float> A(256 * 64), B(256 * 64);
  for (auto& x : A) x = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = bf((rand() % 2001 - 1000) / 1000.f);
  auto upload = [](const std::vector<float>& h) {
    std::vector<__nv_bfloat16> t(h.size());
    for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
    __nv_bfloat16* d;
    cudaMalloc(&d, t.size() * 2);
    cudaMemcpy(d, t.data(), t.size() * 2, cudaMemcpyHostToDevice);
    return d;
  };
  __nv_bfloat16 *dA = upload(A), *dB = upload(B);
  auto mk2d = [&](void* base, uint64_t rows, uint32_t boxrows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, boxrows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  float* dout;
  cudaMalloc(&dout, 256 * 256 * 4);
  const int smem = 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ns[3] = {64, 128, 256};
  for (int t = 0; t < 3; ++t) {
    const int N = Ns[t];
    CUtensorMap tmA = mk2d(dA, 256, 128), tmB = mk2d(dB, N, N / 2);
    Params p{N, dout};
    cudaMemset(dout, 0, 256 * 256 * 4);
    probe2_kernel<<<2, 128, smem>>>(tmA, tmB, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(256 * N);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < 256; ++i)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += (double)A[i * 64 + k] * B[n * 64 + k];
        maxerr = fmax(maxerr, fabs(acc - out[(size_t)i * N + n]));
        maxref = fmax(maxref, fabs(acc));
      }
    printf("PROBE2 cta_group::2 M=256 N=%d: max_err=%.4f (max_ref %.2f)  %s\n", N, maxerr, maxref,
           maxerr < 1e-2 * maxref ? "MATCH" : "mismatch");
  }
  return 0;
}
