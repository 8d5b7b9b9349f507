// Hardware probe (not part of the product library): how does tcgen05.mma read a SWIZZLE_128B operand
// whose start address is NOT 1024-byte aligned and whose 8-row groups are spaced by SBO = 1280 B?
// This decides whether a 3x3 convolution can reuse ONE smem halo patch [(TH+2) x (TW+2) pixels x 64 ch]
// for all 9 taps (rows of the MMA operand = a shifted window of the patch).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I joligen_b200/csrc -o gpurun_out/umma_probe tools/umma_probe.cu
//
// Halo patch: 18 x 10 pixels (TH=16, TW=8), pixel p at byte p*128 (64 bf16), TMA-written with SWIZZLE_128B.
// K-major probe (fwd conv A operand): MMA row i = (th=i/8, tw=i%8) reads patch pixel (th+r)*10 + (tw+s).
// MN-major probe (wgrad X operand): MMA K index k = (th=k/8, tw=k%8) reads the same pixel; M = 64 channels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <math.h>

#include "ptx.cuh"
using namespace jg;

constexpr int PW = 10, PH = 18, NPIX = PW * PH;  // 180 pixels

struct Params {
  int r, s;          // tap
  int base_offset;   // descriptor base_offset field
  int mode;          // 0 = K-major A shifted window, 1 = MN-major A (M = channels, K = pixels) shifted window
  float* out;        // [128][64] (mode 0) or [128 (64 used)][64] (mode 1)
};

__device__ __forceinline__ uint64_t desc_with_base(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = make_smem_desc_sw128(addr, lbo, sbo);
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmHalo, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmDY, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sHalo = smem;                 // 180 * 128 = 23040 B  (padded to 23552)
  uint8_t* sB = smem + 23552;            // mode 0: B [64 n][64 k] K-major 8 KB; mode 1: dY [128 pix][64 ch] 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 23552 + 16384);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t bbytes = p.mode == 0 ? 8192 : 16384;
    mbar_arrive_expect_tx(&bars[0], NPIX * 128 + bbytes);
    tma_load_2d(sHalo, &tmHalo, &bars[0], 0, 0);
    if (p.mode == 0) tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    else tma_load_2d(sB, &tmDY, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sHalo) + (p.r * PW + p.s) * 128;
    if (p.mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_with_base(a0 + k * 32, 16, PW * 128, p.base_offset);
        const uint64_t bd = make_smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024);
        umma_bf16(tmem, ad, bd, idesc, k != 0);
      }
    } else {
      // D[m = channel of X][n = channel of dY] = sum_k X[pix(k)][m] * dY[k][n];  M = 128 (rows 64.. read garbage: ignored)
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 8; ++k) {  // 16 pixels (2 groups of 8) per MMA
        const uint64_t ad = desc_with_base(a0 + k * 2 * PW * 128, 0 /*second 64-ch block: same data*/, PW * 128,
                                           p.base_offset);
        const uint64_t bd = make_smem_desc_sw128(smem_u32(sB) + k * 2048, 8192, 1024);
        umma_bf16(tmem, ad, bd, idesc, k != 0);
      }
    }
    umma_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
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
  std::vector<float> halo(NPIX * 64), B(64 * 64), dY(128 * 64);
  srand(1);
  for (auto& x : halo) x = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : dY) x = bf((rand() % 2001 - 1000) / 1000.f);
  auto upload = [](const std::vector<float>& h) {
    std::vector<__nv_bfloat16> t(h.size());
    for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
    __nv_bfloat16* d;
    cudaMalloc(&d, t.size() * 2);
    cudaMemcpy(d, t.data(), t.size() * 2, cudaMemcpyHostToDevice);
    return d;
  };
  __nv_bfloat16 *dH = upload(halo), *dB = upload(B), *dDY = upload(dY);
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
  CUtensorMap tmH = mk2d(dH, NPIX, NPIX), tmB = mk2d(dB, 64, 64), tmDY = mk2d(dDY, 128, 128);
  float* dout;
  cudaMalloc(&dout, 128 * 64 * 4);
  const int smem = 23552 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * 64);
  const int taps[5][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}, {2, 2}};
  for (int mode = 0; mode < 2; ++mode)
    for (int t = 0; t < 5; ++t)
      for (int variant = 0; variant < 2; ++variant) {
        const int r = taps[t][0], s = taps[t][1];
        Params p;
        p.r = r; p.s = s; p.mode = mode; p.out = dout;
        const int start_row = r * PW + s;
        p.base_offset = variant == 0 ? 0 : (start_row & 7);
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(tmH, tmB, tmDY, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d tap (%d,%d) variant %d: CUDA error %s\n", mode, r, s, variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        if (mode == 0) {
          for (int i = 0; i < 128; ++i) {
            const int pix = (i / 8 + r) * PW + (i % 8 + s);
            for (int n = 0; n < 64; ++n) {
              double acc = 0;
              for (int k = 0; k < 64; ++k) acc += (double)halo[pix * 64 + k] * B[n * 64 + k];
              maxerr = fmax(maxerr, fabs(acc - out[i * 64 + n]));
              maxref = fmax(maxref, fabs(acc));
            }
          }
        } else {
          for (int m = 0; m < 64; ++m)
            for (int n = 0; n < 64; ++n) {
              double acc = 0;
              for (int k = 0; k < 128; ++k) {
                const int pix = (k / 8 + r) * PW + (k % 8 + s);
                acc += (double)halo[pix * 64 + m] * dY[k * 64 + n];
              }
              maxerr = fmax(maxerr, fabs(acc - out[m * 64 + n]));
              maxref = fmax(maxref, fabs(acc));
            }
        }
        printf("PROBE mode=%s tap=(%d,%d) start_row=%d base_offset=%d  max_err=%.4f (max_ref %.2f)  %s\n",
               mode == 0 ? "K-major" : "MN-major", r, s, start_row, p.base_offset, maxerr, maxref,
               maxerr < 1e-2 * maxref ? "MATCH" : "mismatch");
      }
  return 0;
}
