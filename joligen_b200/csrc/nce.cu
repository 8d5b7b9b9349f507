// CUT contrastive path: PatchSampleF's position gather, the L2 normalisation of the
// pooled features and the PatchNCE loss.  Reference: models/modules/cut_networks.py:38-73,
// models/modules/NCE/base_NCE.py:17-77.  The two Linear layers of PatchSampleF's MLP run as 1x1 tcgen05 convolutions
// (conv_igemm.cu) on the gathered rows; everything here is small fp32 / bf16 row work:
//   rows = B * P (P = 256 patches), feature width D = 256  ->  4096 x 256 per NCE layer at batch 16.
//
// STATUS: written at the end of round 1; one run on a B200 (profiles/r01_cut_tests_first_run.log): gather / scatter bit
// exact, L2 normalisation and PatchNCE forward / backward within 1e-4 of oracle/cut_oracle.py
// (tests/test_gpu_widen_cut.py).  Not tuned: one warp per row, keys re-read from L2.
#include "common.cuh"

namespace jg {
namespace {

constexpr int kNceMaxPerLane = 16;  // feature width D <= 32 * 16 = 512, D % 32 == 0

__device__ __forceinline__ float warp_sum_nce(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 32;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---- gather / scatter of spatial positions (the same positions for every image of the batch) ------------------------
// dst[(b * P + p)][c] = src[(b * HW + ids[p])][c]      (8-channel bf16 vectors)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, int lds, const long long* __restrict__ ids,
                                   __nv_bfloat16* __restrict__ dst, int ldd, int B, int HW, int P, int C) {
  const int vecs = C / 8;
  const long long total = (long long)B * P * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long row = i / vecs;
    const int p = (int)(row % P);
    const int b = (int)(row / P);
    const long long s = (long long)b * HW + ids[p];
    *reinterpret_cast<uint4*>(dst + row * ldd + v * 8) = *reinterpret_cast<const uint4*>(src + s * lds + v * 8);
  }
}

// dsrc[(b * HW + ids[p])][c] = ddst[(b * P + p)][c]; every other row of dsrc is zero (memset by the caller side);
// ids are distinct (a prefix of a permutation), so there are no collisions.
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ ddst, int ldd, const long long* __restrict__ ids,
                                    __nv_bfloat16* __restrict__ dsrc, int lds, int B, int HW, int P, int C) {
  const int vecs = C / 8;
  const long long total = (long long)B * P * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long row = i / vecs;
    const int p = (int)(row % P);
    const int b = (int)(row / P);
    const long long s = (long long)b * HW + ids[p];
    *reinterpret_cast<uint4*>(dsrc + s * lds + v * 8) = *reinterpret_cast<const uint4*>(ddst + row * ldd + v * 8);
  }
}

// ---- F.normalize(x, dim=1, eps): y = x / max(||x||_2, eps), one warp per row, bf16 in -> fp32 out -------------------
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, float* __restrict__ y, float* __restrict__ norms,
                  long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  const int per = D / 32;
  for (long long row = warp0; row < rows; row += nwarps) {
    float f[kNceMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e) {
      if (e < per) {
        f[e] = __bfloat162float(x[row * ldx + lane + 32 * e]);
        s = fmaf(f[e], f[e], s);
      }
    }
    const float n = sqrtf(warp_sum_nce(s));
    const float inv = 1.f / fmaxf(n, eps);
    if (lane == 0) norms[row] = n;
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) y[row * D + lane + 32 * e] = f[e] * inv;
  }
}

// dx = (dy - y * <y, dy>) / ||x||   when ||x|| > eps (the clamp is inactive);  dx = dy / eps otherwise
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ norms,
                  __nv_bfloat16* __restrict__ dx, int lddx, long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  const int per = D / 32;
  for (long long row = warp0; row < rows; row += nwarps) {
    float fy[kNceMaxPerLane], fd[kNceMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e) {
      if (e < per) {
        fy[e] = y[row * D + lane + 32 * e];
        fd[e] = dy[row * D + lane + 32 * e];
        s = fmaf(fy[e], fd[e], s);
      }
    }
    const float dot = warp_sum_nce(s);
    const float n = norms[row];
    const bool clamped = n <= eps;
    const float inv = 1.f / fmaxf(n, eps);
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e) {
      if (e < per) {
        const float g = clamped ? fd[e] * inv : (fd[e] - fy[e] * dot) * inv;
        dx[row * lddx + lane + 32 * e] = __float2bfloat16(g);
      }
    }
  }
}

// ---- PatchNCE -------------------------------------------------------------------------------------------------------
// rows = G * P (G groups: the images of the batch, or ONE group holding the whole minibatch); for row i of group g:
//   out_i = [ <q_i, k_i>, <q_i, k_j> for j in the group with the diagonal j == i replaced by -10 ] / T
//   loss_i = logsumexp(out_i) - out_i[0]
// One warp per query row; the row's q lives in registers (D / 32 per lane), keys stream from L2.

__device__ __forceinline__ float nce_dot(const float* qreg, const float* __restrict__ krow, float* kreg, int per,
                                         int lane) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e) {
    if (e < per) {
      kreg[e] = krow[lane + 32 * e];
      s = fmaf(qreg[e], kreg[e], s);
    }
  }
  return warp_sum_nce(s);
}

__global__ void __launch_bounds__(128)
patch_nce_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, int G, int P, int D, float invT,
                     float* __restrict__ loss, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)G * P) return;  // warp-uniform
  const int per = D / 32;
  const long long g0 = (row / P) * P;   // first row of this group
  const int i = (int)(row - g0);
  float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) qr[e] = q[row * D + lane + 32 * e];
  const float pos = nce_dot(qr, k + row * D, kr, per, lane) * invT;
  float m = pos, s = 1.f;  // running max / sum of exp(. - m), the positive logit first
  for (int j = 0; j < P; ++j) {
    float l = nce_dot(qr, k + (g0 + j) * D, kr, per, lane);
    l = (j == i ? -10.f : l) * invT;
    const float mn = fmaxf(m, l);
    s = s * __expf(m - mn) + __expf(l - mn);
    m = mn;
  }
  const float lz = m + __logf(s);
  if (lane == 0) {
    lse[row] = lz;
    loss[row] = lz - pos;
  }
}

// dq_i = g_i / T * [ (p_i0 - 1) k_i + sum_{j != i} p_ij k_j ],  p = softmax(out_i)   (k is detached in the positive)
__global__ void __launch_bounds__(128)
patch_nce_bwd_q_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ lse,
                       const float* __restrict__ gout, int G, int P, int D, float invT, float* __restrict__ dq) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)G * P) return;
  const int per = D / 32;
  const long long g0 = (row / P) * P;
  const int i = (int)(row - g0);
  const float lz = lse[row];
  float qr[kNceMaxPerLane], kr[kNceMaxPerLane], acc[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) qr[e] = q[row * D + lane + 32 * e];
  const float pos = nce_dot(qr, k + row * D, kr, per, lane) * invT;
  const float c0 = __expf(pos - lz) - 1.f;
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) acc[e] = c0 * kr[e];
  for (int j = 0; j < P; ++j) {
    if (j == i) continue;  // the diagonal entry is the constant -10: no gradient
    const float l = nce_dot(qr, k + (g0 + j) * D, kr, per, lane) * invT;
    const float p = __expf(l - lz);
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) acc[e] = fmaf(p, kr[e], acc[e]);
  }
  const float sc = gout[row] * invT;
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) dq[row * D + lane + 32 * e] = acc[e] * sc;
}

// dk_j = 1 / T * sum_{i != j} g_i p_ij q_i   (the negatives are NOT detached in the reference, base_NCE.py:61-77)
__global__ void __launch_bounds__(128)
patch_nce_bwd_k_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ lse,
                       const float* __restrict__ gout, int G, int P, int D, float invT, float* __restrict__ dk) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)G * P) return;
  const int per = D / 32;
  const long long g0 = (row / P) * P;
  const int j = (int)(row - g0);
  float kreg[kNceMaxPerLane], qr[kNceMaxPerLane], acc[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e) {
    if (e < per) {
      kreg[e] = k[row * D + lane + 32 * e];
      acc[e] = 0.f;
    }
  }
  for (int i = 0; i < P; ++i) {
    if (i == j) continue;
    const float l = nce_dot(kreg, q + (g0 + i) * D, qr, per, lane) * invT;  // <q_i, k_j>, q_i left in qr
    const float w = gout[g0 + i] * __expf(l - lse[g0 + i]);
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) acc[e] = fmaf(w, qr[e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) dk[row * D + lane + 32 * e] = acc[e] * invT;
}

}  // namespace
}  // namespace jg

using namespace jg;

extern "C" int jg_gather_rows(const void* src, int lds, const int64_t* ids, void* dst, int ldd, int B, int HW, int P,
                              int C, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && ids && dst && B > 0 && HW > 0 && P > 0 && P <= HW, JG_ERR_INVALID, "gather_rows: bad args");
  JG_CHECK(C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && lds >= C && ldd >= C, JG_ERR_INVALID,
           "gather_rows: C=%d lds=%d ldd=%d must be multiples of 8", C, lds, ldd);
  gather_rows_kernel<<<grid_for((long long)B * P * (C / 8), 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(src), lds, reinterpret_cast<const long long*>(ids),
      static_cast<__nv_bfloat16*>(dst), ldd, B, HW, P, C);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_gather_rows_bwd(const void* ddst, int ldd, const int64_t* ids, void* dsrc, int lds, int B, int HW,
                                  int P, int C, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(ddst && ids && dsrc && B > 0 && HW > 0 && P > 0 && P <= HW, JG_ERR_INVALID, "gather_rows_bwd: bad args");
  JG_CHECK(C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && lds >= C && ldd >= C, JG_ERR_INVALID,
           "gather_rows_bwd: C=%d lds=%d ldd=%d must be multiples of 8", C, lds, ldd);
  JG_CUDA(cudaMemsetAsync(dsrc, 0, sizeof(__nv_bfloat16) * (size_t)B * HW * lds, stream));
  scatter_rows_kernel<<<grid_for((long long)B * P * (C / 8), 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(ddst), ldd, reinterpret_cast<const long long*>(ids),
      static_cast<__nv_bfloat16*>(dsrc), lds, B, HW, P, C);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_l2norm_fwd(const void* x, int ldx, float* y, float* norms, int64_t rows, int D, float eps,
                             jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && y && norms && rows > 0, JG_ERR_INVALID, "l2norm_fwd: null pointer / no rows");
  JG_CHECK(D > 0 && D % 32 == 0 && D <= 32 * kNceMaxPerLane && ldx >= D, JG_ERR_INVALID, "l2norm_fwd: D=%d ldx=%d", D,
           ldx);
  l2norm_fwd_kernel<<<grid_for(rows, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, y, norms, rows,
                                                          D, eps);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_l2norm_bwd(const float* y, const float* dy, const float* norms, void* dx, int lddx, int64_t rows,
                             int D, float eps, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(y && dy && norms && dx && rows > 0, JG_ERR_INVALID, "l2norm_bwd: null pointer / no rows");
  JG_CHECK(D > 0 && D % 32 == 0 && D <= 32 * kNceMaxPerLane && lddx >= D, JG_ERR_INVALID, "l2norm_bwd: D=%d lddx=%d",
           D, lddx);
  l2norm_bwd_kernel<<<grid_for(rows, 8), 256, 0, stream>>>(y, dy, norms, static_cast<__nv_bfloat16*>(dx), lddx, rows,
                                                          D, eps);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

static int nce_check(const float* q, const float* k, int G, int P, int D, float T, const char* what) {
  JG_CHECK(q && k && G > 0 && P > 0 && T > 0.f, JG_ERR_INVALID, "%s: bad args", what);
  JG_CHECK(D > 0 && D % 32 == 0 && D <= 32 * kNceMaxPerLane, JG_ERR_INVALID, "%s: D=%d must be a multiple of 32, <= %d",
           what, D, 32 * kNceMaxPerLane);
  return JG_OK;
}

extern "C" int jg_patch_nce_fwd(const float* q, const float* k, int G, int P, int D, float T, float* loss, float* lse,
                                jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = nce_check(q, k, G, P, D, T, "patch_nce_fwd")) return rc;
  JG_CHECK(loss && lse, JG_ERR_INVALID, "patch_nce_fwd: null output");
  const long long rows = (long long)G * P;
  patch_nce_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(q, k, G, P, D, 1.f / T, loss, lse);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_patch_nce_bwd(const float* q, const float* k, const float* lse, const float* grad_loss, int G, int P,
                                int D, float T, float* dq, float* dk, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = nce_check(q, k, G, P, D, T, "patch_nce_bwd")) return rc;
  JG_CHECK(lse && grad_loss && (dq || dk), JG_ERR_INVALID, "patch_nce_bwd: null pointer");
  const long long rows = (long long)G * P;
  const unsigned grid = (unsigned)((rows + 3) / 4);
  if (dq) {
    patch_nce_bwd_q_kernel<<<grid, 128, 0, stream>>>(q, k, lse, grad_loss, G, P, D, 1.f / T, dq);
    JG_LAUNCH_CHECK();
  }
  if (dk) {
    patch_nce_bwd_k_kernel<<<grid, 128, 0, stream>>>(q, k, lse, grad_loss, G, P, D, 1.f / T, dk);
    JG_LAUNCH_CHECK();
  }
  return JG_OK;
}
