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


// ---- MoNCE: PatchNCE whose negatives are re-weighted by Sinkhorn optimal-transport weights ---------------------------
// models/modules/NCE/monce.py:12-33 + sinkhorn.py (cost "hard", eps 1, 50 iterations):
//   C_ij = <q_i, k_j>, diagonal -10;  K = exp(C);  u = v = 1;  repeat: u_i = 1 / sum_j K_ij v_j,  v_j = 1 / sum_i u_i K_ij
//   f_ij = u_i K_ij v_j (popt - 1) + 1e-8;   out_i = [ <q_i, k_i> / T,  C_ij / T + log f_ij (diagonal: -10 / T) ]
//   loss_i = logsumexp(out_i) - out_i[0]
// The reference differentiates through the iterations w.r.t. q (k is detached inside the OT): the backward runs the
// reverse sweep over the stored u^t, v^t.  C, K (and, backward, the adjoint of K and the direct softmax weights) live
// in an L2-resident workspace of P x P floats per group (image).
//
// Launch structure: the first version ran ONE CTA per image through all 50 iterations (two dependent P x P mat-vec
// products each, K streamed from L2 by 256 threads): 5.4 ms forward and 13.3 ms backward per call, 187 of the 229 ms of
// a CUT step (profiles/r02_bench_cfg3.json).  Now every phase is its own grid-wide launch — a warp per matrix row for
// the row products, a block per 32 columns for the column products — so each of the ~100 sequential half-iterations
// costs one small launch (~3 us) instead of a latency-bound sweep by a single CTA, and the backward no longer updates
// the P x P adjoint in every iteration: Kbar = Kbar_0 + sum_t (u^t (x) sbar^t + rbar^t (x) v^{t-1}) is assembled once,
// in the final pass that consumes it, from the stored vectors.
constexpr int kMonceThreads = 256;
constexpr int kMonceWarps = kMonceThreads / 32;

// C = q k^T (diagonal kept: the loss pass overrides it), K = exp(C) with exp(-10) on the diagonal; V[0] = 1.
// grid (ceil(P / warps), G): one warp per row i.
__global__ void __launch_bounds__(kMonceThreads)
monce_ck_kernel(const float* __restrict__ q, const float* __restrict__ k, int P, int D, int iters,
                float* __restrict__ Cmat, float* __restrict__ Kmat, float* __restrict__ V) {
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kMonceWarps + warp;
  const int per = D / 32;
  if (i >= P) return;
  const float* qg = q + (size_t)g * P * D;
  const float* kg = k + (size_t)g * P * D;
  float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) qr[e] = qg[(size_t)i * D + lane + 32 * e];
  float* Cg = Cmat + ((size_t)g * P + i) * P;
  float* Kg = Kmat + ((size_t)g * P + i) * P;
  for (int j = 0; j < P; ++j) {
    const float c = nce_dot(qr, kg + (size_t)j * D, kr, per, lane);
    if (lane == 0) {
      Cg[j] = c;
      Kg[j] = __expf(j == i ? -10.f : c);
    }
  }
  if (lane == 0) V[(size_t)g * (iters + 1) * P + i] = 1.f;
}

// out[g][i] = post( sum_j M[g][i][j] * vec[g][j] ), one warp per row.  mode 0 (forward u-step): out = 1 / sum.
// mode 1 (backward): vec_j = -vbar_j * vt_j^2 (= sbar^t, also stored), out = rbar_i = -(ubar_i + sum) * ut_i^2.
__global__ void __launch_bounds__(kMonceThreads)
monce_rows_kernel(const float* __restrict__ Kmat, int P, int mode, const float* __restrict__ vec,
                  const float* __restrict__ vt, const float* __restrict__ ut, const float* __restrict__ ubar,
                  float* __restrict__ sbar_out, float* __restrict__ out, size_t vec_stride, size_t vt_stride,
                  size_t ut_stride, size_t out_stride) {
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kMonceWarps + warp;
  if (i >= P) return;
  const float* row = Kmat + ((size_t)g * P + i) * P;
  const float* v = vec + (size_t)g * vec_stride;
  float s = 0.f;
  if (mode == 0) {
    for (int j = lane; j < P; j += 32) s = fmaf(row[j], v[j], s);
    s = warp_sum_nce(s);
    if (lane == 0) out[(size_t)g * out_stride + i] = 1.f / s;
  } else {
    const float* vtg = vt + (size_t)g * vt_stride;
    for (int j = lane; j < P; j += 32) {
      const float sb = -v[j] * vtg[j] * vtg[j];
      if (i == 0) sbar_out[(size_t)g * out_stride + j] = sb;  // (row 0's warp records sbar^t for the final pass)
      s = fmaf(row[j], sb, s);
    }
    s = warp_sum_nce(s);
    if (lane == 0) {
      const float ui = ut[(size_t)g * ut_stride + i];
      const float ub = ubar ? ubar[(size_t)g * P + i] : 0.f;
      out[(size_t)g * out_stride + i] = -(ub + s) * ui * ui;
    }
  }
}

// out[g][j] = post( sum_i vec[g][i] * M[g][i][j] ): block (32 columns x 8 row slices).  mode 0: out = 1 / sum; 1: sum.
__global__ void __launch_bounds__(kMonceThreads)
monce_cols_kernel(const float* __restrict__ Kmat, int P, int mode, const float* __restrict__ vec, float* __restrict__ out,
                  size_t vec_stride, size_t out_stride) {
  __shared__ float part[kMonceWarps][32];
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const float* M = Kmat + (size_t)g * P * P;
  const float* v = vec + (size_t)g * vec_stride;
  float s = 0.f;
  if (j < P)
    for (int i = slice; i < P; i += kMonceWarps) s = fmaf(v[i], M[(size_t)i * P + j], s);
  part[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && j < P) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kMonceWarps; ++w) t += part[w][lane];
    out[(size_t)g * out_stride + j] = mode == 0 ? 1.f / t : t;
  }
}

// loss_i = logsumexp([pos_i, C_ij / T + log f_ij (j != i), -10 / T (j == i)]) - pos_i, f = u_i K_ij v_j (popt-1) + 1e-8
__global__ void __launch_bounds__(kMonceThreads)
monce_loss_kernel(const float* __restrict__ q, const float* __restrict__ k, int P, int D, float invT, float popt1,
                  const float* __restrict__ Cmat, const float* __restrict__ Kmat, const float* __restrict__ uF,
                  const float* __restrict__ vF, size_t u_stride, size_t v_stride, float* __restrict__ loss,
                  float* __restrict__ lse) {
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kMonceWarps + warp;
  const int per = D / 32;
  if (i >= P) return;
  const float* qg = q + (size_t)g * P * D;
  const float* kg = k + (size_t)g * P * D;
  const float* Cg = Cmat + ((size_t)g * P + i) * P;
  const float* Kg = Kmat + ((size_t)g * P + i) * P;
  const float* sv = vF + (size_t)g * v_stride;
  float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) qr[e] = qg[(size_t)i * D + lane + 32 * e];
  const float pos = nce_dot(qr, kg + (size_t)i * D, kr, per, lane) * invT;
  const float ui = uF[(size_t)g * u_stride + i];
  float m = pos;
  for (int j = lane; j < P; j += 32) {
    const float f = ui * Kg[j] * sv[j] * popt1 + 1e-8f;
    const float l = (j == i) ? -10.f * invT : Cg[j] * invT + __logf(f);
    m = fmaxf(m, l);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < P; j += 32) {
    const float f = ui * Kg[j] * sv[j] * popt1 + 1e-8f;
    const float l = (j == i) ? -10.f * invT : Cg[j] * invT + __logf(f);
    s += __expf(l - m);
  }
  s = warp_sum_nce(s) + __expf(pos - m);
  const float lz = m + __logf(s);
  if (lane == 0) {
    lse[(size_t)g * P + i] = lz;
    loss[(size_t)g * P + i] = lz - pos;
  }
}

// Backward, first pass (warp per row): direct softmax weights W_ij = g_i p_ij / T, Kbar_0 = Fbar_ij u_i v_j with
// Fbar_ij = g_i p_ij (popt-1) / f_ij, ubar_i = sum_j Fbar_ij K_ij v_j, vbar_j += Fbar_ij u_i K_ij (global atomics),
// p0_i = exp(pos_i - lse_i).
__global__ void __launch_bounds__(kMonceThreads)
monce_bwd_init_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ lse,
                      const float* __restrict__ gout, int P, int D, float invT, float popt1,
                      const float* __restrict__ Cmat, const float* __restrict__ Kmat, const float* __restrict__ uF,
                      const float* __restrict__ vF, size_t u_stride, size_t v_stride, float* __restrict__ Kbar,
                      float* __restrict__ Wmat, float* __restrict__ ubar, float* __restrict__ vbar,
                      float* __restrict__ p0s) {
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kMonceWarps + warp;
  const int per = D / 32;
  if (i >= P) return;
  const float* qg = q + (size_t)g * P * D;
  const float* kg = k + (size_t)g * P * D;
  const size_t rowoff = ((size_t)g * P + i) * P;
  const float* sv = vF + (size_t)g * v_stride;
  float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) qr[e] = qg[(size_t)i * D + lane + 32 * e];
  const float pos = nce_dot(qr, kg + (size_t)i * D, kr, per, lane) * invT;
  const float ui = uF[(size_t)g * u_stride + i], gi = gout[(size_t)g * P + i], lzi = lse[(size_t)g * P + i];
  float ub = 0.f;
  for (int j = lane; j < P; j += 32) {
    const float kij = Kmat[rowoff + j];
    const float f = ui * kij * sv[j] * popt1 + 1e-8f;
    float w = 0.f, fb = 0.f;
    if (j != i) {
      const float pij = __expf(Cmat[rowoff + j] * invT + __logf(f) - lzi);
      w = gi * pij * invT;
      fb = gi * pij * popt1 / f;
    }
    Wmat[rowoff + j] = w;
    Kbar[rowoff + j] = fb * ui * sv[j];
    ub = fmaf(fb * kij, sv[j], ub);
    atomicAdd(&vbar[(size_t)g * P + j], fb * ui * kij);
  }
  ub = warp_sum_nce(ub);
  if (lane == 0) {
    ubar[(size_t)g * P + i] = ub;
    p0s[(size_t)g * P + i] = __expf(pos - lzi);
  }
}

// Backward, last pass (warp per row i): Kbar_ij = Kbar_0 + sum_t (u^t_i sbar^t_j + rbar^t_i v^{t}_j),
// dq_i = g_i (p0_i - 1) / T k_i + sum_{j != i} (W_ij + Kbar_ij K_ij) k_j
__global__ void __launch_bounds__(kMonceThreads)
monce_bwd_dq_kernel(const float* __restrict__ k, const float* __restrict__ gout, int P, int D, float invT, int iters,
                    const float* __restrict__ Kmat, const float* __restrict__ Kbar, const float* __restrict__ Wmat,
                    const float* __restrict__ U, const float* __restrict__ V, const float* __restrict__ SB,
                    const float* __restrict__ RB, const float* __restrict__ p0s, float* __restrict__ dq) {
  extern __shared__ float sm[];  // per warp: u^t_i and rbar^t_i for all t: 2 * iters floats
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kMonceWarps + warp;
  const int per = D / 32;
  if (i >= P) return;
  float* ui_t = sm + (size_t)warp * 2 * iters;
  float* rb_t = ui_t + iters;
  const float* Ug = U + (size_t)g * iters * P;
  const float* Vg = V + (size_t)g * (iters + 1) * P;
  const float* SBg = SB + (size_t)g * iters * P;
  const float* RBg = RB + (size_t)g * iters * P;
  for (int t = lane; t < iters; t += 32) {
    ui_t[t] = Ug[(size_t)t * P + i];
    rb_t[t] = RBg[(size_t)t * P + i];
  }
  __syncwarp();
  const float* kg = k + (size_t)g * P * D;
  const size_t rowoff = ((size_t)g * P + i) * P;
  float kr[kNceMaxPerLane], acc[kNceMaxPerLane];
  const float c0 = gout[(size_t)g * P + i] * (p0s[(size_t)g * P + i] - 1.f) * invT;
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) acc[e] = c0 * kg[(size_t)i * D + lane + 32 * e];
  // each lane assembles the coefficient of 1 of every 32 columns, then the warp walks the columns together
  for (int j0 = 0; j0 < P; j0 += 32) {
    const int jl = j0 + lane;
    float c = 0.f;
    if (jl < P && jl != i) {
      float kb = Kbar[rowoff + jl];
      for (int t = 0; t < iters; ++t)
        kb = fmaf(ui_t[t], SBg[(size_t)t * P + jl], fmaf(rb_t[t], Vg[(size_t)t * P + jl], kb));
      c = Wmat[rowoff + jl] + kb * Kmat[rowoff + jl];
    }
    for (int l = 0; l < 32 && j0 + l < P; ++l) {
      const float cj = __shfl_sync(0xffffffffu, c, l);
      if (cj == 0.f) continue;
#pragma unroll
      for (int e = 0; e < kNceMaxPerLane; ++e) {
        if (e < per) {
          kr[e] = kg[(size_t)(j0 + l) * D + lane + 32 * e];
          acc[e] = fmaf(cj, kr[e], acc[e]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) dq[((size_t)g * P + i) * D + lane + 32 * e] = acc[e];
}

// dk_j = sum_i W_ij q_i (warp per column j)
__global__ void __launch_bounds__(kMonceThreads)
monce_bwd_dk_kernel(const float* __restrict__ q, int P, int D, const float* __restrict__ Wmat, float* __restrict__ dk) {
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * kMonceWarps + warp;
  const int per = D / 32;
  if (j >= P) return;
  const float* qg = q + (size_t)g * P * D;
  const float* Wg = Wmat + (size_t)g * P * P;
  float acc[kNceMaxPerLane];
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e) acc[e] = 0.f;
  for (int i = 0; i < P; ++i) {
    const float w = Wg[(size_t)i * P + j];
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) acc[e] = fmaf(w, qg[(size_t)i * D + lane + 32 * e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < kNceMaxPerLane; ++e)
    if (e < per) dk[((size_t)g * P + j) * D + lane + 32 * e] = acc[e];
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


static int monce_check(const float* q, const float* k, int G, int P, int D, float T, int iters, const char* what) {
  if (int rc = nce_check(q, k, G, P, D, T, what)) return rc;
  JG_CHECK(P <= 4096 && iters > 0 && iters <= 1000, JG_ERR_UNSUPPORTED, "%s: P=%d iters=%d", what, P, iters);
  return JG_OK;
}

// ws layout: C [G,P,P] | K [G,P,P] | U [G,iters,P] | V [G,iters+1,P] | (backward) Kbar [G,P,P] | W [G,P,P] |
//            SB [G,iters,P] | RB [G,iters,P] | ubar [G,P] | vbar [G,P] | p0 [G,P]
extern "C" size_t jg_monce_ws_floats(int G, int P, int iters, int backward) {
  const size_t pp = (size_t)G * P * P;
  return (backward ? 4 : 2) * pp + (size_t)G * (2 * iters + 1) * P + (backward ? (size_t)G * (2 * iters + 3) * P : 0);
}

extern "C" int jg_monce_fwd(const float* q, const float* k, int G, int P, int D, float T, int num_patches_opt, int iters,
                            float* ws, float* loss, float* lse, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = monce_check(q, k, G, P, D, T, iters, "monce_fwd")) return rc;
  JG_CHECK(ws && loss && lse, JG_ERR_INVALID, "monce_fwd: null pointer");
  const size_t pp = (size_t)G * P * P;
  float* Cm = ws;
  float* Km = ws + pp;
  float* U = ws + 2 * pp;
  float* V = U + (size_t)G * iters * P;
  const dim3 rows(ceil_div(P, kMonceWarps), G), cols(ceil_div(P, 32), G);
  const size_t us = (size_t)iters * P, vs = (size_t)(iters + 1) * P;
  monce_ck_kernel<<<rows, kMonceThreads, 0, stream>>>(q, k, P, D, iters, Cm, Km, V);
  JG_LAUNCH_CHECK();
  for (int t = 0; t < iters; ++t) {
    // u^t = 1 / (K v^t);  v^{t+1} = 1 / (K^T u^t)
    monce_rows_kernel<<<rows, kMonceThreads, 0, stream>>>(Km, P, 0, V + (size_t)t * P, nullptr, nullptr, nullptr, nullptr,
                                                         U + (size_t)t * P, vs, 0, 0, us);
    JG_LAUNCH_CHECK();
    monce_cols_kernel<<<cols, kMonceThreads, 0, stream>>>(Km, P, 0, U + (size_t)t * P, V + (size_t)(t + 1) * P, us, vs);
    JG_LAUNCH_CHECK();
  }
  monce_loss_kernel<<<rows, kMonceThreads, 0, stream>>>(q, k, P, D, 1.f / T, (float)(num_patches_opt - 1), Cm, Km,
                                                       U + (size_t)(iters - 1) * P, V + (size_t)iters * P, us, vs, loss,
                                                       lse);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_monce_bwd(const float* q, const float* k, const float* lse, const float* grad_loss, int G, int P, int D,
                            float T, int num_patches_opt, int iters, float* ws, float* dq, float* dk,
                            jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = monce_check(q, k, G, P, D, T, iters, "monce_bwd")) return rc;
  JG_CHECK(ws && lse && grad_loss && dq, JG_ERR_INVALID, "monce_bwd: null pointer");
  const size_t pp = (size_t)G * P * P;
  float* Cm = ws;
  float* Km = ws + pp;
  float* U = ws + 2 * pp;
  float* V = U + (size_t)G * iters * P;
  float* Kb = V + (size_t)G * (iters + 1) * P;
  float* Wm = Kb + pp;
  float* SB = Wm + pp;
  float* RB = SB + (size_t)G * iters * P;
  float* ubar = RB + (size_t)G * iters * P;
  float* vbar = ubar + (size_t)G * P;
  float* p0s = vbar + (size_t)G * P;
  const dim3 rows(ceil_div(P, kMonceWarps), G), cols(ceil_div(P, 32), G);
  const size_t us = (size_t)iters * P, vs = (size_t)(iters + 1) * P;
  JG_CUDA(cudaMemsetAsync(vbar, 0, sizeof(float) * (size_t)G * P, stream));
  monce_bwd_init_kernel<<<rows, kMonceThreads, 0, stream>>>(q, k, lse, grad_loss, P, D, 1.f / T,
                                                           (float)(num_patches_opt - 1), Cm, Km,
                                                           U + (size_t)(iters - 1) * P, V + (size_t)iters * P, us, vs, Kb,
                                                           Wm, ubar, vbar, p0s);
  JG_LAUNCH_CHECK();
  // reverse sweep: sbar^t = -vbar v^{t+1}^2;  rbar^t = -(ubar + K sbar^t) u^t^2 (ubar only in the first step);
  // vbar <- K^T rbar^t.  The rank-1 updates of Kbar are applied in the final pass from the stored sbar^t / rbar^t.
  for (int t = iters - 1; t >= 0; --t) {
    monce_rows_kernel<<<rows, kMonceThreads, 0, stream>>>(Km, P, 1, vbar, V + (size_t)(t + 1) * P, U + (size_t)t * P,
                                                         t == iters - 1 ? ubar : nullptr, SB + (size_t)t * P,
                                                         RB + (size_t)t * P, (size_t)P, vs, us, us);
    JG_LAUNCH_CHECK();
    monce_cols_kernel<<<cols, kMonceThreads, 0, stream>>>(Km, P, 1, RB + (size_t)t * P, vbar, us, (size_t)P);
    JG_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)kMonceWarps * 2 * iters * sizeof(float);
  if (smem > 48 * 1024)
    JG_CUDA(cudaFuncSetAttribute(monce_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  monce_bwd_dq_kernel<<<rows, kMonceThreads, smem, stream>>>(k, grad_loss, P, D, 1.f / T, iters, Km, Kb, Wm, U, V, SB, RB,
                                                           p0s, dq);
  JG_LAUNCH_CHECK();
  if (dk) {
    monce_bwd_dk_kernel<<<rows, kMonceThreads, 0, stream>>>(q, P, D, Wm, dk);
    JG_LAUNCH_CHECK();
  }
  return JG_OK;
}
