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
// The reference differentiates through the iterations w.r.t. q (k is detached inside the OT): the backward kernel runs
// the reverse sweep over the stored u^t, v^t.  One CTA per group (image); C, K (and, backward, the adjoint of K and the
// direct softmax weights) live in an L2-resident workspace of P x P floats each.  STATUS: compiled, NOT yet run on
// hardware; the algorithm (forward + reverse sweep) was checked against autograd on the CPU in fp64.
constexpr int kMonceThreads = 256;

__device__ __forceinline__ float row_dot_K(const float* __restrict__ row, const float* __restrict__ vec, int P, int lane) {
  float s = 0.f;
  for (int j = lane; j < P; j += 32) s = fmaf(row[j], vec[j], s);
  return warp_sum_nce(s);
}

__global__ void __launch_bounds__(kMonceThreads)
monce_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, int P, int D, float invT, float popt1,
                 int iters, float* __restrict__ Cmat, float* __restrict__ Kmat, float* __restrict__ U,
                 float* __restrict__ V, float* __restrict__ loss, float* __restrict__ lse) {
  extern __shared__ float sm[];  // u [P], v [P]
  float* su = sm;
  float* sv = sm + P;
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kMonceThreads / 32;
  const int per = D / 32;
  const float* qg = q + (size_t)g * P * D;
  const float* kg = k + (size_t)g * P * D;
  float* Cg = Cmat + (size_t)g * P * P;
  float* Kg = Kmat + (size_t)g * P * P;
  float* Ug = U + (size_t)g * iters * P;
  float* Vg = V + (size_t)g * (iters + 1) * P;
  // C and K
  for (int i = warp; i < P; i += nwarps) {
    float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) qr[e] = qg[(size_t)i * D + lane + 32 * e];
    for (int j = 0; j < P; ++j) {
      const float c = nce_dot(qr, kg + (size_t)j * D, kr, per, lane);
      if (lane == 0) {
        Cg[(size_t)i * P + j] = c;
        Kg[(size_t)i * P + j] = __expf(j == i ? -10.f : c);
      }
    }
  }
  for (int j = threadIdx.x; j < P; j += kMonceThreads) {
    su[j] = 1.f;
    sv[j] = 1.f;
    Vg[j] = 1.f;
  }
  __syncthreads();
  // Sinkhorn scalings
  for (int t = 0; t < iters; ++t) {
    for (int i = warp; i < P; i += nwarps) {
      const float r = row_dot_K(Kg + (size_t)i * P, sv, P, lane);
      if (lane == 0) {
        su[i] = 1.f / r;
        Ug[(size_t)t * P + i] = 1.f / r;
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < P; j += kMonceThreads) {
      float s = 0.f;
      for (int i = 0; i < P; ++i) s = fmaf(su[i], Kg[(size_t)i * P + j], s);
      sv[j] = 1.f / s;
      Vg[(size_t)(t + 1) * P + j] = 1.f / s;
    }
    __syncthreads();
  }
  // loss
  for (int i = warp; i < P; i += nwarps) {
    float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) qr[e] = qg[(size_t)i * D + lane + 32 * e];
    const float pos = nce_dot(qr, kg + (size_t)i * D, kr, per, lane) * invT;
    const float ui = su[i];
    float m = pos;
    for (int j = lane; j < P; j += 32) {
      const float f = ui * Kg[(size_t)i * P + j] * sv[j] * popt1 + 1e-8f;
      const float l = (j == i) ? -10.f * invT : Cg[(size_t)i * P + j] * invT + __logf(f);
      m = fmaxf(m, l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < P; j += 32) {
      const float f = ui * Kg[(size_t)i * P + j] * sv[j] * popt1 + 1e-8f;
      const float l = (j == i) ? -10.f * invT : Cg[(size_t)i * P + j] * invT + __logf(f);
      s += __expf(l - m);
    }
    s = warp_sum_nce(s) + __expf(pos - m);
    const float lz = m + __logf(s);
    if (lane == 0) {
      lse[(size_t)g * P + i] = lz;
      loss[(size_t)g * P + i] = lz - pos;
    }
  }
}

__global__ void __launch_bounds__(kMonceThreads)
monce_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ lse,
                 const float* __restrict__ gout, int P, int D, float invT, float popt1, int iters,
                 const float* __restrict__ Cmat, const float* __restrict__ Kmat, const float* __restrict__ U,
                 const float* __restrict__ V, float* __restrict__ Kbar, float* __restrict__ Wmat,
                 float* __restrict__ dq, float* __restrict__ dk) {
  extern __shared__ float sm[];  // ubar, vbar, sbar, rbar, p0 : 5 * P
  float* ubar = sm;
  float* vbar = sm + P;
  float* sbar = sm + 2 * P;
  float* rbar = sm + 3 * P;
  float* p0s = sm + 4 * P;
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kMonceThreads / 32;
  const int per = D / 32;
  const float* qg = q + (size_t)g * P * D;
  const float* kg = k + (size_t)g * P * D;
  const float* Cg = Cmat + (size_t)g * P * P;
  const float* Kg = Kmat + (size_t)g * P * P;
  const float* Ug = U + (size_t)g * iters * P;
  const float* Vg = V + (size_t)g * (iters + 1) * P;
  float* Kb = Kbar + (size_t)g * P * P;
  float* Wg = Wmat + (size_t)g * P * P;
  const float* lz = lse + (size_t)g * P;
  const float* go = gout + (size_t)g * P;
  const float* uF = Ug + (size_t)(iters - 1) * P;  // final u, v
  const float* vF = Vg + (size_t)iters * P;
  for (int j = threadIdx.x; j < P; j += kMonceThreads) vbar[j] = 0.f;
  __syncthreads();
  // softmax weights: W_ij = g_i p_ij / T (direct term), Fbar_ij = g_i p_ij (popt-1) / f_ij; ubar, Kbar initialised
  for (int i = warp; i < P; i += nwarps) {
    float qr[kNceMaxPerLane], kr[kNceMaxPerLane];
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) qr[e] = qg[(size_t)i * D + lane + 32 * e];
    const float pos = nce_dot(qr, kg + (size_t)i * D, kr, per, lane) * invT;
    const float ui = uF[i], gi = go[i], lzi = lz[i];
    float ub = 0.f;
    for (int j = lane; j < P; j += 32) {
      const float kij = Kg[(size_t)i * P + j];
      const float f = ui * kij * vF[j] * popt1 + 1e-8f;
      float w = 0.f, fb = 0.f;
      if (j != i) {
        const float pij = __expf(Cg[(size_t)i * P + j] * invT + __logf(f) - lzi);
        w = gi * pij * invT;
        fb = gi * pij * popt1 / f;
      }
      Wg[(size_t)i * P + j] = w;
      Kb[(size_t)i * P + j] = fb * ui * vF[j];
      ub = fmaf(fb * kij, vF[j], ub);
      atomicAdd(&vbar[j], fb * ui * kij);
    }
    ub = warp_sum_nce(ub);
    if (lane == 0) {
      ubar[i] = ub;
      p0s[i] = __expf(pos - lzi);
    }
  }
  __syncthreads();
  // reverse sweep over the Sinkhorn iterations
  for (int t = iters - 1; t >= 0; --t) {
    const float* ut = Ug + (size_t)t * P;
    const float* vt = Vg + (size_t)(t + 1) * P;
    const float* vp = Vg + (size_t)t * P;
    for (int j = threadIdx.x; j < P; j += kMonceThreads) sbar[j] = -vbar[j] * vt[j] * vt[j];
    __syncthreads();
    for (int i = warp; i < P; i += nwarps) {
      const float ui = ut[i];
      float acc = 0.f;
      for (int j = lane; j < P; j += 32) {
        const float sb = sbar[j];
        acc = fmaf(Kg[(size_t)i * P + j], sb, acc);
        Kb[(size_t)i * P + j] += ui * sb;
      }
      acc = warp_sum_nce(acc);
      if (lane == 0) {
        const float ubt = ubar[i] + acc;
        rbar[i] = -ubt * ui * ui;
        ubar[i] = 0.f;
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < P; j += kMonceThreads) {
      const float vpj = vp[j];
      float acc = 0.f;
      for (int i = 0; i < P; ++i) {
        const float rb = rbar[i];
        acc = fmaf(Kg[(size_t)i * P + j], rb, acc);
        Kb[(size_t)i * P + j] += rb * vpj;
      }
      vbar[j] = acc;
    }
    __syncthreads();
  }
  // dq_i = g_i (p0 - 1) / T k_i + sum_j (W_ij + Kbar_ij K_ij) k_j   (the Sinkhorn branch reaches q only)
  for (int i = warp; i < P; i += nwarps) {
    float kr[kNceMaxPerLane], acc[kNceMaxPerLane];
    const float c0 = go[i] * (p0s[i] - 1.f) * invT;
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) acc[e] = c0 * kg[(size_t)i * D + lane + 32 * e];
    for (int j = 0; j < P; ++j) {
      if (j == i) continue;
      const float c = Wg[(size_t)i * P + j] + Kb[(size_t)i * P + j] * Kg[(size_t)i * P + j];
#pragma unroll
      for (int e = 0; e < kNceMaxPerLane; ++e) {
        if (e < per) {
          kr[e] = kg[(size_t)j * D + lane + 32 * e];
          acc[e] = fmaf(c, kr[e], acc[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < kNceMaxPerLane; ++e)
      if (e < per) dq[((size_t)g * P + i) * D + lane + 32 * e] = acc[e];
  }
  // dk_j = sum_i W_ij q_i
  if (dk) {
    for (int j = warp; j < P; j += nwarps) {
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
  }
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
  JG_CHECK(P <= 1024 && iters > 0 && iters <= 1000, JG_ERR_UNSUPPORTED,
           "%s: P=%d (one CTA per group: at most 1024 patches) iters=%d", what, P, iters);
  return JG_OK;
}

extern "C" size_t jg_monce_ws_floats(int G, int P, int iters, int backward) {
  const size_t pp = (size_t)G * P * P;
  return (backward ? 4 : 2) * pp + (size_t)G * (2 * iters + 1) * P;
}

// ws layout: C [G,P,P] | K [G,P,P] | U [G,iters,P] | V [G,iters+1,P] | (backward only) Kbar [G,P,P] | W [G,P,P]
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
  monce_fwd_kernel<<<G, kMonceThreads, 2 * P * sizeof(float), stream>>>(q, k, P, D, 1.f / T,
                                                                        (float)(num_patches_opt - 1), iters, Cm, Km, U,
                                                                        V, loss, lse);
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
  monce_bwd_kernel<<<G, kMonceThreads, 5 * P * sizeof(float), stream>>>(q, k, lse, grad_loss, P, D, 1.f / T,
                                                                        (float)(num_patches_opt - 1), iters, Cm, Km, U,
                                                                        V, Kb, Wm, dq, dk);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
