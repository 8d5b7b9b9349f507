// GroupNorm (+FiLM scale/shift) (+SiLU) forward / backward on NHWC bf16 with fp32 statistics.
// HBM-bound: 16-byte vector loads, per-thread channel accumulators, shared-memory then global fp32
// atomics for the per-(image, channel) sums.  Also covers the attention block's InstanceNorm1d
// (groups == C, no affine, no activation).
//
// Reference arithmetic:
//   GroupNorm wrapper (fp32 compute)     unet_attn_utils.py:42-48
//   ResBlock FiLM: out_norm(h)*(1+scale)+shift -> SiLU    unet_generator_attn.py:250-258
//   in_layers: GroupNorm -> SiLU         unet_generator_attn.py:186-189
//   normalization1d = InstanceNorm1d     unet_attn_utils.py:60-66,116-117
//
// forward:   y = act(x*a[n,c] + b[n,c]),   a = rstd*gamma*(1+scale),  b = (beta - mean*rstd*gamma)*(1+scale) + shift
// backward:  du = dy*act'(u);  A[n,c] = sum du, B[n,c] = sum du*x;  everything else (dgamma, dbeta, dscale,
//            dshift, the two group means) is a function of A,B;  dx = k1[n,c]*du + k2[n,g]*x + k3[n,g].
#include "act.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

constexpr int kNormThreads = 256;

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// streaming 16-byte load (read once: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream(const __nv_bfloat16* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
// Packs eight floats to bf16 and returns the packed vector (the values that are actually stored).
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
#define JG_ACT_DISPATCH(act, ...)                                                    \
  switch (act) {                                                                     \
    case JG_ACT_SILU: { constexpr int ACT = JG_ACT_SILU; __VA_ARGS__; } break;       \
    case JG_ACT_RELU: { constexpr int ACT = JG_ACT_RELU; __VA_ARGS__; } break;       \
    case JG_ACT_LRELU02: { constexpr int ACT = JG_ACT_LRELU02; __VA_ARGS__; } break; \
    default: { constexpr int ACT = JG_ACT_NONE; __VA_ARGS__; } break;                \
  }

// Block-level sum over the row lanes of per-thread 8-channel partial sums, without shared-memory atomics (fp32
// shared atomics are compare-and-swap loops: with 32 row lanes per channel they cost several microseconds per block).
// part: [Q][rstep][C] scratch, out: [Q][C].  Every thread of the block must call it.
constexpr int kPartFloats = kNormThreads * 8;  // rstep * C <= 256 * 8
template <int Q>
__device__ __forceinline__ void block_channel_sums(const float (&vals)[Q][8], bool active, int rl, int v, int rstep, int C,
                                                   float* part, float* out) {
  if (active) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float* dst = part + ((size_t)q * rstep + rl) * C + v * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(vals[q][0], vals[q][1], vals[q][2], vals[q][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(vals[q][4], vals[q][5], vals[q][6], vals[q][7]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Q * C; i += kNormThreads) {
    const int q = i / C, c = i - q * C;
    const float* src = part + (size_t)q * rstep * C + c;
    float acc = 0.f;
    for (int r = 0; r < rstep; ++r) acc += src[(size_t)r * C];
    out[i] = acc;
  }
  __syncthreads();
}

// ---- pass 1 (fwd): per-(n, group) sum and sum of squares --------------------------------------
// grid (chunks, N); smem: kPartFloats*2 + 2*C floats.  Eight rows (8 x 16 B) are in flight per thread.
__global__ void __launch_bounds__(kNormThreads, 4)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int HW, int C, int groups, int rows_per_block,
                float* __restrict__ sums /*[N][groups][2], or [N][C][2] when groups == 0 (per-channel sums)*/) {
  extern __shared__ float sm[];
  float* part = sm;
  float* csum = sm + 2 * kPartFloats;  // [2][C]: sum, sum of squares
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int rstep = kNormThreads / vecs;
  const int v = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  const bool active = rl < rstep;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, HW);
  float acc[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
  if (active) {
    const __nv_bfloat16* base = x + ((size_t)n * HW) * ldx + v * 8;
    auto reduce = [&](const uint4& u) {
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += f[j];
        acc[1][j] = fmaf(f[j], f[j], acc[1][j]);
      }
    };
    int r = r0 + rl;
    for (; r + 7 * rstep < r1; r += 8 * rstep) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = ldg_stream(base + (size_t)(r + k * rstep) * ldx);
#pragma unroll
      for (int k = 0; k < 8; ++k) reduce(u[k]);
    }
    for (; r < r1; r += rstep) reduce(ldg_stream(base + (size_t)r * ldx));
  }
  block_channel_sums<2>(acc, active, rl, v, rstep, C, part, csum);
  if (groups == 0) {  // per-channel sums (the layout the convolution epilogues accumulate into)
    for (int c = threadIdx.x; c < C; c += kNormThreads) {
      atomicAdd(&sums[((size_t)n * C + c) * 2 + 0], csum[c]);
      atomicAdd(&sums[((size_t)n * C + c) * 2 + 1], csum[C + c]);
    }
    return;
  }
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += kNormThreads) {
    float s = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      s += csum[c];
      q += csum[C + c];
    }
    atomicAdd(&sums[((size_t)n * groups + g) * 2 + 0], s);
    atomicAdd(&sums[((size_t)n * groups + g) * 2 + 1], q);
  }
}

// ---- pass 2 (fwd): mean/rstd and the per-(n,c) affine coefficients -----------------------------
// grid N, threads over C.
// chan_sums != 0: `sums` holds per-(n, channel) sums [N][C][2] (produced by a convolution epilogue or by
// gn_stats_kernel in per-channel mode) instead of per-group sums.
__global__ void gn_finalize_fwd_kernel(const float* __restrict__ sums, int chan_sums, int HW, int C, int groups, float eps,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ film /*[N][2C]*/, float* __restrict__ stats,
                                       float* __restrict__ ab /*[N][C][2]*/) {
  const int n = blockIdx.x;
  const int cpg = C / groups;
  const float cnt = (float)cpg * (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float s, q;
    if (chan_sums) {
      s = q = 0.f;
      for (int k = g * cpg; k < (g + 1) * cpg; ++k) {
        s += sums[((size_t)n * C + k) * 2 + 0];
        q += sums[((size_t)n * C + k) * 2 + 1];
      }
    } else {
      s = sums[((size_t)n * groups + g) * 2 + 0];
      q = sums[((size_t)n * groups + g) * 2 + 1];
    }
    const float mean = s / cnt;
    const float var = fmaxf(q / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    if (c == g * cpg) {
      stats[((size_t)n * groups + g) * 2 + 0] = mean;
      stats[((size_t)n * groups + g) * 2 + 1] = rstd;
    }
    const float ga = gamma ? gamma[c] : 1.f;
    const float be = beta ? beta[c] : 0.f;
    float a = rstd * ga;
    float b = be - mean * a;
    if (film) {
      const float sc = 1.f + film[(size_t)n * 2 * C + c];
      const float sh = film[(size_t)n * 2 * C + C + c];
      a *= sc;
      b = b * sc + sh;
    }
    ab[((size_t)n * C + c) * 2 + 0] = a;
    ab[((size_t)n * C + c) * 2 + 1] = b;
  }
}

// ---- pass 3 (fwd): y = act(x*a + b) -------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(kNormThreads, 3)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy, int HW, int C,
                int rows_per_block, const float* __restrict__ ab) {
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int rstep = kNormThreads / vecs;
  const int v = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  if (rl >= rstep) return;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = ab[((size_t)n * C + v * 8 + j) * 2 + 0];
    b[j] = ab[((size_t)n * C + v * 8 + j) * 2 + 1];
  }
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, HW);
  const __nv_bfloat16* xb = x + ((size_t)n * HW) * ldx + v * 8;
  __nv_bfloat16* yb = y + ((size_t)n * HW) * ldy + v * 8;
  int r = r0 + rl;
  for (; r + 7 * rstep < r1; r += 8 * rstep) {
    uint4 u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = ldg_stream(xb + (size_t)(r + k * rstep) * ldx);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = act_f<ACT>(fmaf(f[j], a[j], b[j]));
      *reinterpret_cast<uint4*>(yb + (size_t)(r + k * rstep) * ldy) = pack8(f);
    }
  }
  for (; r < r1; r += rstep) {
    float f[8];
    unpack8(ldg_stream(xb + (size_t)r * ldx), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_f<ACT>(fmaf(f[j], a[j], b[j]));
    *reinterpret_cast<uint4*>(yb + (size_t)r * ldy) = pack8(f);
  }
}

// ---- bwd pass 1: A[n,c] = sum du, B[n,c] = sum du*x ------------------------------------------
// grid (chunks, N); smem kPartFloats*2 + 2*C floats (block-level reduction before the global atomics).
template <int ACT>
__global__ void __launch_bounds__(kNormThreads, 2)
gn_bwd_sums_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy, int lddy,
                   int HW, int C, int rows_per_block, const float* __restrict__ ab,
                   float* __restrict__ AB /*[N][C][2]*/) {
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int rstep = kNormThreads / vecs;
  const int v = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  float* part = sm;
  float* tot = sm + 2 * kPartFloats;  // [2][C]: A, B
  const bool active = rl < rstep;
  float acc[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
  if (active) {
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = ab[((size_t)n * C + v * 8 + j) * 2 + 0];
      b[j] = ab[((size_t)n * C + v * 8 + j) * 2 + 1];
    }
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, HW);
    const __nv_bfloat16* xb = x + ((size_t)n * HW) * ldx + v * 8;
    const __nv_bfloat16* db = dy + ((size_t)n * HW) * lddy + v * 8;
    int r = r0 + rl;
    for (; r + 3 * rstep < r1; r += 4 * rstep) {
      uint4 ux[4], ud[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ux[k] = ldg_stream(xb + (size_t)(r + k * rstep) * ldx);
        ud[k] = ldg_stream(db + (size_t)(r + k * rstep) * lddy);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8], d[8];
        unpack8(ux[k], f);
        unpack8(ud[k], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float du = d[j];
          if (ACT != JG_ACT_NONE) du *= act_grad<ACT>(fmaf(f[j], a[j], b[j]));
          acc[0][j] += du;
          acc[1][j] = fmaf(du, f[j], acc[1][j]);
        }
      }
    }
    for (; r < r1; r += rstep) {
      float f[8], d[8];
      unpack8(ldg_stream(xb + (size_t)r * ldx), f);
      unpack8(ldg_stream(db + (size_t)r * lddy), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float du = d[j];
        if (ACT != JG_ACT_NONE) du *= act_grad<ACT>(fmaf(f[j], a[j], b[j]));
        acc[0][j] += du;
        acc[1][j] = fmaf(du, f[j], acc[1][j]);
      }
    }
  }
  block_channel_sums<2>(acc, active, rl, v, rstep, C, part, tot);
  for (int i = threadIdx.x; i < 2 * C; i += kNormThreads) {
    const int q = i / C, c = i - q * C;
    atomicAdd(&AB[((size_t)n * C + c) * 2 + q], tot[i]);
  }
}

// ---- bwd pass 2: coefficients, dFiLM ---------------------------------------------------------------
// grid N; smem 2*C floats.  K: [N][C] k1, then [N][groups][2] (k2,k3).
__global__ void gn_finalize_bwd_kernel(const float* __restrict__ AB, int HW, int C, int groups,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ film, const float* __restrict__ stats,
                                       float* __restrict__ k1, float* __restrict__ k23,
                                       float* __restrict__ dfilm /*[N][2C] or null*/,
                                       float* __restrict__ gAB /*[N][C][2]: (1+scale)*A, (1+scale)*Bhat*/) {
  extern __shared__ float sm[];
  float* t1 = sm;      // gamma_eff * A
  float* t2 = sm + C;  // gamma_eff * Bhat
  const int n = blockIdx.x;
  const int cpg = C / groups;
  const float cnt = (float)cpg * (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = stats[((size_t)n * groups + g) * 2 + 0];
    const float rstd = stats[((size_t)n * groups + g) * 2 + 1];
    const float A = AB[((size_t)n * C + c) * 2 + 0];
    const float Bx = AB[((size_t)n * C + c) * 2 + 1];
    const float Bh = rstd * (Bx - mean * A);
    const float ga = gamma ? gamma[c] : 1.f;
    const float be = beta ? beta[c] : 0.f;
    const float sc = film ? 1.f + film[(size_t)n * 2 * C + c] : 1.f;
    const float ge = ga * sc;
    t1[c] = ge * A;
    t2[c] = ge * Bh;
    k1[(size_t)n * C + c] = rstd * ge;
    if (dfilm) {
      dfilm[(size_t)n * 2 * C + c] = ga * Bh + be * A;  // d scale
      dfilm[(size_t)n * 2 * C + C + c] = A;             // d shift
    }
    gAB[((size_t)n * C + c) * 2 + 0] = sc * A;
    gAB[((size_t)n * C + c) * 2 + 1] = sc * Bh;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      s1 += t1[c];
      s2 += t2[c];
    }
    const float mean = stats[((size_t)n * groups + g) * 2 + 0];
    const float rstd = stats[((size_t)n * groups + g) * 2 + 1];
    const float m1 = s1 / cnt, m2 = s2 / cnt;
    k23[((size_t)n * groups + g) * 2 + 0] = -rstd * rstd * m2;
    k23[((size_t)n * groups + g) * 2 + 1] = rstd * rstd * m2 * mean - rstd * m1;
  }
}

// dgamma[c] = sum_n (1+scale)*Bhat, dbeta[c] = sum_n (1+scale)*A
__global__ void gn_param_grad_kernel(const float* __restrict__ gAB, int N, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sa = 0.f, sb = 0.f;
  for (int n = 0; n < N; ++n) {
    sa += gAB[((size_t)n * C + c) * 2 + 0];
    sb += gAB[((size_t)n * C + c) * 2 + 1];
  }
  if (dbeta) dbeta[c] = sa;
  if (dgamma) dgamma[c] = sb;
}

// ---- bwd pass 3: dx = k1*du + k2*x + k3 (+ addend (+ addend2)), optionally colsum[c] += sum over rows of dx ---------
// U rows are in flight per thread.  The per-channel constants take 40 registers: two blocks per SM (128 registers)
// with U = 4 / 3 / 2 (0 / 1 / 2 addends) keep ~50 KB in flight per SM without spilling.
// smem: kPartFloats + C floats (column sums).  NADD = number of extra gradients of x summed in (other consumers of x).
template <int NADD, int ACT, int U>
__global__ void __launch_bounds__(kNormThreads, 2)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy, int lddy,
                    __nv_bfloat16* __restrict__ dx, int lddx, int HW, int C, int groups, int rows_per_block,
                    const float* __restrict__ ab, const float* __restrict__ k1, const float* __restrict__ k23,
                    const __nv_bfloat16* __restrict__ addend, int ldadd, const __nv_bfloat16* __restrict__ addend2,
                    int ldadd2, float* __restrict__ colsum) {
  constexpr bool HAS_ADD = NADD >= 1;
  constexpr bool HAS_ADD2 = NADD >= 2;
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int vecs = C / 8;
  const int rstep = kNormThreads / vecs;
  const int v = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  const bool active = rl < rstep;
  float cs[1][8] = {{0, 0, 0, 0, 0, 0, 0, 0}};
  if (active) {
    const int cpg = C / groups;
    float a[8], b[8], c1[8], c2[8], c3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const int g = c / cpg;
      a[j] = ab[((size_t)n * C + c) * 2 + 0];
      b[j] = ab[((size_t)n * C + c) * 2 + 1];
      c1[j] = k1[(size_t)n * C + c];
      c2[j] = k23[((size_t)n * groups + g) * 2 + 0];
      c3[j] = k23[((size_t)n * groups + g) * 2 + 1];
    }
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, HW);
    const __nv_bfloat16* xb = x + ((size_t)n * HW) * ldx + v * 8;
    const __nv_bfloat16* db = dy + ((size_t)n * HW) * lddy + v * 8;
    __nv_bfloat16* ob = dx + ((size_t)n * HW) * lddx + v * 8;
    const __nv_bfloat16* eb = HAS_ADD ? addend + ((size_t)n * HW) * ldadd + v * 8 : nullptr;
    const __nv_bfloat16* eb2 = HAS_ADD2 ? addend2 + ((size_t)n * HW) * ldadd2 + v * 8 : nullptr;
    // one row: o = c1*du + c2*x + c3 (+ e (+ e2)), stored as bf16; the column sums are taken in fp32 before the rounding
    auto row = [&](const uint4& ux, const uint4& ud, const uint4& ua, const uint4& ua2, __nv_bfloat16* dst) {
      float f[8], d[8], o[8], e[8], e2[8];
      unpack8(ux, f);
      unpack8(ud, d);
      if (HAS_ADD) unpack8(ua, e);
      if (HAS_ADD2) unpack8(ua2, e2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float du = d[j];
        if (ACT != JG_ACT_NONE) du *= act_grad<ACT>(fmaf(f[j], a[j], b[j]));
        o[j] = fmaf(c1[j], du, fmaf(c2[j], f[j], c3[j]));
        if (HAS_ADD) o[j] += e[j];
        if (HAS_ADD2) o[j] += e2[j];
      }
      *reinterpret_cast<uint4*>(dst) = pack8(o);
      if (colsum) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[0][j] += o[j];
      }
    };
    int r = r0 + rl;
    for (; r + (U - 1) * rstep < r1; r += U * rstep) {
      uint4 ux[U], ud[U], ua[U], ua2[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        ux[k] = ldg_stream(xb + (size_t)(r + k * rstep) * ldx);
        ud[k] = ldg_stream(db + (size_t)(r + k * rstep) * lddy);
        if (HAS_ADD) ua[k] = ldg_stream(eb + (size_t)(r + k * rstep) * ldadd);
        if (HAS_ADD2) ua2[k] = ldg_stream(eb2 + (size_t)(r + k * rstep) * ldadd2);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) row(ux[k], ud[k], ua[k], ua2[k], ob + (size_t)(r + k * rstep) * lddx);
    }
    for (; r < r1; r += rstep) {
      const uint4 ux = ldg_stream(xb + (size_t)r * ldx);
      const uint4 ud = ldg_stream(db + (size_t)r * lddy);
      uint4 ua = ux, ua2 = ux;
      if (HAS_ADD) ua = ldg_stream(eb + (size_t)r * ldadd);
      if (HAS_ADD2) ua2 = ldg_stream(eb2 + (size_t)r * ldadd2);
      row(ux, ud, ua, ua2, ob + (size_t)r * lddx);
    }
  }
  if (colsum) {  // block-uniform
    float* tot = sm + kPartFloats;
    block_channel_sums<1>(cs, active, rl, v, rstep, C, sm, tot);
    for (int i = threadIdx.x; i < C; i += kNormThreads) atomicAdd(&colsum[i], tot[i]);
  }
}

// Rows per block for a grid of about per_sm blocks per SM (at least 32 rows per block).  The kernels that end in a
// block-level reduction run as ONE wave of co-resident blocks: measured on B200, twice the blocks cost 25-70% more time.
static int rows_per_block_for(int HW, int N, int per_sm) {
  const int target_blocks = num_sms() * per_sm;
  int chunks = target_blocks / (N > 0 ? N : 1);
  if (chunks < 1) chunks = 1;
  int rpb = (HW + chunks - 1) / chunks;
  if (rpb < 32) rpb = 32;
  return rpb;
}

static int check_norm_args(int N, int HW, int C, int groups, int ldx) {
  JG_CHECK(N > 0 && HW > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C, JG_ERR_INVALID,
           "groupnorm: bad dims N=%d HW=%d C=%d ld=%d", N, HW, C, ldx);
  JG_CHECK(groups > 0 && C % groups == 0, JG_ERR_INVALID, "groupnorm: C=%d not divisible by groups=%d", C, groups);
  JG_CHECK(C / 8 <= kNormThreads, JG_ERR_INVALID, "groupnorm: C=%d too large", C);
  return JG_OK;
}

}  // namespace jg

using namespace jg;

extern "C" size_t jg_groupnorm_fwd_ws_floats(int N, int C, int groups) { return (size_t)N * groups * 2; }
extern "C" size_t jg_groupnorm_bwd_ws_floats(int N, int C, int groups) {
  return (size_t)N * C * 2 /*AB*/ + (size_t)N * C /*k1*/ + (size_t)N * groups * 2 /*k23*/ + (size_t)N * C * 2 /*gAB*/;
}

namespace jg {
// Stand-alone forms of the two reductions that the convolution epilogues fuse (conv_common.cuh): used by the conv
// launchers for kernels whose epilogue cannot do it, and exported for tests.
int launch_chan_stats(const void* x, int ldx, int N, int HW, int C, float* stats, cudaStream_t stream) {
  int rc = check_norm_args(N, HW, C, 1, ldx);
  if (rc) return rc;
  const int rpb = rows_per_block_for(HW, N, 4);
  dim3 grid((HW + rpb - 1) / rpb, N);
  const size_t smem = (2 * kPartFloats + 2 * C) * sizeof(float);
  gn_stats_kernel<<<grid, kNormThreads, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, HW, C, 0, rpb, stats);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
int launch_gn_bwd_sums(const void* x, int ldx, const void* dy, int lddy, int N, int HW, int C, const float* ab, int act,
                       float* AB, cudaStream_t stream) {
  int rc = check_norm_args(N, HW, C, 1, ldx);
  if (rc) return rc;
  const int rpb = rows_per_block_for(HW, N, 2);
  dim3 grid((HW + rpb - 1) / rpb, N);
  JG_ACT_DISPATCH(act, gn_bwd_sums_kernel<ACT><<<grid, kNormThreads, (2 * kPartFloats + 2 * C) * sizeof(float), stream>>>(
                           static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(dy), lddy, HW, C,
                           rpb, ab, AB));
  JG_LAUNCH_CHECK();
  return JG_OK;
}
}  // namespace jg

extern "C" int jg_chan_stats(const void* x, int ldx, int N, int HW, int C, float* stats, jg_stream_t stream_) {
  JG_CHECK(x && stats, JG_ERR_INVALID, "chan_stats: null pointer");
  return launch_chan_stats(x, ldx, N, HW, C, stats, static_cast<cudaStream_t>(stream_));
}

extern "C" int jg_groupnorm_fwd(const void* x, int ldx, void* y, int ldy, int N, int HW, int C, int groups, float eps,
                                const float* gamma, const float* beta, const float* film, int act, float* stats,
                                float* ab, float* ws, const float* chan_stats, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_norm_args(N, HW, C, groups, ldx);
  if (rc) return rc;
  JG_CHECK(x && y && stats && ab && ws, JG_ERR_INVALID, "groupnorm_fwd: null pointer");
  JG_CHECK(ldy % 8 == 0 && ldy >= C, JG_ERR_INVALID, "groupnorm_fwd: bad ldy");
  JG_CHECK(act == JG_ACT_NONE || act == JG_ACT_SILU || act == JG_ACT_RELU || act == JG_ACT_LRELU02, JG_ERR_INVALID,
           "groupnorm_fwd: act %d unsupported", act);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  if (!chan_stats) {
    JG_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)N * groups * 2, stream));
    const int rpb = rows_per_block_for(HW, N, 4);
    dim3 grid((HW + rpb - 1) / rpb, N);
    const size_t smem = (2 * kPartFloats + 2 * C) * sizeof(float);
    gn_stats_kernel<<<grid, kNormThreads, smem, stream>>>(xb, ldx, HW, C, groups, rpb, ws);
    JG_LAUNCH_CHECK();
  }
  // chan_stats: the producer of x (a convolution epilogue, jg_conv_epilogue.stats) already summed it per channel
  gn_finalize_fwd_kernel<<<N, 256, 0, stream>>>(chan_stats ? chan_stats : ws, chan_stats != nullptr, HW, C, groups, eps,
                                                gamma, beta, film, stats, ab);
  JG_LAUNCH_CHECK();
  const int rpb = rows_per_block_for(HW, N, 8);
  dim3 grid((HW + rpb - 1) / rpb, N);
  JG_ACT_DISPATCH(act, gn_apply_kernel<ACT><<<grid, kNormThreads, 0, stream>>>(
                           xb, ldx, static_cast<__nv_bfloat16*>(y), ldy, HW, C, rpb, ab));
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_groupnorm_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, const void* addend,
                                int ldadd, const void* addend2, int ldadd2, int N, int HW, int C, int groups,
                                const float* gamma, const float* beta,
                                const float* film, int act, const float* stats, const float* ab, float* dgamma,
                                float* dbeta, float* dfilm, float* dx_colsum, float* ws, const float* sums_pre,
                                jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_norm_args(N, HW, C, groups, ldx);
  if (rc) return rc;
  JG_CHECK(x && dy && dx && stats && ab && ws, JG_ERR_INVALID, "groupnorm_bwd: null pointer");
  JG_CHECK(lddy % 8 == 0 && lddy >= C && lddx % 8 == 0 && lddx >= C, JG_ERR_INVALID, "groupnorm_bwd: bad ld");
  JG_CHECK(addend == nullptr || (ldadd % 8 == 0 && ldadd >= C), JG_ERR_INVALID, "groupnorm_bwd: bad ldadd");
  JG_CHECK(addend2 == nullptr || (addend != nullptr && ldadd2 % 8 == 0 && ldadd2 >= C), JG_ERR_INVALID,
           "groupnorm_bwd: addend2 needs addend and a valid ldadd2");
  float* AB = ws;
  float* k1 = AB + (size_t)N * C * 2;
  float* k23 = k1 + (size_t)N * C;
  float* gAB = k23 + (size_t)N * groups * 2;
  // sums_pre: A / B were accumulated by the dgrad that produced dy (jg_conv_epilogue.gn_sums): no reduction pass
  if (sums_pre) AB = const_cast<float*>(sums_pre);
  else JG_CUDA(cudaMemsetAsync(AB, 0, sizeof(float) * (size_t)N * C * 2, stream));
  if (dx_colsum) JG_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * C, stream));
  const __nv_bfloat16* xb0 = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* dyb0 = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* addb0 = static_cast<const __nv_bfloat16*>(addend);
  const __nv_bfloat16* addb20 = static_cast<const __nv_bfloat16*>(addend2);
  __nv_bfloat16* dxb0 = static_cast<__nv_bfloat16*>(dx);
  // Image chunks: the sums pass and the apply pass both read x and dy.  When the whole batch is several times the L2
  // (268 MB for 32 x 256^2 x 64 against 126 MB) the second read comes from HBM again; run in chunks of images whose
  // x + dy fit the L2 budget and the apply pass of a chunk finds what the sums pass just read.  Everything between the
  // two passes is per image, so a chunk is the same three launches on offset pointers.  JG_GN_BWD_L2_MB (default 0 =
  // one chunk).
  static const int l2_mb = getenv("JG_GN_BWD_L2_MB") ? atoi(getenv("JG_GN_BWD_L2_MB")) : 0;
  int NC = N;
  if (l2_mb > 0 && !sums_pre) {
    const size_t per_image = (size_t)HW * C * 2 * 2;
    NC = (int)std::max<size_t>(1, ((size_t)l2_mb << 20) / per_image);
    if (NC >= N) NC = N;
    else NC = (N + (N + NC - 1) / NC - 1) / ((N + NC - 1) / NC);  // equal chunks
  }
  for (int n0 = 0; n0 < N; n0 += NC) {
    const int nc = std::min(NC, N - n0);
    const int rpb1 = rows_per_block_for(HW, nc, 2);
    const int rpb = rows_per_block_for(HW, nc, 2);
    dim3 grid1((HW + rpb1 - 1) / rpb1, nc);
    dim3 grid((HW + rpb - 1) / rpb, nc);
    const __nv_bfloat16* xb = xb0 + (size_t)n0 * HW * ldx;
    const __nv_bfloat16* dyb = dyb0 + (size_t)n0 * HW * lddy;
    const __nv_bfloat16* addb = addb0 ? addb0 + (size_t)n0 * HW * ldadd : nullptr;
    const __nv_bfloat16* addb2 = addb20 ? addb20 + (size_t)n0 * HW * ldadd2 : nullptr;
    __nv_bfloat16* dxb = dxb0 + (size_t)n0 * HW * lddx;
    const float* ab_c = ab + (size_t)n0 * C * 2;
    float* AB_c = AB + (size_t)n0 * C * 2;
    float* k1_c = k1 + (size_t)n0 * C;
    float* k23_c = k23 + (size_t)n0 * groups * 2;
    if (!sums_pre) {
      JG_ACT_DISPATCH(act, gn_bwd_sums_kernel<ACT><<<grid1, kNormThreads, (2 * kPartFloats + 2 * C) * sizeof(float), stream>>>(
                               xb, ldx, dyb, lddy, HW, C, rpb1, ab_c, AB_c));
      JG_LAUNCH_CHECK();
    }
    gn_finalize_bwd_kernel<<<nc, 256, 2 * C * sizeof(float), stream>>>(
        AB_c, HW, C, groups, gamma, beta, film ? film + (size_t)n0 * 2 * C : nullptr, stats + (size_t)n0 * groups * 2, k1_c,
        k23_c, dfilm ? dfilm + (size_t)n0 * 2 * C : nullptr, gAB + (size_t)n0 * C * 2);
    JG_LAUNCH_CHECK();
    const size_t smem = dx_colsum ? (kPartFloats + C) * sizeof(float) : 0;
    if (addend2) {
      JG_ACT_DISPATCH(act, gn_bwd_apply_kernel<2, ACT, 2><<<grid, kNormThreads, smem, stream>>>(
                               xb, ldx, dyb, lddy, dxb, lddx, HW, C, groups, rpb, ab_c, k1_c, k23_c, addb, ldadd, addb2,
                               ldadd2, dx_colsum));
    } else if (addend) {
      JG_ACT_DISPATCH(act, gn_bwd_apply_kernel<1, ACT, 3><<<grid, kNormThreads, smem, stream>>>(
                               xb, ldx, dyb, lddy, dxb, lddx, HW, C, groups, rpb, ab_c, k1_c, k23_c, addb, ldadd, nullptr, 0,
                               dx_colsum));
    } else {
      JG_ACT_DISPATCH(act, gn_bwd_apply_kernel<0, ACT, 4><<<grid, kNormThreads, smem, stream>>>(
                               xb, ldx, dyb, lddy, dxb, lddx, HW, C, groups, rpb, ab_c, k1_c, k23_c, nullptr, 0, nullptr, 0,
                               dx_colsum));
    }
    JG_LAUNCH_CHECK();
  }
  if (dgamma || dbeta) {
    gn_param_grad_kernel<<<(C + 127) / 128, 128, 0, stream>>>(gAB, N, C, dgamma, dbeta);
    JG_LAUNCH_CHECK();
  }
  return JG_OK;
}
