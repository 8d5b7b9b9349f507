// Kernels of the video UNet's MotionModule (unet_generator_attn_vid.py:374-590, 862-1054), NHWC bf16 tokens:
//   LayerNorm over channels (+ sinusoidal frame positional encoding)   TemporalTransformerBlock.norms / ff_norm,
//                                                                      PositionalEncoding (:932-947)
//   temporal self-attention over the F frames of one pixel             VersatileAttention (:950-1054) / _attention (:758)
//   GEGLU  a * gelu(gate)                                              GEGLU (:908-929)
// The Linear layers around them are 1x1 convolutions on the same NHWC tensors (conv_igemm.cu).  All three are
// HBM-bound elementwise / tiny-reduction kernels; the token count of config 5 is small (8 x 128^2 at the top level),
// so these are written for clarity: one warp per token (LayerNorm), one thread per (pixel, head) (attention).
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

__device__ __forceinline__ void unpack8v(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8v(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLnMaxVec = 4;  // 8-channel vectors per lane: C <= 32 * 4 * 8 = 1024

// sum over the G-lane group of the calling lane (G a power of two; groups are aligned)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- LayerNorm forward: G lanes per token, 32 / G tokens per warp ----------------------------------------------------
// y = (x - mean) * rstd * gamma + beta (+ pe[frame][c]),  frame = (row / HW) % F;  stats[row] = (mean, rstd)
// (the MotionModule's widths are 64 ... 512: a full warp per 64-channel token would leave 24 lanes idle)
template <int G, int KV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                     long long rows, int C, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ pe, int HW, int F, float* __restrict__ stats) {
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31;
  const int sub = lane % G;
  const int vecs = C / 8;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  for (long long base = warp0 * R; base < rows; base += nwarps * R) {
    const long long row = base + lane / G;
    const bool live = row < rows;
    float f[KV][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int v = sub + G * k;
      if (live && v < vecs) {
        unpack8v(*reinterpret_cast<const uint4*>(x + row * ldx + v * 8), f[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[k][j];
      }
    }
    const float mean = group_sum<G>(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      if (live && sub + G * k < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[k][j] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rstd = rsqrtf(group_sum<G>(q) / (float)C + eps);
    if (!live) continue;
    if (sub == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
    const float* per = pe ? pe + (size_t)((row / HW) % F) * C : nullptr;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int v = sub + G * k;
      if (v < vecs) {
        float gm[8], bt[8], o[8];
        *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(gamma + v * 8);
        *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        *reinterpret_cast<float4*>(bt) = *reinterpret_cast<const float4*>(beta + v * 8);
        *reinterpret_cast<float4*>(bt + 4) = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
        if (per) {
          const float4 p0 = *reinterpret_cast<const float4*>(per + v * 8);
          const float4 p1 = *reinterpret_cast<const float4*>(per + v * 8 + 4);
          bt[0] += p0.x; bt[1] += p0.y; bt[2] += p0.z; bt[3] += p0.w;
          bt[4] += p1.x; bt[5] += p1.y; bt[6] += p1.z; bt[7] += p1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((f[k][j] - mean) * rstd, gm[j], bt[j]);
        *reinterpret_cast<uint4*>(y + row * ldy + v * 8) = pack8v(o);
      }
    }
  }
}

// ---- LayerNorm backward ----------------------------------------------------------------------------------------------
// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)) (+ addend),  g = dy * gamma;  dgamma += dy * xhat, dbeta += dy
// (per-lane register accumulators over the lane's rows, block reduction in shared memory, one atomic per channel)
// addend: the gradient of the transformer's residual branch (x feeds the norm AND the residual add: both gradients meet
// here instead of in an add kernel).  CS: dx_colsum[c] += sum over rows of dx — the bias gradient of the Linear that
// produced x, for free.
template <int G, int KV, bool CS>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy, int lddy,
                     __nv_bfloat16* __restrict__ dx, int lddx, const __nv_bfloat16* __restrict__ addend, int ldadd,
                     long long rows, int C, const float* __restrict__ gamma, const float* __restrict__ stats,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum) {
  extern __shared__ float sm[];  // (2 + CS) * C floats
  constexpr int NCS = CS ? KV : 1;
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31;
  const int sub = lane % G;
  const int vecs = C / 8;
  for (int i = threadIdx.x; i < (CS ? 3 : 2) * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float ag[KV][8], ab[KV][8], gm[KV][8], ac[NCS][8];
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    const int v = sub + G * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ag[k][j] = ab[k][j] = 0.f;
      if (CS) ac[k][j] = 0.f;
      gm[k][j] = v < vecs ? gamma[v * 8 + j] : 0.f;
    }
  }
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  for (long long base = warp0 * R; base < rows; base += nwarps * R) {
    const long long row = base + lane / G;
    const bool live = row < rows;
    float mean = 0.f, rstd = 0.f;
    if (live) {
      const float2 st = *reinterpret_cast<const float2*>(stats + row * 2);
      mean = st.x;
      rstd = st.y;
    }
    float xh[KV][8], g[KV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int v = sub + G * k;
      if (live && v < vecs) {
        float fx[8], fd[8];
        unpack8v(*reinterpret_cast<const uint4*>(x + row * ldx + v * 8), fx);
        unpack8v(*reinterpret_cast<const uint4*>(dy + row * lddy + v * 8), fd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[k][j] = (fx[j] - mean) * rstd;
          g[k][j] = fd[j] * gm[k][j];
          s1 += g[k][j];
          s2 = fmaf(g[k][j], xh[k][j], s2);
          ag[k][j] = fmaf(fd[j], xh[k][j], ag[k][j]);
          ab[k][j] += fd[j];
        }
      }
    }
    const float m1 = group_sum<G>(s1) / (float)C, m2 = group_sum<G>(s2) / (float)C;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int v = sub + G * k;
      if (live && v < vecs) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[k][j] - m1 - xh[k][j] * m2);
        if (addend) {
          float e[8];
          unpack8v(*reinterpret_cast<const uint4*>(addend + row * ldadd + v * 8), e);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += e[j];
        }
        if (CS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ac[k][j] += o[j];
        }
        *reinterpret_cast<uint4*>(dx + row * lddx + v * 8) = pack8v(o);
      }
    }
  }
  // the R row groups of a warp hold the same channels: fold them with shuffles before touching shared memory
#pragma unroll
  for (int k = 0; k < KV; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = G; o < 32; o <<= 1) {
        ag[k][j] += __shfl_xor_sync(0xffffffffu, ag[k][j], o);
        ab[k][j] += __shfl_xor_sync(0xffffffffu, ab[k][j], o);
        if (CS) ac[k][j] += __shfl_xor_sync(0xffffffffu, ac[k][j], o);
      }
    }
    const int v = sub + G * k;
    if (lane < G && v < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sm[v * 8 + j], ag[k][j]);
        atomicAdd(&sm[C + v * 8 + j], ab[k][j]);
        if (CS) atomicAdd(&sm[2 * C + v * 8 + j], ac[k][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], sm[i]);
    atomicAdd(&dbeta[i], sm[C + i]);
    if (CS) atomicAdd(&dx_colsum[i], sm[2 * C + i]);
  }
}

// (G, KV) for a width: the smallest power-of-two lane group that covers C / 8 vectors, then 2 / 4 vectors per lane
#define JG_LN_DISPATCH(vecs, CALL)            \
  do {                                        \
    if ((vecs) <= 8) { CALL(8, 1); }          \
    else if ((vecs) <= 16) { CALL(16, 1); }   \
    else if ((vecs) <= 32) { CALL(32, 1); }   \
    else if ((vecs) <= 64) { CALL(32, 2); }   \
    else { CALL(32, 4); }                     \
  } while (0)

static inline int ln_rows_per_warp(int vecs) { return vecs <= 8 ? 4 : vecs <= 16 ? 2 : 1; }

// ---- temporal self-attention over F <= 8 frames ----------------------------------------------------------------------
// qkv: [B*F][HW][ldqkv] with channels (q | k | v), each heads*ch wide (the three Linear layers packed as one GEMM);
// thread = (b, pixel, head): S = scale * Q K^T (F x F), P = softmax_j(S), O = P V.  Channels are streamed in 8-wide
// vectors, the F x F matrices live in registers.
constexpr int kTF = 8;

template <bool BWD>
__global__ void __launch_bounds__(128)
temporal_attn_kernel(const __nv_bfloat16* __restrict__ qkv, int ldqkv, const __nv_bfloat16* __restrict__ d_out,
                     int lddo, __nv_bfloat16* __restrict__ out /*fwd: o, bwd: dqkv*/, int ldout, int B, int F, int HW,
                     int heads, int ch, float scale) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * HW * heads;
  if (t >= total) return;
  const int h = (int)(t % heads);
  const int d = (int)((t / heads) % HW);
  const int b = (int)(t / ((long long)heads * HW));
  const int C = heads * ch;
  const size_t fstride_in = (size_t)HW * ldqkv;
  const __nv_bfloat16* qb = qkv + ((size_t)b * F * HW + d) * ldqkv + h * ch;
  float S[kTF][kTF];
#pragma unroll
  for (int i = 0; i < kTF; ++i)
#pragma unroll
    for (int j = 0; j < kTF; ++j) S[i][j] = 0.f;
  for (int c = 0; c < ch; c += 8) {
    float q[kTF][8], k[kTF][8];
#pragma unroll
    for (int i = 0; i < kTF; ++i) {
      if (i < F) {
        unpack8v(*reinterpret_cast<const uint4*>(qb + i * fstride_in + c), q[i]);
        unpack8v(*reinterpret_cast<const uint4*>(qb + i * fstride_in + C + c), k[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kTF; ++i)
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (i < F && j < F) {
#pragma unroll
          for (int e = 0; e < 8; ++e) S[i][j] = fmaf(q[i][e], k[j][e], S[i][j]);
        }
  }
  // softmax over j (fp32)
#pragma unroll
  for (int i = 0; i < kTF; ++i) {
    if (i < F) {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (j < F) m = fmaxf(m, S[i][j] * scale);
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (j < F) {
          S[i][j] = __expf(S[i][j] * scale - m);
          den += S[i][j];
        }
      const float inv = 1.f / den;
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (j < F) S[i][j] *= inv;
    }
  }
  if (!BWD) {
    __nv_bfloat16* ob = out + ((size_t)b * F * HW + d) * ldout + h * ch;
    const size_t fstride_out = (size_t)HW * ldout;
    for (int c = 0; c < ch; c += 8) {
      float v[kTF][8];
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (j < F) unpack8v(*reinterpret_cast<const uint4*>(qb + j * fstride_in + 2 * C + c), v[j]);
#pragma unroll
      for (int i = 0; i < kTF; ++i) {
        if (i < F) {
          float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int j = 0; j < kTF; ++j)
            if (j < F) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = fmaf(S[i][j], v[j][e], o[e]);
            }
          *reinterpret_cast<uint4*>(ob + i * fstride_out + c) = pack8v(o);
        }
      }
    }
    return;
  }
  // ---- backward: dP = dO V^T, D_i = sum_j P dP, dS = P (dP - D) * scale; dQ = dS K, dK = dS^T Q, dV = P^T dO
  const __nv_bfloat16* dob = d_out + ((size_t)b * F * HW + d) * lddo + h * ch;
  const size_t fstride_do = (size_t)HW * lddo;
  float dS[kTF][kTF];
#pragma unroll
  for (int i = 0; i < kTF; ++i)
#pragma unroll
    for (int j = 0; j < kTF; ++j) dS[i][j] = 0.f;
  for (int c = 0; c < ch; c += 8) {
    float g[kTF][8], v[kTF][8];
#pragma unroll
    for (int i = 0; i < kTF; ++i) {
      if (i < F) {
        unpack8v(*reinterpret_cast<const uint4*>(dob + i * fstride_do + c), g[i]);
        unpack8v(*reinterpret_cast<const uint4*>(qb + i * fstride_in + 2 * C + c), v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kTF; ++i)
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (i < F && j < F) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dS[i][j] = fmaf(g[i][e], v[j][e], dS[i][j]);
        }
  }
#pragma unroll
  for (int i = 0; i < kTF; ++i) {
    if (i < F) {
      float D = 0.f;
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (j < F) D = fmaf(S[i][j], dS[i][j], D);
#pragma unroll
      for (int j = 0; j < kTF; ++j)
        if (j < F) dS[i][j] = S[i][j] * (dS[i][j] - D) * scale;
    }
  }
  __nv_bfloat16* gb = out + ((size_t)b * F * HW + d) * ldout + h * ch;
  const size_t fstride_g = (size_t)HW * ldout;
  for (int c = 0; c < ch; c += 8) {
    float q[kTF][8], k[kTF][8], g[kTF][8];
#pragma unroll
    for (int i = 0; i < kTF; ++i) {
      if (i < F) {
        unpack8v(*reinterpret_cast<const uint4*>(qb + i * fstride_in + c), q[i]);
        unpack8v(*reinterpret_cast<const uint4*>(qb + i * fstride_in + C + c), k[i]);
        unpack8v(*reinterpret_cast<const uint4*>(dob + i * fstride_do + c), g[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kTF; ++i) {
      if (i < F) {
        float dq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < kTF; ++j)
          if (j < F) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              dq[e] = fmaf(dS[i][j], k[j][e], dq[e]);
              dk[e] = fmaf(dS[j][i], q[j][e], dk[e]);
              dv[e] = fmaf(S[j][i], g[j][e], dv[e]);
            }
          }
        *reinterpret_cast<uint4*>(gb + i * fstride_g + c) = pack8v(dq);
        *reinterpret_cast<uint4*>(gb + i * fstride_g + C + c) = pack8v(dk);
        *reinterpret_cast<uint4*>(gb + i * fstride_g + 2 * C + c) = pack8v(dv);
      }
    }
  }
}

// ---- GEGLU: y = a * gelu(g), (a | g) = the two halves of the projection (exact erf GELU = F.gelu default) -----------
__device__ __forceinline__ float gelu_f(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float g) {
  return 0.5f * (1.f + erff(g * 0.70710678118654752f)) + g * 0.3989422804014327f * __expf(-0.5f * g * g);
}

__global__ void geglu_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                 long long rows, int Cout) {
  const int vecs = Cout / 8;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += gridDim.x * (long long)blockDim.x) {
    const long long row = i / vecs;
    const int v = (int)(i % vecs);
    float a[8], g[8], o[8];
    unpack8v(*reinterpret_cast<const uint4*>(x + row * ldx + v * 8), a);
    unpack8v(*reinterpret_cast<const uint4*>(x + row * ldx + Cout + v * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_f(g[j]);
    *reinterpret_cast<uint4*>(y + row * ldy + v * 8) = pack8v(o);
  }
}

__global__ void geglu_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy,
                                 int lddy, __nv_bfloat16* __restrict__ dx, int lddx, long long rows, int Cout) {
  const int vecs = Cout / 8;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += gridDim.x * (long long)blockDim.x) {
    const long long row = i / vecs;
    const int v = (int)(i % vecs);
    float a[8], g[8], d[8], da[8], dg[8];
    unpack8v(*reinterpret_cast<const uint4*>(x + row * ldx + v * 8), a);
    unpack8v(*reinterpret_cast<const uint4*>(x + row * ldx + Cout + v * 8), g);
    unpack8v(*reinterpret_cast<const uint4*>(dy + row * lddy + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      da[j] = d[j] * gelu_f(g[j]);
      dg[j] = d[j] * a[j] * gelu_grad(g[j]);
    }
    *reinterpret_cast<uint4*>(dx + row * lddx + v * 8) = pack8v(da);
    *reinterpret_cast<uint4*>(dx + row * lddx + Cout + v * 8) = pack8v(dg);
  }
}

static int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace jg

using namespace jg;

extern "C" int jg_layernorm_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int C, float eps,
                                const float* gamma, const float* beta, const float* pe, int HW, int F, float* stats,
                                jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && y && gamma && beta && stats && rows > 0, JG_ERR_INVALID, "layernorm_fwd: null pointer / no rows");
  JG_CHECK(C % 8 == 0 && C > 0 && C <= 32 * kLnMaxVec * 8 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C,
           JG_ERR_INVALID, "layernorm_fwd: bad dims C=%d ldx=%d ldy=%d", C, ldx, ldy);
  JG_CHECK(pe == nullptr || (HW > 0 && F > 0), JG_ERR_INVALID, "layernorm_fwd: positional encoding needs HW, F");
  // two tokens per lane group: enough loads in flight per SM without a long serial loop
  const int grid = grid_for(rows, 8 * ln_rows_per_warp(C / 8) * 2);
#define JG_LN_FWD(G, KV)                                                                                           \
  layernorm_fwd_kernel<G, KV><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx,                 \
                                                        static_cast<__nv_bfloat16*>(y), ldy, rows, C, eps, gamma,  \
                                                        beta, pe, HW > 0 ? HW : 1, F > 0 ? F : 1, stats)
  JG_LN_DISPATCH(C / 8, JG_LN_FWD);
#undef JG_LN_FWD
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_layernorm_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx,
                                const void* addend, int ldadd, int64_t rows, int C, const float* gamma,
                                const float* stats, float* dgamma, float* dbeta, float* dx_colsum,
                                jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && dy && dx && gamma && stats && dgamma && dbeta && rows > 0, JG_ERR_INVALID,
           "layernorm_bwd: null pointer / no rows");
  JG_CHECK(C % 8 == 0 && C > 0 && C <= 32 * kLnMaxVec * 8 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0,
           JG_ERR_INVALID, "layernorm_bwd: bad dims C=%d", C);
  JG_CHECK(addend == nullptr || (ldadd % 8 == 0 && ldadd >= C), JG_ERR_INVALID, "layernorm_bwd: bad addend stride %d",
           ldadd);
  if (dx_colsum) JG_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * C, stream));
  JG_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, stream));
  JG_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, stream));
  int grid = grid_for(rows, 8 * ln_rows_per_warp(C / 8) * 8);  // >= 8 rows per lane group: few blocks, few atomics
  if (grid > num_sms() * 4) grid = num_sms() * 4;
#define JG_LN_BWD_CS(G, KV, CS)                                                                            \
  layernorm_bwd_kernel<G, KV, CS><<<grid, 256, (CS ? 3 : 2) * C * sizeof(float), stream>>>(                 \
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(dy), lddy,               \
      static_cast<__nv_bfloat16*>(dx), lddx, static_cast<const __nv_bfloat16*>(addend), ldadd, rows, C, gamma, \
      stats, dgamma, dbeta, dx_colsum)
#define JG_LN_BWD(G, KV)                                 \
  do {                                                   \
    if (dx_colsum) JG_LN_BWD_CS(G, KV, true);            \
    else JG_LN_BWD_CS(G, KV, false);                     \
  } while (0)
  JG_LN_DISPATCH(C / 8, JG_LN_BWD);
#undef JG_LN_BWD
#undef JG_LN_BWD_CS
  JG_LAUNCH_CHECK();
  return JG_OK;
}

static int check_tattn(int B, int F, int HW, int heads, int ch, int ldqkv) {
  JG_CHECK(B > 0 && HW > 0 && heads > 0 && F >= 1 && F <= kTF, JG_ERR_UNSUPPORTED,
           "temporal attention: 1..%d frames supported (F=%d)", kTF, F);
  JG_CHECK(ch % 8 == 0 && ch > 0 && ldqkv % 8 == 0 && ldqkv >= 3 * heads * ch, JG_ERR_INVALID,
           "temporal attention: bad dims heads=%d ch=%d ldqkv=%d", heads, ch, ldqkv);
  return JG_OK;
}

extern "C" int jg_temporal_attn_fwd(const void* qkv, int ldqkv, void* out, int ldo, int B, int F, int HW, int heads,
                                    int ch, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_tattn(B, F, HW, heads, ch, ldqkv);
  if (rc) return rc;
  JG_CHECK(qkv && out && ldo % 8 == 0 && ldo >= heads * ch, JG_ERR_INVALID, "temporal_attn_fwd: bad output");
  const long long total = (long long)B * HW * heads;
  temporal_attn_kernel<false><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), ldqkv, nullptr, 0, static_cast<__nv_bfloat16*>(out), ldo, B, F, HW,
      heads, ch, 1.f / sqrtf((float)ch));
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_temporal_attn_bwd(const void* qkv, int ldqkv, const void* d_out, int lddo, void* dqkv, int lddqkv,
                                    int B, int F, int HW, int heads, int ch, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_tattn(B, F, HW, heads, ch, ldqkv);
  if (rc) return rc;
  JG_CHECK(qkv && d_out && dqkv && lddo % 8 == 0 && lddqkv % 8 == 0 && lddqkv >= 3 * heads * ch, JG_ERR_INVALID,
           "temporal_attn_bwd: bad gradient buffers");
  const long long total = (long long)B * HW * heads;
  temporal_attn_kernel<true><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), ldqkv, static_cast<const __nv_bfloat16*>(d_out), lddo,
      static_cast<__nv_bfloat16*>(dqkv), lddqkv, B, F, HW, heads, ch, 1.f / sqrtf((float)ch));
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_geglu_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int Cout, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && y && rows > 0 && Cout % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= 2 * Cout && ldy >= Cout,
           JG_ERR_INVALID, "geglu_fwd: bad args");
  geglu_fwd_kernel<<<grid_for(rows * (Cout / 8), 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, rows, Cout);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_geglu_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int64_t rows,
                            int Cout, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && dy && dx && rows > 0 && Cout % 8 == 0 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 &&
               ldx >= 2 * Cout && lddx >= 2 * Cout && lddy >= Cout,
           JG_ERR_INVALID, "geglu_bwd: bad args");
  geglu_bwd_kernel<<<grid_for(rows * (Cout / 8), 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(dy), lddy,
      static_cast<__nv_bfloat16*>(dx), lddx, rows, Cout);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
