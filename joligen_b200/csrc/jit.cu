// Memory-bound kernels of the b2b video backbone (JiTViD, /root/reference/models/modules/vit/vit_vid.py; SURVEY.md 8(f)
// rank 2).  Tokens are rows of bf16 [rows][ld] tensors, rows = N * T (N = B * F frames, T tokens per frame, image n =
// row / T); per-frame modulation vectors are fp32 [N][.] slices of the adaLN Linear's output (row stride ldm).
//
//   jg_rmsnorm_mod_*     RMSNorm (util/model_util.py:165-179, fp32 statistics, eps 1e-6) + adaLN modulate
//                        x * (1 + scale) + shift (vit_vid.py:47-48, :270-279, FinalLayer :283-308)
//   jg_qknorm_rope_*     per-head RMSNorm of q and k + 2-D rotary embedding (Attention.forward :205-231,
//                        VisionRotaryEmbeddingFast util/model_util.py:97-162: rotate_half on interleaved pairs)
//   jg_attn_small_*      softmax attention of one frame's <= 128 tokens (64 patches + 32 in-context tokens at 128^2 /
//                        patch 16: no multiple of the flash kernels' tiles), fp32, one CTA per (frame, head)
//   jg_swiglu_*          SwiGLUFFN (:234-246): silu(x1) * x2 on the halves of the w12 output
//   jg_gated_residual_*  x + gate * branch (:270-279)
// The Linears around them are 1x1 tcgen05 convolutions on the same token tensors (jg_conv2d_*).
#include "act.cuh"
#include "common.cuh"

namespace jg {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float b2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---- RMSNorm + modulate ------------------------------------------------------------------------------------------------
// one warp per row
__global__ void rmsnorm_mod_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                       long long rows, int C, int T, float eps, const float* __restrict__ w,
                                       const float* __restrict__ shift, const float* __restrict__ scale, int ldm,
                                       float* __restrict__ rstd) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + row * ldx;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = b2f(xr[c]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum_f(ss);
  const float r = rsqrtf(ss / (float)C + eps);
  if (lane == 0) rstd[row] = r;
  const long long n = row / T;
  for (int c = lane; c < C; c += 32) {
    float v = w[c] * (b2f(xr[c]) * r);
    if (scale) v = fmaf(v, 1.f + scale[n * ldm + c], shift[n * ldm + c]);
    y[row * ldy + c] = __float2bfloat16(v);
  }
}

// dx: one warp per row.  g = dy * (1 + scale) * w;  dx = r * g - x * r^3 * mean(g * x)
__global__ void rmsnorm_mod_bwd_dx_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy,
                                          int lddy, __nv_bfloat16* __restrict__ dx, int lddx, long long rows, int C, int T,
                                          const float* __restrict__ w, const float* __restrict__ scale, int ldm,
                                          const float* __restrict__ rstd) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long n = row / T;
  const float r = rstd[row];
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float g = b2f(dy[row * lddy + c]) * (scale ? 1.f + scale[n * ldm + c] : 1.f) * w[c];
    dot = fmaf(g, b2f(x[row * ldx + c]), dot);
  }
  dot = warp_sum_f(dot) / (float)C;
  const float k = r * r * r * dot;
  for (int c = lane; c < C; c += 32) {
    const float g = b2f(dy[row * lddy + c]) * (scale ? 1.f + scale[n * ldm + c] : 1.f) * w[c];
    dx[row * lddx + c] = __float2bfloat16(fmaf(r, g, -k * b2f(x[row * ldx + c])));
  }
}

// parameter / modulation gradients: thread = (image n, channel c), loop over the image's T rows.
//   dshift[n][c] = sum dy; dscale[n][c] = sum dy * w * xhat; dw[c] += sum dy * (1 + scale) * xhat   (atomic over n)
__global__ void rmsnorm_mod_bwd_par_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy,
                                           int lddy, int N, int C, int T, const float* __restrict__ w,
                                           const float* __restrict__ scale, int ldm, const float* __restrict__ rstd,
                                           float* __restrict__ dw, float* __restrict__ dshift, float* __restrict__ dscale,
                                           int lddm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (c >= C) return;
  float a_shift = 0.f, a_xh = 0.f;
  for (int t = 0; t < T; ++t) {
    const long long row = (long long)n * T + t;
    const float d = b2f(dy[row * lddy + c]);
    a_shift += d;
    a_xh = fmaf(d, b2f(x[row * ldx + c]) * rstd[row], a_xh);
  }
  if (dshift) {
    dshift[(long long)n * lddm + c] = a_shift;
    dscale[(long long)n * lddm + c] = a_xh * w[c];
  }
  atomicAdd(&dw[c], a_xh * (scale ? 1.f + scale[(long long)n * ldm + c] : 1.f));
}

// ---- per-head RMSNorm of q, k + rotary ---------------------------------------------------------------------------------
// thread = (row, head, which in {q, k}); qkv channel of element i: which * D + h * hd + i
template <int HD>
__global__ void qknorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ldq, __nv_bfloat16* __restrict__ out,
                                       int ldo, long long rows, int T, int heads, float eps, const float* __restrict__ wq,
                                       const float* __restrict__ wk, const float* __restrict__ cosb,
                                       const float* __restrict__ sinb, float* __restrict__ rstd) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * heads * 2) return;
  const int which = (int)(idx % 2);
  const int h = (int)((idx / 2) % heads);
  const long long row = idx / (2 * heads);
  const int t = (int)(row % T);
  const int D = heads * HD;
  const __nv_bfloat16* src = qkv + row * ldq + which * D + h * HD;
  const float* w = which ? wk : wq;
  float v[HD];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < HD; ++i) {
    v[i] = b2f(src[i]);
    ss = fmaf(v[i], v[i], ss);
  }
  const float r = rsqrtf(ss / (float)HD + eps);
  rstd[idx] = r;
  __nv_bfloat16* dst = out + row * ldo + which * D + h * HD;
  const float* cs = cosb + (size_t)t * HD;
  const float* sn = sinb + (size_t)t * HD;
#pragma unroll
  for (int j = 0; j < HD; j += 2) {
    const float a = w[j] * v[j] * r, b = w[j + 1] * v[j + 1] * r;
    dst[j] = __float2bfloat16(a * cs[j] - b * sn[j]);
    dst[j + 1] = __float2bfloat16(b * cs[j + 1] + a * sn[j + 1]);
  }
}

template <int HD>
__global__ void qknorm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ldq, const __nv_bfloat16* __restrict__ dout,
                                       int lddo, __nv_bfloat16* __restrict__ dqkv, int lddq, long long rows, int T,
                                       int heads, const float* __restrict__ wq, const float* __restrict__ wk,
                                       const float* __restrict__ cosb, const float* __restrict__ sinb,
                                       const float* __restrict__ rstd, float* __restrict__ dwq, float* __restrict__ dwk) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * heads * 2) return;
  const int which = (int)(idx % 2);
  const int h = (int)((idx / 2) % heads);
  const long long row = idx / (2 * heads);
  const int t = (int)(row % T);
  const int D = heads * HD;
  const __nv_bfloat16* src = qkv + row * ldq + which * D + h * HD;
  const __nv_bfloat16* dsrc = dout + row * lddo + which * D + h * HD;
  const float* w = which ? wk : wq;
  float* dwp = which ? dwk : dwq;
  const float r = rstd[idx];
  const float* cs = cosb + (size_t)t * HD;
  const float* sn = sinb + (size_t)t * HD;
  float dn[HD], xv[HD];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < HD; j += 2) {
    const float d0 = b2f(dsrc[j]), d1 = b2f(dsrc[j + 1]);
    dn[j] = d0 * cs[j] + d1 * sn[j + 1];
    dn[j + 1] = d1 * cs[j + 1] - d0 * sn[j];
  }
  // dw: lanes of equal parity share `which` (idx % 2): sum them by shuffles first — per-thread atomics onto the 2 * HD
  // addresses (rows * heads * 2 threads) serialised in L2 and cost more than the rest of the block's backward
  const bool tail_warp = ((idx | 31) >= rows * heads * 2);   // (a partially filled last warp keeps per-thread atomics)
#pragma unroll
  for (int i = 0; i < HD; ++i) {
    xv[i] = b2f(src[i]);
    dot = fmaf(dn[i] * w[i], xv[i], dot);
    float g = dn[i] * xv[i] * r;
    if (!tail_warp) {
#pragma unroll
      for (int o = 2; o < 32; o <<= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if ((threadIdx.x & 31) < 2) atomicAdd(&dwp[i], g);
    } else {
      atomicAdd(&dwp[i], g);
    }
  }
  const float k = r * r * r * dot / (float)HD;
  __nv_bfloat16* dst = dqkv + row * lddq + which * D + h * HD;
#pragma unroll
  for (int i = 0; i < HD; ++i) dst[i] = __float2bfloat16(fmaf(r, dn[i] * w[i], -k * xv[i]));
}

// ---- small attention -----------------------------------------------------------------------------------------------------
// grid (heads, N), 128 threads, dynamic smem: K | V as fp32 [T][HD].  Thread t < T owns query t.
template <int HD>
__global__ void __launch_bounds__(128)
attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                      const __nv_bfloat16* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ o, int ldo,
                      float* __restrict__ lse, int T, float scale) {
  extern __shared__ float sm[];
  float* sK = sm;
  float* sV = sm + (size_t)T * HD;
  const int h = blockIdx.x, n = blockIdx.y, heads = gridDim.x;
  const long long row0 = (long long)n * T;
  for (int i = threadIdx.x; i < T * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD;
    sK[i] = b2f(k[(row0 + t) * ldk + h * HD + c]);
    sV[i] = b2f(v[(row0 + t) * ldv + h * HD + c]);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= T) return;
  float qv[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) qv[c] = b2f(q[(row0 + t) * ldq + h * HD + c]) * scale;
  float m = -INFINITY;
  for (int j = 0; j < T; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(qv[c], sK[j * HD + c], s);
    m = fmaxf(m, s);
  }
  float l = 0.f, acc[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) acc[c] = 0.f;
  for (int j = 0; j < T; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(qv[c], sK[j * HD + c], s);
    const float p = __expf(s - m);
    l += p;
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = fmaf(p, sV[j * HD + c], acc[c]);
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < HD; ++c) o[(row0 + t) * ldo + h * HD + c] = __float2bfloat16(acc[c] * inv);
  lse[((long long)n * heads + h) * T + t] = m + __logf(l);
}

// backward: smem Q | K | V | dO as fp32 [T][HD], then lse[T], Dv[T].  Phase A (thread = query): D, dq.  Phase B
// (thread = key): dk, dv.
template <int HD>
__global__ void __launch_bounds__(128)
attn_small_bwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                      const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ o, int ldo,
                      const __nv_bfloat16* __restrict__ d_o, int lddo, const float* __restrict__ lse,
                      __nv_bfloat16* __restrict__ dq, int lddq, __nv_bfloat16* __restrict__ dk, int lddk,
                      __nv_bfloat16* __restrict__ dv, int lddv, int T, float scale) {
  extern __shared__ float sm[];
  float* sQ = sm;
  float* sK = sQ + (size_t)T * HD;
  float* sV = sK + (size_t)T * HD;
  float* sD = sV + (size_t)T * HD;   // dO
  float* sL = sD + (size_t)T * HD;
  float* sDv = sL + T;
  const int h = blockIdx.x, n = blockIdx.y, heads = gridDim.x;
  const long long row0 = (long long)n * T;
  for (int i = threadIdx.x; i < T * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD;
    sQ[i] = b2f(q[(row0 + t) * ldq + h * HD + c]);
    sK[i] = b2f(k[(row0 + t) * ldk + h * HD + c]);
    sV[i] = b2f(v[(row0 + t) * ldv + h * HD + c]);
    sD[i] = b2f(d_o[(row0 + t) * lddo + h * HD + c]);
  }
  const int t = threadIdx.x;
  if (t < T) {
    sL[t] = lse[((long long)n * heads + h) * T + t];
    float dsum = 0.f;
    for (int c = 0; c < HD; ++c)
      dsum = fmaf(b2f(d_o[(row0 + t) * lddo + h * HD + c]), b2f(o[(row0 + t) * ldo + h * HD + c]), dsum);
    sDv[t] = dsum;
  }
  __syncthreads();
  if (t < T) {
    // phase A: dq_t = scale * sum_j dS_tj k_j
    float acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    const float lt = sL[t], dt = sDv[t];
    for (int j = 0; j < T; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        s = fmaf(sQ[t * HD + c], sK[j * HD + c], s);
        dp = fmaf(sD[t * HD + c], sV[j * HD + c], dp);
      }
      const float ds = __expf(s * scale - lt) * (dp - dt);
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] = fmaf(ds, sK[j * HD + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[(row0 + t) * lddq + h * HD + c] = __float2bfloat16(acc[c] * scale);
    // phase B: key t
    float ak[HD], av[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) ak[c] = av[c] = 0.f;
    for (int i = 0; i < T; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        s = fmaf(sQ[i * HD + c], sK[t * HD + c], s);
        dp = fmaf(sD[i * HD + c], sV[t * HD + c], dp);
      }
      const float p = __expf(s * scale - sL[i]);
      const float ds = p * (dp - sDv[i]);
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        av[c] = fmaf(p, sD[i * HD + c], av[c]);
        ak[c] = fmaf(ds, sQ[i * HD + c], ak[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      dk[(row0 + t) * lddk + h * HD + c] = __float2bfloat16(ak[c] * scale);
      dv[(row0 + t) * lddv + h * HD + c] = __float2bfloat16(av[c]);
    }
  }
}

// ---- SwiGLU --------------------------------------------------------------------------------------------------------------
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                  long long rows, int H) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * H) return;
  const long long r = idx / H;
  const int c = (int)(idx % H);
  const float a = b2f(x[r * ldx + c]), b = b2f(x[r * ldx + H + c]);
  y[r * ldy + c] = __float2bfloat16(act_f<JG_ACT_SILU>(a) * b);
}
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy, int lddy,
                                  __nv_bfloat16* __restrict__ dx, int lddx, long long rows, int H) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * H) return;
  const long long r = idx / H;
  const int c = (int)(idx % H);
  const float a = b2f(x[r * ldx + c]), b = b2f(x[r * ldx + H + c]), d = b2f(dy[r * lddy + c]);
  dx[r * lddx + c] = __float2bfloat16(d * b * act_grad<JG_ACT_SILU>(a));
  dx[r * lddx + H + c] = __float2bfloat16(d * act_f<JG_ACT_SILU>(a));
}

// ---- gated residual ------------------------------------------------------------------------------------------------------
__global__ void gated_residual_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ y,
                                          int ldy, const float* __restrict__ gate, int ldm, __nv_bfloat16* __restrict__ out,
                                          int ldo, long long rows, int C, int T) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const long long r = idx / C;
  const int c = (int)(idx % C);
  out[r * ldo + c] = __float2bfloat16(fmaf(gate[(r / T) * ldm + c], b2f(y[r * ldy + c]), b2f(x[r * ldx + c])));
}
// dy_branch = gate * d (elementwise);  dgate[n][c] = sum_t d * y (thread = (n, c))
__global__ void gated_residual_bwd_kernel(const __nv_bfloat16* __restrict__ d, int ldd, const __nv_bfloat16* __restrict__ y,
                                          int ldy, const float* __restrict__ gate, int ldm, __nv_bfloat16* __restrict__ dy,
                                          int lddy, float* __restrict__ dgate, int lddm, int N, int C, int T) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (c >= C) return;
  const float g = gate[(long long)n * ldm + c];
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    const long long r = (long long)n * T + t;
    const float dv = b2f(d[r * ldd + c]);
    acc = fmaf(dv, b2f(y[r * ldy + c]), acc);
    dy[r * lddy + c] = __float2bfloat16(g * dv);
  }
  dgate[(long long)n * lddm + c] = acc;
}

static unsigned blocks_for(long long total, int block) { return (unsigned)((total + block - 1) / block); }

}  // namespace
}  // namespace jg

using namespace jg;

extern "C" int jg_rmsnorm_mod_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int C, int T, float eps,
                                  const float* w, const float* shift, const float* scale, int ldm, float* rstd,
                                  jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && y && w && rstd && rows > 0 && C > 0 && T > 0 && rows % T == 0, JG_ERR_INVALID, "rmsnorm_mod_fwd: bad args");
  JG_CHECK((shift == nullptr) == (scale == nullptr), JG_ERR_INVALID, "rmsnorm_mod_fwd: shift and scale go together");
  rmsnorm_mod_fwd_kernel<<<blocks_for(rows * 32, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, rows, C, T, eps, w, shift, scale, ldm,
      rstd);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_rmsnorm_mod_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int C,
                                  int T, const float* w, const float* scale, int ldm, const float* rstd, float* dw,
                                  float* dshift, float* dscale, int lddm, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && dy && dx && w && rstd && dw && rows > 0 && rows % T == 0, JG_ERR_INVALID, "rmsnorm_mod_bwd: bad args");
  JG_CHECK((dshift == nullptr) == (dscale == nullptr) && (scale != nullptr || dshift == nullptr), JG_ERR_INVALID,
           "rmsnorm_mod_bwd: modulation gradients need the modulation");
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* db = static_cast<const __nv_bfloat16*>(dy);
  rmsnorm_mod_bwd_dx_kernel<<<blocks_for(rows * 32, 256), 256, 0, stream>>>(xb, ldx, db, lddy,
                                                                           static_cast<__nv_bfloat16*>(dx), lddx, rows, C, T,
                                                                           w, scale, ldm, rstd);
  JG_LAUNCH_CHECK();
  JG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C, stream));
  const int N = (int)(rows / T);
  rmsnorm_mod_bwd_par_kernel<<<dim3((C + 127) / 128, N), 128, 0, stream>>>(xb, ldx, db, lddy, N, C, T, w, scale, ldm, rstd,
                                                                         dw, dshift, dscale, lddm);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

#define JG_HD_DISPATCH(hd, ...)                                                   \
  switch (hd) {                                                                   \
    case 16: { constexpr int HD = 16; __VA_ARGS__; break; }                       \
    case 32: { constexpr int HD = 32; __VA_ARGS__; break; }                       \
    case 64: { constexpr int HD = 64; __VA_ARGS__; break; }                       \
    default: JG_CHECK(false, JG_ERR_UNSUPPORTED, "head dim %d (16, 32, 64)", hd); \
  }

extern "C" int jg_qknorm_rope_fwd(const void* qkv, int ldq, void* out, int ldo, int64_t rows, int T, int heads, int hd,
                                  float eps, const float* wq, const float* wk, const float* cosb, const float* sinb,
                                  float* rstd, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(qkv && out && wq && wk && cosb && sinb && rstd && rows > 0 && rows % T == 0, JG_ERR_INVALID,
           "qknorm_rope_fwd: bad args");
  const long long total = rows * heads * 2;
  JG_HD_DISPATCH(hd, (qknorm_rope_fwd_kernel<HD><<<blocks_for(total, 128), 128, 0, stream>>>(
                         static_cast<const __nv_bfloat16*>(qkv), ldq, static_cast<__nv_bfloat16*>(out), ldo, rows, T, heads,
                         eps, wq, wk, cosb, sinb, rstd)));
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_qknorm_rope_bwd(const void* qkv, int ldq, const void* dout, int lddo, void* dqkv, int lddq, int64_t rows,
                                  int T, int heads, int hd, const float* wq, const float* wk, const float* cosb,
                                  const float* sinb, const float* rstd, float* dwq, float* dwk, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(qkv && dout && dqkv && rstd && dwq && dwk && rows > 0 && rows % T == 0, JG_ERR_INVALID,
           "qknorm_rope_bwd: bad args");
  JG_CUDA(cudaMemsetAsync(dwq, 0, sizeof(float) * hd, stream));
  JG_CUDA(cudaMemsetAsync(dwk, 0, sizeof(float) * hd, stream));
  const long long total = rows * heads * 2;
  JG_HD_DISPATCH(hd, (qknorm_rope_bwd_kernel<HD><<<blocks_for(total, 128), 128, 0, stream>>>(
                         static_cast<const __nv_bfloat16*>(qkv), ldq, static_cast<const __nv_bfloat16*>(dout), lddo,
                         static_cast<__nv_bfloat16*>(dqkv), lddq, rows, T, heads, wq, wk, cosb, sinb, rstd, dwq, dwk)));
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_attn_small_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                                 float* lse, int N, int T, int heads, int hd, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(q && k && v && o && lse && N > 0 && T > 0 && T <= 128 && heads > 0, JG_ERR_INVALID, "attn_small_fwd: bad args");
  const float scale = 1.f / sqrtf((float)hd);
  const size_t smem = (size_t)2 * T * hd * sizeof(float);
  JG_HD_DISPATCH(hd, {
    if (smem > 48 * 1024)
      JG_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_small_fwd_kernel<HD><<<dim3(heads, N), 128, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
        static_cast<const __nv_bfloat16*>(v), ldv, static_cast<__nv_bfloat16*>(o), ldo, lse, T, scale);
  });
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_attn_small_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o,
                                 int ldo, const void* d_o, int lddo, const float* lse, void* dq, int lddq, void* dk, int lddk,
                                 void* dv, int lddv, int N, int T, int heads, int hd, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(q && k && v && o && d_o && lse && dq && dk && dv && N > 0 && T > 0 && T <= 128, JG_ERR_INVALID,
           "attn_small_bwd: bad args");
  const float scale = 1.f / sqrtf((float)hd);
  const size_t smem = ((size_t)4 * T * hd + 2 * T) * sizeof(float);
  JG_HD_DISPATCH(hd, {
    if (smem > 48 * 1024)
      JG_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_small_bwd_kernel<HD><<<dim3(heads, N), 128, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
        static_cast<const __nv_bfloat16*>(v), ldv, static_cast<const __nv_bfloat16*>(o), ldo,
        static_cast<const __nv_bfloat16*>(d_o), lddo, lse, static_cast<__nv_bfloat16*>(dq), lddq,
        static_cast<__nv_bfloat16*>(dk), lddk, static_cast<__nv_bfloat16*>(dv), lddv, T, scale);
  });
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_swiglu_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int H, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && y && rows > 0 && H > 0 && ldx >= 2 * H && ldy >= H, JG_ERR_INVALID, "swiglu_fwd: bad args");
  swiglu_fwd_kernel<<<blocks_for(rows * H, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                                  static_cast<__nv_bfloat16*>(y), ldy, rows, H);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_swiglu_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int H,
                             jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && dy && dx && rows > 0 && H > 0 && ldx >= 2 * H && lddx >= 2 * H && lddy >= H, JG_ERR_INVALID,
           "swiglu_bwd: bad args");
  swiglu_bwd_kernel<<<blocks_for(rows * H, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                                  static_cast<const __nv_bfloat16*>(dy), lddy,
                                                                  static_cast<__nv_bfloat16*>(dx), lddx, rows, H);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_gated_residual_fwd(const void* x, int ldx, const void* y, int ldy, const float* gate, int ldm, void* out,
                                     int ldo, int64_t rows, int C, int T, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && y && gate && out && rows > 0 && rows % T == 0, JG_ERR_INVALID, "gated_residual_fwd: bad args");
  gated_residual_fwd_kernel<<<blocks_for(rows * C, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(y), ldy, gate, ldm,
      static_cast<__nv_bfloat16*>(out), ldo, rows, C, T);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_gated_residual_bwd(const void* d, int ldd, const void* y, int ldy, const float* gate, int ldm, void* dy,
                                     int lddy, float* dgate, int lddm, int64_t rows, int C, int T, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(d && y && gate && dy && dgate && rows > 0 && rows % T == 0, JG_ERR_INVALID, "gated_residual_bwd: bad args");
  const int N = (int)(rows / T);
  gated_residual_bwd_kernel<<<dim3((C + 127) / 128, N), 128, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(d), ldd, static_cast<const __nv_bfloat16*>(y), ldy, gate, ldm,
      static_cast<__nv_bfloat16*>(dy), lddy, dgate, lddm, N, C, T);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
