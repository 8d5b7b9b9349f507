// Spatial self-attention of the UNet AttentionBlock, flash style (the T x T logits are never
// materialised), forward + backward, bf16 in / fp32 softmax.
//
//   QKVAttentionLegacy.forward   unet_generator_attn.py:331-347
//     qkv [N, heads*3*ch, T] is split per head into (q | k | v) of ch channels each ("split heads
//     before split qkv"); weight = softmax_fp32((q*s)^T (k*s)), s = ch^-1/4; a = weight @ v.
//
// Layout here: qkv is NHWC [N][T][3C] bf16 (row stride ldqkv), head h owns channels
// [h*3*ch, (h+1)*3*ch) = q|k|v; output a is [N][T][C] with head h at channels [h*ch, (h+1)*ch).
//
// This op is ~1% of the step's FLOPs (one 1024-token block in the Palette UNet); it runs on
// mma.sync.m16n8k16 (HMMA) with 64x64 tiles rather than tcgen05 — see DESIGN.md.
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

int launch_attn_fwd_tc(const void* qkv, int ldqkv, void* out, int ldo, float* lse, int N, int T, int heads, int ch,
                       int hstride, int koff, int voff, float scale_log2, cudaStream_t stream);

int launch_attn_bwd_tc(const void* qkv, int ldqkv, const void* d_out, int lddo, const float* lse, const float* D,
                       void* dqkv, int lddqkv, int N, int T, int heads, int ch, int hstride, int koff, int voff,
                       float scale_log2, float scale, cudaStream_t stream);

constexpr int kAttnThreads = 128;
constexpr int kBM = 64;  // rows per CTA (4 warps x 16)
constexpr int kBN = 64;  // columns per inner block

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(p)));
}

// smem tile: 64 rows x HD bf16, row pitch HD + 8 (16-byte pad: conflict-free ldmatrix).
template <int HD>
struct Tile {
  static constexpr int PITCH = HD + 8;
  __nv_bfloat16 d[64 * PITCH];
  __device__ __forceinline__ __nv_bfloat16* row(int r) { return d + r * PITCH; }
  // cooperative load of 64 rows (row stride ld elements) by 128 threads
  __device__ __forceinline__ void load(const __nv_bfloat16* g, int ld) {
    constexpr int VPR = HD / 8;
    for (int i = threadIdx.x; i < 64 * VPR; i += kAttnThreads) {
      const int r = i / VPR, v = i % VPR;
      *reinterpret_cast<uint4*>(row(r) + v * 8) = *reinterpret_cast<const uint4*>(g + (size_t)r * ld + v * 8);
    }
  }
  // A fragment (16 rows starting at r0, k columns [k0, k0+16))
  __device__ __forceinline__ void a_frag(uint32_t (&a)[4], int r0, int k0) {
    const int l = threadIdx.x & 31;
    ldsm_x4(a, row(r0 + (l & 15)) + k0 + (l >> 4) * 8);
  }
  // B fragment when the tile is stored [n][k] (k contiguous): n rows [n0, n0+8), k [k0, k0+16)
  __device__ __forceinline__ void b_frag_nk(uint32_t& b0, uint32_t& b1, int n0, int k0) {
    const int l = threadIdx.x & 31;
    ldsm_x2(b0, b1, row(n0 + (l & 7)) + k0 + ((l >> 3) & 1) * 8);
  }
  // B fragment when the tile is stored [k][n] (n contiguous): k rows [k0, k0+16), n [n0, n0+8)
  __device__ __forceinline__ void b_frag_kn(uint32_t& b0, uint32_t& b1, int k0, int n0) {
    const int l = threadIdx.x & 31;
    ldsm_x2_trans(b0, b1, row(k0 + (l & 15)) + n0);
  }
};

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ldqkv, __nv_bfloat16* __restrict__ out, int ldo,
                float* __restrict__ lse, int T, int heads, float scale_log2, int hstride, int koff, int voff) {
  __shared__ __align__(16) Tile<HD> sQ, sK, sV;
  const int bh = blockIdx.y;
  const int n = bh / heads, h = bh % heads;
  const int q0 = blockIdx.x * kBM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const __nv_bfloat16* base = qkv + (size_t)n * T * ldqkv + h * hstride;

  sQ.load(base + (size_t)q0 * ldqkv, ldqkv);
  __syncthreads();
  uint32_t qa[HD / 16][4];
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) sQ.a_frag(qa[kk], warp * 16, kk * 16);

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < T; k0 += kBN) {
    __syncthreads();
    sK.load(base + koff + (size_t)k0 * ldqkv, ldqkv);
    sV.load(base + voff + (size_t)k0 * ldqkv, ldqkv);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t b0, b1;
        sK.b_frag_nk(b0, b1, nt * 8, kk * 16);
        mma_bf16_16816(s[nt], qa[kk], b0, b1);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = quad_max(mx0) * scale_log2;
    mx1 = quad_max(mx1) * scale_log2;
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * scale_log2 - mn0);
      s[nt][1] = exp2f(s[nt][1] * scale_log2 - mn0);
      s[nt][2] = exp2f(s[nt][2] * scale_log2 - mn1);
      s[nt][3] = exp2f(s[nt][3] * scale_log2 - mn1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        uint32_t b0, b1;
        sV.b_frag_kn(b0, b1, kk * 16, i * 8);
        mma_bf16_16816(o[i], pa, b0, b1);
      }
    }
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __nv_bfloat16* ob = out + (size_t)n * T * ldo + h * HD;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * ldo + i * 8 + t4 * 2) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
    *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * ldo + i * 8 + t4 * 2) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
  }
  if (t4 == 0) {
    lse[(size_t)bh * T + r0] = m0 + log2f(l0);
    lse[(size_t)bh * T + r1] = m1 + log2f(l1);
  }
}

// D[bh][t] = sum_c dO[t][c] * O[t][c]
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, int ldo, const __nv_bfloat16* __restrict__ d_o,
                                     int lddo, float* __restrict__ D, int T, int heads, int HD, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // over N*heads*T
  if (i >= total) return;
  const int t = (int)(i % T);
  const int bh = (int)(i / T);
  const int n = bh / heads, h = bh % heads;
  const __nv_bfloat16* po = o + ((size_t)n * T + t) * ldo + h * HD;
  const __nv_bfloat16* pd = d_o + ((size_t)n * T + t) * lddo + h * HD;
  float acc = 0.f;
  for (int c = 0; c < HD; c += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(po + c);
    const uint4 b = *reinterpret_cast<const uint4*>(pd + c);
    const uint32_t* pa = &a.x;
    const uint32_t* pb = &b.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(pa[j]), y = unpack_bf16x2(pb[j]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  D[i] = acc;
}

// ------------------------------------------------------------------------------------------------
// backward, dQ: CTA = 64 queries, loop over key blocks
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, int ldqkv, const __nv_bfloat16* __restrict__ d_o, int lddo,
                   const float* __restrict__ lse, const float* __restrict__ D, __nv_bfloat16* __restrict__ dqkv,
                   int lddqkv, int T, int heads, float scale_log2, float scale, int hstride, int koff, int voff) {
  __shared__ __align__(16) Tile<HD> sQ, sK, sV;  // sQ is reused for dO
  const int bh = blockIdx.y;
  const int n = bh / heads, h = bh % heads;
  const int q0 = blockIdx.x * kBM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const __nv_bfloat16* base = qkv + (size_t)n * T * ldqkv + h * hstride;
  const __nv_bfloat16* dob = d_o + (size_t)n * T * lddo + h * HD;

  uint32_t qa[HD / 16][4], da[HD / 16][4];
  sQ.load(base + (size_t)q0 * ldqkv, ldqkv);
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) sQ.a_frag(qa[kk], warp * 16, kk * 16);
  __syncthreads();
  sQ.load(dob + (size_t)q0 * lddo, lddo);
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) sQ.a_frag(da[kk], warp * 16, kk * 16);

  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float L0 = lse[(size_t)bh * T + r0], L1 = lse[(size_t)bh * T + r1];
  const float D0 = D[(size_t)bh * T + r0], D1 = D[(size_t)bh * T + r1];

  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

  for (int k0 = 0; k0 < T; k0 += kBN) {
    __syncthreads();
    sK.load(base + koff + (size_t)k0 * ldqkv, ldqkv);
    sV.load(base + voff + (size_t)k0 * ldqkv, ldqkv);
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t b0, b1;
        sK.b_frag_nk(b0, b1, nt * 8, kk * 16);
        mma_bf16_16816(s[nt], qa[kk], b0, b1);
        sV.b_frag_nk(b0, b1, nt * 8, kk * 16);
        mma_bf16_16816(dp[nt], da[kk], b0, b1);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] * scale_log2 - L0), p1 = exp2f(s[nt][1] * scale_log2 - L0);
      const float p2 = exp2f(s[nt][2] * scale_log2 - L1), p3 = exp2f(s[nt][3] * scale_log2 - L1);
      s[nt][0] = p0 * (dp[nt][0] - D0);
      s[nt][1] = p1 * (dp[nt][1] - D0);
      s[nt][2] = p2 * (dp[nt][2] - D1);
      s[nt][3] = p3 * (dp[nt][3] - D1);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        uint32_t b0, b1;
        sK.b_frag_kn(b0, b1, kk * 16, i * 8);
        mma_bf16_16816(dq[i], pa, b0, b1);
      }
    }
  }
  __nv_bfloat16* ob = dqkv + (size_t)n * T * lddqkv + h * hstride;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * lddqkv + i * 8 + t4 * 2) =
        pack_bf16x2(dq[i][0] * scale, dq[i][1] * scale);
    *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * lddqkv + i * 8 + t4 * 2) =
        pack_bf16x2(dq[i][2] * scale, dq[i][3] * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, dK and dV: CTA = 64 keys, loop over query blocks (transposed problem: rows = keys)
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, int ldqkv, const __nv_bfloat16* __restrict__ d_o, int lddo,
                    const float* __restrict__ lse, const float* __restrict__ D, __nv_bfloat16* __restrict__ dqkv,
                    int lddqkv, int T, int heads, float scale_log2, float scale, int hstride, int koff, int voff) {
  __shared__ __align__(16) Tile<HD> sK, sQ, sDO;  // sK is reused for V while building fragments
  __shared__ float sL[kBN], sD[kBN];
  const int bh = blockIdx.y;
  const int n = bh / heads, h = bh % heads;
  const int k0 = blockIdx.x * kBM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const __nv_bfloat16* base = qkv + (size_t)n * T * ldqkv + h * hstride;
  const __nv_bfloat16* dob = d_o + (size_t)n * T * lddo + h * HD;

  uint32_t ka[HD / 16][4], va[HD / 16][4];
  sK.load(base + koff + (size_t)k0 * ldqkv, ldqkv);
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) sK.a_frag(ka[kk], warp * 16, kk * 16);
  __syncthreads();
  sK.load(base + voff + (size_t)k0 * ldqkv, ldqkv);
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) sK.a_frag(va[kk], warp * 16, kk * 16);

  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }

  for (int q0 = 0; q0 < T; q0 += kBN) {
    __syncthreads();
    sQ.load(base + (size_t)q0 * ldqkv, ldqkv);
    sDO.load(dob + (size_t)q0 * lddo, lddo);
    if (threadIdx.x < kBN) {
      sL[threadIdx.x] = lse[(size_t)bh * T + q0 + threadIdx.x];
      sD[threadIdx.x] = D[(size_t)bh * T + q0 + threadIdx.x];
    }
    __syncthreads();
    // S^T[key][query] and dP^T[key][query]
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t b0, b1;
        sQ.b_frag_nk(b0, b1, nt * 8, kk * 16);
        mma_bf16_16816(s[nt], ka[kk], b0, b1);
        sDO.b_frag_nk(b0, b1, nt * 8, kk * 16);
        mma_bf16_16816(dp[nt], va[kk], b0, b1);
      }
    }
    // columns are queries: thread holds columns nt*8 + 2*t4 + {0,1}
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + t4 * 2;
      const float La = sL[c], Lb = sL[c + 1], Da = sD[c], Db = sD[c + 1];
      const float p0 = exp2f(s[nt][0] * scale_log2 - La), p1 = exp2f(s[nt][1] * scale_log2 - Lb);
      const float p2 = exp2f(s[nt][2] * scale_log2 - La), p3 = exp2f(s[nt][3] * scale_log2 - Lb);
      s[nt][0] = p0; s[nt][1] = p1; s[nt][2] = p2; s[nt][3] = p3;
      dp[nt][0] = p0 * (dp[nt][0] - Da);
      dp[nt][1] = p1 * (dp[nt][1] - Db);
      dp[nt][2] = p2 * (dp[nt][2] - Da);
      dp[nt][3] = p3 * (dp[nt][3] - Db);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4], sa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      sa[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
      sa[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
      sa[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      sa[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        uint32_t b0, b1;
        sDO.b_frag_kn(b0, b1, kk * 16, i * 8);  // dV += P^T dO
        mma_bf16_16816(dv[i], pa, b0, b1);
        sQ.b_frag_kn(b0, b1, kk * 16, i * 8);  // dK += dS^T Q
        mma_bf16_16816(dk[i], sa, b0, b1);
      }
    }
  }
  const int r0 = k0 + warp * 16 + g, r1 = r0 + 8;
  __nv_bfloat16* ob = dqkv + (size_t)n * T * lddqkv + h * hstride;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(ob + koff + (size_t)r0 * lddqkv + i * 8 + t4 * 2) =
        pack_bf16x2(dk[i][0] * scale, dk[i][1] * scale);
    *reinterpret_cast<uint32_t*>(ob + koff + (size_t)r1 * lddqkv + i * 8 + t4 * 2) =
        pack_bf16x2(dk[i][2] * scale, dk[i][3] * scale);
    *reinterpret_cast<uint32_t*>(ob + voff + (size_t)r0 * lddqkv + i * 8 + t4 * 2) = pack_bf16x2(dv[i][0], dv[i][1]);
    *reinterpret_cast<uint32_t*>(ob + voff + (size_t)r1 * lddqkv + i * 8 + t4 * 2) = pack_bf16x2(dv[i][2], dv[i][3]);
  }
}

static int check_attn(int N, int T, int heads, int ch, int ldqkv) {
  JG_CHECK(N > 0 && T > 0 && heads > 0, JG_ERR_INVALID, "attention: bad dims");
  JG_CHECK(ch == 16 || ch == 32 || ch == 64, JG_ERR_INVALID, "attention: head channels %d unsupported (16/32/64)", ch);
  JG_CHECK(T % 64 == 0, JG_ERR_INVALID, "attention: T=%d must be a multiple of 64", T);
  JG_CHECK(ldqkv % 8 == 0 && ldqkv >= 3 * heads * ch, JG_ERR_INVALID, "attention: bad ldqkv %d", ldqkv);
  return JG_OK;
}

}  // namespace jg

using namespace jg;

// layout 0: per-head (q|k|v) interleave = QKVAttentionLegacy; layout 1: (q | k | v), each heads*ch wide = QKVAttention
#define JG_ATTN_LAYOUT(layout, heads, ch)                         \
  const int hstride = (layout) == 0 ? 3 * (ch) : (ch);            \
  const int koff = (layout) == 0 ? (ch) : (heads) * (ch);         \
  const int voff = (layout) == 0 ? 2 * (ch) : 2 * (heads) * (ch);

extern "C" int jg_attn_fwd(const void* qkv, int ldqkv, void* out, int ldo, float* lse, int N, int T, int heads, int ch,
                           int layout, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_attn(N, T, heads, ch, ldqkv);
  if (rc) return rc;
  JG_CHECK(qkv && out && lse && ldo % 8 == 0 && ldo >= heads * ch, JG_ERR_INVALID, "attn_fwd: bad args");
  JG_CHECK(layout == 0 || layout == 1, JG_ERR_INVALID, "attn_fwd: layout %d", layout);
  JG_ATTN_LAYOUT(layout, heads, ch)
  const float scale = 1.f / sqrtf((float)ch);  // (ch^-1/4)^2
  const float scale_log2 = scale * 1.4426950408889634f;
  // tcgen05 forward (attention_tc.cu) where the shape qualifies (ch 32 / 64, T a multiple of 256); JG_ATTN_TC=0: off
  static const bool use_tc = getenv("JG_ATTN_TC") == nullptr || atoi(getenv("JG_ATTN_TC")) != 0;
  if (use_tc) {
    rc = launch_attn_fwd_tc(qkv, ldqkv, out, ldo, lse, N, T, heads, ch, hstride, koff, voff, scale_log2, stream);
    if (rc != JG_ERR_UNSUPPORTED) return rc;
  }
  dim3 grid(T / kBM, N * heads);
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (ch == 16) attn_fwd_kernel<16><<<grid, kAttnThreads, 0, stream>>>(q, ldqkv, o, ldo, lse, T, heads, scale_log2, hstride, koff, voff);
  else if (ch == 32) attn_fwd_kernel<32><<<grid, kAttnThreads, 0, stream>>>(q, ldqkv, o, ldo, lse, T, heads, scale_log2, hstride, koff, voff);
  else attn_fwd_kernel<64><<<grid, kAttnThreads, 0, stream>>>(q, ldqkv, o, ldo, lse, T, heads, scale_log2, hstride, koff, voff);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_attn_bwd(const void* qkv, int ldqkv, const void* out, int ldo, const void* d_out, int lddo,
                           const float* lse, void* dqkv, int lddqkv, float* ws /* N*heads*T floats */, int N, int T,
                           int heads, int ch, int layout, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_attn(N, T, heads, ch, ldqkv);
  if (rc) return rc;
  JG_CHECK(qkv && out && d_out && lse && dqkv && ws, JG_ERR_INVALID, "attn_bwd: null pointer");
  JG_CHECK(layout == 0 || layout == 1, JG_ERR_INVALID, "attn_bwd: layout %d", layout);
  JG_ATTN_LAYOUT(layout, heads, ch)
  JG_CHECK(ldo % 8 == 0 && lddo % 8 == 0 && lddqkv % 8 == 0 && lddqkv >= 3 * heads * ch, JG_ERR_INVALID,
           "attn_bwd: bad ld");
  const float scale = 1.f / sqrtf((float)ch);
  const float scale_log2 = scale * 1.4426950408889634f;
  const long long total = (long long)N * heads * T;
  attn_bwd_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(out), ldo, static_cast<const __nv_bfloat16*>(d_out), lddo, ws, T, heads, ch,
      total);
  JG_LAUNCH_CHECK();
  static const bool use_tc = getenv("JG_ATTN_TC") == nullptr || atoi(getenv("JG_ATTN_TC")) != 0;
  if (use_tc) {
    rc = launch_attn_bwd_tc(qkv, ldqkv, d_out, lddo, lse, ws, dqkv, lddqkv, N, T, heads, ch, hstride, koff, voff,
                            scale_log2, scale, stream);
    if (rc != JG_ERR_UNSUPPORTED) return rc;
  }
  dim3 grid(T / kBM, N * heads);
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(d_out);
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
#define JG_ATTN_BWD(HD)                                                                                         \
  attn_bwd_dq_kernel<HD><<<grid, kAttnThreads, 0, stream>>>(q, ldqkv, d, lddo, lse, ws, dq, lddqkv, T, heads,    \
                                                            scale_log2, scale, hstride, koff, voff);             \
  attn_bwd_dkv_kernel<HD><<<grid, kAttnThreads, 0, stream>>>(q, ldqkv, d, lddo, lse, ws, dq, lddqkv, T, heads,   \
                                                             scale_log2, scale, hstride, koff, voff);
  if (ch == 16) { JG_ATTN_BWD(16) } else if (ch == 32) { JG_ATTN_BWD(32) } else { JG_ATTN_BWD(64) }
#undef JG_ATTN_BWD
  JG_LAUNCH_CHECK();
  return JG_OK;
}
