// Shared pieces of the implicit-GEMM convolution kernels (conv_igemm.cu, conv_halo.cu).
#pragma once
#include "act.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kWgradThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16

struct ConvFwdParams {
  int N, Ho, Wo;
  int Cout;
  int RS, S, pad, stride;
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles;    // ceil(Cout / BLOCK_N)
  int kc_blocks;  // ceil(Cin / 64)
  int total_tiles;
  int ldy, ldres;
  int act;
  float res_scale;
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  // ---- fused GroupNorm work (jg_conv_epilogue; TMA-store epilogue with one image per tile only) ----
  // stats: per-(image, channel) sum and sum of squares of the STORED (bf16-rounded) output, [N][Cout][2] fp32,
  //        accumulated with red.global.add: the statistics pass of the GroupNorm that consumes y.
  float* stats;
  // gn_sums: this launch is the dgrad of the conv that follows act(a*x+b) (a, b = the GroupNorm's fused per-(n,c)
  //        coefficients): its output is dy.  A[n,c] += sum du, B[n,c] += sum du*x with du = dy*act'(a*x+b),
  //        [N][Cout][2] fp32.  x travels as the `res` operand (res_mode = 1: it is NOT added to the output).
  float* gn_sums;
  const float* gn_ab;  // [N][Cout][2]
  int gn_act;
  int res_mode;        // 0: y += res_scale * res;  1: res is the GroupNorm input x of the gn_sums mode
};

// Column sums over the 32 lanes of a warp of 32 per-lane values each: on return lane l holds sum_lanes v[l].
// Transposing butterfly: at every step a lane keeps one half of its values and hands the other half to its
// partner (31 shuffles for 32 columns instead of 32 x 5).
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ float act_grad_rt(float u, int act) {
  switch (act) {
    case JG_ACT_SILU: return act_grad<JG_ACT_SILU>(u);
    case JG_ACT_RELU: return act_grad<JG_ACT_RELU>(u);
    case JG_ACT_LRELU02: return act_grad<JG_ACT_LRELU02>(u);
    default: return 1.f;
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case JG_ACT_RELU: return v > 0.f ? v : 0.f;
    case JG_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
    case JG_ACT_TANH: return tanhf(v);
    case JG_ACT_SILU: return v / (1.f + __expf(-v));
    default: return v;
  }
}


// Epilogue of one 128 x BLOCK_N output tile held in TMEM: bias + res_scale*residual + activation -> bf16 ->
// 16-byte global stores.  Called by the 8 epilogue warps: q = TMEM lane quarter of the calling warp, half
// = which of the two warps of that quarter (they interleave 32-column chunks).  The residual chunk (DRAM
// latency) is prefetched one chunk ahead; call conv_epilogue_prefetch before waiting for the accumulator.
struct EpiPrefetch {
  uint4 r[4];
};

// All threads: copy bias[0, Cout) (fp32, Cout a multiple of 8) into shared memory, followed by 64 zeros (the TMA-store
// epilogue processes whole 64-channel slabs without per-group bounds checks).  Call before __syncthreads.
// widest layer on the path: the video UNet's GEGLU projection at the bottleneck (512 * 4 * 2 output features)
constexpr int kMaxCout = 4096;

__device__ __forceinline__ void conv_stage_bias(const ConvFwdParams& p, float* s_bias) {
  if (p.bias)
    for (int i = threadIdx.x; i < p.Cout + 64; i += blockDim.x) s_bias[i] = i < p.Cout ? p.bias[i] : 0.f;
}

template <int BLOCK_N>
__device__ __forceinline__ void conv_epilogue_prefetch(const ConvFwdParams& p, EpiPrefetch& pf, int half, int n_tile,
                                                       bool valid, size_t pix) {
  const int c = half * 32;
  const int co0 = n_tile * BLOCK_N + c;
  if (p.res && valid && c < BLOCK_N) {
    const __nv_bfloat16* rp = p.res + pix * p.ldres + co0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (co0 + g * 8 < p.Cout) pf.r[g] = *reinterpret_cast<const uint4*>(rp + g * 8);
  }
}

// s_bias: the bias vector staged in shared memory by conv_stage_bias (nullptr when the conv has no bias).
template <int BLOCK_N>
__device__ __forceinline__ void conv_epilogue_tile(const ConvFwdParams& p, EpiPrefetch& pf, const float* s_bias,
                                                   uint32_t t_acc, int q, int half, int n_tile, bool valid,
                                                   size_t pix) {
  const uint32_t t_row = t_acc + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
  for (int c = half * 32; c < BLOCK_N; c += 64) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c, v);
    const int co0 = n_tile * BLOCK_N + c;
    const bool do_store = valid && co0 < p.Cout;
    uint4 rcur[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) rcur[g] = pf.r[g];
    // prefetch the next chunk's residual while this one is processed
    if (p.res && valid && c + 64 < BLOCK_N) {
      const __nv_bfloat16* rn = p.res + pix * p.ldres + co0 + 64;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (co0 + 64 + g * 8 < p.Cout) pf.r[g] = *reinterpret_cast<const uint4*>(rn + g * 8);
    }
    tmem_ld_wait();
    if (do_store) {
      __nv_bfloat16* yp = p.y + pix * p.ldy + co0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {  // 8 channels per 16-byte store
        if (co0 + g * 8 < p.Cout) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
          if (s_bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + co0 + g * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + co0 + g * 8 + 4);
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (p.res) {
            const uint4 rv = rcur[g];
            const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y);
            const float2 r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
            f[0] += p.res_scale * r0.x; f[1] += p.res_scale * r0.y;
            f[2] += p.res_scale * r1.x; f[3] += p.res_scale * r1.y;
            f[4] += p.res_scale * r2.x; f[5] += p.res_scale * r2.y;
            f[6] += p.res_scale * r3.x; f[7] += p.res_scale * r3.y;
          }
          if (p.act != JG_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], p.act);
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(yp + g * 8) = o;
        }
      }
    }
  }
}

// Epilogue through shared memory + TMA stores.  The direct epilogue above has every lane write 64 contiguous bytes of
// its own pixel row: a warp-level 16-byte store touches 32 different 128-byte lines, and ncu showed the L1/LSU store
// path 65% busy on the 64-channel 256^2 layers (the tile's MMAs were waiting for the epilogue).  Here the 8 epilogue
// warps write each 64-channel slab of the tile (128 pixel rows x 128 bytes, 128B-swizzled exactly as TMA expects, which
// also makes the 16-byte shared stores conflict-free) into one of two 16 KB staging buffers and one thread hands the
// slab to the TMA engine.  One named barrier per slab: before it, the issuing thread has waited for the previous
// slab's store to finish reading the other buffer.
constexpr int kStageBytes = 128 * 128;
constexpr int kEpiThreads = 256;

template <int BLOCK_N>
__device__ __forceinline__ void conv_epilogue_tile_tma(const ConvFwdParams& p, EpiPrefetch& pf, const float* s_bias,
                                                       uint32_t t_acc, int q, int half, int n_tile, bool valid,
                                                       size_t pix, uint8_t* stage, int& stage_idx,
                                                       const CUtensorMap* tmY, int c1, int c2, int c3, bool issuer,
                                                       const uint8_t* res_tile = nullptr, int img = 0,
                                                       const float* s_ab = nullptr) {
  static_assert(BLOCK_N % 64 == 0, "TMA-store epilogue works on 64-channel slabs");
  const uint32_t t_row = t_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int row = q * 32 + (threadIdx.x & 31);
  const int swz = row & 7;
#pragma unroll 1
  for (int c = half * 32; c < BLOCK_N; c += 64) {
    const int co0 = n_tile * BLOCK_N + c;
    const int slab_co = co0 - half * 32;         // first channel of this 64-channel slab
    if (slab_co >= p.Cout) break;                // uniform over all 8 warps: the rest of the tile is channel padding
    uint32_t v[32];
    tmem_ld_32x32(t_row + c, v);
    // While the TMEM load is in flight: bias (shared memory, zero-padded past Cout) and residual.  Straight-line code:
    // the four 8-channel groups are independent, the two epilogue warps of an SM sub-partition have little else to
    // hide latency with.  Channels >= Cout hold don't-care values that the TMA store clips.
    float4 bias4[8];
    if (s_bias) {
#pragma unroll
      for (int g = 0; g < 8; ++g) bias4[g] = *reinterpret_cast<const float4*>(s_bias + co0 + g * 4);
    }
    uint4 rcur[4];
    const bool has_res = p.res != nullptr;
    if (res_tile) {
      // residual slab staged in shared memory by TMA (same 128B-swizzled layout as the output staging)
#pragma unroll
      for (int g = 0; g < 4; ++g)
        rcur[g] = *reinterpret_cast<const uint4*>(res_tile + row * 128 + (((half * 4 + g) ^ swz) << 4));
    } else if (has_res) {
#pragma unroll
      for (int g = 0; g < 4; ++g) rcur[g] = pf.r[g];
      if (valid && c + 64 < BLOCK_N) {
        const __nv_bfloat16* rn = p.res + pix * p.ldres + co0 + 64;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (co0 + 64 + g * 8 < p.Cout) pf.r[g] = *reinterpret_cast<const uint4*>(rn + g * 8);
      }
    }
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (s_bias) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        f[g * 4 + 0] += bias4[g].x; f[g * 4 + 1] += bias4[g].y;
        f[g * 4 + 2] += bias4[g].z; f[g * 4 + 3] += bias4[g].w;
      }
    }
    if (has_res && p.res_mode == 0 && (res_tile || valid)) {
      const float rs = p.res_scale;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t w4[4] = {rcur[g].x, rcur[g].y, rcur[g].z, rcur[g].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 r = unpack_bf16x2(w4[k]);
          f[g * 8 + 2 * k] = fmaf(rs, r.x, f[g * 8 + 2 * k]);
          f[g * 8 + 2 * k + 1] = fmaf(rs, r.y, f[g * 8 + 2 * k + 1]);
        }
      }
    }
    if (p.act != JG_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
    }
    uint8_t* buf = stage + stage_idx * kStageBytes + row * 128;
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // 8 channels per 16-byte chunk; 128B swizzle: chunk index XOR (row mod 8)
      uint4 o;
      o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
      o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
      o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
      o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
      *reinterpret_cast<uint4*>(buf + (((half * 4 + g) ^ swz) << 4)) = o;
      if (p.stats || p.gn_sums) {  // keep the values that were actually stored (bf16-rounded)
        const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 r = unpack_bf16x2(w4[k]);
          f[g * 8 + 2 * k] = r.x;
          f[g * 8 + 2 * k + 1] = r.y;
        }
      }
    }
    if (p.stats) {
      // statistics of the stored tile for the GroupNorm that reads y next: per-channel sum / sum of squares over the
      // warp's 32 pixel rows, one 8-byte red per lane (lane l <-> channel co0 + l)
      float sq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float r = valid ? f[j] : 0.f;  // rows outside a ragged output contribute nothing
        sq[j] = r * r;
        f[j] = r;
      }
      const float S = warp_colsum32(f);
      const float Q = warp_colsum32(sq);
      const int co = co0 + (threadIdx.x & 31);
      if (co < p.Cout) red_add_v2(p.stats + (static_cast<size_t>(img) * p.Cout + co) * 2, S, Q);
    } else if (p.gn_sums) {  // (never together with stats: f is consumed by the column sums)
      // GroupNorm-backward sums: du = dy * act'(a*x + b), A = sum du, B = sum du*x over the warp's 32 pixel rows
      float dux[32];
      const bool live = res_tile != nullptr || valid;  // x of rows outside a ragged output was never loaded
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t w4[4] = {rcur[g].x, rcur[g].y, rcur[g].z, rcur[g].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 xv = unpack_bf16x2(w4[k]);
          const int j = g * 8 + 2 * k;
          const float4 ab = *reinterpret_cast<const float4*>(s_ab + 2 * (c + j));  // a_j, b_j, a_j+1, b_j+1
          const float du0 = live ? f[j] * act_grad_rt(fmaf(xv.x, ab.x, ab.y), p.gn_act) : 0.f;
          const float du1 = live ? f[j + 1] * act_grad_rt(fmaf(xv.y, ab.z, ab.w), p.gn_act) : 0.f;
          f[j] = du0;
          f[j + 1] = du1;
          dux[j] = live ? du0 * xv.x : 0.f;
          dux[j + 1] = live ? du1 * xv.y : 0.f;
        }
      }
      const float A = warp_colsum32(f);
      const float B = warp_colsum32(dux);
      const int co = co0 + (threadIdx.x & 31);
      if (co < p.Cout) red_add_v2(p.gn_sums + (static_cast<size_t>(img) * p.Cout + co) * 2, A, B);
    }
    fence_proxy_async();                 // generic-proxy smem writes -> visible to the TMA (async proxy)
    if (issuer) bulk_wait_read<0>();     // the previous slab's store has released the other buffer
    bar_sync(1, kEpiThreads);
    if (issuer) {
      tma_store_4d(tmY, stage + stage_idx * kStageBytes, slab_co, c1, c2, c3);
      bulk_commit();
    }
    stage_idx ^= 1;
  }
}

// Launch of the halo-reuse 3x3 kernel (conv_halo.cu); returns JG_ERR_UNSUPPORTED when the shape does not qualify.
// e (may be NULL): fused GroupNorm work; *fused is set when the launched kernel's epilogue did it.
int launch_conv_halo(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                     const float* bias, const void* residual, void* y, cudaStream_t stream, bool* fused);

// Fills the fused-GroupNorm fields of the kernel parameters from the C-ABI structs; returns the operand that travels
// through the residual plumbing (the residual itself, or the GroupNorm input x of the gn_sums mode).
inline const void* conv_apply_epilogue(ConvFwdParams& p, const jg_conv_desc* d, const jg_conv_epilogue* e,
                                       const void* residual) {
  p.stats = nullptr; p.gn_sums = nullptr; p.gn_ab = nullptr; p.gn_act = 0; p.res_mode = 0;
  if (!e) return residual;
  p.stats = e->stats;
  if (e->gn_sums) {
    p.gn_sums = e->gn_sums; p.gn_ab = e->gn_ab; p.gn_act = e->gn_act; p.res_mode = 1;
    p.ldres = e->ldgx;
    return e->gn_x;
  }
  return residual;
}

// Stand-alone forms of the fused reductions (norm.cu).
int launch_chan_stats(const void* x, int ldx, int N, int HW, int C, float* stats, cudaStream_t stream);
int launch_gn_bwd_sums(const void* x, int ldx, const void* dy, int lddy, int N, int HW, int C, const float* ab, int act,
                       float* AB, cudaStream_t stream);

// The (a, b) coefficients of one tile's image and channel block, staged in shared memory for the gn_sums epilogue by
// the 256 epilogue threads (double-buffered by the caller; named barrier 2).
template <int BLOCK_N>
__device__ __forceinline__ void conv_stage_gn_ab(const ConvFwdParams& p, float* dst, int img, int n_tile) {
  const int e = threadIdx.x - 64;
  const int c0 = n_tile * BLOCK_N;
  for (int i = e; i < 2 * BLOCK_N; i += kEpiThreads)
    dst[i] = (c0 + (i >> 1) < p.Cout) ? p.gn_ab[(static_cast<size_t>(img) * p.Cout + c0) * 2 + i] : 0.f;
  bar_sync(2, kEpiThreads);
}

// wgrad with halo reuse (conv_halo.cu): zeroes ws, accumulates, writes OIHW; JG_ERR_UNSUPPORTED if not eligible.
int launch_wgrad_halo(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* ws, float* dw_oihw,
                      float beta, cudaStream_t stream);

}  // namespace jg
