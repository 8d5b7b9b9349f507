// Shared pieces of the implicit-GEMM convolution kernels (conv_igemm.cu, conv_halo.cu).
#pragma once
#include <stdlib.h>

#include "act.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kWgradThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kStageBytes = 128 * 128;  // one 64-channel slab of an output tile (TMA-store epilogue)
constexpr int kEpiThreads = 256;
constexpr int kMaxFusedCout = 1024;     // widest output whose GroupNorm sums fit the shared-memory accumulators

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
  int dbg;             // JG_DBG_EPI bits (timing experiments only): 1 = plain stores instead of red.global.add
};

__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// Tile schedule of the persistent forward kernels.  Default: static round robin (tile = cta, cta + grid, ...): CTAs
// that run together work on neighbouring tiles.  With fused GroupNorm reductions every CTA takes ONE CONTIGUOUS chunk of
// tiles instead: its tiles then stay inside one or two images, so the per-(image, channel) sums can be accumulated in
// shared memory and reach global memory once per image.  (Measured on B200: with round robin, the 148 CTAs of a wave all
// hit the same image's 2*Cout addresses with red.global.add at the same time and the conv slows down by up to 50 %.)
struct TileRange {
  int begin, end, step;
};
__device__ __forceinline__ TileRange conv_tile_range(const ConvFwdParams& p) {
  if ((p.stats || p.gn_sums) && !(p.dbg & 2)) {
    const int per = (p.total_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int b = static_cast<int>(blockIdx.x) * per;
    return TileRange{b, min(p.total_tiles, b + per), 1};
  }
  return TileRange{static_cast<int>(blockIdx.x), p.total_tiles, static_cast<int>(gridDim.x)};
}

// Epilogue threads (256): add the CTA's shared-memory sums of image `img` into the global [N][Cout][2] array and clear
// them.  Named barrier 2 on both sides (every epilogue thread must call it at the same point of its tile loop).
__device__ __forceinline__ void conv_flush_sums(const ConvFwdParams& p, float* s_acc, int img) {
  bar_sync(2, kEpiThreads);
  float* dst = (p.stats ? p.stats : p.gn_sums) + static_cast<size_t>(img) * p.Cout * 2;
  const uint32_t acc_a = smem_u32(s_acc);
  for (int i = (static_cast<int>(threadIdx.x) - 64) * 4; i < 2 * p.Cout; i += kEpiThreads * 4) {
    const float4 v = lds_f4(acc_a + i * 4);
    sts_v4(acc_a + i * 4, make_uint4(0u, 0u, 0u, 0u));
    if (p.dbg & 1) *reinterpret_cast<float4*>(dst + i) = v;  // (wrong sums: timing experiments only)
    else red_add_v4f(dst + i, v.x, v.y, v.z, v.w);
  }
  bar_sync(2, kEpiThreads);
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
  if (p.res && p.res_mode == 0 && valid && c < BLOCK_N) {
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
// RAGGED: output tiles may hang over the image border (generic kernel); the halo kernel tiles its output exactly.
// Residual slabs fetched by TMA INTO the output staging buffers (BLOCK_N >= 128 halo kernels: no shared memory left for
// separate residual tiles).  A thread reads the residual of exactly the 16-byte chunks it overwrites with its outputs, so
// the buffer needs no extra hand-over: the load of slab s+1 goes into the other staging buffer as soon as the store of
// slab s-1 has finished reading it (the wait the issuer does anyway), i.e. right behind the barrier of slab s; the
// load for the first slab of the NEXT tile is issued behind the last slab's barrier and flies during the mainloop.
struct ResInplace {
  const CUtensorMap* tmR;
  uint64_t* rfull;      // one mbarrier per staging buffer
  uint32_t phase;       // bit b = parity to wait for on buffer b
  bool has_next;        // coordinates of the next tile's first slab
  int nco, n1, n2, n3;
};

template <int BLOCK_N>
__device__ __forceinline__ void conv_res_prefetch_rest(const ConvFwdParams& p, const CUtensorMap* tmR, int co, int a1,
                                                       int a2, int a3) {
#pragma unroll
  for (int j = 1; j < BLOCK_N / 64; ++j)
    if (co + 64 * j < p.Cout) tma_prefetch_l2_4d(tmR, co + 64 * j, a1, a2, a3);
}

template <int BLOCK_N, bool RAGGED = true>
__device__ __forceinline__ void conv_epilogue_tile_tma(const ConvFwdParams& p, EpiPrefetch& pf, const float* s_bias,
                                                       uint32_t t_acc, int q, int half, int n_tile, bool valid,
                                                       size_t pix, uint8_t* stage, int& stage_idx,
                                                       const CUtensorMap* tmY, int c1, int c2, int c3, bool issuer,
                                                       const uint8_t* res_tile = nullptr, int tile_w = 8,
                                                       float* s_acc = nullptr, ResInplace* rin = nullptr) {
  static_assert(BLOCK_N % 64 == 0, "TMA-store epilogue works on 64-channel slabs");
  const uint32_t t_row = t_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int row = q * 32 + (threadIdx.x & 31);
  const int swz = row & 7;
  // every shared-memory access below goes by shared-space address (LDS / STS, see ptx.cuh)
  const uint32_t stage_a = smem_u32(stage);
  const uint32_t bias_a = s_bias ? smem_u32(s_bias) : 0u;
  const uint32_t res_a = res_tile ? smem_u32(res_tile) + row * 128 : 0u;
  const uint32_t acc_a = s_acc ? smem_u32(s_acc) : 0u;
  // Fused GroupNorm reductions run as a COLUMN pass over the staged slab after the slab barrier.  Epilogue warp w8
  // owns the channel pairs 4*w8 .. 4*w8+3 of the slab; lane = (row lane rl, pair pp) reads rows rl, rl+8, ..: the 32
  // lanes of a load hit 32 different banks (the 128B swizzle XORs the 16-byte chunk index w8 with row mod 8 = rl).
  const int w8 = (static_cast<int>(threadIdx.x) - 64) >> 5;
  const int pp = threadIdx.x & 3, rl = (threadIdx.x & 31) >> 2;
#pragma unroll 1
  for (int c = half * 32; c < BLOCK_N; c += 64) {
    const int co0 = n_tile * BLOCK_N + c;
    const int slab_co = co0 - half * 32;         // first channel of this 64-channel slab
    if (slab_co >= p.Cout) break;                // uniform over all 8 warps: the rest of the tile is channel padding
    uint32_t v[32];
    tmem_ld_32x32(t_row + c, v);
    // While the TMEM load is in flight: the first half of the bias (shared memory, zero-padded past Cout) and the
    // residual.  Channels >= Cout hold don't-care values that the TMA store clips.
    float4 bias4[4];
    if (s_bias) {
#pragma unroll
      for (int g = 0; g < 4; ++g) bias4[g] = lds_f4(bias_a + (co0 + g * 4) * 4);
    }
    uint4 rcur[4];
    const bool has_res = p.res != nullptr;
    if (rin) {
      // residual slab fetched into THIS slab's staging buffer (see ResInplace)
      mbar_wait(&rin->rfull[stage_idx], (rin->phase >> stage_idx) & 1u);
      rin->phase ^= 1u << stage_idx;
      const uint32_t ra = stage_a + stage_idx * kStageBytes + row * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) rcur[g] = lds_v4(ra + (((half * 4 + g) ^ swz) << 4));
    } else if (res_tile && p.res_mode == 0) {
      // residual slab staged in shared memory by TMA (same 128B-swizzled layout as the output staging)
#pragma unroll
      for (int g = 0; g < 4; ++g) rcur[g] = lds_v4(res_a + (((half * 4 + g) ^ swz) << 4));
    } else if (has_res && p.res_mode == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) rcur[g] = pf.r[g];
      if (valid && c + 64 < BLOCK_N) {
        const __nv_bfloat16* rn = p.res + pix * p.ldres + co0 + 64;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (co0 + 64 + g * 8 < p.Cout) pf.r[g] = *reinterpret_cast<const uint4*>(rn + g * 8);
      }
    }
    const int colco = slab_co + 2 * (4 * w8 + pp);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (s_bias) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        f[g * 4 + 0] += bias4[g].x; f[g * 4 + 1] += bias4[g].y;
        f[g * 4 + 2] += bias4[g].z; f[g * 4 + 3] += bias4[g].w;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) bias4[g] = lds_f4(bias_a + (co0 + 16 + g * 4) * 4);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        f[16 + g * 4 + 0] += bias4[g].x; f[16 + g * 4 + 1] += bias4[g].y;
        f[16 + g * 4 + 2] += bias4[g].z; f[16 + g * 4 + 3] += bias4[g].w;
      }
    }
    if (has_res && p.res_mode == 0 && (rin || res_tile || valid)) {
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
    const uint32_t buf_a = stage_a + stage_idx * kStageBytes;
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // 8 channels per 16-byte chunk; 128B swizzle: chunk index XOR (row mod 8)
      uint4 o;
      o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
      o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
      o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
      o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
      sts_v4(buf_a + row * 128 + (((half * 4 + g) ^ swz) << 4), o);
    }
    // x words of the first XB rows, issued here (the accumulator registers are dead by now) so that the loads fly
    // during the fence / slab barrier / store issue below.  XB = 16 rows at once where registers allow (BLOCK_N = 64,
    // the kernels with the least time per tile), two batches of 8 otherwise.
    constexpr int XB = BLOCK_N == 64 ? 16 : 8;
    uint32_t xw[XB];
    float4 ab4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_x = [&](int k0) {
#pragma unroll
      for (int k = 0; k < XB; ++k) {
        const int rr = rl + 8 * (k0 + k);
        const int pw = c1 + rr % tile_w, ph = c2 + rr / tile_w;
        xw[k] = (pw < p.Wo && ph < p.Ho)
                    ? *reinterpret_cast<const uint32_t*>(p.res + ((static_cast<size_t>(c3) * p.Ho + ph) * p.Wo + pw) * p.ldres + colco)
                    : 0u;
      }
    };
    // gn_sums with res_tile: the GroupNorm input tile was fetched by TMA (64-channel kernels) — read it like the slab
    const bool x_smem = res_tile != nullptr && p.res_mode == 1;
    if (p.gn_sums && colco < p.Cout) {
      ab4 = *reinterpret_cast<const float4*>(p.gn_ab + (static_cast<size_t>(c3) * p.Cout + colco) * 2);
      if (!x_smem) load_x(0);
    }
    fence_proxy_async();                 // generic-proxy smem writes -> visible to the TMA (async proxy)
    if (issuer) bulk_wait_read<0>();     // the previous slab's store has released the other buffer
    bar_sync(1, kEpiThreads);
    if (issuer) {
      if (rin) {  // the other buffer is free (wait above): fetch the next slab's residual into it
        const bool more = c + 64 < BLOCK_N && slab_co + 64 < p.Cout;
        if (more || rin->has_next) {
          uint64_t* rb = &rin->rfull[stage_idx ^ 1];
          mbar_arrive_expect_tx(rb, kStageBytes);
          tma_load_4d(stage + (stage_idx ^ 1) * kStageBytes, rin->tmR, rb, more ? slab_co + 64 : rin->nco,
                      more ? c1 : rin->n1, more ? c2 : rin->n2, more ? c3 : rin->n3);
          // The slabs AFTER the first of the next tile are loaded one slab ahead only (their buffer is busy until then):
          // too short for DRAM latency.  Pull them into L2 now, a whole mainloop ahead.
          if (!more) conv_res_prefetch_rest<BLOCK_N>(p, rin->tmR, rin->nco, rin->n1, rin->n2, rin->n3);
        }
      }
      tma_store_4d(tmY, stage + stage_idx * kStageBytes, slab_co, c1, c2, c3);
      bulk_commit();
    }
    if ((p.stats || p.gn_sums) && colco < p.Cout) {
      // (this buffer is rewritten two slabs from now, behind the next slab's barrier: every thread has left by then)
      // The two modes are separate loops (the mode test is NOT inside the row loop): only the code of the active mode
      // is ever fetched — with both interleaved, ncu showed the kernel stalling on instruction-cache misses.
      const uint32_t col_a = buf_a + rl * 128 + (((w8 ^ rl) << 4) | (pp << 2));
      float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
      if (p.stats) {
        // statistics of the STORED values for the GroupNorm that reads y next: sum, sum of squares.  All 16 row words
        // are loaded before the first add (16 independent LDS in flight), two accumulator chains per sum.
        uint32_t yw[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) yw[k] = lds_u32(col_a + k * 1024);
        float s0b = 0.f, s1b = 0.f, t0b = 0.f, t1b = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          float2 ya = unpack_bf16x2(yw[k]), yb = unpack_bf16x2(yw[k + 1]);
          if (RAGGED) {
            const int ra = rl + 8 * k, rb = ra + 8;
            if (!((c1 + ra % tile_w < p.Wo) && (c2 + ra / tile_w < p.Ho))) ya = make_float2(0.f, 0.f);
            if (!((c1 + rb % tile_w < p.Wo) && (c2 + rb / tile_w < p.Ho))) yb = make_float2(0.f, 0.f);
          }
          s0 += ya.x; t0 = fmaf(ya.x, ya.x, t0);
          s1 += ya.y; t1 = fmaf(ya.y, ya.y, t1);
          s0b += yb.x; t0b = fmaf(yb.x, yb.x, t0b);
          s1b += yb.y; t1b = fmaf(yb.y, yb.y, t1b);
        }
        s0 += s0b; t0 += t0b; s1 += s1b; t1 += t1b;
      } else if (x_smem) {
        // GroupNorm-backward sums, x tile in shared memory (same swizzled layout as the slab)
        const uint32_t xcol_a = smem_u32(res_tile) + rl * 128 + (((w8 ^ rl) << 4) | (pp << 2));
        uint32_t yw[16], xs[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          yw[k] = lds_u32(col_a + k * 1024);
          xs[k] = lds_u32(xcol_a + k * 1024);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 yv = unpack_bf16x2(yw[k]), xv = unpack_bf16x2(xs[k]);
          const float du0 = yv.x * act_grad_rt(fmaf(xv.x, ab4.x, ab4.y), p.gn_act);
          const float du1 = yv.y * act_grad_rt(fmaf(xv.y, ab4.z, ab4.w), p.gn_act);
          s0 += du0; t0 = fmaf(du0, xv.x, t0);
          s1 += du1; t1 = fmaf(du1, xv.y, t1);
        }
      } else {
        // GroupNorm-backward sums: y is dL/d(act output); du = y * act'(a*x + b); A = sum du, B = sum du * x
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += XB) {
          if (k0 > 0) load_x(k0);
#pragma unroll
          for (int k = 0; k < XB; ++k) {
            const int rr = rl + 8 * (k0 + k);
            const float2 yv = unpack_bf16x2(lds_u32(col_a + (k0 + k) * 1024));
            const float2 xv = unpack_bf16x2(xw[k]);
            const bool live = !RAGGED || ((c1 + rr % tile_w < p.Wo) && (c2 + rr / tile_w < p.Ho));
            const float du0 = live ? yv.x * act_grad_rt(fmaf(xv.x, ab4.x, ab4.y), p.gn_act) : 0.f;
            const float du1 = live ? yv.y * act_grad_rt(fmaf(xv.y, ab4.z, ab4.w), p.gn_act) : 0.f;
            s0 += du0; t0 = fmaf(du0, xv.x, t0);
            s1 += du1; t1 = fmaf(du1, xv.y, t1);
          }
        }
      }
      // Sum over the 8 row lanes (lane bits 2..4) as a transpose-reduction: 4 shuffles instead of 12.  After the xor-16
      // step a lane keeps one channel's (sum, sum2), after xor-8 one of the two, xor-4 finishes it: the four lanes
      // (bit 4, bit 3) with bit 2 clear each own one of the pair's four sums and are the only threads that ever touch
      // that shared-memory word (no atomics: fp32 shared atomics are CAS loops).
      const bool b4 = (threadIdx.x & 16) != 0, b3 = (threadIdx.x & 8) != 0;
      const float ka = (b4 ? s1 : s0) + __shfl_xor_sync(0xffffffffu, b4 ? s0 : s1, 16);
      const float kb = (b4 ? t1 : t0) + __shfl_xor_sync(0xffffffffu, b4 ? t0 : t1, 16);
      float val = (b3 ? kb : ka) + __shfl_xor_sync(0xffffffffu, b3 ? ka : kb, 8);
      val += __shfl_xor_sync(0xffffffffu, val, 4);
      if ((threadIdx.x & 4) == 0) {  // CTA-local sums of the current image; conv_flush_sums ships them once per image
        const uint32_t a = acc_a + (2 * colco + (b4 ? 2 : 0) + (b3 ? 1 : 0)) * 4;
        sts_f32(a, lds_f32(a) + val);
      }
    }
    stage_idx ^= 1;
  }
}

// Launch of the halo-reuse 3x3 kernel (conv_halo.cu); returns JG_ERR_UNSUPPORTED when the shape does not qualify.
// e (may be NULL): fused GroupNorm work; *fused is set when the launched kernel's epilogue did it.
int launch_conv_halo(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                     const float* bias, const void* residual, void* y, cudaStream_t stream, bool* fused);

// CTA-pair (tcgen05.mma.cta_group::2) variant of the halo kernel (conv_halo2.cu); same contract.
int launch_conv_halo2(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                      const float* bias, const void* residual, void* y, cudaStream_t stream, bool* fused);

// Weight gradient of 3x3 convolutions on CTA pairs (conv_halo2.cu): raw accumulation into acc [9][Cin][Cout].
int launch_wgrad_halo2(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* acc, cudaStream_t stream);
// [R*S][Cin][Cout] accumulator -> OIHW gradient, dst = beta * dst + src (conv_halo.cu)
int launch_unpack_hwio(const float* src, float* dst, int Cout, int Cin, int RS, float beta, cudaStream_t stream);

// Fills the fused-GroupNorm fields of the kernel parameters from the C-ABI structs; returns the operand that travels
// through the residual plumbing (the residual itself, or the GroupNorm input x of the gn_sums mode).
inline const void* conv_apply_epilogue(ConvFwdParams& p, const jg_conv_desc* d, const jg_conv_epilogue* e,
                                       const void* residual) {
  p.stats = nullptr; p.gn_sums = nullptr; p.gn_ab = nullptr; p.gn_act = 0; p.res_mode = 0;
  static const int dbg = getenv("JG_DBG_EPI") ? atoi(getenv("JG_DBG_EPI")) : 0;
  p.dbg = dbg;
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


// wgrad with halo reuse (conv_halo.cu): zeroes ws, accumulates, writes OIHW; JG_ERR_UNSUPPORTED if not eligible.
int launch_wgrad_halo(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* ws, float* dw_oihw,
                      float beta, cudaStream_t stream);

}  // namespace jg
