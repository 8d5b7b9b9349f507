// Flash-attention FORWARD on tcgen05 (QKVAttentionLegacy / QKVAttention, unet_generator_attn.py:331-347,
// unet_generator_attn_vid.py:334-363): S = Q K^T and O += P V on the 5th-generation tensor cores with the accumulators
// in TMEM, the softmax in fp32 by two warpgroups that ping-pong on the tensor pipe.  Same contract as attn_fwd_kernel
// (attention.cu): qkv NHWC [N][T][ldqkv] bf16, head h owns `ch` channels of q / k / v at offsets
// (h*hstride, +koff, +voff); out [N][T][ldo] bf16; lse fp32 [N*heads][T] in the log2 domain.
//
// One CTA = one (image, head) and TWO 128-query tiles (256 queries), 320 threads:
//   warp 0      TMA producer: Q0, Q1 once; then (K_j, V_j) 128-key blocks through a 3-stage ring.  Every tile is a
//               (64 channels x 128 rows) box of the [3C] x [N*T] tensor map, SWIZZLE_128B: for ch = 32 the box is twice
//               as wide as the head — the MMAs simply never read (Q, K: only the first K = ch of the 128-byte rows
//               is stepped through) or never use (V: accumulator columns >= ch) the neighbour's channels; channels
//               past 3C are the TMA unit's zero fill.  That keeps every operand in the K-major / MN-major
//               SWIZZLE_128B forms the convolution kernels established on hardware (tools/umma_probe.cu).
//   warp 1      MMA issuer: S_g = Q_g K_j^T (M128 N128 K=ch, both K-major), O_g += P_g V_j (M128 N64 K128, A = P
//               K-major from shared memory, B = V MN-major).  Issue order S0 S1 | PV0 S0' PV1 S1' | ...: while
//               warpgroup 0 runs the softmax of block j+1 the tensor pipe does PV1_j and S1_{j+1}.
//   warps 2..9  two softmax warpgroups (g = 0, 1), one query row per thread: tcgen05.ld of the 128 scores, running
//               max / sum in the exp2 domain, P as bf16 into the 128B-swizzled K-major tile (conflict-free 16-byte
//               stores, as the conv epilogue's staging), fence.proxy.async, mbarrier.  The O accumulator is rescaled
//               in TMEM only when a row's max grew by more than 2^8 since the last rescale (the stale max stays a
//               valid softmax reference; l and the final 1/l use the same one).
// TMEM: S0 | S1 (128 fp32 columns each), O0 | O1 (64 each) = 384 of 512 columns.
// Shared memory: Q 2 x 16 KB, K/V ring 3 x 32 KB, P 2 x 32 KB = 192 KB.
#include "common.cuh"
#include "ptx.cuh"

namespace jg {
namespace {

constexpr int kTcThreads = 320;
constexpr int kKvStages = 3;
constexpr int kTileBytes = 128 * 128;  // 128 rows x 128 B
constexpr float kRescaleThreshold = 8.f;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct AttnTcParams {
  __nv_bfloat16* out;
  float* lse;
  int ldo, T, heads, hstride, koff, voff;
  float scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcParams p) {
  static_assert(HD == 32 || HD == 64, "head dim 32 or 64");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 tiles
  uint8_t* sKV = sQ + 2 * kTileBytes;                   // kKvStages x (K tile, V tile)
  uint8_t* sP = sKV + kKvStages * 2 * kTileBytes;       // 2 x (2 slabs of 64 keys)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kTileBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // kKvStages
  uint64_t* kv_empty = kv_full + kKvStages;
  uint64_t* s_full = kv_empty + kKvStages; // 2
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* pv_done = p_full + 2;          // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int n = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.x * 256;
  const int nb = p.T / 128;
  const int row0 = n * p.T;               // first row of this image in the [N*T] dimension
  const int cq = h * p.hstride, ck = cq + p.koff, cv = cq + p.voff;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);   // the four warps of a softmax warpgroup
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tS[2] = {tmem, tmem + 128};
  const uint32_t tO[2] = {tmem + 256, tmem + 320};

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      tma_load_2d(sQ, &tmQKV, q_full, cq, row0 + q0);
      tma_load_2d(sQ + kTileBytes, &tmQKV, q_full, cq, row0 + q0 + 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], 2 * kTileBytes);
        uint8_t* st = sKV + stage * 2 * kTileBytes;
        tma_load_2d(st, &tmQKV, &kv_full[stage], ck, row0 + j * 128);
        tma_load_2d(st + kTileBytes, &tmQKV, &kv_full[stage], cv, row0 + j * 128);
        if (++stage == kKvStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t q_desc[2] = {make_smem_desc_sw128(smem_u32(sQ), 16, 1024),
                                make_smem_desc_sw128(smem_u32(sQ + kTileBytes), 16, 1024)};
    const uint64_t p_desc[2] = {make_smem_desc_sw128(smem_u32(sP), 16, 1024),
                                make_smem_desc_sw128(smem_u32(sP + 2 * kTileBytes), 16, 1024)};
    auto issue_s = [&](int g, int stage) {
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sKV + stage * 2 * kTileBytes), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tS[g], q_desc[g] + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[g]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int g, int stage, bool accumulate, bool release_stage) {
      // B = V tile [128 keys][64 channels] MN-major: SBO = 8-key group stride, 16 keys per MMA = 2048 B
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sKV + stage * 2 * kTileBytes + kTileBytes), 8192, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = P [128 rows][128 keys] K-major: two 64-key slabs of 16 KB, 32 B per 16 keys inside a slab
          const uint64_t a = p_desc[g] + ((k >> 2) * (kTileBytes >> 4)) + 2 * (k & 3);
          umma_bf16(tO[g], a, v_desc + k * 128, idesc_pv, (accumulate || k != 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[g]);
        if (release_stage) umma_commit(&kv_empty[stage]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    int stage = 0;
    uint32_t phase = 0;       // of kv_full[stage]
    uint32_t pphase = 0;      // of p_full[g] at block j
    for (int j = 0; j < nb; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kKvStages) {
        nstage = 0;
        nphase ^= 1;
      }
      mbar_wait(&p_full[0], pphase);
      tc_fence_after();
      issue_pv(0, stage, j > 0, false);
      if (j + 1 < nb) {
        mbar_wait(&kv_full[nstage], nphase);
        tc_fence_after();
        issue_s(0, nstage);
      }
      mbar_wait(&p_full[1], pphase);
      tc_fence_after();
      issue_pv(1, stage, j > 0, true);
      if (j + 1 < nb) issue_s(1, nstage);
      stage = nstage;
      phase = nphase;
      pphase ^= 1;
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                // query row within the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t p_a = smem_u32(sP + g * 2 * kTileBytes) + row * 128;
    const int swz = row & 7;
    float m_used = -INFINITY, l = 0.f;
    float c_pend = 1.f;   // rescale of O_g decided in block j, applied at the start of block j+1 (or in the epilogue)
    uint32_t ph = 0;
    auto rescale_o = [&](float c) {  // whole warp; c per row
      uint32_t o[32];
#pragma unroll
      for (int cc = 0; cc < HD / 32; ++cc) {
        tmem_ld_32x32(tO[g] + lane_off + cc * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * c);
        tmem_st_32x32(tO[g] + lane_off + cc * 32, o);
      }
      tmem_st_wait();
    };
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      if (j > 0) {
        // P_g / O_g of block j-1 are consumed / final (PV_g(j-1) was issued before S_g(j): this wait does not stall)
        mbar_wait(&pv_done[g], ph ^ 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, c_pend != 1.f)) rescale_o(c_pend);
        c_pend = 1.f;
      } else {
        // first block: the exact row max (a second pass over TMEM; later blocks use the running reference)
        float mx8[8];
        uint32_t v[32];
        tmem_ld_32x32(tS[g] + lane_off, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(v[i]);
#pragma unroll
        for (int i = 8; i < 32; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(v[i]));
#pragma unroll
        for (int c = 1; c < 4; ++c) {
          tmem_ld_32x32(tS[g] + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(v[i]));
        }
        m_used = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                       fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]))) * p.scale_log2;
      }
      // One pass: p = 2^(s * scale - m_used) against the RUNNING reference m_used (any reference gives the same
      // softmax as long as O and l use the same one), row sum, row max relative to the reference, bf16 P into the
      // swizzled K-major tile.  Independent accumulator chains (8 sums / 8 maxima); the TMEM load of chunk c+1 flies
      // while chunk c is processed.
      float rs, mrel;
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        const float nm = -m_used;
        float rs8[8], mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rs8[i] = 0.f;
          mx8[i] = -INFINITY;
        }
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tS[g] + lane_off, va);
        auto chunk = [&](const uint32_t (&v)[32], int c) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {  // 8 keys per 16-byte chunk
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float t = fmaf(__uint_as_float(v[i4 * 8 + i]), p.scale_log2, nm);
              mx8[i] = fmaxf(mx8[i], t);
              e[i] = fast_exp2(t);
              rs8[i] += e[i];
            }
            uint4 o;
            o.x = pack_bf16x2(e[0], e[1]);
            o.y = pack_bf16x2(e[2], e[3]);
            o.z = pack_bf16x2(e[4], e[5]);
            o.w = pack_bf16x2(e[6], e[7]);
            const int c8 = c * 4 + i4;
            sts_v4(p_a + (c8 >> 3) * kTileBytes + (((c8 & 7) ^ swz) << 4), o);
          }
        };
        tmem_ld_wait();
        tmem_ld_32x32(tS[g] + lane_off + 32, vb);
        chunk(va, 0);
        tmem_ld_wait();
        tmem_ld_32x32(tS[g] + lane_off + 64, va);
        chunk(vb, 1);
        tmem_ld_wait();
        tmem_ld_32x32(tS[g] + lane_off + 96, vb);
        chunk(va, 2);
        tmem_ld_wait();
        chunk(vb, 3);
        rs = ((rs8[0] + rs8[1]) + (rs8[2] + rs8[3])) + ((rs8[4] + rs8[5]) + (rs8[6] + rs8[7]));
        mrel = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                     fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        // a row whose scores outgrew the reference by more than 2^64 would overflow: adopt the new max NOW (rescale
        // O_g, which nobody is accumulating into at this point) and redo the pass.  Practically never taken.
        if (attempt == 0 && __any_sync(0xffffffffu, mrel > 64.f)) {
          const float c = mrel > 64.f ? fast_exp2(-mrel) : 1.f;
          if (j > 0) rescale_o(c);
          l *= c;
          if (mrel > 64.f) m_used += mrel;
          continue;
        }
        break;
      }
      l += rs;
      if (mrel > kRescaleThreshold) {  // adopt the larger max; O_g is rescaled when PV_g(j) has landed (next block)
        c_pend = fast_exp2(-mrel);
        l *= c_pend;
        m_used += mrel;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      ph ^= 1;
    }
    // epilogue: O / l -> bf16, lse
    mbar_wait(&pv_done[g], ph ^ 1);
    tc_fence_after();
    const float inv = c_pend / l;   // (a rescale still pending from the last block folds into the normalisation)
    const int t = q0 + g * 128 + row;
    __nv_bfloat16* op = p.out + (static_cast<size_t>(n) * p.T + t) * p.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(tO[g] + lane_off + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o[i * 8 + 0]) * inv, __uint_as_float(o[i * 8 + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o[i * 8 + 2]) * inv, __uint_as_float(o[i * 8 + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o[i * 8 + 4]) * inv, __uint_as_float(o[i * 8 + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o[i * 8 + 6]) * inv, __uint_as_float(o[i * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(op + c * 32 + i * 8) = w;
      }
    }
    p.lse[static_cast<size_t>(bh) * p.T + t] = m_used + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// ----------------------------------------------------------------------------------------------------------------------
// Backward on tcgen05.  With P_ij = 2^(s_ij * scale_log2 - lse_i), dP = dO V^T, dS = P o (dP - D_i) (D = rowsum(dO o O),
// attn_bwd_prep_kernel):   dV = P^T dO,   dK = scale * dS^T Q,   dQ = scale * dS K.
// Two kernels, as the mma.sync version (no atomics): dK / dV per 128-key tile looping over the query blocks, dQ per
// 128-query tile looping over the key blocks.  576 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17
// elementwise: thread = (accumulator row, column half) — two warps share a TMEM lane quarter and split the 128 columns.
// Per block two "score" MMAs (K = ch) fill S and dP in TMEM, the elementwise warps turn them into bf16 P / dS tiles in
// shared memory (128B-swizzled K-major, the conflict-free 16-byte stores of the forward), two (dK/dV) or one (dQ)
// accumulation MMAs (K = 128) consume them.  Every global operand tile is the same (64 channels x 128 rows) TMA box as in
// the forward and serves as K-major operand of the score MMAs AND as MN-major operand of the accumulation MMAs.
// ----------------------------------------------------------------------------------------------------------------------
// 16 elementwise warps: thread = (accumulator row, column quarter of 32) — four warps per SM sub-partition hide the
// MUFU / TMEM-load latencies that two could not (ncu on the forward: issue slots 36 % busy with 8 softmax warps).
constexpr int kBwdThreads = 64 + 16 * 32;

struct AttnBwdParams {
  const float* lse;   // [N*heads][T], log2 domain
  const float* D;     // [N*heads][T]
  __nv_bfloat16* dqkv;
  int lddqkv, T, heads, hstride, koff, voff;
  float scale_log2, scale;
};

// P / dS tile writer: row r of a [128 rows][128 cols] bf16 K-major tile (two 64-column slabs), columns [hq*32 + c4*8, +8)
__device__ __forceinline__ void store_row_quarter(uint32_t tile_a, int r, int hq, int c4, const float (&e)[8]) {
  const int hc = hq >> 1, c = (hq & 1) * 4 + c4;  // slab of 64 columns, 16-byte chunk inside the slab's 128-byte row
  uint4 o;
  o.x = pack_bf16x2(e[0], e[1]);
  o.y = pack_bf16x2(e[2], e[3]);
  o.z = pack_bf16x2(e[4], e[5]);
  o.w = pack_bf16x2(e[6], e[7]);
  sts_v4(tile_a + hc * kTileBytes + r * 128 + ((c ^ (r & 7)) << 4), o);
}

// ---- dK, dV: CTA = (128-key tile, image, head); loop over query blocks ------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                       const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTileBytes;
  uint8_t* sQD = smem + 2 * kTileBytes;            // 2 stages x (Q_i tile, dO_i tile)
  uint8_t* sPT = sQD + 4 * kTileBytes;             // P^T  [keys][queries], 2 slabs
  uint8_t* sDS = sPT + 2 * kTileBytes;             // dS^T [keys][queries], 2 slabs
  float* sVec = reinterpret_cast<float*>(sDS + 2 * kTileBytes);  // 2 x (lse[128] | D[128])
  uint64_t* bars = reinterpret_cast<uint64_t*>(sVec + 2 * 256);
  uint64_t* kv_full = bars;          // 1
  uint64_t* st_full = bars + 1;      // 2
  uint64_t* st_empty = bars + 3;     // 2
  uint64_t* sdp_full = bars + 5;     // 1
  uint64_t* pds_full = bars + 6;     // 1 (8 warps)
  uint64_t* acc_done = bars + 7;     // 1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int n = bh / p.heads, h = bh % p.heads;
  const int k0 = blockIdx.x * 128;
  const int nq = p.T / 128;
  const int row0 = n * p.T;
  const int cq = h * p.hstride, ck = cq + p.koff, cv = cq + p.voff;
  const int cdo = h * HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 16);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * kTileBytes);
      tma_load_2d(sK, &tmQKV, kv_full, ck, row0 + k0);
      tma_load_2d(sV, &tmQKV, kv_full, cv, row0 + k0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nq; ++i) {
        mbar_wait(&st_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&st_full[stage], 2 * kTileBytes);
        uint8_t* st = sQD + stage * 2 * kTileBytes;
        tma_load_2d(st, &tmQKV, &st_full[stage], cq, row0 + i * 128);
        tma_load_2d(st + kTileBytes, &tmDO, &st_full[stage], cdo, row0 + i * 128);
        if (++stage == 2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
    const uint64_t pt_desc = make_smem_desc_sw128(smem_u32(sPT), 16, 1024);
    const uint64_t ds_desc = make_smem_desc_sw128(smem_u32(sDS), 16, 1024);
    auto issue_scores = [&](int stage) {
      const uint32_t st = smem_u32(sQD + stage * 2 * kTileBytes);
      const uint64_t q_desc = make_smem_desc_sw128(st, 16, 1024);
      const uint64_t do_desc = make_smem_desc_sw128(st + kTileBytes, 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tS, k_desc + 2 * k, q_desc + 2 * k, idesc_s, k != 0);    // S^T = K Q^T
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tDP, v_desc + 2 * k, do_desc + 2 * k, idesc_s, k != 0);  // dP^T = V dO^T
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    mbar_wait(&st_full[0], 0);
    tc_fence_after();
    issue_scores(0);
    int stage = 0;
    uint32_t phase = 0, eph = 0;
    for (int i = 0; i < nq; ++i) {
      mbar_wait(pds_full, eph);
      tc_fence_after();
      const uint32_t st = smem_u32(sQD + stage * 2 * kTileBytes);
      const uint64_t q_mn = make_smem_desc_sw128(st, 8192, 1024);                 // Q_i  [queries][ch] as MN-major B
      const uint64_t do_mn = make_smem_desc_sw128(st + kTileBytes, 8192, 1024);   // dO_i
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t off = ((k >> 2) * (kTileBytes >> 4)) + 2 * (k & 3);
          umma_bf16(tDV, pt_desc + off, do_mn + k * 128, idesc_acc, (i > 0 || k != 0) ? 1u : 0u);  // dV += P^T dO
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t off = ((k >> 2) * (kTileBytes >> 4)) + 2 * (k & 3);
          umma_bf16(tDK, ds_desc + off, q_mn + k * 128, idesc_acc, (i > 0 || k != 0) ? 1u : 0u);   // dK += dS^T Q
        }
        umma_commit(acc_done);
        umma_commit(&st_empty[stage]);
      }
      __syncwarp();
      if (++stage == 2) {
        stage = 0;
        phase ^= 1;
      }
      eph ^= 1;
      if (i + 1 < nq) {
        mbar_wait(&st_full[stage], phase);
        tc_fence_after();
        issue_scores(stage);
      }
    }
  } else {
    // elementwise warps: thread = (key row r, query-column quarter hq)
    const int q = warp & 3;
    const int hq = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int vi = threadIdx.x - 64;  // 0..511: threads 0..255 stage one float of (lse | D) each
    const float* vsrc = (vi < 128 ? p.lse : p.D) + static_cast<size_t>(bh) * p.T + (vi & 127);
    const uint32_t pt_a = smem_u32(sPT), ds_a = smem_u32(sDS);
    const uint32_t vec_a = smem_u32(sVec);
    uint32_t ph = 0;
    float myv = vi < 256 ? vsrc[0] : 0.f;  // (fetched one block ahead: ncu had 10 % of the samples on this store's load)
    for (int i = 0; i < nq; ++i) {
      if (vi < 256) sts_f32(vec_a + ((i & 1) * 256 + vi) * 4, myv);
      if (vi < 256 && i + 1 < nq) myv = vsrc[(i + 1) * 128];
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      if (i > 0) {  // the accumulation MMAs of block i-1 have read P^T / dS^T (they were issued before S, dP of block i)
        mbar_wait(acc_done, ph ^ 1);
      }
      bar_sync(1, 512);  // lse / D of this query block are staged
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tS + lane_off + hq * 32, sv);
      tmem_ld_32x32(tDP + lane_off + hq * 32, dv);
      const uint32_t va = vec_a + ((i & 1) * 256 + hq * 32) * 4;
      tmem_ld_wait();
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float pe[8], de[8];
        const float4 l0 = lds_f4(va + i4 * 32), l1 = lds_f4(va + i4 * 32 + 16);
        const float4 d0 = lds_f4(va + 512 + i4 * 32), d1 = lds_f4(va + 512 + i4 * 32 + 16);
        const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          pe[k] = fast_exp2(fmaf(__uint_as_float(sv[i4 * 8 + k]), p.scale_log2, -ls[k]));
          de[k] = pe[k] * (__uint_as_float(dv[i4 * 8 + k]) - dd[k]);
        }
        store_row_quarter(pt_a, r, hq, i4, pe);
        store_row_quarter(ds_a, r, hq, i4, de);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      ph ^= 1;
    }
    const int hc = hq;  // epilogue roles: quarter 0 writes dV, quarter 1 dK; 2 and 3 are done
    // epilogue: warps of half 0 write dV, half 1 write dK (* scale)
    mbar_wait(acc_done, ph ^ 1);
    tc_fence_after();
    const float mul = hc == 0 ? 1.f : p.scale;
    __nv_bfloat16* op = p.dqkv + (static_cast<size_t>(n) * p.T + k0 + r) * p.lddqkv + (hc == 0 ? cv : ck);
#pragma unroll
    for (int c = 0; c < (hc < 2 ? HD / 32 : 0); ++c) {
      uint32_t o[32];
      tmem_ld_32x32((hc == 0 ? tDV : tDK) + lane_off + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o[i * 8 + 0]) * mul, __uint_as_float(o[i * 8 + 1]) * mul);
        w.y = pack_bf16x2(__uint_as_float(o[i * 8 + 2]) * mul, __uint_as_float(o[i * 8 + 3]) * mul);
        w.z = pack_bf16x2(__uint_as_float(o[i * 8 + 4]) * mul, __uint_as_float(o[i * 8 + 5]) * mul);
        w.w = pack_bf16x2(__uint_as_float(o[i * 8 + 6]) * mul, __uint_as_float(o[i * 8 + 7]) * mul);
        *reinterpret_cast<uint4*>(op + c * 32 + i * 8) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- dQ: CTA = (128-query tile, image, head); loop over key blocks -----------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                      const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + kTileBytes;
  uint8_t* sKV = smem + 2 * kTileBytes;            // 2 stages x (K_j, V_j)
  uint8_t* sDS = sKV + 4 * kTileBytes;             // dS [queries][keys], 2 slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * kTileBytes);
  uint64_t* qd_full = bars;          // 1
  uint64_t* st_full = bars + 1;      // 2
  uint64_t* st_empty = bars + 3;     // 2
  uint64_t* sdp_full = bars + 5;
  uint64_t* ds_full = bars + 6;      // 8 warps
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int n = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.x * 128;
  const int nk = p.T / 128;
  const int row0 = n * p.T;
  const int cq = h * p.hstride, ck = cq + p.koff, cv = cq + p.voff;
  const int cdo = h * HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(qd_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(ds_full, 16);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tS = tmem, tDP = tmem + 128, tDQ = tmem + 256;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(qd_full, 2 * kTileBytes);
      tma_load_2d(sQ, &tmQKV, qd_full, cq, row0 + q0);
      tma_load_2d(sDO, &tmDO, qd_full, cdo, row0 + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nk; ++j) {
        mbar_wait(&st_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&st_full[stage], 2 * kTileBytes);
        uint8_t* st = sKV + stage * 2 * kTileBytes;
        tma_load_2d(st, &tmQKV, &st_full[stage], ck, row0 + j * 128);
        tma_load_2d(st + kTileBytes, &tmQKV, &st_full[stage], cv, row0 + j * 128);
        if (++stage == 2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t do_desc = make_smem_desc_sw128(smem_u32(sDO), 16, 1024);
    const uint64_t ds_desc = make_smem_desc_sw128(smem_u32(sDS), 16, 1024);
    auto issue_scores = [&](int stage) {
      const uint32_t st = smem_u32(sKV + stage * 2 * kTileBytes);
      const uint64_t k_desc = make_smem_desc_sw128(st, 16, 1024);
      const uint64_t v_desc = make_smem_desc_sw128(st + kTileBytes, 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tS, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);     // S = Q K^T
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tDP, do_desc + 2 * k, v_desc + 2 * k, idesc_s, k != 0);   // dP = dO V^T
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    mbar_wait(qd_full, 0);
    mbar_wait(&st_full[0], 0);
    tc_fence_after();
    issue_scores(0);
    int stage = 0;
    uint32_t phase = 0, eph = 0;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(ds_full, eph);
      tc_fence_after();
      const uint64_t k_mn = make_smem_desc_sw128(smem_u32(sKV + stage * 2 * kTileBytes), 8192, 1024);  // K_j as MN-major B
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t off = ((k >> 2) * (kTileBytes >> 4)) + 2 * (k & 3);
          umma_bf16(tDQ, ds_desc + off, k_mn + k * 128, idesc_acc, (j > 0 || k != 0) ? 1u : 0u);  // dQ += dS K
        }
        umma_commit(acc_done);
        umma_commit(&st_empty[stage]);
      }
      __syncwarp();
      if (++stage == 2) {
        stage = 0;
        phase ^= 1;
      }
      eph ^= 1;
      if (j + 1 < nk) {
        mbar_wait(&st_full[stage], phase);
        tc_fence_after();
        issue_scores(stage);
      }
    }
  } else {
    const int q = warp & 3;
    const int hq = (warp - 2) >> 2;   // column quarter
    const int hc = hq;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float nl = -p.lse[static_cast<size_t>(bh) * p.T + q0 + r];
    const float Dr = p.D[static_cast<size_t>(bh) * p.T + q0 + r];
    const uint32_t ds_a = smem_u32(sDS);
    uint32_t ph = 0;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      if (j > 0) mbar_wait(acc_done, ph ^ 1);
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tS + lane_off + hq * 32, sv);
      tmem_ld_32x32(tDP + lane_off + hq * 32, dv);
      tmem_ld_wait();
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float de[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float pe = fast_exp2(fmaf(__uint_as_float(sv[i4 * 8 + k]), p.scale_log2, nl));
          de[k] = pe * (__uint_as_float(dv[i4 * 8 + k]) - Dr);
        }
        store_row_quarter(ds_a, r, hq, i4, de);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      ph ^= 1;
    }
    mbar_wait(acc_done, ph ^ 1);
    tc_fence_after();
    if (hc == 0) {
      __nv_bfloat16* op = p.dqkv + (static_cast<size_t>(n) * p.T + q0 + r) * p.lddqkv + cq;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32(tDQ + lane_off + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[i * 8 + 0]) * p.scale, __uint_as_float(o[i * 8 + 1]) * p.scale);
          w.y = pack_bf16x2(__uint_as_float(o[i * 8 + 2]) * p.scale, __uint_as_float(o[i * 8 + 3]) * p.scale);
          w.z = pack_bf16x2(__uint_as_float(o[i * 8 + 4]) * p.scale, __uint_as_float(o[i * 8 + 5]) * p.scale);
          w.w = pack_bf16x2(__uint_as_float(o[i * 8 + 6]) * p.scale, __uint_as_float(o[i * 8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(op + c * 32 + i * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// Returns JG_ERR_UNSUPPORTED when the shape does not qualify (the caller falls back to the mma.sync kernel).
int launch_attn_fwd_tc(const void* qkv, int ldqkv, void* out, int ldo, float* lse, int N, int T, int heads, int ch,
                       int hstride, int koff, int voff, float scale_log2, cudaStream_t stream) {
  if (!(ch == 32 || ch == 64) || T % 256 != 0 || ldqkv % 8 != 0 || ldo % 8 != 0) return JG_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return JG_ERR_UNSUPPORTED;
  CUtensorMap tm;
  {
    // channels x rows.  dims[0] is the LOGICAL width (3 * heads * ch), not the row stride: the operand may be a channel
    // slice of a wider buffer, and a 64-channel box of the last head must end in the TMA unit's zero fill, not in the
    // neighbour's channels (or, on the last row, past the allocation)
    uint64_t dims[2] = {(uint64_t)3 * heads * ch, (uint64_t)N * T};
    uint64_t strides[1] = {(uint64_t)ldqkv * 2};
    uint32_t box[2] = {64, 128};
    uint32_t es[2] = {1, 1};
    int rc = make_tmap_bf16(&tm, qkv, 2, dims, strides, box, es);
    if (rc) return rc;
  }
  AttnTcParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.ldo = ldo; p.T = T; p.heads = heads; p.hstride = hstride; p.koff = koff; p.voff = voff;
  p.scale_log2 = scale_log2;
  const int smem = (2 + 2 * kKvStages + 4) * kTileBytes + 16 * 8 + 1024;
  dim3 grid(T / 256, N * heads);
  if (ch == 32) {
    static bool attr = false;
    if (!attr) {
      JG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    attn_fwd_tc_kernel<32><<<grid, kTcThreads, smem, stream>>>(tm, p);
  } else {
    static bool attr = false;
    if (!attr) {
      JG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr = true;
    }
    attn_fwd_tc_kernel<64><<<grid, kTcThreads, smem, stream>>>(tm, p);
  }
  JG_LAUNCH_CHECK();
  return JG_OK;
}

// Backward (dq, dk, dv into dqkv) after attn_bwd_prep_kernel has filled D; JG_ERR_UNSUPPORTED -> mma.sync kernels.
int launch_attn_bwd_tc(const void* qkv, int ldqkv, const void* d_out, int lddo, const float* lse, const float* D,
                       void* dqkv, int lddqkv, int N, int T, int heads, int ch, int hstride, int koff, int voff,
                       float scale_log2, float scale, cudaStream_t stream) {
  if (!(ch == 32 || ch == 64) || T % 128 != 0 || ldqkv % 8 != 0 || lddo % 8 != 0 || lddqkv % 8 != 0)
    return JG_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv) & 15))
    return JG_ERR_UNSUPPORTED;
  CUtensorMap tmQ, tmD;
  {
    uint64_t dims[2] = {(uint64_t)3 * heads * ch, (uint64_t)N * T};   // logical widths (see the forward)
    uint64_t strides[1] = {(uint64_t)ldqkv * 2};
    uint32_t box[2] = {64, 128};
    uint32_t es[2] = {1, 1};
    int rc = make_tmap_bf16(&tmQ, qkv, 2, dims, strides, box, es);
    if (rc) return rc;
    uint64_t dims2[2] = {(uint64_t)heads * ch, (uint64_t)N * T};
    uint64_t strides2[1] = {(uint64_t)lddo * 2};
    rc = make_tmap_bf16(&tmD, d_out, 2, dims2, strides2, box, es);
    if (rc) return rc;
  }
  AttnBwdParams p;
  p.lse = lse; p.D = D; p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.lddqkv = lddqkv; p.T = T; p.heads = heads; p.hstride = hstride; p.koff = koff; p.voff = voff;
  p.scale_log2 = scale_log2; p.scale = scale;
  const int smem_kv = 10 * kTileBytes + 2 * 256 * 4 + 16 * 8 + 1024;
  const int smem_q = 8 * kTileBytes + 16 * 8 + 1024;
  dim3 grid(T / 128, N * heads);
  static bool attr = false;
  if (!attr) {
    JG_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    JG_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
    JG_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    JG_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
    attr = true;
  }
  if (ch == 32) {
    attn_bwd_dq_tc_kernel<32><<<grid, kBwdThreads, smem_q, stream>>>(tmQ, tmD, p);
    attn_bwd_dkv_tc_kernel<32><<<grid, kBwdThreads, smem_kv, stream>>>(tmQ, tmD, p);
  } else {
    attn_bwd_dq_tc_kernel<64><<<grid, kBwdThreads, smem_q, stream>>>(tmQ, tmD, p);
    attn_bwd_dkv_tc_kernel<64><<<grid, kBwdThreads, smem_kv, stream>>>(tmQ, tmD, p);
  }
  JG_LAUNCH_CHECK();
  return JG_OK;
}

}  // namespace jg
