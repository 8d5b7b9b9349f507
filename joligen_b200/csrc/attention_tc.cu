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
    uint32_t ph = 0;
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      // pass 1 over the 128 scores of this row (TMEM reads are cheap; keeping all 128 in registers is not): the max
      float mx;
      {
        uint32_t v[32];
        tmem_ld_32x32(tS[g] + lane_off, v);
        tmem_ld_wait();
        mx = __uint_as_float(v[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
#pragma unroll
        for (int c = 1; c < 4; ++c) {
          tmem_ld_32x32(tS[g] + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        mx *= p.scale_log2;  // into the exp2 domain (scale > 0)
      }
      // P_g and O_g of block j-1 must be consumed / final before they are touched
      if (j > 0) {
        mbar_wait(&pv_done[g], ph ^ 1);
        tc_fence_after();
      }
      const bool grow = mx > m_used + kRescaleThreshold;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_used;
        const float c = (j > 0) ? fast_exp2(m_used - m_new) : 0.f;
        if (j > 0) {
          uint32_t o[32];
          tmem_ld_32x32(tO[g] + lane_off, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * c);
          tmem_st_32x32(tO[g] + lane_off, o);
          if (HD == 64) {
            tmem_ld_32x32(tO[g] + lane_off + 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * c);
            tmem_st_32x32(tO[g] + lane_off + 32, o);
          }
          tmem_st_wait();
        }
        l *= c;
        m_used = m_new;
      }
      // pass 2: p = 2^(s * scale - m), row sum, bf16 P into the swizzled K-major tile
      float rs = 0.f;
      const float nm = -m_used;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tS[g] + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {  // 8 keys per 16-byte chunk
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = fast_exp2(fmaf(__uint_as_float(v[i4 * 8 + i]), p.scale_log2, nm));
            rs += e[i];
          }
          uint4 o;
          o.x = pack_bf16x2(e[0], e[1]);
          o.y = pack_bf16x2(e[2], e[3]);
          o.z = pack_bf16x2(e[4], e[5]);
          o.w = pack_bf16x2(e[6], e[7]);
          const int c8 = c * 4 + i4;
          sts_v4(p_a + (c8 >> 3) * kTileBytes + (((c8 & 7) ^ swz) << 4), o);
        }
      }
      l += rs;
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      ph ^= 1;
    }
    // epilogue: O / l -> bf16, lse
    mbar_wait(&pv_done[g], ph ^ 1);
    tc_fence_after();
    const float inv = 1.f / l;
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

}  // namespace

// Returns JG_ERR_UNSUPPORTED when the shape does not qualify (the caller falls back to the mma.sync kernel).
int launch_attn_fwd_tc(const void* qkv, int ldqkv, void* out, int ldo, float* lse, int N, int T, int heads, int ch,
                       int hstride, int koff, int voff, float scale_log2, cudaStream_t stream) {
  if (!(ch == 32 || ch == 64) || T % 256 != 0 || ldqkv % 8 != 0 || ldo % 8 != 0) return JG_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return JG_ERR_UNSUPPORTED;
  CUtensorMap tm;
  {
    // channels x rows; a box may reach past the last channel of the buffer row (zero fill) but never past ldqkv's
    // logical width: dims[0] is the row width in channels
    uint64_t dims[2] = {(uint64_t)ldqkv, (uint64_t)N * T};
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

}  // namespace jg
