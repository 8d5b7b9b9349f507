// CTA-PAIR variant of the halo-reuse 3x3 implicit GEMM (conv_halo.cu): tcgen05.mma.cta_group::2, M = 256.
//
// Why: ncu on conv_halo_kernel<128, ...> (profiles/r02_gn_fusion_ab.md) shows the single-CTA kernel bound by the
// shared-memory port while the tensor pipe is busy — per M128 x N x K16 MMA it reads A (4 KB) AND the whole weight
// tile B (N x 32 B) from shared memory, and every CTA streams every weight tile from L2 by TMA.  A pair of CTAs on
// the two SMs of a TPC computes two horizontally adjacent 8x16-pixel output tiles with ONE instruction stream:
//   * CTA r stages its own halo patch (A rows [128 r, 128 r + 128)) and HALF of each weight tile (rows
//     [N/2 r, N/2 r + N/2)) in its own shared memory: B traffic per CTA halves, both from L2 (TMA) and from shared
//     memory (operand reads);
//   * the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 (SASS UTCHMMA.2CTA); the accumulator rows of a
//     CTA's pixels live in that CTA's TMEM, so the epilogue (bias / residual / activation / TMA store / fused
//     GroupNorm statistics: conv_common.cuh) is unchanged and local.
// Barrier protocol (tools/umma2cta_probe.cu established the pieces on B200):
//   a_full / b_full   the LEADER's barriers; both CTAs' TMA loads (.cta_group::2) complete_tx on them, the leader's
//                     producer expects the bytes of both;
//   a_empty / b_empty / tfull   one per CTA at the same shared-memory offset; tcgen05.commit multicasts to both;
//   tempty            the leader's; the peer's epilogue warps arrive remotely (mapa + mbarrier.arrive.cluster).
#include "conv_common.cuh"

#include <stdlib.h>

namespace jg {

constexpr int kH2TW = 8, kH2TH = 16;

struct Halo2Params {
  ConvFwdParams c;  // tile geometry of ONE CTA (8 x 16 pixels); total_tiles counts PAIR tiles
  int R, PW, PH;
  int a_stage_bytes;
  int pair_w;       // pair tiles per image row (= tiles_w / 2)
};

// B_RESIDENT: all R*S*ceil(Cin/64) half weight tiles of the (single) N tile stay in shared memory for the whole kernel
// (SB is then unused): with the tile split over the pair, 128 -> 128 3x3 (2 x 9 x 8 KB = 144 KB per CTA) fits, which no
// single CTA can hold — its weights are then read from L2 once per CTA instead of once per output tile.
template <int BLOCK_N, int SA, int SB, int KS, bool B_RESIDENT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                  const Halo2Params hp) {
  const ConvFwdParams& p = hp.c;
  constexpr int B_HALF = BLOCK_N / 2;          // weight rows staged by one CTA
  constexpr int B_BYTES = B_HALF * 128;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  constexpr bool RES_TMA = BLOCK_N == 64;
  static_assert(BLOCK_N >= 64, "pair kernel: TMA-store epilogue only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + SA * hp.a_stage_bytes;
  const int b_tiles = B_RESIDENT ? p.RS * p.kc_blocks : SB;
  uint8_t* stage = smB + static_cast<size_t>(b_tiles) * B_BYTES;
  uint8_t* res_stage = stage + 2 * kStageBytes;
  const bool res_tma = RES_TMA && p.res != nullptr && p.res_mode == 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(res_stage + (res_tma ? 2 * kStageBytes : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + SA;
  uint64_t* b_full = bars + 2 * SA;
  uint64_t* b_empty = bars + 2 * SA + SB;
  uint64_t* tfull = bars + 2 * SA + 2 * SB;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rfull + 2);
  float* s_bias = reinterpret_cast<float*>(rfull + 4);
  float* s_acc = s_bias + ((p.Cout + 64 + 31) / 32) * 32;
  conv_stage_bias(p, s_bias);
  if (p.stats || p.gn_sums)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) s_acc[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (res_tma || (BLOCK_N >= 128 && p.res && p.res_mode == 0)) tma_prefetch_desc(&tmR);
    for (int i = 0; i < SA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);  // 8 epilogue warps of each CTA of the pair (the leader's copy is the one waited on)
      mbar_init(&rfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // pair tile -> (n_tile, pair column, tile row, image); this CTA's tile is column 2*pc + rank
  auto decode = [&](int pt, int& n_tile, int& tw, int& th, int& tn) {
    n_tile = pt % p.n_tiles;
    const int pm = pt / p.n_tiles;
    tw = (pm % hp.pair_w) * 2 + static_cast<int>(rank);
    th = (pm / hp.pair_w) % p.tiles_h;
    tn = pm / (hp.pair_w * p.tiles_h);
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; all bytes are counted on the leader's barriers) =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      const uint32_t a_bytes = static_cast<uint32_t>(hp.PW * hp.PH * 128);
      if (B_RESIDENT) {
        // every (slab, tap) half tile of this CTA, once; all bytes of the pair are counted on the leader's b_full[0]
        if (rank == 0) mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(2 * p.RS * p.kc_blocks * B_BYTES));
        const uint32_t lead = mapa_u32(smem_u32(&b_full[0]), 0);
        for (int kc = 0; kc < p.kc_blocks; ++kc)
          for (int tap = 0; tap < p.RS; ++tap)
            tma_load_3d_2cta(smB + (kc * p.RS + tap) * B_BYTES, &tmB, lead, kc * 64, tap, static_cast<int>(rank) * B_HALF);
      }
      for (int pt = pair_id; pt < p.total_tiles; pt += num_pairs) {
        int n_tile, tw, th, tn;
        decode(pt, n_tile, tw, th, tn);
        const int w0 = tw * kH2TW - p.pad;
        const int h0 = th * kH2TH - p.pad;
        for (int kc = 0; kc < p.kc_blocks; ++kc) {
          mbar_wait(&a_empty[sa], pha ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * a_bytes);
          tma_load_4d_2cta(smA + sa * hp.a_stage_bytes, &tmA, mapa_u32(smem_u32(&a_full[sa]), 0), kc * 64, w0, h0, tn);
          if (++sa == SA) {
            sa = 0;
            pha ^= 1;
          }
          if (!B_RESIDENT) {
            for (int tap = 0; tap < p.RS; ++tap) {
              mbar_wait(&b_empty[sb], phb ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2 * B_BYTES);
              tma_load_3d_2cta(smB + sb * B_BYTES, &tmB, mapa_u32(smem_u32(&b_full[sb]), 0), kc * 64, tap,
                               n_tile * BLOCK_N + static_cast<int>(rank) * B_HALF);
              if (++sb == SB) {
                sb = 0;
                phb ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader only =====================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);
      const uint32_t sbo_a = static_cast<uint32_t>(hp.PW * 128);
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smA), 16, sbo_a);
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smB), 16, 1024);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t row_skip16 = static_cast<uint32_t>((hp.PW - p.S) * 128) >> 4;
      if (B_RESIDENT) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      for (int pt = pair_id; pt < p.total_tiles; pt += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint64_t b_res = b_desc0;  // resident weights: tiles are consecutive in (slab, tap) order
        for (int kc = 0; kc < p.kc_blocks; ++kc) {
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          if (elect_one()) {
            uint64_t a_desc = a_desc0 + (static_cast<uint32_t>(sa * hp.a_stage_bytes) >> 4);
            int sbl = sb;
            uint32_t phl = phb;
            uint64_t b_desc = b_res;
            auto tap_mmas = [&](int tap) {
              if (!B_RESIDENT) {
                mbar_wait(&b_full[sbl], phl);
                tc_fence_after();
                b_desc = b_desc0 + (static_cast<uint32_t>(sbl * B_BYTES) >> 4);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              if (B_RESIDENT) {
                b_desc += B_BYTES >> 4;
              } else {
                umma_commit_2cta(&b_empty[sbl], 0x3);
                if (++sbl == SB) {
                  sbl = 0;
                  phl ^= 1;
                }
              }
            };
            if (KS > 0) {
#pragma unroll
              for (int r = 0; r < KS; ++r) {
#pragma unroll
                for (int c = 0; c < KS; ++c) {
                  tap_mmas(r * KS + c);
                  a_desc += 8;
                }
                a_desc += row_skip16;
              }
            } else {
              int sx = 0;
              for (int tap = 0; tap < p.RS; ++tap) {
                tap_mmas(tap);
                a_desc += 8;
                if (++sx == p.S) {
                  sx = 0;
                  a_desc += row_skip16;
                }
              }
            }
            umma_commit_2cta(&a_empty[sa], 0x3);
          }
          __syncwarp();
          if (B_RESIDENT) {
            b_res += static_cast<uint64_t>(p.RS) * (B_BYTES >> 4);
          } else {
            const int adv = sb + p.RS;
            phb ^= static_cast<uint32_t>(adv / SB) & 1u;
            sb = adv % SB;
          }
          if (++sa == SA) {
            sa = 0;
            pha ^= 1;
          }
        }
        if (elect_one()) umma_commit_2cta(&tfull[acc], 0x3);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of BOTH CTAs: each CTA owns its 128 accumulator lanes) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = threadIdx.x == 64;
    EpiPrefetch pf;
    int acc = 0;
    int stage_idx = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_lead[2] = {mapa_u32(smem_u32(&tempty[0]), 0), mapa_u32(smem_u32(&tempty[1]), 0)};
    const bool res_inplace = BLOCK_N >= 128 && p.res != nullptr && p.res_mode == 0 && !(p.dbg & 4);
    ResInplace rin{&tmR, rfull, 0u, false, 0, 0, 0, 0};
    if (res_inplace && issuer && pair_id < p.total_tiles) {
      int n_tile, tw, th, tn;
      decode(pair_id, n_tile, tw, th, tn);
      mbar_arrive_expect_tx(&rfull[0], kStageBytes);
      tma_load_4d(stage, &tmR, &rfull[0], n_tile * BLOCK_N, tw * kH2TW, th * kH2TH, tn);
      conv_res_prefetch_rest<BLOCK_N>(p, &tmR, n_tile * BLOCK_N, tw * kH2TW, th * kH2TH, tn);
    }
    auto load_res_tile = [&](int pt, int buf) {
      int n_tile, tw, th, tn;
      decode(pt, n_tile, tw, th, tn);
      mbar_arrive_expect_tx(&rfull[buf], kStageBytes);
      tma_load_4d(res_stage + buf * kStageBytes, &tmR, &rfull[buf], n_tile * BLOCK_N, tw * kH2TW, th * kH2TH, tn);
    };
    if (res_tma && issuer) {
      if (pair_id < p.total_tiles) load_res_tile(pair_id, 0);
      if (pair_id + num_pairs < p.total_tiles) load_res_tile(pair_id + num_pairs, 1);
    }
    int rbuf = 0;
    uint32_t rphase = 0;
    const bool fused = p.stats || p.gn_sums;
    int cur_img = -1;
    for (int pt = pair_id; pt < p.total_tiles; pt += num_pairs) {
      int n_tile, tw, th, tn;
      decode(pt, n_tile, tw, th, tn);
      const int pw = tw * kH2TW + (row % kH2TW);
      const int ph = th * kH2TH + (row / kH2TW);
      const bool valid = (pw < p.Wo) && (ph < p.Ho);
      const size_t pix = (static_cast<size_t>(tn) * p.Ho + ph) * p.Wo + pw;
      if (!res_tma && !res_inplace) conv_epilogue_prefetch<BLOCK_N>(p, pf, half, n_tile, valid, pix);
      if (res_inplace) {
        rin.has_next = pt + num_pairs < p.total_tiles;
        if (rin.has_next) {
          int n2, tw2, th2, tn2;
          decode(pt + num_pairs, n2, tw2, th2, tn2);
          rin.nco = n2 * BLOCK_N; rin.n1 = tw2 * kH2TW; rin.n2 = th2 * kH2TH; rin.n3 = tn2;
        }
      }
      if (fused && tn != cur_img) {
        if (cur_img >= 0) conv_flush_sums(p, s_acc, cur_img);
        cur_img = tn;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (res_tma) mbar_wait(&rfull[rbuf], rphase);
      conv_epilogue_tile_tma<BLOCK_N, false>(p, pf, p.bias ? s_bias : nullptr, tmem_base + acc * BLOCK_N, q, half,
                                             n_tile, valid, pix, stage, stage_idx, &tmY, tw * kH2TW, th * kH2TH, tn, issuer,
                                             res_tma ? res_stage + rbuf * kStageBytes : nullptr, kH2TW, s_acc,
                                             res_inplace ? &rin : nullptr);
      if (res_tma) {
        if (issuer && pt + 2 * num_pairs < p.total_tiles) load_res_tile(pt + 2 * num_pairs, rbuf);
        if (++rbuf == 2) {
          rbuf = 0;
          rphase ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // "this warp has read its accumulator lanes": the hazard is TMEM only (the leader's next MMAs overwrite the
        // peer's accumulator), ordered by tcgen05.wait::ld + the fence above — no memory release needed.  With
        // .release.cluster ncu showed 8.7 % of all stall samples on the MEMBAR.ALL / ERRBAR pair in front of it.
        if (rank == 0) mbar_arrive(&tempty[acc]);
        else mbar_arrive_cluster_relaxed(tempty_lead[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (fused && cur_img >= 0) conv_flush_sums(p, s_acc, cur_img);
    if (issuer) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // nobody leaves (and frees TMEM / shared memory) while the peer may still read or signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int SA, int SB, bool B_RESIDENT>
static int halo2_smem_bytes(const Halo2Params& hp) {
  const int b_tiles = B_RESIDENT ? hp.c.RS * hp.c.kc_blocks : SB;
  return SA * hp.a_stage_bytes + b_tiles * (BLOCK_N / 2) * 128 + 2 * kStageBytes +
         ((BLOCK_N == 64 && hp.c.res && hp.c.res_mode == 0) ? 2 * kStageBytes : 0) + (2 * SA + 2 * SB + 8) * 8 +
         (((hp.c.Cout + 64) * 4 + 127) / 128) * 128 + ((hp.c.stats || hp.c.gn_sums) ? 2 * hp.c.Cout * 4 : 0) + 1024;
}

template <int BLOCK_N, int SA, int SB, int KS, bool B_RESIDENT>
static int launch_halo2_ks(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                           const Halo2Params& hp, cudaStream_t stream) {
  const int smem = halo2_smem_bytes<BLOCK_N, SA, SB, B_RESIDENT>(hp);
  JG_CHECK(smem <= 232448, JG_ERR_INVALID, "conv_halo2: smem %d too large", smem);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    JG_CUDA(cudaFuncSetAttribute(conv_halo2_kernel<BLOCK_N, SA, SB, KS, B_RESIDENT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  int pairs = num_sms() / 2;
  if (pairs > hp.c.total_tiles) pairs = hp.c.total_tiles;
  conv_halo2_kernel<BLOCK_N, SA, SB, KS, B_RESIDENT><<<2 * pairs, kThreads, smem, stream>>>(tmA, tmB, tmY, tmR, hp);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

template <int BLOCK_N, int SA, int SB, bool B_RESIDENT>
static int launch_halo2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                        const Halo2Params& hp, cudaStream_t stream) {
  if (hp.R == 3 && hp.c.S == 3) return launch_halo2_ks<BLOCK_N, SA, SB, 3, B_RESIDENT>(tmA, tmB, tmY, tmR, hp, stream);
  return launch_halo2_ks<BLOCK_N, SA, SB, 0, B_RESIDENT>(tmA, tmB, tmY, tmR, hp, stream);
}

// Returns JG_ERR_UNSUPPORTED when the shape does not qualify (the caller falls back to the single-CTA kernels).
int launch_conv_halo2(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                      const float* bias, const void* residual_in, void* y, cudaStream_t stream, bool* fused) {
  if (d->stride != 1 || d->R * d->S == 1 || d->R > 5 || d->S > 5) return JG_ERR_UNSUPPORTED;
  // N = 64 tiles stay on the single-CTA kernel: an M256 x N64 pair MMA is only ~32 tensor cycles long and the
  // cta_group::2 issue / completion overhead dominates (measured: 64 -> 64 @256^2 983 -> 568 TF/s as a pair)
  if (d->Wo % (2 * kH2TW) != 0 || d->Ho % kH2TH != 0 || d->Cout <= 64) return JG_ERR_UNSUPPORTED;
  // one 64-channel slab of K (Cin <= 64: the dgrads of the 64-channel convolutions): too little MMA work per tile
  // for the pair protocol to pay (measured in the step: 64 -> 128 @256^2 +7 %, 64 -> 192 +18 % slower as pairs)
  static const int min_cin = getenv("JG_PAIR_MIN_CIN") ? atoi(getenv("JG_PAIR_MIN_CIN")) : 65;
  if (d->Cin < min_cin) return JG_ERR_UNSUPPORTED;
  Halo2Params hp{};
  ConvFwdParams& p = hp.c;
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
  p.RS = d->R * d->S; p.S = d->S; p.pad = d->pad; p.stride = 1;
  p.TW = kH2TW; p.TH = kH2TH; p.TN = 1;
  p.tiles_w = d->Wo / kH2TW;
  p.tiles_h = d->Ho / kH2TH;
  p.tiles_n = d->N;
  hp.pair_w = p.tiles_w / 2;
  const int block_n = d->Cout > 128 ? 256 : d->Cout > 64 ? 128 : 64;
  // the pair splits the weight tile in halves of block_n / 2 rows: the channel padding of a ragged last tile must
  // not leave a half without a valid row start (TMA clips rows >= Cout; a wholly out-of-range half is still legal)
  p.n_tiles = ceil_div(d->Cout, block_n);
  p.kc_blocks = ceil_div(d->Cin, 64);
  p.total_tiles = hp.pair_w * p.tiles_h * p.tiles_n * p.n_tiles;  // PAIR tiles
  p.ldy = d->ldy; p.ldres = d->ldres; p.act = d->act; p.res_scale = d->res_scale;
  p.bias = bias;
  const bool fuse = e != nullptr && d->Cout <= kMaxFusedCout;
  const void* residual = conv_apply_epilogue(p, d, fuse ? e : nullptr, residual_in);
  p.dbg |= 2;  // the pair kernel keeps its round-robin tile order (sums are flushed whenever the image changes)
  if (fused) *fused = fuse;
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  hp.R = d->R;
  hp.PW = kH2TW + d->S - 1;
  hp.PH = kH2TH + d->R - 1;
  hp.a_stage_bytes = (hp.PW * hp.PH * 128 + 1023) / 1024 * 1024;

  CUtensorMap tmA, tmB, tmY;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldx * 2, (uint64_t)d->W * d->ldx * 2, (uint64_t)d->H * d->W * d->ldx * 2};
    uint32_t box[4] = {64, (uint32_t)hp.PW, (uint32_t)hp.PH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmA, x, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    const int cin8 = (d->Cin + 7) / 8 * 8;
    uint64_t dims[3] = {(uint64_t)cin8, (uint64_t)p.RS, (uint64_t)d->Cout};
    uint64_t strides[2] = {(uint64_t)cin8 * 2, (uint64_t)p.RS * cin8 * 2};
    uint32_t box[3] = {64, 1, (uint32_t)(block_n / 2)};
    uint32_t es[3] = {1, 1, 1};
    rc = make_tmap_bf16(&tmB, w_packed, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldy * 2, (uint64_t)d->Wo * d->ldy * 2, (uint64_t)d->Ho * d->Wo * d->ldy * 2};
    uint32_t box[4] = {64, (uint32_t)kH2TW, (uint32_t)kH2TH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmY, y, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  CUtensorMap tmR = tmA;
  if (block_n >= 64 && residual && p.res_mode == 0) {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)p.ldres * 2, (uint64_t)d->Wo * p.ldres * 2, (uint64_t)d->Ho * d->Wo * p.ldres * 2};
    uint32_t box[4] = {64, (uint32_t)kH2TW, (uint32_t)kH2TH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmR, residual, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  // resident weights when the pair can hold every (slab, tap) half tile of its single N tile next to the A ring
  const bool one = p.n_tiles == 1;
  switch (block_n) {
    case 256:
      return launch_halo2<256, 3, 6, false>(tmA, tmB, tmY, tmR, hp, stream);
    case 128:
      if (one && halo2_smem_bytes<128, 3, 1, true>(hp) <= 232448) return launch_halo2<128, 3, 1, true>(tmA, tmB, tmY, tmR, hp, stream);
      if (one && halo2_smem_bytes<128, 2, 1, true>(hp) <= 232448) return launch_halo2<128, 2, 1, true>(tmA, tmB, tmY, tmR, hp, stream);
      return launch_halo2<128, 3, 12, false>(tmA, tmB, tmY, tmR, hp, stream);
    default:
      if (one && halo2_smem_bytes<64, 3, 1, true>(hp) <= 232448) return launch_halo2<64, 3, 1, true>(tmA, tmB, tmY, tmR, hp, stream);
      return launch_halo2<64, 3, 9, false>(tmA, tmB, tmY, tmR, hp, stream);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on CTA pairs (3x3 filters)
// ------------------------------------------------------------------------------------------------
// dW[tap][ci][co] += sum_pixels X[pixel + tap][ci] * dY[pixel][co]        (cf. wgrad_halo_kernel, conv_halo.cu)
// Single-CTA limit this kernel removes: with a 128-channel co block (the MMA shape that runs at the full tensor rate)
// a 3x3 filter needs 5 tap-pair accumulators x 128 columns = 640 > 512 TMEM columns, i.e. TWO passes over the
// activations, the short second one L2-bound (profiles/r01_wgrad_taps.log).  A CTA pair has 2 x 512 columns.
//
// One tcgen05.mma.cta_group::2 carries ONE A descriptor that each CTA applies to its own shared memory.  So the two
// CTAs stage the SAME 10 x 10 pixel halo-patch shape, but CTA 1's patch starts ONE IMAGE ROW LOWER: the window that is
// filter tap (r, s) in CTA 0 is tap (r + 1, s) in CTA 1.  Per 64-pixel k-block, three M = 256 groups
//     group     CTA 0 rows 0..63 | 64..127        CTA 1 rows 0..63 | 64..127      window origin, second-window distance
//       0       (0,0)   (0,1)                      (1,0)   (1,1)                   0,  1 pixel
//       1       (0,2)   (1,2)                      (1,2)*  (2,2)                   2,  10 pixels (one patch row)
//       2       (2,0)   (2,1)                      (3,0)-  (3,1)-                  20, 1 pixel
//   (* computed twice, CTA 1's copy is dropped; - outside the filter, dropped) cover all 9 taps in ONE pass with
//   3 x 128 = 384 TMEM columns per CTA.  Each CTA stages HALF of the dY tile (64 of the 128 output channels: the B
//   operand of cta_group::2 is split over the pair) — 20.8 KB per k-block and CTA instead of 28.8 KB.
// Work item = (64-channel ci block, 128-channel co block, pixel range); split-K over pixel blocks; fp32
// red.global.add.v4 into the [R*S][Cin][Cout] accumulator (layout code 0, like wgrad_halo_kernel).
struct WgradHalo2Params {
  int Cin, Cout, pad;
  int x_stage_bytes;
  int tiles_w, tiles_h, pix_blocks;
  int cib, cob;
  int ksplit, kb_per_split, total_items;
  float* acc;
};

constexpr int kW2PW = 10, kW2PH = 10;  // 8 x 8 pixel k-block + the 3x3 halo

template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgradThreads, 1)
wgrad_halo2_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                   const WgradHalo2Params p) {
  constexpr int NCO = 128;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int DY_BYTES = 8192;  // 64 pixels x 64 channels (this CTA's half of the 128-channel dY tile)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = DY_BYTES + p.x_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(STAGES) * stage_bytes);
  uint64_t* full = bars;            // the leader's are waited on
  uint64_t* empty = bars + STAGES;  // per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 1;     // the leader's: 4 epilogue warps of each CTA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = static_cast<uint32_t>(DY_BYTES + kW2PW * kW2PH * 128);
      for (int item = pair_id; item < p.total_items; item += num_pairs) {
        const int split = item % p.ksplit;
        const int cob = (item / p.ksplit) % p.cob;
        const int cib = item / (p.ksplit * p.cob);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.pix_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tn = kb / (p.tiles_w * p.tiles_h);
          mbar_wait(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * bytes);
          const uint32_t lead = mapa_u32(smem_u32(&full[stage]), 0);
          uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
          tma_load_4d_2cta(st, &tmDY, lead, cob * NCO + static_cast<int>(rank) * 64, tw * 8, th * 8, tn);
          // CTA 1's patch starts one image row lower (see the header): same shape, same shared-memory offsets
          tma_load_4d_2cta(st + DY_BYTES, &tmX, lead, cib * 64, tw * 8 - p.pad, th * 8 - p.pad + static_cast<int>(rank), tn);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, NCO, 1, 1);
      const uint32_t sbo_x = static_cast<uint32_t>(kW2PW * 128);
      const uint64_t dy_desc0 = make_smem_desc_sw128(smem_u32(smem), 8192, 1024);
      const uint64_t x_desc0 = make_smem_desc_sw128(smem_u32(smem) + DY_BYTES, 0, sbo_x);
      // (window origin, distance to the second window) of the three groups, in pixels of the patch
      const int g_off[3] = {0, 2, 2 * kW2PW};
      const int g_lbo[3] = {1, kW2PW, 1};
      uint64_t group_delta[3];
#pragma unroll
      for (int g = 0; g < 3; ++g)
        group_delta[g] = static_cast<uint64_t>(g_off[g] * 8) | (static_cast<uint64_t>((g_lbo[g] * 8) & 0x3FFF) << 16);
      const uint32_t kstep_x = (2 * sbo_x) >> 4;  // 16 pixels = two patch rows
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = pair_id; item < p.total_items; item += num_pairs) {
        const int split = item % p.ksplit;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.pix_blocks);
        mbar_wait(tempty, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t st16 = static_cast<uint32_t>(stage * stage_bytes) >> 4;
          const uint64_t dy_desc = dy_desc0 + st16;
          const uint64_t x_desc = x_desc0 + st16;
          const uint32_t first = kb > kb0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              const uint64_t a_desc = x_desc + group_delta[g];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2cta(tmem_base + g * NCO, a_desc + k * kstep_x, dy_desc + k * 128, idesc,
                               (first | k) != 0 ? 1u : 0u);
            }
            umma_commit_2cta(&empty[stage], 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_2cta(tfull, 0x3);
        __syncwarp();
        acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t acc_phase = 0;
    const uint32_t tempty_lead = mapa_u32(smem_u32(tempty), 0);
    // filter tap (r*3 + s) held by (group, cluster rank, row half); -1 = dropped (duplicate / outside the filter)
    const int tap_of[3][2][2] = {{{0, 1}, {3, 4}}, {{2, 5}, {-1, 8}}, {{6, 7}, {-1, -1}}};
    for (int item = pair_id; item < p.total_items; item += num_pairs) {
      const int split = item % p.ksplit;
      const int cob = (item / p.ksplit) % p.cob;
      const int cib = item / (p.ksplit * p.cob);
      const bool has_work = split * p.kb_per_split < p.pix_blocks;
      const int ci = cib * 64 + (row & 63);
      const int co0 = cob * NCO;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int tap = tap_of[g][rank][row >> 6];
        const bool ok = has_work && tap >= 0 && ci < p.Cin;
        // (warp-uniform: rows 0..63 / 64..127 are whole warps) skip the TMEM reads of dropped windows altogether
        if (tap_of[g][rank][0] < 0 && tap_of[g][rank][1] < 0) continue;
        float* dst = p.acc + (static_cast<size_t>(tap < 0 ? 0 : tap) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
        for (int c = 0; c < NCO; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * NCO + c, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int gg = 0; gg < 8; ++gg) {
              if (co0 + c + gg * 4 < p.Cout)
                red_add_v4f(dst + c + gg * 4, __uint_as_float(v[gg * 4]), __uint_as_float(v[gg * 4 + 1]),
                            __uint_as_float(v[gg * 4 + 2]), __uint_as_float(v[gg * 4 + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty);
        else mbar_arrive_cluster_relaxed(tempty_lead);
      }
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// Raw accumulation only (jg_conv2d_wgrad_acc semantics: acc is the caller's [9][Cin][Cout] fp32 accumulator, neither
// zeroed nor unpacked here).  JG_ERR_UNSUPPORTED when the shape does not qualify.
int launch_wgrad_halo2(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* acc,
                       cudaStream_t stream) {
  if (d->stride != 1 || d->R != 3 || d->S != 3 || d->pad != 1) return JG_ERR_UNSUPPORTED;
  if (d->Wo % 8 != 0 || d->Ho % 8 != 0 || d->Cout < 128) return JG_ERR_UNSUPPORTED;
  WgradHalo2Params p{};
  p.Cin = d->Cin; p.Cout = d->Cout; p.pad = d->pad;
  p.x_stage_bytes = (kW2PW * kW2PH * 128 + 1023) / 1024 * 1024;
  p.tiles_w = d->Wo / 8;
  p.tiles_h = d->Ho / 8;
  p.pix_blocks = p.tiles_w * p.tiles_h * d->N;
  p.cib = ceil_div(d->Cin, 64);
  p.cob = ceil_div(d->Cout, 128);
  const int items = p.cib * p.cob;
  const int pairs = num_sms() / 2;
  int ksplit = pairs / items;
  if (ksplit < 1) ksplit = 1;
  const int max_split = p.pix_blocks / 8 > 0 ? p.pix_blocks / 8 : 1;
  if (ksplit > max_split) ksplit = max_split;
  p.kb_per_split = ceil_div(p.pix_blocks, ksplit);
  p.ksplit = ceil_div(p.pix_blocks, p.kb_per_split);
  p.total_items = items * p.ksplit;
  p.acc = acc;

  CUtensorMap tmDY, tmX;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)lddy * 2, (uint64_t)d->Wo * lddy * 2, (uint64_t)d->Ho * d->Wo * lddy * 2};
    uint32_t box[4] = {64, 8, 8, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmDY, dy, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldx * 2, (uint64_t)d->W * d->ldx * 2, (uint64_t)d->H * d->W * d->ldx * 2};
    uint32_t box[4] = {64, (uint32_t)kW2PW, (uint32_t)kW2PH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmX, x, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  constexpr int STAGES = 9;
  const int smem = STAGES * (8192 + p.x_stage_bytes) + (2 * STAGES + 2) * 8 + 16 + 1024;
  JG_CHECK(smem <= 232448, JG_ERR_INVALID, "wgrad_halo2: smem %d too large", smem);
  static bool attr_done = false;
  if (!attr_done) {
    JG_CUDA(cudaFuncSetAttribute(wgrad_halo2_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  int grid_pairs = pairs < p.total_items ? pairs : p.total_items;
  wgrad_halo2_kernel<STAGES><<<2 * grid_pairs, kWgradThreads, smem, stream>>>(tmDY, tmX, p);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

}  // namespace jg
