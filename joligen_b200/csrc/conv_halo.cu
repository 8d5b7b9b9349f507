// 3x3 (generally RxS, stride 1) implicit-GEMM convolution with shared-memory HALO REUSE across taps.
//
// conv_igemm.cu stages the A operand once per (tap, 64-channel slab): the same input patch is pulled
// from L2 R*S times.  Here ONE 4-D TMA box of (TH+R-1) x (TW+S-1) pixels x 64 channels is staged per
// slab, and the R*S tap operands are shifted WINDOWS of that patch: the UMMA descriptor for tap (r,s)
// starts at byte ((r*PW + s) * 128) of the patch with SBO = PW*128 (8-pixel row groups of a TW = 8 wide
// tile are PW = TW+S-1 pixels apart).  tools/umma_probe.cu established on B200 that tcgen05.mma applies
// the 128B swizzle to absolute shared-memory address bits, so a window whose start is only 128-byte
// aligned and whose 8-row groups are 1280 bytes apart reads exactly what TMA wrote (base_offset = 0).
// L2->SM traffic for A drops by R*S*128/180 = 6.4x for 3x3.
//
// B (weights): streamed per (slab, tap) through its own smem ring, or — when all R*S*ceil(Cin/64) tiles
// fit (Cin = 64 layers, the 256^2-level convs) — loaded ONCE per CTA and kept resident.
//
// Same warp roles / TMEM double buffering / epilogue as conv_igemm.cu.
#include "conv_common.cuh"

#include <stdlib.h>

namespace jg {

constexpr int kHaloTW = 8, kHaloTH = 16;

struct HaloParams {
  ConvFwdParams c;  // TW/TH/TN = 8/16/1
  int R, PW, PH;    // filter rows, patch width / height in pixels
  int a_stage_bytes;
};

// KS: compile-time filter size (3 = 3x3: the tap loop is unrolled, descriptor offsets become immediates); 0 = runtime R, S.
template <int BLOCK_N, int SA, int SB, bool B_RESIDENT, int KS>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                 const HaloParams hp) {
  const ConvFwdParams& p = hp.c;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  constexpr bool TMA_STORE = BLOCK_N >= 64;  // output tiles leave through shared memory + TMA (conv_common.cuh)
  // Residual tiles arrive through TMA as well when the tile is one 64-channel slab: per-lane 64-byte residual loads
  // made the 64-channel 256^2 conv2 layers 80% slower than the same conv without a residual (L1/LSU bound).
  constexpr bool RES_TMA = BLOCK_N == 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + SA * hp.a_stage_bytes;
  const int k_slabs = p.kc_blocks;
  const int b_tiles = B_RESIDENT ? p.RS * k_slabs : SB;
  uint8_t* stage = smB + static_cast<size_t>(b_tiles) * B_BYTES;  // 2 x 16 KB output staging (1024-byte aligned)
  uint8_t* res_stage = stage + (TMA_STORE ? 2 * kStageBytes : 0);  // 2 x 16 KB residual tiles (RES_TMA)
  uint64_t* bars = reinterpret_cast<uint64_t*>(res_stage + ((RES_TMA && p.res) ? 2 * kStageBytes : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + SA;
  uint64_t* b_full = bars + 2 * SA;           // SB entries (entry 0 only when resident)
  uint64_t* b_empty = bars + 2 * SA + SB;
  uint64_t* tfull = bars + 2 * SA + 2 * SB;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rfull + 2);
  float* s_bias = reinterpret_cast<float*>(rfull + 4);
  float* s_acc = s_bias + ((p.Cout + 64 + 31) / 32) * 32;  // [Cout][2] GroupNorm sums of the current image (fused modes)
  conv_stage_bias(p, s_bias);
  if (p.stats || p.gn_sums)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) s_acc[i] = 0.f;
  const TileRange tr = conv_tile_range(p);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMA_STORE) tma_prefetch_desc(&tmY);
    if (p.res && (RES_TMA || (BLOCK_N >= 128 && p.res_mode == 0))) tma_prefetch_desc(&tmR);
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
      mbar_init(&tempty[i], 8);
      mbar_init(&rfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      if (B_RESIDENT) {
        // all weight tiles of this CTA's (single) N tile, once
        mbar_arrive_expect_tx(&b_full[0], static_cast<uint32_t>(p.RS * k_slabs * B_BYTES));
        for (int kc = 0; kc < k_slabs; ++kc)
          for (int tap = 0; tap < p.RS; ++tap)
            tma_load_3d(smB + (kc * p.RS + tap) * B_BYTES, &tmB, &b_full[0], kc * 64, tap, 0);
      }
      for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * kHaloTW - p.pad;
        const int h0 = th * kHaloTH - p.pad;
        for (int kc = 0; kc < k_slabs; ++kc) {
          mbar_wait(&a_empty[sa], pha ^ 1);
          mbar_arrive_expect_tx(&a_full[sa], static_cast<uint32_t>(hp.PW * hp.PH * 128));
          tma_load_4d(smA + sa * hp.a_stage_bytes, &tmA, &a_full[sa], kc * 64, w0, h0, tn);
          if (++sa == SA) {
            sa = 0;
            pha ^= 1;
          }
          if (!B_RESIDENT) {
            for (int tap = 0; tap < p.RS; ++tap) {
              mbar_wait(&b_empty[sb], phb ^ 1);
              mbar_arrive_expect_tx(&b_full[sb], B_BYTES);
              tma_load_3d(smB + sb * B_BYTES, &tmB, &b_full[sb], kc * 64, tap, n_tile * BLOCK_N);
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
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    const uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
    const uint32_t sbo_a = static_cast<uint32_t>(hp.PW * 128);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smA), 16, sbo_a);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smB), 16, 1024);
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (B_RESIDENT) {
      mbar_wait(&b_full[0], 0);
      tc_fence_after();
    }
    const uint32_t row_skip16 = static_cast<uint32_t>((hp.PW - p.S) * 128) >> 4;  // window origin step at a filter-row end
    for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      uint64_t b_res = b_desc0;  // resident weights: tiles are consecutive in (slab, tap) order
      for (int kc = 0; kc < k_slabs; ++kc) {
        mbar_wait(&a_full[sa], pha);
        tc_fence_after();
        // One elected lane issues the whole slab (R*S taps x 4 k-steps).  The tensor pipe accepts only a couple of MMAs
        // ahead, so every instruction between two tcgen05.mma of the issuing thread is potential pipe idle time:
        // the tap loop is unrolled for 3x3 and nothing but the weight-tile barrier wait sits between the MMAs.
        if (elect_one()) {
          uint64_t a_desc = a_desc0 + (static_cast<uint32_t>(sa * hp.a_stage_bytes) >> 4);
          uint64_t b_desc = b_res;
          int sbl = sb;
          uint32_t phl = phb;
          auto tap_mmas = [&](int tap) {
            if (!B_RESIDENT) {
              mbar_wait(&b_full[sbl], phl);
              tc_fence_after();
              b_desc = b_desc0 + (static_cast<uint32_t>(sbl * B_BYTES) >> 4);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
            if (B_RESIDENT) {
              b_desc += B_BYTES >> 4;
            } else {
              umma_commit(&b_empty[sbl]);
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
                a_desc += 8;  // one pixel (128 B) to the right
              }
              a_desc += row_skip16;  // start of the next filter row
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
          umma_commit(&a_empty[sa]);
        }
        __syncwarp();
        // ring state, advanced identically by all lanes
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
      if (elect_one()) umma_commit(&tfull[acc]);
      __syncwarp();
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = threadIdx.x == 64;  // warp 2, lane 0
    EpiPrefetch pf;
    int acc = 0;
    int stage_idx = 0;
    uint32_t acc_phase = 0;
    // (res_mode 1: the tile is the GroupNorm input x of the fused backward sums, not a residual)
    const bool res_tma = RES_TMA && p.res != nullptr;
    // BLOCK_N >= 128: residual slabs travel through the output staging buffers (ResInplace, conv_common.cuh)
    const bool res_inplace = TMA_STORE && BLOCK_N >= 128 && p.res != nullptr && p.res_mode == 0 && !(p.dbg & 4);
    ResInplace rin{&tmR, rfull, 0u, false, 0, 0, 0, 0};
    auto tile_coords = [&](int tile, int& co, int& a1, int& a2, int& a3) {
      const int m_tile = tile / p.n_tiles;
      co = (tile % p.n_tiles) * BLOCK_N;
      a1 = (m_tile % p.tiles_w) * kHaloTW;
      a2 = ((m_tile / p.tiles_w) % p.tiles_h) * kHaloTH;
      a3 = m_tile / (p.tiles_w * p.tiles_h);
    };
    if (res_inplace && issuer && tr.begin < tr.end) {
      int co, a1, a2, a3;
      tile_coords(tr.begin, co, a1, a2, a3);
      mbar_arrive_expect_tx(&rfull[0], kStageBytes);
      tma_load_4d(stage, &tmR, &rfull[0], co, a1, a2, a3);
      conv_res_prefetch_rest<BLOCK_N>(p, &tmR, co, a1, a2, a3);
    }
    // residual tile of this CTA's (i)th tile -> res_stage[i & 1]; issued two tiles ahead by the issuer thread
    auto load_res_tile = [&](int tile, int buf) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      mbar_arrive_expect_tx(&rfull[buf], kStageBytes);
      tma_load_4d(res_stage + buf * kStageBytes, &tmR, &rfull[buf], n_tile * BLOCK_N, tw * kHaloTW, th * kHaloTH, tn);
    };
    if (res_tma && issuer) {
      if (tr.begin < tr.end) load_res_tile(tr.begin, 0);
      if (tr.begin + tr.step < tr.end) load_res_tile(tr.begin + tr.step, 1);
    }
    int rbuf = 0;
    uint32_t rphase = 0;
    const bool fused = TMA_STORE && (p.stats || p.gn_sums);
    int cur_img = -1;
    for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      const int pw = tw * kHaloTW + (row % kHaloTW);
      const int ph = th * kHaloTH + (row / kHaloTW);
      const bool valid = (pw < p.Wo) && (ph < p.Ho);
      const size_t pix = (static_cast<size_t>(tn) * p.Ho + ph) * p.Wo + pw;
      if (!res_tma && !res_inplace) conv_epilogue_prefetch<BLOCK_N>(p, pf, half, n_tile, valid, pix);
      if (res_inplace) {
        rin.has_next = tile + tr.step < tr.end;
        if (rin.has_next) tile_coords(tile + tr.step, rin.nco, rin.n1, rin.n2, rin.n3);
      }
      if (fused && tn != cur_img) {  // the chunk of tiles moved on to the next image: ship the finished one's sums
        if (cur_img >= 0) conv_flush_sums(p, s_acc, cur_img);
        cur_img = tn;
      }

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if constexpr (TMA_STORE) {
        if (res_tma) mbar_wait(&rfull[rbuf], rphase);
        conv_epilogue_tile_tma<BLOCK_N, false>(p, pf, p.bias ? s_bias : nullptr, tmem_base + acc * BLOCK_N, q, half,
                                               n_tile, valid, pix, stage, stage_idx, &tmY, tw * kHaloTW, th * kHaloTH, tn, issuer,
                                        res_tma ? res_stage + rbuf * kStageBytes : nullptr, kHaloTW, s_acc,
                                        res_inplace ? &rin : nullptr);
        if (res_tma) {
          // the slab barrier inside the epilogue ordered every thread's reads of this residual buffer before here
          if (issuer && tile + 2 * tr.step < tr.end) load_res_tile(tile + 2 * tr.step, rbuf);
          if (++rbuf == 2) {
            rbuf = 0;
            rphase ^= 1;
          }
        }
      } else
        conv_epilogue_tile<BLOCK_N>(p, pf, p.bias ? s_bias : nullptr, tmem_base + acc * BLOCK_N, q, half, n_tile,
                                    valid, pix);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (fused && cur_img >= 0) conv_flush_sums(p, s_acc, cur_img);
    if (TMA_STORE && issuer) bulk_wait_read<0>();  // the staging buffers live until the last store has read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int SA, int SB, bool B_RESIDENT>
static int halo_smem_bytes(const HaloParams& hp) {
  const int b_tiles = B_RESIDENT ? hp.c.RS * hp.c.kc_blocks : SB;
  return SA * hp.a_stage_bytes + b_tiles * BLOCK_N * 128 + (BLOCK_N >= 64 ? 2 * kStageBytes : 0) +
         ((BLOCK_N == 64 && hp.c.res) ? 2 * kStageBytes : 0) + (2 * SA + 2 * SB + 8) * 8 +
         (((hp.c.Cout + 64) * 4 + 127) / 128) * 128 + ((hp.c.stats || hp.c.gn_sums) ? 2 * hp.c.Cout * 4 : 0) + 1024;
}

template <int BLOCK_N, int SA, int SB, bool B_RESIDENT, int KS>
static int launch_halo_ks(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                          const HaloParams& hp, cudaStream_t stream) {
  const int smem = halo_smem_bytes<BLOCK_N, SA, SB, B_RESIDENT>(hp);
  JG_CHECK(smem <= 232448, JG_ERR_INVALID, "conv_halo: smem %d too large", smem);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    JG_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N, SA, SB, B_RESIDENT, KS>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int grid = hp.c.total_tiles < num_sms() ? hp.c.total_tiles : num_sms();
  conv_halo_kernel<BLOCK_N, SA, SB, B_RESIDENT, KS><<<grid, kThreads, smem, stream>>>(tmA, tmB, tmY, tmR, hp);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

template <int BLOCK_N, int SA, int SB, bool B_RESIDENT>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                       const HaloParams& hp, cudaStream_t stream) {
  if (hp.R == 3 && hp.c.S == 3) return launch_halo_ks<BLOCK_N, SA, SB, B_RESIDENT, 3>(tmA, tmB, tmY, tmR, hp, stream);
  return launch_halo_ks<BLOCK_N, SA, SB, B_RESIDENT, 0>(tmA, tmB, tmY, tmR, hp, stream);
}

int launch_conv_halo(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                     const float* bias, const void* residual_in, void* y, cudaStream_t stream, bool* fused) {
  // qualification: stride 1, a real spatial filter, "same"-style geometry handled generally via pad,
  // tiles of 8 x 16 output pixels must tile the output exactly (ragged sizes go to the generic kernel)
  if (d->stride != 1 || d->R * d->S == 1 || d->R > 5 || d->S > 5) return JG_ERR_UNSUPPORTED;
  if (d->Wo % kHaloTW != 0 || d->Ho % kHaloTH != 0) return JG_ERR_UNSUPPORTED;
  HaloParams hp{};
  ConvFwdParams& p = hp.c;
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
  p.RS = d->R * d->S; p.S = d->S; p.pad = d->pad; p.stride = 1;
  p.TW = kHaloTW; p.TH = kHaloTH; p.TN = 1;
  p.tiles_w = d->Wo / kHaloTW;
  p.tiles_h = d->Ho / kHaloTH;
  p.tiles_n = d->N;
  const int block_n = d->Cout > 128 ? 256 : d->Cout > 64 ? 128 : d->Cout > 32 ? 64 : 32;
  p.n_tiles = ceil_div(d->Cout, block_n);
  p.kc_blocks = ceil_div(d->Cin, 64);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  p.ldy = d->ldy; p.ldres = d->ldres; p.act = d->act; p.res_scale = d->res_scale;
  p.bias = bias;
  // fused GroupNorm work needs the TMA-store epilogue (Cout > 32); otherwise the caller runs the stand-alone reduction
  const bool fuse = e != nullptr && block_n >= 64 && d->Cout <= kMaxFusedCout;
  const void* residual = conv_apply_epilogue(p, d, fuse ? e : nullptr, residual_in);
  if (fused) *fused = fuse;
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  hp.R = d->R;
  hp.PW = kHaloTW + d->S - 1;
  hp.PH = kHaloTH + d->R - 1;
  hp.a_stage_bytes = (hp.PW * hp.PH * 128 + 1023) / 1024 * 1024;

  CUtensorMap tmA, tmB;
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
    uint32_t box[3] = {64, 1, (uint32_t)block_n};
    uint32_t es[3] = {1, 1, 1};
    rc = make_tmap_bf16(&tmB, w_packed, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  CUtensorMap tmY = tmA;  // only read by the kernels that store through TMA (block_n >= 64)
  if (block_n >= 64) {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldy * 2, (uint64_t)d->Wo * d->ldy * 2, (uint64_t)d->Ho * d->Wo * d->ldy * 2};
    uint32_t box[4] = {64, (uint32_t)kHaloTW, (uint32_t)kHaloTH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmY, y, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  CUtensorMap tmR = tmA;  // residual tiles (64-channel kernels with a residual)
  if (block_n >= 64 && residual && (p.res_mode == 0 || block_n == 64)) {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)p.ldres * 2, (uint64_t)d->Wo * p.ldres * 2,
                           (uint64_t)d->Ho * d->Wo * p.ldres * 2};
    uint32_t box[4] = {64, (uint32_t)kHaloTW, (uint32_t)kHaloTH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmR, residual, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  // resident weights: one N tile and all (slab, tap) tiles fit next to the A ring and the staging buffers (the
  // residual staging exists only when there is a residual; two A stages are enough when a slab's 36 MMAs cover the
  // load latency, which lets the 128 -> 64 layers at 256^2 keep their 147 KB of weights resident too)
  const bool resident = p.n_tiles == 1 && (block_n == 64 ? halo_smem_bytes<64, 3, 1, true>(hp)
                                                         : halo_smem_bytes<32, 4, 1, true>(hp)) <= 232448;
  const bool resident2 = p.n_tiles == 1 && block_n == 64 && halo_smem_bytes<64, 2, 1, true>(hp) <= 232448;
  switch (block_n) {
    case 256: return launch_halo<256, 2, 4, false>(tmA, tmB, tmY, tmR, hp, stream);
    case 128: return launch_halo<128, 3, 6, false>(tmA, tmB, tmY, tmR, hp, stream);
    case 64:
      if (resident) return launch_halo<64, 3, 1, true>(tmA, tmB, tmY, tmR, hp, stream);
      if (resident2) return launch_halo<64, 2, 1, true>(tmA, tmB, tmY, tmR, hp, stream);
      return launch_halo<64, 3, 7, false>(tmA, tmB, tmY, tmR, hp, stream);
    default:
      return resident ? launch_halo<32, 4, 1, true>(tmA, tmB, tmY, tmR, hp, stream)
                      : launch_halo<32, 4, 8, false>(tmA, tmB, tmY, tmR, hp, stream);
  }
}


// ------------------------------------------------------------------------------------------------
// wgrad with halo reuse
// ------------------------------------------------------------------------------------------------
// dW[tap][ci][co] += sum_pixels X[pixel + tap][ci] * dY[pixel][co]
// Work item = (64-channel ci block, 64-channel co block, pixel range).  Per 64-pixel k-block (an 8x8 patch)
// ONE X halo patch ((8+S-1) x (8+R-1) pixels x 64 ci) and ONE dY tile (64 pixels x 64 co) are staged; the
// R*S tap operands are windows of the halo (MN-major, SBO = PW*128).  Two taps share one M = 128 MMA: the
// second 64 rows of A are the next tap's window (LBO = distance between the two window origins), so a 3x3
// filter needs 5 MMA groups x N = 64 -> 320 TMEM columns; 9/10 of the issued MMA work is useful.
// Accumulator layout: fp32 [R*S][Cin][Cout] (HWIO), split-K partials added with red.global.add.v4.f32.
struct WgradHaloParams {
  int Cin, Cout, RS, S, pad;
  int PW, PH, x_stage_bytes;
  int tiles_w, tiles_h, pix_blocks;
  int cib, cob;
  int tap0, ntaps, npairs;  // this launch covers taps [tap0, tap0 + ntaps) as npairs = ceil(ntaps / 2) MMA groups
  int ksplit, kb_per_split, total_items;
  float* acc;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// NCO = output-channel block = N of the MMA (64 or 128).  tools/umma_bench.cu: an M128 x N64 MMA is shared-memory
// bound (48 cycles instead of 32: 6 KB of operands per MMA), N = 128 runs at the full tensor rate — but 5 tap
// pairs x 128 columns exceed the 512 TMEM columns, so with NCO = 128 a 3x3 filter is covered by two launches
// (taps 0..7 as 4 pairs, then tap 8), while NCO = 64 covers all 9 taps in one.
template <int STAGES, int NCO>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const WgradHaloParams p) {
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int DY_BYTES = 8192 * (NCO / 64);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = DY_BYTES + p.x_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(STAGES) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // item -> (ci block, co block, split): split fastest
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
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
          mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(DY_BYTES + p.PW * p.PH * 128));
          uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
#pragma unroll
          for (int h = 0; h < NCO / 64; ++h)
            tma_load_4d(st + h * 8192, &tmDY, &full[stage], cob * NCO + h * 64, tw * 8, th * 8, tn);
          tma_load_4d(st + DY_BYTES, &tmX, &full[stage], cib * 64, tw * 8 - p.pad, th * 8 - p.pad, tn);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, NCO, 1, 1);
    const uint32_t sbo_x = static_cast<uint32_t>(p.PW * 128);
    const uint64_t dy_desc0 = make_smem_desc_sw128(smem_u32(smem), 8192, 1024);
    // X-window descriptor of stage 0 / window origin 0 with LBO = 0; per tap pair only the start address (low
    // bits) and the LBO field (bits 16..29 = distance between the two taps' window origins) differ: those
    // deltas are precomputed so that the issue loop is a handful of 64-bit adds per MMA.
    const uint64_t x_desc0 = make_smem_desc_sw128(smem_u32(smem) + DY_BYTES, 0, sbo_x);
    uint64_t pair_delta[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t0 = p.tap0 + 2 * j;
      const int t1 = (2 * j + 1 < p.ntaps) ? t0 + 1 : t0;
      const int off0 = (t0 / p.S) * p.PW + (t0 % p.S);
      const int off1 = (t1 / p.S) * p.PW + (t1 % p.S);
      pair_delta[j] = static_cast<uint64_t>(off0 * 8) | (static_cast<uint64_t>(((off1 - off0) * 8) & 0x3FFF) << 16);
    }
    const uint32_t kstep_x = (2 * sbo_x) >> 4;  // 16 pixels = two PW-pixel rows of the halo, in 16-byte units
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
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
          for (int j = 0; j < 8; ++j) {
            if (j < p.npairs) {
              const uint64_t a_desc = x_desc + pair_delta[j];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + j * NCO, a_desc + k * kstep_x, dy_desc + k * 128, idesc,
                          (first | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
      acc_phase ^= 1;
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int split = item % p.ksplit;
      const int cob = (item / p.ksplit) % p.cob;
      const int cib = item / (p.ksplit * p.cob);
      const bool has_work = split * p.kb_per_split < p.pix_blocks;
      const int ci = cib * 64 + (row & 63);
      const int co0 = cob * NCO;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      for (int j = 0; j < p.npairs; ++j) {
        const int tl = 2 * j + (row >> 6);  // tap index within this launch
        const int tap = p.tap0 + tl;
        const bool ok = has_work && tl < p.ntaps && ci < p.Cin;
        float* dst = p.acc + (static_cast<size_t>(tap) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
        for (int c = 0; c < NCO; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * NCO + c, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (co0 + c + g * 4 < p.Cout)
                red_add_v4(dst + c + g * 4, __uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                           __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// acc [RS][Cin][Cout] (HWIO) -> OIHW
__global__ void unpack_hwio_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int RS,
                                   float beta) {
  const long long total = (long long)Cout * Cin * RS;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tap = (int)(i % RS);
  const int ci = (int)((i / RS) % Cin);
  const int co = (int)(i / ((long long)RS * Cin));
  const float v = src[((long long)tap * Cin + ci) * Cout + co];
  dst[i] = beta == 0.f ? v : beta * dst[i] + v;
}

int launch_unpack_hwio(const float* src, float* dst, int Cout, int Cin, int RS, float beta, cudaStream_t stream) {
  const long long total = (long long)Cout * Cin * RS;
  unpack_hwio_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, Cout, Cin, RS, beta);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

template <int STAGES, int NCO>
static int launch_wgrad_halo_one(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradHaloParams& p,
                                 cudaStream_t stream) {
  const int smem = STAGES * (8192 * (NCO / 64) + p.x_stage_bytes) + (2 * STAGES + 2) * 8 + 16 + 1024;
  JG_CHECK(smem <= 232448, JG_ERR_INVALID, "wgrad_halo: smem %d too large", smem);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    JG_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<STAGES, NCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int grid = p.total_items < num_sms() ? p.total_items : num_sms();
  wgrad_halo_kernel<STAGES, NCO><<<grid, kWgradThreads, smem, stream>>>(tmDY, tmX, p);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

int launch_wgrad_halo(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* ws, float* dw_oihw,
                      float beta, cudaStream_t stream) {
  if (d->stride != 1 || d->R * d->S == 1 || d->R * d->S > 32 || d->R > 5 || d->S > 5) return JG_ERR_UNSUPPORTED;
  if (d->Wo % 8 != 0 || d->Ho % 8 != 0) return JG_ERR_UNSUPPORTED;
  static const bool force64 = getenv("JG_WGRAD_N64") != nullptr;
  // N = 128 runs the MMA at full rate (N = 64 is shared-memory bound at 2/3 rate) but holds only 8 taps in TMEM, so a 3x3
  // filter takes two passes over the activations; past 9 taps the extra passes cost more than the MMA rate buys.
  const int nco = (d->Cout >= 128 && d->R * d->S <= 9 && !force64) ? 128 : 64;
  // 512 TMEM columns / nco columns per pair * 2 taps per pair.  With NCO = 128 a 3x3 filter needs two passes over the
  // activations; 6 + 3 taps (3 + 2 MMA pairs) instead of 8 + 1 keeps the second pass from being a pure L2 stream
  // (28.8 KB of operands per 256 MMA cycles per SM is twice what the L2 sustains chip-wide).
  // Measured (profiles/r01_wgrad_taps.log): 6 + 3 is 4-10 % faster than 8 + 1 on every 3x3 layer with Cout >= 128.
  const int taps_per_launch = nco == 128 ? (d->R * d->S == 9 ? 6 : 8) : 16;
  WgradHaloParams p{};
  p.Cin = d->Cin; p.Cout = d->Cout; p.RS = d->R * d->S; p.S = d->S; p.pad = d->pad;
  p.PW = 8 + d->S - 1;
  p.PH = 8 + d->R - 1;
  p.x_stage_bytes = (p.PW * p.PH * 128 + 1023) / 1024 * 1024;
  p.tiles_w = d->Wo / 8;
  p.tiles_h = d->Ho / 8;
  p.pix_blocks = p.tiles_w * p.tiles_h * d->N;
  p.cib = ceil_div(d->Cin, 64);
  // work items of one pass: (ci block, co block of `n` channels, pixel range); split-K over the pixel blocks so that
  // every SM gets an item where possible
  auto plan = [&](int n) {
    p.cob = ceil_div(d->Cout, n);
    const int pairs = p.cib * p.cob;
    int ksplit = num_sms() / pairs;
    if (ksplit < 1) ksplit = 1;
    const int max_split = p.pix_blocks / 8 > 0 ? p.pix_blocks / 8 : 1;
    if (ksplit > max_split) ksplit = max_split;
    p.kb_per_split = ceil_div(p.pix_blocks, ksplit);
    p.ksplit = ceil_div(p.pix_blocks, p.kb_per_split);
    p.total_items = pairs * p.ksplit;
  };
  p.acc = ws;
  // dw_oihw == nullptr: raw mode (jg_conv2d_wgrad_acc): the caller owns a persistent [R*S][Cin][Cout] accumulator that
  // is neither zeroed nor unpacked here
  if (dw_oihw) JG_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)p.RS * d->Cin * d->Cout, stream));

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
    uint32_t box[4] = {64, (uint32_t)p.PW, (uint32_t)p.PH, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmX, x, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  for (int t0 = 0; t0 < p.RS; t0 += taps_per_launch) {
    p.tap0 = t0;
    p.ntaps = p.RS - t0 < taps_per_launch ? p.RS - t0 : taps_per_launch;
    p.npairs = (p.ntaps + 1) / 2;
    // The short second pass of a 3x3 filter (3 taps = 2 MMA pairs) is L2-bound: 28.8 KB of operands per 512 MMA cycles
    // per SM.  With 256 output channels per item the same X patch serves twice the MMA work (44.8 KB per 1 024 cycles,
    // 2 pairs x 256 TMEM columns), i.e. 22 % less L2 traffic for that pass: +2..8 % on the Cout >= 256 layers
    // (profiles/r01_wgrad_n256.log).
    if (nco == 128 && p.npairs <= 2 && d->Cout >= 256) {
      plan(256);
      rc = launch_wgrad_halo_one<4, 256>(tmDY, tmX, p, stream);
    } else if (nco == 128) {
      plan(128);
      rc = launch_wgrad_halo_one<7, 128>(tmDY, tmX, p, stream);
    } else {
      plan(64);
      rc = launch_wgrad_halo_one<8, 64>(tmDY, tmX, p, stream);
    }
    if (rc) return rc;
  }
  if (dw_oihw) {
    const long long total = (long long)d->Cout * d->Cin * p.RS;
    unpack_hwio_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(ws, dw_oihw, d->Cout, d->Cin, p.RS, beta);
    JG_LAUNCH_CHECK();
  }
  return JG_OK;
}

}  // namespace jg
