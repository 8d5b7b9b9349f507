// Implicit-GEMM convolution on tcgen05 tensor cores for sm_100a.
//
//   forward / dgrad :  Y[pixel][co] = sum_{tap,ci} X[pixel + tap][ci] * Wp[co][tap][ci]
//       A (M = 128 output pixels = a TWxTHxTN patch) is staged by ONE 4-D TMA box load per
//       (tap, 64-channel slice): the box origin is shifted by the tap offset, out-of-bounds
//       coordinates (the zero padding, ragged tiles, channel tails) are zero-filled by the TMA
//       unit, and the smem image is a K-major SWIZZLE_128B operand.  B = packed weights, K-major.
//   wgrad :  dW[co][tap][ci] = sum_pixel dY[pixel][co] * X[pixel + tap][ci]
//       Both operands are MN-major (the reduction index, pixels, is the smem row); 64-pixel
//       k-blocks; split-K over pixel blocks with fp32 atomics into an OHWI accumulator.
//
// Kernel shape (both): persistent CTAs, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected thread), warps 2..5 = epilogue (TMEM -> registers -> global).  smem ring of
// STAGES {A,B} tiles with full/empty mbarriers, TMEM accumulator double-buffered with
// tmem_full/tmem_empty mbarriers so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Replaces the cuDNN calls behind nn.Conv2d in the reference:
//   models/modules/unet_generator_attn/unet_generator_attn.py:186-190,208-220,481-483,639-643
#include "conv_common.cuh"

#include <stdlib.h>

namespace jg {

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmY, const ConvFwdParams p) {
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr bool TMA_STORE = BLOCK_N >= 64;  // output tiles leave through shared memory + TMA (conv_common.cuh)
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of 2");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * kABytes;
  uint8_t* stage_out = smB + STAGES * B_BYTES;  // 2 x 16 KB output staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + (TMA_STORE ? 2 * kStageBytes : 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* s_bias = reinterpret_cast<float*>(bars + 2 * STAGES + 6);
  float* s_acc = s_bias + kMaxCout + 64;  // [Cout <= kMaxFusedCout][2] GroupNorm sums of the current image
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
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
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

  const int k_blocks = p.RS * p.kc_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.TW * p.stride - p.pad;
        const int h0 = th * p.TH * p.stride - p.pad;
        const int n0 = tn * p.TN;
        for (int tap = 0; tap < p.RS; ++tap) {
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          for (int kc = 0; kc < p.kc_blocks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], kABytes + B_BYTES);
            tma_load_4d(smA + stage * kABytes, &tmA, &full[stage], kc * 64, w0 + s, h0 + r, n0);
            tma_load_3d(smB + stage * B_BYTES, &tmB, &full[stage], kc * 64, tap, n_tile * BLOCK_N);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers;
    // one elected lane issues tcgen05.mma / tcgen05.commit.
    const uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smA), 16, 1024);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smB), 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      // one elected lane issues the whole tile (see conv_halo.cu: instructions between two MMAs are pipe idle time)
      if (elect_one()) {
        int st = stage;
        uint32_t ph = phase;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint64_t a_desc = desc_advance(a_desc0, st * kABytes);
          const uint64_t b_desc = desc_advance(b_desc0, st * B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, desc_advance(a_desc, k * 32), desc_advance(b_desc, k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[st]);
          if (++st == STAGES) {
            st = 0;
            ph ^= 1;
          }
        }
      }
      __syncwarp();
      {
        const int adv = stage + k_blocks;
        phase ^= static_cast<uint32_t>(adv / STAGES) & 1u;
        stage = adv % STAGES;
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
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = threadIdx.x == 64;  // warp 2, lane 0
    EpiPrefetch pf;
    int acc = 0;
    int stage_idx = 0;
    uint32_t acc_phase = 0;
    const bool fused = TMA_STORE && (p.stats || p.gn_sums);
    int cur_img = -1;
    for (int tile = tr.begin; tile < tr.end; tile += tr.step) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      if (fused && tn != cur_img) {  // (TN == 1 in the fused modes: tn is the image)
        if (cur_img >= 0) conv_flush_sums(p, s_acc, cur_img);
        cur_img = tn;
      }
      const int pw = tw * p.TW + (row % p.TW);
      const int ph = th * p.TH + ((row / p.TW) % p.TH);
      const int pn = tn * p.TN + row / (p.TW * p.TH);
      const bool valid = (pw < p.Wo) && (ph < p.Ho) && (pn < p.N);
      const size_t pix = (static_cast<size_t>(pn) * p.Ho + ph) * p.Wo + pw;

      conv_epilogue_prefetch<BLOCK_N>(p, pf, half, n_tile, valid, pix);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if constexpr (TMA_STORE)
        conv_epilogue_tile_tma<BLOCK_N>(p, pf, p.bias ? s_bias : nullptr, tmem_base + acc * BLOCK_N, q, half, n_tile,
                                        valid, pix, stage_out, stage_idx, &tmY, tw * p.TW, th * p.TH, tn * p.TN,
                                        issuer, nullptr, p.TW, s_acc);  // (fused GroupNorm sums: TN == 1)
      else
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

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
struct ConvWgradParams {
  int Cin, Cout;
  int RS, S, pad, stride;
  int KW, KH, KN;  // 64-pixel k-block patch
  int tiles_w, tiles_h, tiles_n;
  int pix_blocks;  // tiles_w * tiles_h * tiles_n
  int cblocks;     // ceil(Cin / 64)
  int co_tiles;    // ceil(Cout / 128)
  int nb_tiles;    // RS * cblocks / NB
  int ksplit;
  int kb_per_split;
  int total_items;
  float* dw;  // [Cout][RS][Cin] fp32
};

template <int NB, int STAGES>
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const ConvWgradParams p) {
  constexpr int BLOCK_N = 64 * NB;
  constexpr int B_BYTES = NB * 8192;
  constexpr uint32_t TMEM_COLS = (NB == 1) ? 128 : (NB == 2) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + STAGES * B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
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

  // item -> (co_tile, nb_tile, split): split fastest so that CTAs working on the same output tile
  // run concurrently and share X / dY tiles through L2.
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int split = item % p.ksplit;
        const int nb_tile = (item / p.ksplit) % p.nb_tiles;
        const int co_tile = item / (p.ksplit * p.nb_tiles);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.pix_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tn = kb / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.KW, h0 = th * p.KH, n0 = tn * p.KN;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kABytes + B_BYTES);
          uint8_t* a = smA + stage * kABytes;
          tma_load_4d(a, &tmDY, &full[stage], co_tile * 128, w0, h0, n0);
          tma_load_4d(a + 8192, &tmDY, &full[stage], co_tile * 128 + 64, w0, h0, n0);
          uint8_t* b = smB + stage * B_BYTES;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const int bi = nb_tile * NB + j;
            const int tap = bi / p.cblocks;
            const int ci0 = (bi - tap * p.cblocks) * 64;
            const int r = tap / p.S;
            const int s = tap - r * p.S;
            tma_load_4d(b + j * 8192, &tmX, &full[stage], ci0, w0 * p.stride + s - p.pad,
                        h0 * p.stride + r - p.pad, n0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
    // MN-major: 16 K-rows (pixels) per MMA = 2048 B; LBO = next 64-channel block (8192 B), SBO = next
    // 8-pixel group (1024 B).
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smA), 8192, 1024);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smB), 8192, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int split = item % p.ksplit;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.pix_blocks);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = desc_advance(a_desc0, stage * kABytes);
        const uint64_t b_desc = desc_advance(b_desc0, stage * B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, desc_advance(a_desc, k * 2048), desc_advance(b_desc, k * 2048), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
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
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int split = item % p.ksplit;
      const int nb_tile = (item / p.ksplit) % p.nb_tiles;
      const int co_tile = item / (p.ksplit * p.nb_tiles);
      const int kb0 = split * p.kb_per_split;
      const bool has_work = kb0 < p.pix_blocks;
      const int co = co_tile * 128 + row;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c, v);
        tmem_ld_wait();
        const int bi = nb_tile * NB + c / 64;
        const int tap = bi / p.cblocks;
        const int ci0 = (bi - tap * p.cblocks) * 64 + (c & 32);
        if (has_work && co < p.Cout) {
          float* dst = p.dw + (static_cast<size_t>(co) * p.RS + tap) * p.Cin + ci0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (ci0 + j < p.Cin) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int pow2_le(int v, int cap) {
  int p = 1;
  while (p * 2 <= v && p * 2 <= cap) p *= 2;
  return p;
}

// Split `total` (a power of two) pixels-per-tile into a (tw, th, tn) patch of powers of two:
// widest run along W first (<= 16), then H, the remainder over images.  Ragged extents are
// handled by TMA zero fill on load and by masking in the epilogue.
static void pick_patch(int total, int Wo, int Ho, int* tw, int* th, int* tn) {
  const int w = pow2_le(Wo, total < 16 ? total : 16);
  const int h = pow2_le(Ho, total / w);
  *tw = w;
  *th = h;
  *tn = total / (w * h);
}

template <int BLOCK_N, int STAGES>
static int launch_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const ConvFwdParams& p,
                      cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + BLOCK_N * 128) + (BLOCK_N >= 64 ? 2 * kStageBytes : 0) +
                       (2 * STAGES + 6) * 8 + kMaxCout * 4 + 256 + 2 * kMaxFusedCout * 4 + 1024;
  static_assert(smem <= 232448, "conv_fwd_kernel: shared memory budget");
  static bool attr_done = false;
  if (!attr_done) {
    JG_CUDA(cudaFuncSetAttribute(conv_fwd_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem));
    attr_done = true;
  }
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  conv_fwd_kernel<BLOCK_N, STAGES><<<grid, kThreads, smem, stream>>>(tmA, tmB, tmY, p);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

template <int NB, int STAGES>
static int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const ConvWgradParams& p,
                        cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + NB * 8192) + (2 * STAGES + 4) * 8 + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    JG_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<NB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem));
    attr_done = true;
  }
  int grid = p.total_items < num_sms() ? p.total_items : num_sms();
  conv_wgrad_kernel<NB, STAGES><<<grid, kWgradThreads, smem, stream>>>(tmDY, tmX, p);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

static int check_desc(const jg_conv_desc* d, bool forward = true) {
  JG_CHECK(d != nullptr, JG_ERR_INVALID, "conv: null descriptor");
  JG_CHECK(d->N > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0, JG_ERR_INVALID, "conv: bad dims");
  JG_CHECK(d->Cin > 0 && d->Cin % 8 == 0 && d->ldx % 8 == 0 && d->ldx >= d->Cin, JG_ERR_INVALID,
           "conv: Cin=%d ldx=%d must be multiples of 8 with ldx >= Cin", d->Cin, d->ldx);
  JG_CHECK(d->Cout > 0 && d->Cout % 8 == 0 && d->ldy % 8 == 0 && d->ldy >= d->Cout, JG_ERR_INVALID,
           "conv: Cout=%d ldy=%d must be multiples of 8 with ldy >= Cout", d->Cout, d->ldy);
  JG_CHECK(d->R > 0 && d->S > 0 && d->R * d->S <= 64, JG_ERR_INVALID, "conv: bad filter %dx%d", d->R, d->S);
  // (the forward kernels stage the bias vector in shared memory; the weight gradient has no such buffer)
  JG_CHECK(!forward || d->Cout <= kMaxCout, JG_ERR_INVALID, "conv: Cout %d > %d (bias staging buffer)", d->Cout, kMaxCout);
  JG_CHECK(d->stride == 1 || d->stride == 2, JG_ERR_INVALID, "conv: stride %d unsupported", d->stride);
  JG_CHECK(d->up2x == 0, JG_ERR_INVALID, "conv: up2x is reserved");
  const int ho = (d->H + 2 * d->pad - d->R) / d->stride + 1;
  const int wo = (d->W + 2 * d->pad - d->S) / d->stride + 1;
  // Ho/Wo may be SMALLER than the full correlation size: a cropped output (transposed convolutions with
  // output_padding read one row/column less than the symmetric-padding formula yields)
  JG_CHECK(d->Ho <= ho && d->Wo <= wo && d->Ho > 0 && d->Wo > 0, JG_ERR_INVALID,
           "conv: output dims %dx%d exceed %dx%d", d->Ho, d->Wo, ho, wo);
  return JG_OK;
}

}  // namespace jg

using namespace jg;

extern "C" int jg_conv2d_fwd(const jg_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                             const void* residual, void* y, jg_stream_t stream_) {
  return jg_conv2d_fwd_ex(d, nullptr, x, w_packed, bias, residual, y, stream_);
}

// The fused-GroupNorm outputs that the launched kernel's epilogue did not produce, as stand-alone reductions over y.
static int conv_epilogue_fallback(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* y, cudaStream_t stream) {
  int rc = JG_OK;
  if (e->stats) rc = launch_chan_stats(y, d->ldy, d->N, d->Ho * d->Wo, d->Cout, e->stats, stream);
  if (rc == JG_OK && e->gn_sums)
    rc = launch_gn_bwd_sums(e->gn_x, e->ldgx, y, d->ldy, d->N, d->Ho * d->Wo, d->Cout, e->gn_ab, e->gn_act, e->gn_sums,
                            stream);
  return rc;
}

extern "C" int jg_conv2d_fwd_ex(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                                const float* bias, const void* residual_in, void* y, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_desc(d);
  if (rc) return rc;
  const void* residual = residual_in;
  if (e && !e->stats && !e->gn_sums) e = nullptr;
  if (e && e->gn_sums) {
    JG_CHECK(residual == nullptr, JG_ERR_INVALID, "conv_fwd_ex: gn_sums cannot be combined with a residual");
    JG_CHECK(e->gn_x && e->gn_ab && e->ldgx % 8 == 0 && e->ldgx >= d->Cout &&
                 (reinterpret_cast<uintptr_t>(e->gn_x) & 15) == 0,
             JG_ERR_INVALID, "conv_fwd_ex: gn_sums needs gn_x (16-byte aligned, ldgx >= Cout) and gn_ab");
    JG_CHECK(d->act == JG_ACT_NONE, JG_ERR_INVALID, "conv_fwd_ex: gn_sums expects a linear epilogue (a dgrad)");
  }
  JG_CHECK(x && w_packed && y, JG_ERR_INVALID, "conv_fwd: null pointer");
  JG_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
           JG_ERR_INVALID, "conv_fwd: pointers must be 16-byte aligned");
  JG_CHECK(residual == nullptr || (d->ldres % 8 == 0 && d->ldres >= d->Cout), JG_ERR_INVALID,
           "conv_fwd: bad ldres %d", d->ldres);
  JG_CHECK(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, JG_ERR_INVALID,
           "conv_fwd: bias must be 16-byte aligned");

  // stride-1 spatial filters on 8x16-tileable outputs: halo-reuse kernel (conv_halo.cu)
  static const bool no_halo = getenv("JG_NO_HALO") != nullptr;
  // CTA pairs (M = 256 across the two SMs of a TPC): JG_CONV_2CTA=0 turns them off
  static const bool pairs = getenv("JG_CONV_2CTA") == nullptr || atoi(getenv("JG_CONV_2CTA")) != 0;
  if (!no_halo && pairs) {
    bool fused = false;
    rc = launch_conv_halo2(d, e, x, w_packed, bias, residual, y, stream, &fused);
    if (rc != JG_ERR_UNSUPPORTED) {
      if (rc == JG_OK && e && !fused) rc = conv_epilogue_fallback(d, e, y, stream);
      return rc;
    }
  }
  if (!no_halo) {
    bool fused = false;
    rc = launch_conv_halo(d, e, x, w_packed, bias, residual, y, stream, &fused);
    if (rc != JG_ERR_UNSUPPORTED) {
      if (rc == JG_OK && e && !fused) rc = conv_epilogue_fallback(d, e, y, stream);
      return rc;
    }
  }

  ConvFwdParams p{};
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
  p.RS = d->R * d->S; p.S = d->S; p.pad = d->pad; p.stride = d->stride;
  pick_patch(128, d->Wo, d->Ho, &p.TW, &p.TH, &p.TN);
  p.tiles_w = ceil_div(d->Wo, p.TW);
  p.tiles_h = ceil_div(d->Ho, p.TH);
  p.tiles_n = ceil_div(d->N, p.TN);
  const int block_n = d->Cout > 128 ? 256 : d->Cout > 64 ? 128 : d->Cout > 32 ? 64 : 32;
  p.n_tiles = ceil_div(d->Cout, block_n);
  p.kc_blocks = ceil_div(d->Cin, 64);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  p.ldy = d->ldy; p.ldres = d->ldres; p.act = d->act; p.res_scale = d->res_scale;
  p.bias = bias;
  // fused GroupNorm work: TMA-store epilogue and one image per tile (per-(image, channel) sums)
  const bool fuse = e != nullptr && block_n >= 64 && p.TN == 1 && d->Cout <= kMaxFusedCout;
  residual = conv_apply_epilogue(p, d, fuse ? e : nullptr, residual);
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);

  // A: x as (C, W, H, N); with stride 2 the box walks every other pixel.
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldx * 2, (uint64_t)d->W * d->ldx * 2, (uint64_t)d->H * d->W * d->ldx * 2};
    uint32_t st = (uint32_t)d->stride;
    uint32_t box[4] = {64, (uint32_t)p.TW * st, (uint32_t)p.TH * st, (uint32_t)p.TN};
    uint32_t es[4] = {1, st, st, 1};
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
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmY, y, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  switch (block_n) {
    case 256: rc = launch_fwd<256, 3>(tmA, tmB, tmY, p, stream); break;
    case 128: rc = launch_fwd<128, 5>(tmA, tmB, tmY, p, stream); break;
    case 64: rc = launch_fwd<64, 7>(tmA, tmB, tmY, p, stream); break;
    default: rc = launch_fwd<32, 8>(tmA, tmB, tmY, p, stream); break;
  }
  if (rc == JG_OK && e && !fuse) rc = conv_epilogue_fallback(d, e, y, stream);
  return rc;
}

// Shared by jg_conv2d_wgrad (zero + accumulate + unpack into dw_oihw) and jg_conv2d_wgrad_acc (dw_oihw == nullptr:
// accumulate into the caller's persistent raw accumulator; *layout = 0: [R*S][Cin][Cout], 1: [Cout][R*S][Cin]).
static int conv2d_wgrad_impl(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* ws, float* dw_oihw,
                             float beta, int* layout, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_desc(d, /*forward=*/false);
  if (rc) return rc;
  JG_CHECK(x && dy && ws, JG_ERR_INVALID, "conv_wgrad: null pointer");
  JG_CHECK(lddy % 8 == 0 && lddy >= d->Cout, JG_ERR_INVALID, "conv_wgrad: bad lddy %d", lddy);
  static const bool no_halo = getenv("JG_NO_HALO") != nullptr;
  // 3x3, Cout >= 128: CTA pairs cover all 9 taps in one pass at the 128-channel MMA shape (JG_WGRAD_2CTA=0: off)
  static const bool wpairs = getenv("JG_WGRAD_2CTA") == nullptr || atoi(getenv("JG_WGRAD_2CTA")) != 0;
  if (!no_halo && wpairs) {
    if (dw_oihw) JG_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)d->R * d->S * d->Cin * d->Cout, stream));
    rc = launch_wgrad_halo2(d, x, dy, lddy, ws, stream);
    if (rc != JG_ERR_UNSUPPORTED) {
      if (layout) *layout = 0;
      if (rc == JG_OK && dw_oihw) rc = launch_unpack_hwio(ws, dw_oihw, d->Cout, d->Cin, d->R * d->S, beta, stream);
      return rc;
    }
  }
  if (!no_halo) {
    rc = launch_wgrad_halo(d, x, dy, lddy, ws, dw_oihw, beta, stream);
    if (rc != JG_ERR_UNSUPPORTED) {
      if (layout) *layout = 0;
      return rc;
    }
  }
  if (layout) *layout = 1;
  float* dw = ws;  // generic kernel: OHWI accumulator
  if (dw_oihw) JG_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)d->R * d->S * d->Cin * d->Cout, stream));

  ConvWgradParams p{};
  p.Cin = d->Cin; p.Cout = d->Cout; p.RS = d->R * d->S; p.S = d->S; p.pad = d->pad; p.stride = d->stride;
  pick_patch(64, d->Wo, d->Ho, &p.KW, &p.KH, &p.KN);
  p.tiles_w = ceil_div(d->Wo, p.KW);
  p.tiles_h = ceil_div(d->Ho, p.KH);
  p.tiles_n = ceil_div(d->N, p.KN);
  p.pix_blocks = p.tiles_w * p.tiles_h * p.tiles_n;
  p.cblocks = ceil_div(d->Cin, 64);
  p.co_tiles = ceil_div(d->Cout, 128);
  const int total_blocks = p.RS * p.cblocks;
  const int nb = (total_blocks % 4 == 0) ? 4 : (total_blocks % 3 == 0) ? 3 : (total_blocks % 2 == 0) ? 2 : 1;
  p.nb_tiles = total_blocks / nb;
  // split-K: aim for >= 2 waves of CTAs while keeping >= 8 k-blocks per item.
  const int out_tiles = p.co_tiles * p.nb_tiles;
  int ksplit = ceil_div(2 * num_sms(), out_tiles);
  const int max_split = p.pix_blocks / 8 > 0 ? p.pix_blocks / 8 : 1;
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  p.kb_per_split = ceil_div(p.pix_blocks, ksplit);
  p.ksplit = ceil_div(p.pix_blocks, p.kb_per_split);
  p.total_items = out_tiles * p.ksplit;
  p.dw = dw;

  CUtensorMap tmDY, tmX;
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)lddy * 2, (uint64_t)d->Wo * lddy * 2, (uint64_t)d->Ho * d->Wo * lddy * 2};
    uint32_t box[4] = {64, (uint32_t)p.KW, (uint32_t)p.KH, (uint32_t)p.KN};
    uint32_t es[4] = {1, 1, 1, 1};
    rc = make_tmap_bf16(&tmDY, dy, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldx * 2, (uint64_t)d->W * d->ldx * 2, (uint64_t)d->H * d->W * d->ldx * 2};
    const uint32_t st = (uint32_t)d->stride;  // stride 2: the box walks every other input pixel
    uint32_t box[4] = {64, (uint32_t)p.KW * st, (uint32_t)p.KH * st, (uint32_t)p.KN};
    uint32_t es[4] = {1, st, st, 1};
    rc = make_tmap_bf16(&tmX, x, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  switch (nb) {
    case 4: rc = launch_wgrad<4, 4>(tmDY, tmX, p, stream); break;
    case 3: rc = launch_wgrad<3, 5>(tmDY, tmX, p, stream); break;
    case 2: rc = launch_wgrad<2, 6>(tmDY, tmX, p, stream); break;
    default: rc = launch_wgrad<1, 8>(tmDY, tmX, p, stream); break;
  }
  if (rc) return rc;
  if (!dw_oihw) return JG_OK;
  return jg_unpack_conv_wgrad(ws, dw_oihw, d->Cout, d->Cin, d->R, d->S, beta, stream_);
}

extern "C" int jg_conv2d_wgrad(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* ws,
                               float* dw_oihw, float beta, jg_stream_t stream_) {
  JG_CHECK(dw_oihw, JG_ERR_INVALID, "conv_wgrad: null pointer");
  return conv2d_wgrad_impl(d, x, dy, lddy, ws, dw_oihw, beta, nullptr, stream_);
}

extern "C" int jg_conv2d_wgrad_acc(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* acc,
                                   int* layout, jg_stream_t stream_) {
  JG_CHECK(layout, JG_ERR_INVALID, "conv_wgrad_acc: null layout pointer");
  return conv2d_wgrad_impl(d, x, dy, lddy, acc, nullptr, 0.f, layout, stream_);
}
