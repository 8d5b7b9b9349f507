// On-GPU input preparation and Haar wavelet ops: what the reference does on the host in
// its data pipeline / with its upfirdn2d CUDA op, as HBM-bound elementwise kernels on fp32 NCHW images.
//
//   fill_mask_with_random          data/online_creation.py:1366-1376
//   ToTensor + Normalize(0.5,0.5)  data/base_dataset.py (get_transform: transforms.ToTensor, transforms.Normalize)
//   conditioning dropout of masks  models/palette_model.py:565-584 (mask <- num_classes-1 for dropped samples)
//   HaarTransform / InverseHaarTransform   models/modules/freq_utils.py:22-59 on upfirdn2d
//                                  (models/modules/op/upfirdn2d.py:167-208 defines the arithmetic; its CUDA kernel is
//                                  models/modules/op/upfirdn2d_kernel.cu:49-200)
//
// Index / mask handling is bit exact; the float arithmetic uses explicitly rounded (non-contracted) operations in the
// reference's order, so images match the reference bit for bit as well.
#include "common.cuh"

namespace jg {

// out = img*(1-m) + noise*m,  m = (mask != 0) if cls == -1 else (mask == cls)   (one mask channel for all C channels)
__global__ void fill_mask_random_kernel(const float* __restrict__ img, const float* __restrict__ maskf,
                                        const long long* __restrict__ maski, const float* __restrict__ noise,
                                        float* __restrict__ out, int C, long long HW, long long total, int cls) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / (C * HW);
  const long long p = i % HW;
  bool sel;
  if (maski) {
    const long long mv = maski[n * HW + p];
    sel = cls == -1 ? (mv != 0) : (mv == cls);
  } else {
    const float mv = maskf[n * HW + p];
    sel = cls == -1 ? (mv != 0.f) : (mv == (float)cls);
  }
  const float m = sel ? 1.f : 0.f;
  out[i] = __fadd_rn(__fmul_rn(img[i], __fsub_rn(1.f, m)), __fmul_rn(noise[i], m));
}

// uint8 HWC [N][H][W][C] -> fp32 NCHW, (x/255 - mean)/std
__global__ void u8_to_f32_norm_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int C, int H,
                                      int W, long long total, float mean, float stdv) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % W);
  const int h = (int)((i / W) % H);
  const int c = (int)((i / ((long long)W * H)) % C);
  const long long n = i / ((long long)W * H * C);
  const float v = (float)src[((n * H + h) * W + w) * C + c];
  dst[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.f), mean), stdv);
}

// mask[n] <- fill where drop_u[n] < prob  (int64 or fp32 mask, [N][HW] per image)
__global__ void mask_dropout_kernel(const float* __restrict__ maskf, const long long* __restrict__ maski,
                                    const float* __restrict__ drop_u, float prob, long long fill, float* __restrict__ outf,
                                    long long* __restrict__ outi, long long per_image, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const bool drop = drop_u[i / per_image] < prob;
  if (maski) outi[i] = drop ? fill : maski[i];
  else outf[i] = drop ? (float)fill : maskf[i];
}

// Haar analysis / synthesis.  k_band[a][b] (a = row, b = column), all entries +-(1/sqrt2)^2:
//   ll = +,+,+,+   lh = -,-,+,+   hl = -,+,-,+   hh = +,-,-,+        (freq_utils.get_haar_wavelet)
// upfirdn2d(down=2) correlates with the FLIPPED kernel: out[i][j] = sum_ab x[2i+a][2j+b] * k[1-a][1-b];
// upfirdn2d(up=2, pad=(1,0,1,0)): out[2i+a][2j+b] = x[i][j] * k[a][b].
__device__ __forceinline__ float haar_k(int band, int a, int b) {
  // sign tables, row-major [a][b]
  const int s = band == 0 ? 0x0 : band == 1 ? 0x3 : band == 2 ? 0x5 : 0x6;  // bit (a*2+b) set = negative
  // the reference builds the taps as fp32 products of 1/sqrt(2): 0.70710677f * 0.70710677f = 0.49999997f, not 0.5f
  const float kk = __fmul_rn(0.70710677f, 0.70710677f);
  return ((s >> (a * 2 + b)) & 1) ? -kk : kk;
}

// mode 0: DWT forward   x [N][C][H][W] -> y [N][4C][H/2][W/2]   (bands ll | lh | hl | hh)
// mode 1: DWT backward  dy [N][4C][H/2][W/2] -> dx [N][C][H][W]
// mode 2: IWT forward   y [N][4C][h][w] -> x [N][C][2h][2w]       (synthesis kernels ll, -lh, -hl, hh)
// mode 3: IWT backward  dx [N][C][2h][2w] -> dy [N][4C][h][w]
// One thread per low-resolution pixel (n, c, i, j): it owns the 2x2 full-resolution block and the 4 band values.
__global__ void haar_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int h, int w, long long total,
                            int mode) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % w);
  const int i = (int)((idx / w) % h);
  const int c = (int)((idx / ((long long)w * h)) % C);
  const long long n = idx / ((long long)w * h * C);
  const long long full = ((n * C + c) * (2LL * h) + 2 * i) * (2LL * w) + 2 * j;  // (2i, 2j) of the full-res plane
  const long long W2 = 2LL * w;
  auto band_at = [&](int band) { return ((n * 4 * C + band * C + c) * (long long)h + i) * w + j; };
  if (mode == 0 || mode == 3) {
    // full resolution -> bands
    float x[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) x[a][b] = src[full + a * W2 + b];
#pragma unroll
    for (int band = 0; band < 4; ++band) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          // DWT: flipped analysis kernel; IWT backward: the synthesis kernel itself (adjoint of the scatter)
          float k = mode == 0 ? haar_k(band, 1 - a, 1 - b) : haar_k(band, a, b);
          if (mode == 3 && (band == 1 || band == 2)) k = -k;
          acc = __fadd_rn(acc, __fmul_rn(x[a][b], k));
        }
      dst[band_at(band)] = acc;
    }
  } else {
    // bands -> full resolution
    float v[4];
#pragma unroll
    for (int band = 0; band < 4; ++band) v[band] = src[band_at(band)];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float acc = 0.f;
#pragma unroll
        for (int band = 0; band < 4; ++band) {
          // IWT: synthesis kernels (ll, -lh, -hl, hh), summed in that order; DWT backward: flipped analysis kernels
          float k = mode == 2 ? haar_k(band, a, b) : haar_k(band, 1 - a, 1 - b);
          if (mode == 2 && (band == 1 || band == 2)) k = -k;
          acc = __fadd_rn(acc, __fmul_rn(v[band], k));
        }
        dst[full + a * W2 + b] = acc;
      }
  }
}

// ---- per-pixel label embedding (PaletteDenoiseFn.compute_cond, palette_denoise_fn.py:118-136) --------------------------
// out[row][col0 + e] = bf16(table[idx[row]][e]) for an NHWC bf16 tensor with channel stride ld: the mask-embedding
// channels that the reference concatenates to the UNet input.  idx is the int64 / fp32 semantic mask.
__global__ void embed_rows_kernel(const float* __restrict__ table, const float* __restrict__ idxf,
                                  const long long* __restrict__ idxi, __nv_bfloat16* __restrict__ out, int ld, int col0,
                                  long long rows, int E, int K) {
  // (col0 = 6 for the Palette input: [y_cond | y_noisy | mask embedding] — only 4-byte aligned: bf16 pairs)
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int e2 = E / 2;
  if (i >= rows * e2) return;
  const long long row = i / e2;
  const int v = (int)(i % e2);
  long long k = idxi ? idxi[row] : (long long)idxf[row];  // (.to(torch.int32) truncates)
  k = k < 0 ? 0 : (k >= K ? K - 1 : k);                   // an out-of-range label is an error upstream; stay in bounds
  const float2 a = *reinterpret_cast<const float2*>(table + (size_t)k * E + v * 2);
  *reinterpret_cast<__nv_bfloat162*>(out + row * ld + col0 + v * 2) = __floats2bfloat162_rn(a.x, a.y);
}

// dtable[k][e] += sum over rows with idx[row] == k of d[row][col0 + e]; counts[k] += number of such rows
// (scale_grad_by_freq divides afterwards).  One warp walks a contiguous chunk of rows; lane <-> channel (e, e+32, ..);
// every warp owns a private [K][E] accumulator in shared memory (no atomics inside the block), one fp32 atomic per
// (class, channel) and block at the end.
constexpr int kEmbWarps = 8;
__global__ void __launch_bounds__(kEmbWarps * 32)
embed_rows_bwd_kernel(const __nv_bfloat16* __restrict__ d, int ld, int col0, const float* __restrict__ idxf,
                      const long long* __restrict__ idxi, long long rows, int E, int K, long long rows_per_warp,
                      float* __restrict__ dtable, float* __restrict__ counts) {
  extern __shared__ float sacc[];  // [warps][K][E] then [warps][K] counts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* acc = sacc + (size_t)warp * K * E;
  float* cnt = sacc + (size_t)kEmbWarps * K * E + warp * K;
  for (int i = lane; i < K * E; i += 32) acc[i] = 0.f;
  for (int i = lane; i < K; i += 32) cnt[i] = 0.f;
  __syncwarp();
  const long long w = blockIdx.x * (long long)kEmbWarps + warp;
  const long long r0 = w * rows_per_warp;
  const long long r1 = r0 + rows_per_warp < rows ? r0 + rows_per_warp : rows;
  for (long long r = r0; r < r1; ++r) {
    long long k = idxi ? idxi[r] : (long long)idxf[r];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    for (int e = lane; e < E; e += 32) acc[k * E + e] += __bfloat162float(d[r * ld + col0 + e]);
    if (lane == 0) cnt[k] += 1.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * E; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kEmbWarps; ++ww) t += sacc[(size_t)ww * K * E + i];
    if (t != 0.f) atomicAdd(&dtable[i], t);
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kEmbWarps; ++ww) t += sacc[(size_t)kEmbWarps * K * E + ww * K + i];
    if (t != 0.f) atomicAdd(&counts[i], t);
  }
}

}  // namespace jg

using namespace jg;

static inline unsigned grid_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

extern "C" int jg_fill_mask_random(const float* img, const float* mask_f32, const int64_t* mask_i64, const float* noise,
                                   float* out, int N, int C, int HW, int cls, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(img && noise && out && (mask_f32 || mask_i64) && N > 0 && C > 0 && HW > 0, JG_ERR_INVALID,
           "fill_mask_random: bad args");
  const long long total = (long long)N * C * HW;
  fill_mask_random_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      img, mask_f32, reinterpret_cast<const long long*>(mask_i64), noise, out, C, HW, total, cls);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_u8_to_f32_normalized(const uint8_t* src_nhwc, float* dst_nchw, int N, int C, int H, int W, float mean,
                                       float stdv, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src_nhwc && dst_nchw && N > 0 && C > 0 && H > 0 && W > 0 && stdv != 0.f, JG_ERR_INVALID,
           "u8_to_f32_normalized: bad args");
  const long long total = (long long)N * C * H * W;
  u8_to_f32_norm_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src_nhwc, dst_nchw, C, H, W, total, mean, stdv);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_mask_class_dropout(const float* mask_f32, const int64_t* mask_i64, const float* drop_u, float prob,
                                     int64_t fill, float* out_f32, int64_t* out_i64, int N, int64_t per_image,
                                     jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(drop_u && N > 0 && per_image > 0 && ((mask_f32 && out_f32) || (mask_i64 && out_i64)), JG_ERR_INVALID,
           "mask_class_dropout: bad args");
  const long long total = (long long)N * per_image;
  mask_dropout_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      mask_f32, reinterpret_cast<const long long*>(mask_i64), drop_u, prob, (long long)fill, out_f32,
      reinterpret_cast<long long*>(out_i64), per_image, total);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_haar(const float* src, float* dst, int N, int C, int h, int w, int mode, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && N > 0 && C > 0 && h > 0 && w > 0 && mode >= 0 && mode <= 3, JG_ERR_INVALID, "haar: bad args");
  const long long total = (long long)N * C * h * w;
  haar_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, dst, C, h, w, total, mode);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_embed_rows(const float* table, const float* idx_f32, const int64_t* idx_i64, void* out, int ld,
                             int col0, int64_t rows, int E, int K, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(table && out && (idx_f32 || idx_i64) && rows > 0 && K > 0, JG_ERR_INVALID, "embed_rows: bad args");
  JG_CHECK(E > 0 && E % 2 == 0 && ld % 8 == 0 && col0 % 2 == 0 && col0 + E <= ld, JG_ERR_INVALID,
           "embed_rows: E=%d and col0=%d must be even, ld=%d a multiple of 8, col0 + E <= ld", E, col0, ld);
  const long long total = (long long)rows * (E / 2);
  embed_rows_kernel<<<grid_for(total, 256), 256, 0, stream>>>(table, idx_f32,
                                                             reinterpret_cast<const long long*>(idx_i64),
                                                             static_cast<__nv_bfloat16*>(out), ld, col0, rows, E, K);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_embed_rows_bwd(const void* d, int ld, int col0, const float* idx_f32, const int64_t* idx_i64,
                                 int64_t rows, int E, int K, float* dtable, float* counts, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(d && (idx_f32 || idx_i64) && dtable && counts && rows > 0 && K > 0 && E > 0, JG_ERR_INVALID,
           "embed_rows_bwd: bad args");
  const size_t smem = (size_t)kEmbWarps * K * (E + 1) * sizeof(float);
  JG_CHECK(smem <= 96 * 1024, JG_ERR_INVALID, "embed_rows_bwd: %d classes x %d channels exceed the accumulator", K, E);
  JG_CUDA(cudaMemsetAsync(dtable, 0, sizeof(float) * (size_t)K * E, stream));
  JG_CUDA(cudaMemsetAsync(counts, 0, sizeof(float) * (size_t)K, stream));
  static bool attr_done = false;
  if (!attr_done) {
    JG_CUDA(cudaFuncSetAttribute(embed_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done = true;
  }
  const int blocks = num_sms() * 2;
  const long long warps = (long long)blocks * kEmbWarps;
  const long long rpw = (rows + warps - 1) / warps;
  embed_rows_bwd_kernel<<<blocks, kEmbWarps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(d), ld, col0, idx_f32,
                                                                 reinterpret_cast<const long long*>(idx_i64), rows, E, K,
                                                                 rpw, dtable, counts);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
