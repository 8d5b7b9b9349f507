// Element-wise / layout kernels specific to the GAN generator / discriminator path (NHWC bf16, HBM-bound):
//   reflection / replication padding  (nn.ReflectionPad2d, resnet_generator.py:52-53, 207, 328) fwd + bwd
//   zero insertion ("dilate 2x")      the stride-2 transposed-convolution / stride-2 dgrad prologue
//   activation backward               tanh (resnet_generator.py:339), LeakyReLU(0.2) (discriminators.py:56)
//   lsgan / hinge losses on the 1-channel PatchGAN logits (loss.py:11-85)
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

__device__ __forceinline__ int reflect_idx(int q, int n) {  // q in [-pad, n+pad), pad < n
  if (q < 0) q = -q;
  if (q >= n) q = 2 * (n - 1) - q;
  return q;
}
__device__ __forceinline__ int clamp_idx(int q, int n) { return q < 0 ? 0 : (q >= n ? n - 1 : q); }

// dst[N, H+2p, W+2p, C] = pad(src[N,H,W,C]); mode 0 reflect, 1 replicate
__global__ void pad2d_fwd_kernel(const __nv_bfloat16* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst,
                                 int ldd, int N, int H, int W, int C, int pad, int mode) {
  const int vecs = C / 8;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long total = (long long)N * Hp * Wp * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long pix = i / vecs;
    const int w = (int)(pix % Wp);
    pix /= Wp;
    const int h = (int)(pix % Hp);
    const int n = (int)(pix / Hp);
    const int sh = mode == 0 ? reflect_idx(h - pad, H) : clamp_idx(h - pad, H);
    const int sw = mode == 0 ? reflect_idx(w - pad, W) : clamp_idx(w - pad, W);
    *reinterpret_cast<uint4*>(dst + (((long long)n * Hp + h) * Wp + w) * ldd + v * 8) =
        *reinterpret_cast<const uint4*>(src + (((long long)n * H + sh) * W + sw) * lds + v * 8);
  }
}

// Replication padding backward: the border rows / columns of the source also feed the `pad` padded rows / columns
// next to them, every other source position feeds its interior copy only.  (Not exercised on hardware yet: the
// reference's default --G_padding_type is reflect.)
__global__ void pad2d_bwd_replicate_kernel(const __nv_bfloat16* __restrict__ dpad, int ldp,
                                           __nv_bfloat16* __restrict__ dsrc, int lds, int N, int H, int W, int C,
                                           int pad) {
  const int vecs = C / 8;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long total = (long long)N * H * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long pix = i / vecs;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int n = (int)(pix / H);
    // padded rows [h0, h1] and columns [w0, w1] that read source (h, w); a 1-wide source is both borders at once
    const int h0 = h == 0 ? 0 : h + pad, h1 = h == H - 1 ? Hp - 1 : h + pad;
    const int w0 = w == 0 ? 0 : w + pad, w1 = w == W - 1 ? Wp - 1 : w + pad;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = h0; a <= h1; ++a)
      for (int b = w0; b <= w1; ++b) {
        const uint4 u = *reinterpret_cast<const uint4*>(dpad + (((long long)n * Hp + a) * Wp + b) * ldp + v * 8);
        const uint32_t* pu = &u.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(pu[j]);
          f[2 * j] += x.x;
          f[2 * j + 1] += x.y;
        }
      }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dsrc + (((long long)n * H + h) * W + w) * lds + v * 8) = o;
  }
}

// dsrc[N,H,W,C] = sum of dpad over the padded positions that read (h,w)   (reflection padding, pad < min(H,W))
__global__ void pad2d_bwd_kernel(const __nv_bfloat16* __restrict__ dpad, int ldp, __nv_bfloat16* __restrict__ dsrc,
                                 int lds, int N, int H, int W, int C, int pad) {
  const int vecs = C / 8;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long total = (long long)N * H * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long pix = i / vecs;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int n = (int)(pix / H);
    // padded rows that map to h: the interior copy, the top mirror (1 <= h <= pad), the bottom mirror
    int hq[3], wq[3], nh = 0, nw = 0;
    hq[nh++] = h + pad;
    if (h >= 1 && h <= pad) hq[nh++] = pad - h;
    if (h <= H - 2 && h >= H - 1 - pad) hq[nh++] = 2 * (H - 1) - h + pad;
    wq[nw++] = w + pad;
    if (w >= 1 && w <= pad) wq[nw++] = pad - w;
    if (w <= W - 2 && w >= W - 1 - pad) wq[nw++] = 2 * (W - 1) - w + pad;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        const uint4 u =
            *reinterpret_cast<const uint4*>(dpad + (((long long)n * Hp + hq[a]) * Wp + wq[b]) * ldp + v * 8);
        const uint32_t* pu = &u.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(pu[j]);
          f[2 * j] += x.x;
          f[2 * j + 1] += x.y;
        }
      }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dsrc + (((long long)n * H + h) * W + w) * lds + v * 8) = o;
  }
}

// dst[N, 2H, 2W, C]: dst[2h][2w] = src[h][w], zeros elsewhere (mode 0);  mode 1: the inverse gather
// dst[N,H,W,C] = src[N,2H,2W,C][2h][2w].
__global__ void dilate2x_kernel(const __nv_bfloat16* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst,
                                int ldd, int N, int Hd, int Wd, int C, int mode) {
  const int vecs = C / 8;
  const long long total = (long long)N * Hd * Wd * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long pix = i / vecs;
    const int w = (int)(pix % Wd);
    pix /= Wd;
    const int h = (int)(pix % Hd);
    const int n = (int)(pix / Hd);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (mode == 0) {
      if (((h | w) & 1) == 0)
        o = *reinterpret_cast<const uint4*>(src + (((long long)n * (Hd / 2) + h / 2) * (Wd / 2) + w / 2) * lds + v * 8);
    } else {
      o = *reinterpret_cast<const uint4*>(src + (((long long)n * (Hd * 2) + 2 * h) * (Wd * 2) + 2 * w) * lds + v * 8);
    }
    *reinterpret_cast<uint4*>(dst + (((long long)n * Hd + h) * Wd + w) * ldd + v * 8) = o;
  }
}

// dx = dy * act'(.) expressed through the OUTPUT y: tanh' = 1 - y^2; LeakyReLU/ReLU keep the sign of x.
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ y, int ldy, const __nv_bfloat16* __restrict__ dy,
                               int lddy, __nv_bfloat16* __restrict__ dx, int lddx, long long rows, int C, int act) {
  const int vecs = C / 8;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vecs;
    const int v = (int)(i - r * vecs);
    const uint4 uy = *reinterpret_cast<const uint4*>(y + r * ldy + v * 8);
    const uint4 ud = *reinterpret_cast<const uint4*>(dy + r * lddy + v * 8);
    const uint32_t* py = &uy.x;
    const uint32_t* pd = &ud.x;
    uint32_t res[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16x2(py[j]), d = unpack_bf16x2(pd[j]);
      float g0, g1;
      if (act == JG_ACT_TANH) {
        g0 = 1.f - a.x * a.x;
        g1 = 1.f - a.y * a.y;
      } else if (act == JG_ACT_LRELU02) {
        g0 = a.x > 0.f ? 1.f : 0.2f;
        g1 = a.y > 0.f ? 1.f : 0.2f;
      } else {  // relu
        g0 = a.x > 0.f ? 1.f : 0.f;
        g1 = a.y > 0.f ? 1.f : 0.f;
      }
      res[j] = pack_bf16x2(d.x * g0, d.y * g1);
    }
    *reinterpret_cast<uint4*>(dx + r * lddx + v * 8) = make_uint4(res[0], res[1], res[2], res[3]);
  }
}

// GANLoss on PatchGAN logits pred[rows][ld] (C real channels):
//   mode 0 lsgan:  mean((pred - target)^2)            (loss.py:70-72, nn.MSELoss against a constant label)
//   mode 1 hinge:  mean(relu(1 - sign*pred))          (loss.py:78-83, "projected", sign = +1 real / -1 fake)
//   mode 2 linear: mean(-sign*pred)                   (loss.py:74-77 wgangp / :84 projected without relu)
// fwd: loss (overwritten by the caller's memset + atomics); bwd: dpred = gout * dloss/dpred.
__global__ void gan_loss_fwd_kernel(const __nv_bfloat16* __restrict__ pred, int ld, long long rows, int C, int mode,
                                    float target, float sign, float inv_count, float* __restrict__ loss) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * C;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float p = __bfloat162float(pred[r * ld + c]);
    if (mode == 0) acc += (p - target) * (p - target);
    else if (mode == 1) acc += fmaxf(1.f - sign * p, 0.f);
    else acc += -sign * p;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    atomicAdd(loss, s * inv_count);
  }
}
__global__ void gan_loss_bwd_kernel(const __nv_bfloat16* __restrict__ pred, int ld, long long rows, int C, int mode,
                                    float target, float sign, float inv_count, const float* __restrict__ gout,
                                    __nv_bfloat16* __restrict__ dpred, int ldd) {
  const float g = (gout ? *gout : 1.f) * inv_count;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * ldd;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ldd;
    const int c = (int)(i - r * ldd);
    float d = 0.f;
    if (c < C) {
      const float p = __bfloat162float(pred[r * ld + c]);
      if (mode == 0) d = 2.f * (p - target);
      else if (mode == 1) d = (1.f - sign * p > 0.f) ? -sign : 0.f;
      else d = -sign;
    }
    dpred[i] = __float2bfloat16(g * d);
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace jg

using namespace jg;

extern "C" int jg_pad2d_fwd(const void* src, int lds, void* dst, int ldd, int N, int H, int W, int C, int pad, int mode,
                            jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && N > 0 && C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, JG_ERR_INVALID,
           "pad2d_fwd: bad args");
  JG_CHECK(pad >= 0 && pad < H && pad < W && (mode == 0 || mode == 1), JG_ERR_INVALID, "pad2d_fwd: pad %d / mode %d",
           pad, mode);
  const long long total = (long long)N * (H + 2 * pad) * (W + 2 * pad) * (C / 8);
  pad2d_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), lds,
                                                            static_cast<__nv_bfloat16*>(dst), ldd, N, H, W, C, pad,
                                                            mode);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_pad2d_bwd(const void* dpad, int ldp, void* dsrc, int lds, int N, int H, int W, int C, int pad,
                            int mode, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(dpad && dsrc && N > 0 && C > 0 && C % 8 == 0 && ldp % 8 == 0 && lds % 8 == 0, JG_ERR_INVALID,
           "pad2d_bwd: bad args");
  JG_CHECK(mode == 0 || mode == 1, JG_ERR_INVALID, "pad2d_bwd: mode %d (0 reflect, 1 replicate)", mode);
  JG_CHECK(pad >= 0 && pad < H && pad < W, JG_ERR_INVALID, "pad2d_bwd: pad %d", pad);
  const long long total = (long long)N * H * W * (C / 8);
  if (mode == 0) {
    pad2d_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dpad), ldp,
                                                              static_cast<__nv_bfloat16*>(dsrc), lds, N, H, W, C, pad);
  } else {
    pad2d_bwd_replicate_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(dpad), ldp, static_cast<__nv_bfloat16*>(dsrc), lds, N, H, W, C, pad);
  }
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_dilate2x(const void* src, int lds, void* dst, int ldd, int N, int Hd, int Wd, int C, int mode,
                           jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && N > 0 && C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && (mode == 0 || mode == 1),
           JG_ERR_INVALID, "dilate2x: bad args");
  JG_CHECK(mode == 1 || (Hd % 2 == 0 && Wd % 2 == 0), JG_ERR_INVALID, "dilate2x: odd output %dx%d", Hd, Wd);
  const long long total = (long long)N * Hd * Wd * (C / 8);
  dilate2x_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), lds,
                                                           static_cast<__nv_bfloat16*>(dst), ldd, N, Hd, Wd, C, mode);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_act_bwd(const void* y, int ldy, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int C,
                          int act, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(y && dy && dx && rows > 0 && C > 0 && C % 8 == 0, JG_ERR_INVALID, "act_bwd: bad args");
  JG_CHECK(act == JG_ACT_TANH || act == JG_ACT_LRELU02 || act == JG_ACT_RELU, JG_ERR_INVALID, "act_bwd: act %d", act);
  const long long total = rows * (C / 8);
  act_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), ldy,
                                                          static_cast<const __nv_bfloat16*>(dy), lddy,
                                                          static_cast<__nv_bfloat16*>(dx), lddx, rows, C, act);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_gan_loss_fwd(const void* pred, int ld, int64_t rows, int C, int mode, float target, float sign,
                               float* loss, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(pred && loss && rows > 0 && C > 0 && ld >= C && mode >= 0 && mode <= 2, JG_ERR_INVALID,
           "gan_loss_fwd: bad args");
  JG_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  const long long total = rows * C;
  gan_loss_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), ld, rows, C,
                                                               mode, target, sign, 1.f / (float)total, loss);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_gan_loss_bwd(const void* pred, int ld, int64_t rows, int C, int mode, float target, float sign,
                               const float* grad_out, void* dpred, int ldd, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(pred && dpred && rows > 0 && C > 0 && ld >= C && ldd >= C && mode >= 0 && mode <= 2, JG_ERR_INVALID,
           "gan_loss_bwd: bad args");
  const long long total = rows * ldd;
  gan_loss_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), ld, rows, C,
                                                               mode, target, sign, 1.f / (float)(rows * C), grad_out,
                                                               static_cast<__nv_bfloat16*>(dpred), ldd);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
