// Small fused kernels around the UNet: timestep-embedding Linear (+SiLU on the input), the
// DiffusionGenerator noising prologue, the Palette masked eps-loss (forward + backward in one pass
// each), and the fused multi-tensor AdamW + EMA update.
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

__device__ __forceinline__ float silu_f(float u) { return u / (1.f + __expf(-u)); }
__device__ __forceinline__ float silu_g(float u) {
  const float s = 1.f / (1.f + __expf(-u));
  return s * (1.f + u * (1.f - s));
}

// y[b][o] = sum_i act(x[b][i]) * w[o][i] + bias[o]        (emb_layers: SiLU -> Linear,
// unet_generator_attn.py:201-207; cond_embed MLP, diffusion_generator.py:72-76)
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ y, int B, int I, int O,
                                  int act_in, int act_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * O) return;
  const int b = idx / O, o = idx % O;
  float acc = bias ? bias[o] : 0.f;
  for (int i = 0; i < I; ++i) {
    float xv = x[b * I + i];
    if (act_in == JG_ACT_SILU) xv = silu_f(xv);
    acc += xv * w[(size_t)o * I + i];
  }
  if (act_out == JG_ACT_SILU) acc = silu_f(acc);
  y[idx] = acc;
}
// dx[b][i] = act'(x[b][i]) * sum_o dy[b][o] w[o][i]
// one block per batch row; the O reduction is split over blockDim.x / 32 slices and combined in smem.
__global__ void __launch_bounds__(256)
linear_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy,
                     float* __restrict__ dx, int B, int I, int O, int act_in, int accumulate) {
  __shared__ float part[8][32];
  const int b = blockIdx.x;
  const int il = threadIdx.x & 31;
  const int slice = threadIdx.x >> 5;
  for (int i0 = 0; i0 < I; i0 += 32) {
    const int i = i0 + il;
    float acc = 0.f;
    if (i < I) {
      // 8 independent partial sums: the loads of 8 consecutive iterations are in flight together
      float a8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int o = slice;
      for (; o + 56 < O; o += 64) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a8[k] = fmaf(dy[(size_t)b * O + o + 8 * k], w[(size_t)(o + 8 * k) * I + i], a8[k]);
      }
      for (; o < O; o += 8) a8[0] = fmaf(dy[(size_t)b * O + o], w[(size_t)o * I + i], a8[0]);
      acc = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
    }
    part[slice][il] = acc;
    __syncthreads();
    if (slice == 0 && i < I) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += part[k][il];
      const int idx = b * I + i;
      if (act_in == JG_ACT_SILU) tot *= silu_g(x[idx]);
      dx[idx] = accumulate ? dx[idx] + tot : tot;
    }
    __syncthreads();
  }
}
// dw[o][i] = sum_b dy[b][o] act(x[b][i]);  db[o] = sum_b dy[b][o]
__global__ void linear_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                     float* __restrict__ db, int B, int I, int O, int act_in) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * I) return;
  const int o = idx / I, i = idx % I;
  float acc = 0.f, accb = 0.f;
#pragma unroll 8
  for (int b = 0; b < B; ++b) {
    float xv = x[b * I + i];
    if (act_in == JG_ACT_SILU) xv = silu_f(xv);
    const float d = dy[(size_t)b * O + o];
    acc = fmaf(d, xv, acc);
    accb += d;
  }
  dw[idx] = acc;
  if (db && i == 0) db[o] = accb;
}

// ---- all emb_layers Linears of a UNet in ONE launch ("the 28 tiny Linears -> one batched GEMM") -------
// Every ResBlock owns Linear(emb_channels -> 2*Cout) applied to the SAME SiLU(emb) [B][I] (unet_generator_attn.py:
// 201-207, 247-258).  Item i writes its own contiguous [B][O_i] block of Y at float offset B * off_i.
// One tile = (item, 64 outputs); grid = total tiles; 256 threads = 64 outputs x 4 batch slices.
struct LinearItem {  // mirror of jg_linear_item
  const float* w;    // [O][I]
  const float* b;    // [O] or null
  int O, off;        // outputs; first output in the concatenation of all items
};

__device__ __forceinline__ int find_item(const int* __restrict__ tile_start, int n, int tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_start[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int kLinMaxI = 128, kLinMaxB = 64;

__global__ void __launch_bounds__(256)
linear_batched_fwd_kernel(const float* __restrict__ x, const LinearItem* __restrict__ items,
                          const int* __restrict__ tile_start, int n, float* __restrict__ Y, int B, int I, int act_in) {
  __shared__ float sx[kLinMaxB * kLinMaxI];
  for (int i = threadIdx.x; i < B * I; i += 256) sx[i] = act_in == JG_ACT_SILU ? silu_f(x[i]) : x[i];
  __syncthreads();
  const int it = find_item(tile_start, n, blockIdx.x);
  const LinearItem item = items[it];
  const int o = (blockIdx.x - tile_start[it]) * 64 + (threadIdx.x & 63);
  if (o >= item.O) return;
  const float* wr = item.w + (size_t)o * I;
  const float bias = item.b ? item.b[o] : 0.f;
  float* y = Y + (size_t)B * item.off;
  for (int b = threadIdx.x >> 6; b < B; b += 4) {
    float acc = bias;
    for (int i = 0; i < I; ++i) acc = fmaf(sx[b * I + i], wr[i], acc);
    y[(size_t)b * item.O + o] = acc;
  }
}

// dW_i[o][k] = sum_b dY_i[b][o] act(x[b][k]);  db_i[o] = sum_b dY_i[b][o];
// dx[b][k] += act'(x[b][k]) * sum_o dY_i[b][o] W_i[o][k]   (summed over ALL items: fp32 atomics into a zeroed dx)
__global__ void __launch_bounds__(256)
linear_batched_bwd_kernel(const float* __restrict__ x, const LinearItem* __restrict__ items,
                          const int* __restrict__ tile_start, int n, const float* __restrict__ dY,
                          float* __restrict__ dW, float* __restrict__ dB, float* __restrict__ dx, int B, int I,
                          int act_in) {
  __shared__ float sx[kLinMaxB * kLinMaxI];   // act(x)
  __shared__ float sdy[kLinMaxB * 64];        // dY tile [B][64]
  for (int i = threadIdx.x; i < B * I; i += 256) sx[i] = act_in == JG_ACT_SILU ? silu_f(x[i]) : x[i];
  const int it = find_item(tile_start, n, blockIdx.x);
  const LinearItem item = items[it];
  const int o0 = (blockIdx.x - tile_start[it]) * 64;
  const float* dy = dY + (size_t)B * item.off;
  for (int i = threadIdx.x; i < B * 64; i += 256) {
    const int b = i >> 6, oo = i & 63;
    sdy[i] = (o0 + oo < item.O) ? dy[(size_t)b * item.O + o0 + oo] : 0.f;
  }
  __syncthreads();
  // weight / bias gradients: thread -> (output oo, input slice)
  for (int idx = threadIdx.x; idx < 64 * I; idx += 256) {
    const int oo = idx / I, k = idx - oo * I;
    if (o0 + oo >= item.O) continue;
    float acc = 0.f, accb = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = sdy[b * 64 + oo];
      acc = fmaf(d, sx[b * I + k], acc);
      accb += d;
    }
    dW[(size_t)(item.off + o0 + oo) * I + k] = acc;
    if (k == 0) dB[item.off + o0 + oo] = accb;
  }
  // input gradient: thread -> (b, k), reduce over the tile's 64 outputs
  if (dx) {
    for (int idx = threadIdx.x; idx < B * I; idx += 256) {
      const int b = idx / I, k = idx - b * I;
      float acc = 0.f;
      for (int oo = 0; oo < 64 && o0 + oo < item.O; ++oo) acc = fmaf(sdy[b * 64 + oo], item.w[(size_t)(o0 + oo) * I + k], acc);
      if (act_in == JG_ACT_SILU) acc *= silu_g(x[idx]);
      atomicAdd(&dx[idx], acc);
    }
  }
}

// DiffusionGenerator.forward prologue (diffusion_generator.py:480-491):
//   y_noisy = sqrt(g)*y0 + sqrt(1-g)*noise;  y_noisy = y_noisy*m + (1-m)*y0, m = clamp(mask,0,1);
//   input = cat([y_cond, y_noisy], dim=1)  ->  NHWC bf16 with channel stride ld (zero padded).
__global__ void noise_pack_kernel(const float* __restrict__ y0, const float* __restrict__ ycond,
                                  const float* __restrict__ noise, const float* __restrict__ maskf,
                                  const long long* __restrict__ maski, const float* __restrict__ gammas,
                                  __nv_bfloat16* __restrict__ out, int B, int C, int HW, int ld) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const float g = gammas[b];
  const float sg = sqrtf(g), s1 = sqrtf(1.f - g);
  float m = 1.f;
  bool has_mask = false;
  if (maskf) {
    m = fminf(fmaxf(maskf[idx], 0.f), 1.f);
    has_mask = true;
  } else if (maski) {
    const long long mv = maski[idx];
    m = mv < 0 ? 0.f : (mv > 1 ? 1.f : (float)mv);
    has_mask = true;
  }
  __nv_bfloat16* o = out + idx * ld;
  for (int c = 0; c < C; ++c) {
    const size_t src = ((size_t)b * C + c) * HW + p;
    const float y = y0[src];
    float yn = sg * y + s1 * noise[src];
    if (has_mask) yn = yn * m + (1.f - m) * y;
    o[c] = __float2bfloat16(ycond[src]);
    o[C + c] = __float2bfloat16(yn);
  }
  for (int c = 2 * C; c < ld; ++c) o[c] = __float2bfloat16(0.f);
}

// One reverse-diffusion step (DiffusionGenerator.p_sample + the mask blend of restoration_ddpm,
// diffusion_generator.py:122-177, 192-283; predict_start_from_noise / q_posterior, diffusion_utils.py:122-137):
//   y0_hat = clamp(c1*y_t - c2*eps, -1, 1);  mean = pm1*y0_hat + pm2*y_t;  y = mean + sigma*noise
//   y = y_0*(1-m) + m*y  (m = clamp(mask, 0, 1))
// and, in the same pass, the NEXT step's UNet input cat([y_cond, y]) as NHWC bf16 (x_next, zero-padded to ld).
// eps: the UNet output, NHWC bf16 with channel stride lde; y_t / y_cond / y_0 / noise / y_next fp32 NCHW;
// coef fp32 [B][5] = (c1, c2, pm1, pm2, sigma) gathered at t; noise NULL = 0 (the t == 0 step).
// ddim != 0: the deterministic DDIM update of ddim_p_sample / ddim_p_mean_variance (:349-456),
//   y = clamp(c1*y_t + c2*clamp(eps, -1, 1), -1, 1),  c1 = sqrt(g_prev/g_t), c2 = coef_eps - sqrt(g_prev*(1-g_t)/g_t).
__global__ void ddpm_step_kernel(const __nv_bfloat16* __restrict__ eps, int lde, const float* __restrict__ yt,
                                 const float* __restrict__ ycond, const float* __restrict__ y0,
                                 const float* __restrict__ maskf, const long long* __restrict__ maski,
                                 const float* __restrict__ noise, const float* __restrict__ coef,
                                 float* __restrict__ ynext, __nv_bfloat16* __restrict__ xnext, int B, int C, int HW,
                                 int ld, int ddim) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  const float c1 = coef[b * 5 + 0], c2 = coef[b * 5 + 1], pm1 = coef[b * 5 + 2], pm2 = coef[b * 5 + 3];
  const float sigma = coef[b * 5 + 4];
  float m = 1.f;
  bool has_mask = false;
  if (maskf) {
    m = fminf(fmaxf(maskf[idx], 0.f), 1.f);
    has_mask = true;
  } else if (maski) {
    const long long mv = maski[idx];
    m = mv < 0 ? 0.f : (mv > 1 ? 1.f : (float)mv);
    has_mask = true;
  }
  __nv_bfloat16* o = xnext ? xnext + idx * ld : nullptr;
  for (int c = 0; c < C; ++c) {
    const size_t src = ((size_t)b * C + c) * HW + p;
    const float y = yt[src];
    const float e = __bfloat162float(eps[idx * lde + c]);
    float out;
    if (ddim) {
      // ddim_p_mean_variance (:389-456): the network output is clamped, then the mean
      const float ec = fminf(fmaxf(e, -1.f), 1.f);
      out = fminf(fmaxf(c1 * y + c2 * ec, -1.f), 1.f);
    } else {
      const float y0h = fminf(fmaxf(c1 * y - c2 * e, -1.f), 1.f);
      out = pm1 * y0h + pm2 * y;
      if (noise) out += sigma * noise[src];
    }
    if (has_mask) out = y0[src] * (1.f - m) + m * out;
    ynext[src] = out;
    if (o) {
      o[c] = __float2bfloat16(ycond[src]);
      o[C + c] = __float2bfloat16(out);
    }
  }
  if (o)
    for (int c = 2 * C; c < ld; ++c) o[c] = __float2bfloat16(0.f);
}

// Palette loss (palette_model.py:596-620): loss = lambda * mean_{b,c,p} (w_b*m*(noise - noise_hat))^2  (MSE)
//                                      or   lambda * mean |w_b*m*(noise - noise_hat)|              (L1)
// noise fp32 NCHW, noise_hat NHWC bf16 (channel stride ld).  One pass: block partial sums -> atomicAdd(loss).
__global__ void palette_loss_fwd_kernel(const float* __restrict__ noise, const __nv_bfloat16* __restrict__ nh, int ld,
                                        const float* __restrict__ maskf, const long long* __restrict__ maski,
                                        const float* __restrict__ wb, int B, int C, int HW, float coef, int l1,
                                        float* __restrict__ loss) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (idx < (long long)B * HW) {
    const int b = (int)(idx / HW);
    const int p = (int)(idx % HW);
    float m = 1.f;
    if (maskf) m = fminf(fmaxf(maskf[idx], 0.f), 1.f);
    else if (maski) {
      const long long mv = maski[idx];
      m = mv < 0 ? 0.f : (mv > 1 ? 1.f : (float)mv);
    }
    const float wm = (wb ? wb[b] : 1.f) * m;
    for (int c = 0; c < C; ++c) {
      const float e = noise[((size_t)b * C + c) * HW + p];
      const float eh = __bfloat162float(nh[idx * ld + c]);
      const float d = wm * (e - eh);
      acc += l1 ? fabsf(d) : d * d;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    atomicAdd(loss, s * coef);
  }
}
// d loss / d noise_hat, NHWC bf16 (channels C..ld-1 zero), scaled by the upstream scalar *gout.
__global__ void palette_loss_bwd_kernel(const float* __restrict__ noise, const __nv_bfloat16* __restrict__ nh, int ld,
                                        const float* __restrict__ maskf, const long long* __restrict__ maski,
                                        const float* __restrict__ wb, int B, int C, int HW, float coef, int l1,
                                        const float* __restrict__ gout, __nv_bfloat16* __restrict__ dnh, int ldd) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const int p = (int)(idx % HW);
  float m = 1.f;
  if (maskf) m = fminf(fmaxf(maskf[idx], 0.f), 1.f);
  else if (maski) {
    const long long mv = maski[idx];
    m = mv < 0 ? 0.f : (mv > 1 ? 1.f : (float)mv);
  }
  const float wm = (wb ? wb[b] : 1.f) * m;
  const float gs = (gout ? *gout : 1.f) * coef;
  for (int c = 0; c < C; ++c) {
    const float e = noise[((size_t)b * C + c) * HW + p];
    const float eh = __bfloat162float(nh[idx * ld + c]);
    const float d = wm * (eh - e);
    const float gr = l1 ? (d > 0.f ? wm : (d < 0.f ? -wm : 0.f)) : 2.f * wm * d;
    dnh[idx * ldd + c] = __float2bfloat16(gs * gr);
  }
  for (int c = C; c < ldd; ++c) dnh[idx * ldd + c] = __float2bfloat16(0.f);
}

// Fused AdamW / Adam + EMA over flat fp32 buffers (torch.optim.AdamW semantics, train.py:57-58;
// ema_step, base_model.py:1284-1297: p_ema = p + beta*(p_ema - p)).
__global__ void step_increment_kernel(int* step) { *step += 1; }

__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, long long n, float lr, float beta1,
                                 float beta2, float eps, float wd, int adamw, int step_host,
                                 const int* __restrict__ step_dev, float grad_scale, float ema_beta, int ema_init) {
  // The step count may live on the device (CUDA-graph replay: host scalars would be frozen at capture).
  const int step = step_dev ? *step_dev : step_host;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  if (step_dev) ema_init = (step == 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pv = p[i];
    float gv = g[i] * grad_scale;
    if (!adamw && wd != 0.f) gv += wd * pv;
    if (adamw) pv *= 1.f - lr * wd;
    const float mv = beta1 * m[i] + (1.f - beta1) * gv;
    const float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
    if (ema) {
      const float e = ema_init ? pv : ema[i];
      ema[i] = pv + ema_beta * (e - pv);
    }
  }
}

}  // namespace jg

using namespace jg;

extern "C" int jg_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int I, int O,
                             int act_in, int act_out, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && w && y && B > 0 && I > 0 && O > 0, JG_ERR_INVALID, "linear_fwd: bad args");
  linear_fwd_kernel<<<(B * O + 127) / 128, 128, 0, stream>>>(x, w, bias, y, B, I, O, act_in, act_out);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_linear_batched_tiles(int O) { return (O + 63) / 64; }

extern "C" int jg_linear_batched_fwd(const float* x, const jg_linear_item* items_dev, const int* tile_start_dev, int n,
                                     int total_tiles, float* Y, int B, int I, int act_in, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && items_dev && tile_start_dev && Y && n > 0 && total_tiles > 0, JG_ERR_INVALID,
           "linear_batched_fwd: bad args");
  JG_CHECK(B > 0 && B <= kLinMaxB && I > 0 && I <= kLinMaxI, JG_ERR_INVALID,
           "linear_batched_fwd: B=%d (<= %d), I=%d (<= %d)", B, kLinMaxB, I, kLinMaxI);
  linear_batched_fwd_kernel<<<total_tiles, 256, 0, stream>>>(x, reinterpret_cast<const LinearItem*>(items_dev),
                                                             tile_start_dev, n, Y, B, I, act_in);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_linear_batched_bwd(const float* x, const jg_linear_item* items_dev, const int* tile_start_dev, int n,
                                     int total_tiles, const float* dY, float* dW, float* dB, float* dx, int B, int I,
                                     int act_in, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && items_dev && tile_start_dev && dY && dW && dB && n > 0 && total_tiles > 0, JG_ERR_INVALID,
           "linear_batched_bwd: bad args");
  JG_CHECK(B > 0 && B <= kLinMaxB && I > 0 && I <= kLinMaxI, JG_ERR_INVALID,
           "linear_batched_bwd: B=%d (<= %d), I=%d (<= %d)", B, kLinMaxB, I, kLinMaxI);
  if (dx) JG_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * I, stream));
  linear_batched_bwd_kernel<<<total_tiles, 256, 0, stream>>>(x, reinterpret_cast<const LinearItem*>(items_dev),
                                                             tile_start_dev, n, dY, dW, dB, dx, B, I, act_in);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_linear_bwd(const float* x, const float* w, const float* dy, float* dx, int dx_accumulate, float* dw,
                             float* db, int B, int I, int O, int act_in, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(x && w && dy && B > 0 && I > 0 && O > 0, JG_ERR_INVALID, "linear_bwd: bad args");
  if (dx) {
    linear_bwd_dx_kernel<<<B, 256, 0, stream>>>(x, w, dy, dx, B, I, O, act_in, dx_accumulate);
    JG_LAUNCH_CHECK();
  }
  if (dw) {
    linear_bwd_dw_kernel<<<(O * I + 127) / 128, 128, 0, stream>>>(x, dy, dw, db, B, I, O, act_in);
    JG_LAUNCH_CHECK();
  }
  return JG_OK;
}

extern "C" int jg_noise_pack_fwd(const float* y0, const float* ycond, const float* noise, const float* mask_f32,
                                 const int64_t* mask_i64, const float* gammas, void* out, int B, int C, int H, int W,
                                 int ld, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(y0 && ycond && noise && gammas && out && B > 0 && C > 0 && ld >= 2 * C && ld % 8 == 0, JG_ERR_INVALID,
           "noise_pack_fwd: bad args");
  const long long total = (long long)B * H * W;
  noise_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      y0, ycond, noise, mask_f32, reinterpret_cast<const long long*>(mask_i64), gammas,
      static_cast<__nv_bfloat16*>(out), B, C, H * W, ld);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_palette_loss_fwd(const float* noise, const void* noise_hat, int ld, const float* mask_f32,
                                   const int64_t* mask_i64, const float* w_b, int B, int C, int HW, float lambda_g,
                                   int l1, float* loss, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(noise && noise_hat && loss && B > 0 && C > 0 && ld >= C, JG_ERR_INVALID, "palette_loss_fwd: bad args");
  JG_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  const long long total = (long long)B * HW;
  const float coef = lambda_g / ((float)B * (float)C * (float)HW);
  palette_loss_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      noise, static_cast<const __nv_bfloat16*>(noise_hat), ld, mask_f32, reinterpret_cast<const long long*>(mask_i64),
      w_b, B, C, HW, coef, l1, loss);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_palette_loss_bwd(const float* noise, const void* noise_hat, int ld, const float* mask_f32,
                                   const int64_t* mask_i64, const float* w_b, int B, int C, int HW, float lambda_g,
                                   int l1, const float* grad_out, void* d_noise_hat, int ldd, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(noise && noise_hat && d_noise_hat && B > 0 && C > 0 && ld >= C && ldd >= C, JG_ERR_INVALID,
           "palette_loss_bwd: bad args");
  const long long total = (long long)B * HW;
  const float coef = lambda_g / ((float)B * (float)C * (float)HW);
  palette_loss_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      noise, static_cast<const __nv_bfloat16*>(noise_hat), ld, mask_f32, reinterpret_cast<const long long*>(mask_i64),
      w_b, B, C, HW, coef, l1, grad_out, static_cast<__nv_bfloat16*>(d_noise_hat), ldd);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int adamw, int step,
                                 int* step_dev, float grad_scale, float ema_beta, int ema_init,
                                 jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(p && g && m && v && n > 0 && (step > 0 || step_dev), JG_ERR_INVALID, "adamw_ema_step: bad args");
  if (step_dev) {
    step_increment_kernel<<<1, 1, 0, stream>>>(step_dev);
    JG_LAUNCH_CHECK();
  }
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  adamw_ema_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay,
                                                         adamw, step, step_dev, grad_scale, ema_beta, ema_init);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_ddpm_step(const void* eps, int lde, const float* y_t, const float* y_cond, const float* y_0,
                            const float* mask_f32, const int64_t* mask_i64, const float* noise, const float* coef,
                            float* y_next, void* x_next, int B, int C, int H, int W, int ld, int ddim,
                            jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(eps && y_t && y_cond && coef && y_next && B > 0 && C > 0 && lde >= C, JG_ERR_INVALID,
           "ddpm_step: null pointer / bad dims");
  JG_CHECK(!(mask_f32 || mask_i64) || y_0, JG_ERR_INVALID, "ddpm_step: a mask needs y_0");
  JG_CHECK(x_next == nullptr || (ld >= 2 * C && ld % 8 == 0), JG_ERR_INVALID, "ddpm_step: bad ld %d", ld);
  const long long total = (long long)B * H * W;
  ddpm_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(eps), lde, y_t, y_cond, y_0, mask_f32,
      reinterpret_cast<const long long*>(mask_i64), noise, coef, y_next, static_cast<__nv_bfloat16*>(x_next), B, C,
      H * W, ld, ddim);
  JG_LAUNCH_CHECK();
  return JG_OK;
}
