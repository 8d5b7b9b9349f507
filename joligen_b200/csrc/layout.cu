// Layout kernels: weight packing (fp32 OIHW master -> bf16 implicit-GEMM operands), wgrad
// unpacking, bias gradient, NCHW fp32 <-> NHWC bf16 boundary conversion, channel-slice copies,
// 2x nearest upsample / 2x2 average pool in NHWC.  All HBM-bound, 16-byte vector accesses.
#include "common.cuh"
#include "ptx.cuh"

namespace jg {

__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int RS, int Cin8,
                                   int Cout8) {
  const long long total_f = (long long)Cout * RS * Cin8;
  const long long total_d = wd ? (long long)Cin * RS * Cout8 : 0;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < total_f) {
    const int ci = (int)(i % Cin8);
    const int tap = (int)((i / Cin8) % RS);
    const int co = (int)(i / ((long long)Cin8 * RS));
    const float v = ci < Cin ? w[((long long)co * Cin + ci) * RS + tap] : 0.f;
    wf[i] = __float2bfloat16(v);
  } else if (i < total_f + total_d) {
    const long long j = i - total_f;
    const int co = (int)(j % Cout8);
    const int tap = (int)((j / Cout8) % RS);
    const int ci = (int)(j / ((long long)Cout8 * RS));
    const float v = co < Cout ? w[((long long)co * Cin + ci) * RS + (RS - 1 - tap)] : 0.f;
    wd[j] = __float2bfloat16(v);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin,
                                    int RS, float beta) {
  const long long total = (long long)Cout * Cin * RS;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes dst (OIHW): co, ci, tap
  const int tap = (int)(i % RS);
  const int ci = (int)((i / RS) % Cin);
  const int co = (int)(i / ((long long)RS * Cin));
  const float v = src[((long long)co * RS + tap) * Cin + ci];
  dst[i] = beta == 0.f ? v : beta * dst[i] + v;
}

// db[c] += sum_rows dy[row][c]: 8 channels per thread, block-level reduction in shared memory, then
// ONE global atomic per channel per block (grid ~ a few waves of the SMs).
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, long long rows, int C, int ld, float* __restrict__ db,
                 long long rows_per_block) {
  extern __shared__ float sm[];  // C floats
  const int vecs = C / 8;
  const int tid = threadIdx.x;
  const int v = tid % vecs;
  const int rlane = tid / vecs;
  const int rstep = blockDim.x / vecs;
  for (int i = tid; i < C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  if (rlane < rstep) {
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long r = r0 + rlane; r < r1; r += rstep) {
      const uint4 u = *reinterpret_cast<const uint4*>(dy + r * ld + v * 8);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sm[v * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = tid; i < C; i += blockDim.x) atomicAdd(db + i, sm[i]);
}

// NCHW fp32 -> NHWC bf16 (channels >= C zero-filled up to ld), 32x32 smem transpose tiles.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C,
                                    int HW, int ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    tile[j][tx] = (c < C && p < HW) ? src[((long long)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    if (p < HW && c < ld) dst[((long long)n * HW + p) * ld + c] = __float2bfloat16(tile[tx][j]);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C,
                                    int HW, int ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    tile[j][tx] = (p < HW && c < C) ? __bfloat162float(src[((long long)n * HW + p) * ld + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    if (c < C && p < HW) dst[((long long)n * C + c) * HW + p] = tile[tx][j];
  }
}

// dst[row][dst_off + c] = src[row][src_off + c] (optionally accumulated), 8 channels per thread.
__global__ void copy_channels_kernel(const __nv_bfloat16* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst,
                                     int ldd, long long rows, int C, int accumulate) {
  const int vecs = C / 8;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vecs;
    const int v = (int)(i - r * vecs);
    uint4 u = *reinterpret_cast<const uint4*>(src + r * lds + v * 8);
    uint4* dp = reinterpret_cast<uint4*>(dst + r * ldd + v * 8);
    if (accumulate) {
      const uint4 o = *dp;
      const uint32_t* a = &u.x;
      const uint32_t* b = &o.x;
      uint32_t res[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(a[j]), y = unpack_bf16x2(b[j]);
        res[j] = pack_bf16x2(x.x + y.x, x.y + y.y);
      }
      u = make_uint4(res[0], res[1], res[2], res[3]);
    }
    *dp = u;
  }
}

// mode 0: nearest 2x upsample  (dst is [N,2H,2W,C]);  mode 1: 2x2 average pool (dst [N,H/2,W/2,C]);
// mode 2: 2x2 sum "un-upsample" = backward of mode 0;  mode 3: backward of avg pool (dst [N,2H,2W,C], 0.25*src)
__global__ void resample2x_kernel(const __nv_bfloat16* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst,
                                  int ldd, int N, int Hs, int Ws, int C, int mode) {
  const int vecs = C / 8;
  const bool up = (mode == 0 || mode == 3);
  const int Hd = up ? Hs * 2 : Hs / 2;
  const int Wd = up ? Ws * 2 : Ws / 2;
  const long long total = (long long)N * Hd * Wd * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long pix = i / vecs;
    const int w = (int)(pix % Wd);
    pix /= Wd;
    const int h = (int)(pix % Hd);
    const int n = (int)(pix / Hd);
    float f[8];
    if (up) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + (((long long)n * Hs + h / 2) * Ws + w / 2) * lds + v * 8);
      const uint32_t* a = &u.x;
      const float sc = mode == 3 ? 0.25f : 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(a[j]);
        f[2 * j] = x.x * sc;
        f[2 * j + 1] = x.y * sc;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
          const uint4 u = *reinterpret_cast<const uint4*>(
              src + (((long long)n * Hs + 2 * h + dh) * Ws + 2 * w + dw) * lds + v * 8);
          const uint32_t* a = &u.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = unpack_bf16x2(a[j]);
            f[2 * j] += x.x;
            f[2 * j + 1] += x.y;
          }
        }
      if (mode == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= 0.25f;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + (((long long)n * Hd + h) * Wd + w) * ldd + v * 8) = o;
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}


// ======================================================================================================
// Batched weight kernels: ONE launch for all convolutions of a model (a step used to spend ~1.5 ms in 75 pack
// launches, 58 unpack launches and 58 memsets, each a tiny transposing kernel).  Work unit = a tile of 32 output
// channels x 32 input channels x up to 9 filter taps staged in shared memory, so that global reads AND writes are
// runs of contiguous elements in every layout involved.  tile_start[i] = first tile of item i (prefix sums).
// ======================================================================================================
constexpr int kWT = 32;        // tile edge in channels
constexpr int kWTaps = 9;      // taps staged per tile

__device__ __forceinline__ int find_item(const int* __restrict__ tile_start, int n, int tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last i with tile_start[i] <= tile
    const int mid = (lo + hi + 1) >> 1;
    if (tile_start[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Thread layout of both kernels: 256 threads = 8 warps; `lane` runs along the contiguous dimension of whatever is being
// read or written, `wrow` = warp index picks one of 8 rows per pass (4 passes cover the 32-wide tile).  All loops have
// compile-time bounds (taps are predicated against nt), so the 36 loads of a phase are independent and in flight together.

// Tile geometry.  R*S > 1: 32 co x 32 ci x up to 9 taps (slot = tap).  1x1 convolutions (every Linear of the video
// UNet's MotionModule) have a single tap, so their tiles take 9 consecutive 32-channel ci blocks instead (slot = ci
// block): the same 9216-element work unit instead of a ninth of it.
struct WTile {
  int co0, ci0, t0, nt, RS, Cout, Cin;
  bool wide;
  __device__ __forceinline__ int ci_of(int slot, int x) const { return wide ? ci0 + slot * kWT + x : ci0 + x; }
  __device__ __forceinline__ int tap_of(int slot) const { return wide ? 0 : t0 + slot; }
};
__device__ __forceinline__ WTile make_tile(int tile, int Cout, int Cin, int RS) {
  WTile g;
  g.RS = RS; g.Cout = Cout; g.Cin = Cin;
  g.wide = RS == 1;
  const int ci_span = g.wide ? kWT * kWTaps : kWT;
  const int tchunks = g.wide ? 1 : (RS + kWTaps - 1) / kWTaps;
  const int ci_tiles = (Cin + ci_span - 1) / ci_span;
  const int tc = tile % tchunks; tile /= tchunks;
  const int cit = tile % ci_tiles;
  g.co0 = (tile / ci_tiles) * kWT;
  g.ci0 = cit * ci_span;
  g.t0 = tc * kWTaps;
  g.nt = g.wide ? min(kWTaps, (Cin - g.ci0 + kWT - 1) / kWT) : min(kWTaps, RS - g.t0);
  return g;
}

// OIHW side of a tile: for a fixed co the (ci, tap) elements of the tile are one contiguous run of 32*nt floats (when
// the chunk covers all taps); pos enumerates that run (tap fastest, as in memory).
struct OihwIter {
  int ci, tl;   // channel within the slot's 32-wide block, slot
  bool ok;
  size_t off;
};
__device__ __forceinline__ OihwIter oihw_pos(const WTile& g, int pos, int co) {
  OihwIter r;
  if (g.wide) {
    r.tl = pos >> 5;
    r.ci = pos & 31;
    r.ok = (g.co0 + co < g.Cout) && (g.ci0 + pos < g.Cin);
    r.off = (size_t)(g.co0 + co) * g.Cin + g.ci0 + pos;
  } else {
    r.ci = pos / g.nt;
    r.tl = pos - r.ci * g.nt;
    r.ok = (g.co0 + co < g.Cout) && (g.ci0 + r.ci < g.Cin) && r.ci < kWT;
    r.off = ((size_t)(g.co0 + co) * g.Cin + g.ci0 + r.ci) * g.RS + g.t0 + r.tl;
  }
  return r;
}

// fp32 OIHW master -> bf16 wf [Cout8][RS][Cin8] and wd [Cin8][RS][Cout8] (taps flipped)
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const jg_pack_item* __restrict__ items, const int* __restrict__ tile_start, int n) {
  __shared__ float sm[kWTaps][kWT][kWT + 1];  // [slot][co][ci]
  const int it = find_item(tile_start, n, blockIdx.x);
  const jg_pack_item p = items[it];
  const WTile g = make_tile(blockIdx.x - tile_start[it], p.Cout, p.Cin, p.RS);
  const int RS = p.RS, nt = g.nt;
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  // phase 1: OIHW runs -> smem
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = wrow + 8 * k;
    float v[kWTaps];
#pragma unroll
    for (int c = 0; c < kWTaps; ++c) {
      const OihwIter q = oihw_pos(g, lane + 32 * c, co);
      v[c] = (c < nt && q.ok) ? p.w[q.off] : 0.f;
    }
#pragma unroll
    for (int c = 0; c < kWTaps; ++c) {
      const OihwIter q = oihw_pos(g, lane + 32 * c, co);
      if (c < nt && q.ci < kWT) sm[q.tl][co][q.ci] = v[c];
    }
  }
  __syncthreads();
  __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(p.wf);
  __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(p.wd);
  // phase 2a: wf[co][tap][ci], lane = ci
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = wrow + 8 * k;
    if (g.co0 + co < p.Cout) {
#pragma unroll
      for (int tl = 0; tl < kWTaps; ++tl)
        if (tl < nt && g.ci_of(tl, lane) < p.Cin)
          wf[((size_t)(g.co0 + co) * RS + g.tap_of(tl)) * p.Cin8 + g.ci_of(tl, lane)] =
              __float2bfloat16(sm[tl][co][lane]);
    }
  }
  // phase 2b: wd[ci][RS-1-tap][co], lane = co
  if (wd) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ci = wrow + 8 * k;
      if (g.co0 + lane < p.Cout) {
#pragma unroll
        for (int tl = 0; tl < kWTaps; ++tl)
          if (tl < nt && g.ci_of(tl, ci) < p.Cin)
            wd[((size_t)g.ci_of(tl, ci) * RS + (RS - 1 - g.tap_of(tl))) * p.Cout8 + g.co0 + lane] =
                __float2bfloat16(sm[tl][lane][ci]);
      }
    }
  }
}

// dw_oihw += acc (layout 0: [RS][Cin][Cout], 1: [Cout][RS][Cin]);  acc = 0 (ready for the next step's split-K sums)
__global__ void __launch_bounds__(256)
wgrad_unpack_batched_kernel(const jg_unpack_item* __restrict__ items, const int* __restrict__ tile_start, int n) {
  __shared__ float sm[kWTaps][kWT][kWT + 1];  // [slot][co][ci]
  const int it = find_item(tile_start, n, blockIdx.x);
  const jg_unpack_item p = items[it];
  const WTile g = make_tile(blockIdx.x - tile_start[it], p.Cout, p.Cin, p.RS);
  const int RS = p.RS, nt = g.nt;
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  // phase 1: raw accumulator -> smem (and zero it)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = wrow + 8 * k;
    float v[kWTaps];
    size_t off[kWTaps];
    bool ok[kWTaps];
#pragma unroll
    for (int tl = 0; tl < kWTaps; ++tl) {
      if (p.layout == 0) {  // [tap][ci][co]: lane = co, row = ci
        ok[tl] = tl < nt && (g.ci_of(tl, row) < p.Cin) && (g.co0 + lane < p.Cout);
        off[tl] = ((size_t)g.tap_of(tl) * p.Cin + g.ci_of(tl, row)) * p.Cout + g.co0 + lane;
      } else {              // [co][tap][ci]: lane = ci, row = co
        ok[tl] = tl < nt && (g.co0 + row < p.Cout) && (g.ci_of(tl, lane) < p.Cin);
        off[tl] = ((size_t)(g.co0 + row) * RS + g.tap_of(tl)) * p.Cin + g.ci_of(tl, lane);
      }
    }
#pragma unroll
    for (int tl = 0; tl < kWTaps; ++tl) v[tl] = ok[tl] ? p.acc[off[tl]] : 0.f;
#pragma unroll
    for (int tl = 0; tl < kWTaps; ++tl) {
      if (tl < nt) {
        if (ok[tl]) p.acc[off[tl]] = 0.f;
        if (p.layout == 0) sm[tl][lane][row] = v[tl]; else sm[tl][row][lane] = v[tl];
      }
    }
  }
  __syncthreads();
  // phase 2: OIHW runs, read-modify-write
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = wrow + 8 * k;
    float v[kWTaps];
#pragma unroll
    for (int c = 0; c < kWTaps; ++c) {
      const OihwIter q = oihw_pos(g, lane + 32 * c, co);
      v[c] = (c < nt && q.ok) ? p.dw[q.off] : 0.f;
    }
#pragma unroll
    for (int c = 0; c < kWTaps; ++c) {
      const OihwIter q = oihw_pos(g, lane + 32 * c, co);
      if (c < nt && q.ok) p.dw[q.off] = v[c] + sm[q.tl][co][q.ci];
    }
  }
}

}  // namespace jg

using namespace jg;

extern "C" int jg_pack_conv_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int R, int S,
                                   jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(w && wf && Cout > 0 && Cin > 0 && R > 0 && S > 0, JG_ERR_INVALID, "pack_conv_weight: bad args");
  const int RS = R * S, Cin8 = (Cin + 7) / 8 * 8, Cout8 = (Cout + 7) / 8 * 8;
  const long long total = (long long)Cout * RS * Cin8 + (wd ? (long long)Cin * RS * Cout8 : 0);
  const int block = 256;
  const long long grid = (total + block - 1) / block;
  pack_weight_kernel<<<(unsigned)grid, block, 0, stream>>>(w, static_cast<__nv_bfloat16*>(wf),
                                                           static_cast<__nv_bfloat16*>(wd), Cout, Cin, RS, Cin8,
                                                           Cout8);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_unpack_conv_wgrad(const float* src, float* dst, int Cout, int Cin, int R, int S, float beta,
                                    jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && Cout > 0 && Cin > 0, JG_ERR_INVALID, "unpack_conv_wgrad: bad args");
  const long long total = (long long)Cout * Cin * R * S;
  const int block = 256;
  unpack_wgrad_kernel<<<(unsigned)((total + block - 1) / block), block, 0, stream>>>(src, dst, Cout, Cin, R * S,
                                                                                    beta);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_bias_grad(const void* dy, int64_t rows, int C, int ld, float* db, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(dy && db && rows > 0 && C > 0 && C % 8 == 0 && ld % 8 == 0 && ld >= C, JG_ERR_INVALID,
           "bias_grad: bad args (C=%d ld=%d)", C, ld);
  JG_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, stream));
  long long blocks = (long long)num_sms() * 8;
  long long rows_per_block = (rows + blocks - 1) / blocks;
  if (rows_per_block < 64) rows_per_block = 64;
  const int grid = (int)((rows + rows_per_block - 1) / rows_per_block);
  // one thread column per 8 channels, 256 columns per block: wider tensors go in 2048-channel slabs
  for (int c0 = 0; c0 < C; c0 += 2048) {
    const int cs = C - c0 < 2048 ? C - c0 : 2048;
    bias_grad_kernel<<<grid, 256, cs * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(dy) + c0, rows, cs,
                                                                ld, db + c0, rows_per_block);
    JG_LAUNCH_CHECK();
  }
  return JG_OK;
}

extern "C" int jg_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int ld,
                                        jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && N > 0 && C > 0 && ld >= C && ld % 8 == 0, JG_ERR_INVALID, "nchw_to_nhwc: bad args");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (ld + 31) / 32, N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), C, HW, ld);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int ld,
                                        jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && N > 0 && C > 0 && ld >= C, JG_ERR_INVALID, "nhwc_to_nchw: bad args");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), dst, C, HW, ld);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_copy_channels(const void* src, int lds, void* dst, int ldd, int64_t rows, int C, int accumulate,
                                jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && rows > 0 && C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, JG_ERR_INVALID,
           "copy_channels: bad args");
  const long long total = rows * (C / 8);
  copy_channels_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), lds,
                                                                 static_cast<__nv_bfloat16*>(dst), ldd, rows, C,
                                                                 accumulate);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_resample2x(const void* src, int lds, void* dst, int ldd, int N, int Hs, int Ws, int C, int mode,
                             jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(src && dst && N > 0 && C > 0 && C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && mode >= 0 && mode <= 3,
           JG_ERR_INVALID, "resample2x: bad args");
  const bool up = (mode == 0 || mode == 3);
  JG_CHECK(up || (Hs % 2 == 0 && Ws % 2 == 0), JG_ERR_INVALID, "resample2x: odd size %dx%d", Hs, Ws);
  const long long total = (long long)N * (up ? Hs * 2 : Hs / 2) * (up ? Ws * 2 : Ws / 2) * (C / 8);
  resample2x_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), lds,
                                                              static_cast<__nv_bfloat16*>(dst), ldd, N, Hs, Ws, C,
                                                              mode);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_pack_conv_weights_batched(const jg_pack_item* items_dev, const int* tile_start_dev, int n,
                                            int total_tiles, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(items_dev && tile_start_dev && n > 0 && total_tiles > 0, JG_ERR_INVALID, "pack_conv_weights_batched: bad args");
  pack_weights_batched_kernel<<<total_tiles, 256, 0, stream>>>(items_dev, tile_start_dev, n);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_wgrad_unpack_batched(const jg_unpack_item* items_dev, const int* tile_start_dev, int n,
                                       int total_tiles, jg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  JG_CHECK(items_dev && tile_start_dev && n > 0 && total_tiles > 0, JG_ERR_INVALID, "wgrad_unpack_batched: bad args");
  wgrad_unpack_batched_kernel<<<total_tiles, 256, 0, stream>>>(items_dev, tile_start_dev, n);
  JG_LAUNCH_CHECK();
  return JG_OK;
}

extern "C" int jg_weight_tiles(int Cout, int Cin, int RS) {
  if (RS == 1) return ((Cout + 31) / 32) * ((Cin + 32 * 9 - 1) / (32 * 9));  // 1x1: 9 ci blocks per tile (make_tile)
  return ((Cout + 31) / 32) * ((Cin + 31) / 32) * ((RS + 8) / 9);
}
