/*
 * jg_b200.h — C ABI of libjg_b200.so: hand-written sm_100a kernels for the joliGEN training
 * inner loop (Palette diffusion UNet + GAN generator/discriminator), see DESIGN.md.
 *
 * Conventions
 *   - every entry point returns 0 on success, <0 on error (JG_ERR_*); the message is available
 *     from jg_last_error() (thread-local).  There is NO CPU fallback: unsupported shapes are errors.
 *   - all pointers are DEVICE pointers unless the name ends in _host; memory is borrowed for the
 *     duration of the call (the caller's allocator owns it); launches are asynchronous on `stream`.
 *   - activations are NHWC bf16 ("channels-last"); `ld*` arguments are channel strides in ELEMENTS
 *     so that an operand may be a channel slice of a wider (concatenated) buffer.  Channel counts,
 *     channel offsets and ld* must be multiples of 8 (16-byte TMA alignment).
 *   - weights: fp32 master copies stay in the reference layout (OIHW, what joliGEN's
 *     state_dict holds); the kernels consume bf16 packed copies produced by jg_pack_conv_weight.
 *
 * Each function cites the reference code (joliGEN @ /root/reference) whose arithmetic it replaces.
 */
#ifndef JG_B200_H
#define JG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* jg_stream_t; /* cudaStream_t */

enum {
  JG_OK = 0,
  JG_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  JG_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  JG_ERR_UNSUPPORTED = -3, /* device is not sm_100 */
};

enum { JG_ACT_NONE = 0, JG_ACT_RELU = 1, JG_ACT_LRELU02 = 2, JG_ACT_TANH = 3, JG_ACT_SILU = 4 };

const char* jg_last_error(void);
int jg_version(void);
/* 0 if the current device can run the library (compute capability 10.x), JG_ERR_UNSUPPORTED otherwise */
int jg_check_device(void);
/* Number of CUDA kernels this library has launched so far in the process (every launch site counts itself);
 * bench.py's `gpu_launches` is the difference over one step. */
unsigned long long jg_kernel_launches(void);

/* ---------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM on tcgen05 tensor cores (TMA-staged NHWC tiles, fp32 accumulation
 * in TMEM).  Replaces nn.Conv2d / nn.Conv1d(k=1) / nn.Linear on feature maps:
 *   models/modules/unet_generator_attn/unet_generator_attn.py:186-190,208-220,481-483,639-643
 *   (ResBlock convs, skip 1x1, UNet in/out convs), :298,:305 (attention qkv / proj_out),
 *   models/modules/resnet_architecture/resnet_generator.py:11-95, models/modules/discriminators.py:10-117.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int N, H, W;      /* input batch / height / width */
  int Cin, ldx;     /* input channels and channel stride of x */
  int Ho, Wo;       /* output height / width */
  int Cout, ldy;    /* output channels and channel stride of y */
  int R, S;         /* filter height / width */
  int stride;       /* 1 or 2 (same in h and w) */
  int pad;          /* zero padding (same in h and w) */
  int up2x;         /* 1: the conv reads a virtual nearest-2x-upsampled x (H,W are the upsampled dims) — reserved */
  int act;          /* JG_ACT_* applied after bias (+ residual) */
  int ldres;        /* channel stride of the residual operand (0 = none) */
  float res_scale;  /* y = act(conv + bias + res_scale * residual) */
} jg_conv_desc;

/* y[N,Ho,Wo,Cout] = act(conv(x, w) + bias + res_scale*residual).
 * w_packed: bf16 [Cout][R*S][Cin8] with Cin8 = round_up(Cin, 8) (jg_pack_conv_weight, fwd layout).
 * bias: fp32 [Cout] or NULL.  residual: bf16 NHWC like y with stride ldres, or NULL.
 * Also used for dgrad of stride-1 convolutions (pass the dgrad-packed weights; Cin/Cout swapped). */
int jg_conv2d_fwd(const jg_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                  const void* residual, void* y, jg_stream_t stream);

/* GroupNorm work fused into the convolution's epilogue ("emit partial sum x, sum x^2 from the
 * producing conv's epilogue so GN becomes a single read-modify-write pass"), all optional (NULL = off):
 *   stats   fp32 [N][Cout][2], ACCUMULATED into (zero it first): per-(image, channel) sum and sum of squares of the
 *           stored (bf16-rounded) output y — the statistics pass of the GroupNorm that reads y next
 *           (unet_attn_utils.py:42-48 computes them in a separate pass over y).
 *   gn_sums fp32 [N][Cout][2], ACCUMULATED into: this call is the DGRAD of the convolution that follows
 *           act(a*x + b), a / b = the fused per-(n,c) GroupNorm(+FiLM) coefficients (jg_groupnorm_fwd's `ab`), so its
 *           output y IS dL/d(act output); A[n,c] += sum_pixels du, B[n,c] += sum_pixels du*x with
 *           du = y * act'(a*x + b): the reduction pass of jg_groupnorm_bwd (pass the result as `sums_pre`).
 *           gn_x = the GroupNorm's input x (bf16 NHWC, channel stride ldgx), gn_ab = [N][Cout][2], gn_act = JG_ACT_*.
 *           Not combinable with a residual operand.
 * Kernels whose epilogue cannot fuse the work (tiles spanning several images, Cout <= 32) run the equivalent
 * stand-alone reduction right after the convolution: the outputs are always filled. */
typedef struct {
  float* stats;
  float* gn_sums;
  const void* gn_x;
  int ldgx;
  const float* gn_ab;
  int gn_act;
} jg_conv_epilogue;
int jg_conv2d_fwd_ex(const jg_conv_desc* d, const jg_conv_epilogue* e, const void* x, const void* w_packed,
                     const float* bias, const void* residual, void* y, jg_stream_t stream);

/* dw_oihw[Cout][Cin][R][S] = beta*dw_oihw + sum over pixels dy (x) x  (fp32, the reference's weight layout).
 * ws: fp32 workspace of R*S*Cin*Cout floats (zeroed by the call; split-K partial tiles are accumulated
 * into it with fp32 atomics, then permuted to OIHW).  d describes the FORWARD conv. */
int jg_conv2d_wgrad(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* ws, float* dw_oihw,
                    float beta, jg_stream_t stream);

/* fp32 OIHW master weight -> bf16 packed copies.
 *   w_fwd   [Cout][R*S][Cin8]   (B operand of the forward implicit GEMM)
 *   w_dgrad [Cin][R*S][Cout8]   taps flipped (B operand of the stride-1 dgrad), may be NULL */
int jg_pack_conv_weight(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int R, int S,
                        jg_stream_t stream);
/* Raw split-K accumulation for a trainer that keeps one persistent fp32 accumulator per convolution: adds this
 * call's partial sums into acc (R*S*Cin*Cout floats; NOT zeroed, NOT permuted) and reports the accumulator layout in
 * *layout (0: [R*S][Cin][Cout], 1: [Cout][R*S][Cin]; fixed for a given descriptor). */
int jg_conv2d_wgrad_acc(const jg_conv_desc* d, const void* x, const void* dy, int lddy, float* acc, int* layout,
                        jg_stream_t stream);

/* Batched weight kernels: ONE launch for all convolutions of a model.  items_dev / tile_start_dev are DEVICE arrays;
 * tile_start_dev[i] = first tile of item i, with jg_weight_tiles(Cout, Cin, R*S) tiles per item.
 *   jg_pack_conv_weights_batched: jg_pack_conv_weight for n convolutions (w_dgrad may be NULL per item).
 *   jg_wgrad_unpack_batched: dw_oihw += acc (permuted from its layout), then acc = 0 — the end of the backward pass
 *   for accumulators filled by jg_conv2d_wgrad_acc. */
typedef struct {
  const float* w;  /* fp32 OIHW */
  void* wf;        /* bf16 [Cout8][R*S][Cin8] */
  void* wd;        /* bf16 [Cin8][R*S][Cout8], taps flipped, or NULL */
  int Cout, Cin, RS, Cin8, Cout8, pad_;
} jg_pack_item;
typedef struct {
  float* acc;  /* raw accumulator */
  float* dw;   /* fp32 OIHW gradient, accumulated into */
  int Cout, Cin, RS, layout;
} jg_unpack_item;
int jg_weight_tiles(int Cout, int Cin, int RS);
int jg_pack_conv_weights_batched(const jg_pack_item* items_dev, const int* tile_start_dev, int n, int total_tiles,
                                 jg_stream_t stream);
int jg_wgrad_unpack_batched(const jg_unpack_item* items_dev, const int* tile_start_dev, int n, int total_tiles,
                            jg_stream_t stream);
/* fp32 OHWI wgrad accumulator -> fp32 OIHW gradient (dst = beta*dst + src). */
int jg_unpack_conv_wgrad(const float* dw_ohwi, float* dw_oihw, int Cout, int Cin, int R, int S, float beta,
                         jg_stream_t stream);
/* db[c] = sum over rows of dy[rows][ld] (bias gradient), fp32, overwritten. */
int jg_bias_grad(const void* dy, int64_t rows, int C, int ld, float* db, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Gradient exchange: a communicator owned by the library, one per process / GPU,
 * bound to NCCL at run time.  Replaces DistributedDataParallel's bucketed, backward-overlapped all-reduce
 * (models/base_model.py:725-737) for the trainers of this package.
 *   jg_comm_unique_id   rank 0 creates the 128-byte id; the caller ships it to the other ranks (torch.distributed
 *                       broadcast, MPI, a file ...).  HOST pointer.
 *   jg_comm_init        collective over all ranks; uses the calling thread's current device.
 *   jg_comm_allreduce_async  in-place SUM of buf[count] (dtype 0 = fp32, 1 = bf16) on the communicator's own stream,
 *                       ordered after everything `compute_stream` has been given so far; returns at once, so the
 *                       caller keeps launching backward kernels while the bucket is in flight.
 *   jg_comm_wait        `compute_stream` waits for every collective issued so far (no host synchronisation).
 *   Both work inside a CUDA-graph capture of `compute_stream` (the collective becomes a forked branch of the graph).
 *   jg_comm_broadcast   in-place byte broadcast from `root`, complete on `compute_stream` order.
 * ------------------------------------------------------------------------------------------- */
typedef struct jg_comm* jg_comm_t;
int jg_comm_unique_id(void* id128_host);
int jg_comm_init(const void* id128_host, int rank, int world, jg_comm_t* out);
int jg_comm_allreduce_async(jg_comm_t c, void* buf, size_t count, int dtype, jg_stream_t compute_stream);
int jg_comm_broadcast(jg_comm_t c, void* buf, size_t bytes, int root, jg_stream_t compute_stream);
int jg_comm_wait(jg_comm_t c, jg_stream_t compute_stream);
int jg_comm_info(jg_comm_t c, int* rank, int* world, unsigned long long* collectives, unsigned long long* bytes,
                 int* nccl_version);
int jg_comm_destroy(jg_comm_t c);

/* ---------------------------------------------------------------------------------------------
 * Boundary / layout kernels (HBM-bound).
 * ------------------------------------------------------------------------------------------- */
/* NCHW fp32 (reference layout, e.g. UNet.compute_feats input `h = input.type(torch.float32)`,
 * unet_generator_attn.py:670) -> NHWC bf16 with channel stride ld (channels C..ld-1 zero-filled). */
int jg_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int H, int W, int ld,
                             jg_stream_t stream);
int jg_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int H, int W, int ld,
                             jg_stream_t stream);
/* dst[row][0..C) (=|+=) src[row][0..C): channel-slice copy; torch.cat([h, hs.pop()], dim=1)
 * (unet_generator_attn.py:687) and its backward. */
int jg_copy_channels(const void* src, int lds, void* dst, int ldd, int64_t rows, int C, int accumulate,
                     jg_stream_t stream);
/* mode 0: F.interpolate(x, scale_factor=2, mode="nearest") (Upsample, unet_generator_attn.py:84-93)
 * mode 1: nn.AvgPool2d(2,2) (Downsample, :125-140); mode 2 / 3: their backward passes. */
int jg_resample2x(const void* src, int lds, void* dst, int ldd, int N, int Hs, int Ws, int C, int mode,
                  jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * On-GPU input preparation and Haar wavelets, fp32 NCHW images.  HBM-bound.
 *   jg_fill_mask_random      data/online_creation.py:1366-1376 fill_mask_with_random(img, mask, cls):
 *                            out = img*(1-m) + noise*m, m = (mask != 0) if cls == -1 else (mask == cls); the mask
 *                            ([N][1][H][W], fp32 or int64 — pass exactly one) is compared exactly.
 *   jg_u8_to_f32_normalized  transforms.ToTensor + transforms.Normalize(mean, std) (data/base_dataset.py
 *                            get_transform): uint8 [N][H][W][C] -> fp32 [N][C][H][W], (x/255 - mean)/std.
 *   jg_mask_class_dropout    models/palette_model.py:565-584: mask <- fill (= num_classes - 1) for the samples with
 *                            drop_u[n] < prob, unchanged elsewhere (per_image = elements of one sample's mask).
 *   jg_haar                  models/modules/freq_utils.py:22-59 on upfirdn2d (models/modules/op/upfirdn2d.py:167-208,
 *                            upfirdn2d_kernel.cu:49-200).  (h, w) is the LOW-resolution size.
 *                            mode 0 HaarTransform: x [N][C][2h][2w] -> [N][4C][h][w] = cat(ll, lh, hl, hh)
 *                            mode 1 its backward;  mode 2 InverseHaarTransform: [N][4C][h][w] -> [N][C][2h][2w]
 *                            mode 3 its backward.
 * ------------------------------------------------------------------------------------------- */
int jg_fill_mask_random(const float* img, const float* mask_f32, const int64_t* mask_i64, const float* noise,
                        float* out, int N, int C, int HW, int cls, jg_stream_t stream);
int jg_u8_to_f32_normalized(const uint8_t* src_nhwc, float* dst_nchw, int N, int C, int H, int W, float mean,
                            float stdv, jg_stream_t stream);
int jg_mask_class_dropout(const float* mask_f32, const int64_t* mask_i64, const float* drop_u, float prob,
                          int64_t fill, float* out_f32, int64_t* out_i64, int N, int64_t per_image,
                          jg_stream_t stream);
int jg_haar(const float* src, float* dst, int N, int C, int h, int w, int mode, jg_stream_t stream);
/* Per-pixel label embedding of PaletteDenoiseFn ("mask" conditioning, palette_denoise_fn.py:118-136):
 * out[row][col0 .. col0+E) = bf16(table[idx[row]][:]) written into an NHWC bf16 tensor of channel stride ld (the
 * embedding channels the reference concatenates to the UNet input); idx = the semantic mask, int64 or fp32 (exactly
 * one).  Backward: dtable [K][E] and counts [K] (both overwritten) = the row-wise scatter-sum of d and the label
 * frequencies (nn.Embedding(scale_grad_by_freq=True) divides by them). */
int jg_embed_rows(const float* table, const float* idx_f32, const int64_t* idx_i64, void* out, int ld, int col0,
                  int64_t rows, int E, int K, jg_stream_t stream);
int jg_embed_rows_bwd(const void* d, int ld, int col0, const float* idx_f32, const int64_t* idx_i64, int64_t rows,
                      int E, int K, float* dtable, float* counts, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GroupNorm (+ FiLM scale/shift) (+ SiLU), NHWC bf16, fp32 statistics.  HBM-bound.
 *   unet_attn_utils.py:42-48 (GroupNorm wrapper, fp32 compute, eps 1e-5),
 *   unet_generator_attn.py:186-189 (GN -> SiLU), :250-258 (out_norm(h)*(1+scale)+shift -> SiLU),
 *   unet_attn_utils.py:60-66,116-117 (attention InstanceNorm1d: groups == C, gamma/beta NULL).
 * film: fp32 [N][2C] = (scale | shift) = emb_layers output (torch.chunk(emb_out, 2, dim=1)) or NULL.
 * stats [N][groups][2] = (mean, rstd) and ab [N][C][2] are outputs of fwd and inputs of bwd.
 * ------------------------------------------------------------------------------------------- */
size_t jg_groupnorm_fwd_ws_floats(int N, int C, int groups);
size_t jg_groupnorm_bwd_ws_floats(int N, int C, int groups);
int jg_groupnorm_fwd(const void* x, int ldx, void* y, int ldy, int N, int HW, int C, int groups, float eps,
                     const float* gamma, const float* beta, const float* film, int act, float* stats, float* ab,
                     float* ws, const float* chan_stats, jg_stream_t stream);
/* chan_stats (above; may be NULL): per-(n, channel) sum / sum of squares of x, fp32 [N][C][2], already accumulated by
 * the producer of x (jg_conv_epilogue.stats) or by jg_chan_stats: the statistics pass over x is skipped. */
int jg_chan_stats(const void* x, int ldx, int N, int HW, int C, float* stats /* accumulated into */,
                  jg_stream_t stream);
/* dx = d/dx (+ addend (+ addend2), NHWC bf16 tensors like dx with strides ldadd / ldadd2, or NULL: the gradients of
 * other consumers of x — the ResBlock's skip path, the decoder's concat (unet_generator_attn.py:687) — summed in the
 * same pass; addend may alias dx; addend2 requires addend);
 * dgamma/dbeta [C] overwritten (may be NULL); dfilm [N][2C] overwritten (may be NULL);
 * dx_colsum [C] (may be NULL) receives sum over (n, pixel) of dx: x is the output of a conv, so this IS that conv's
 * bias gradient (nn.Conv2d bias, unet_generator_attn.py:190,207) and saves the separate pass over dx. */
int jg_groupnorm_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, const void* addend,
                     int ldadd, const void* addend2, int ldadd2, int N,
                     int HW, int C, int groups, const float* gamma, const float* beta, const float* film, int act,
                     const float* stats, const float* ab, float* dgamma, float* dbeta, float* dfilm,
                     float* dx_colsum, float* ws, const float* sums_pre, jg_stream_t stream);
/* sums_pre (may be NULL): fp32 [N][C][2] = (sum du, sum du*x) per (n, channel), already accumulated by the dgrad that
 * produced dy (jg_conv_epilogue.gn_sums): the reduction pass over (x, dy) is skipped. */

/* ---------------------------------------------------------------------------------------------
 * Spatial self-attention (flash style, T x T never materialised), bf16, fp32 softmax.
 *   QKVAttentionLegacy.forward, unet_generator_attn.py:331-347 (per-head interleaved q|k|v channels,
 *   scale ch^-1/4 on q and k, softmax in fp32).  qkv NHWC [N][T][3*heads*ch]; out [N][T][heads*ch];
 *   lse fp32 [N*heads][T] (log2 domain) is saved for the backward.  T % 64 == 0, ch in {16,32,64}.
 * ------------------------------------------------------------------------------------------- */
/* layout 0: per-head (q|k|v) channel interleave (QKVAttentionLegacy, unet_generator_attn.py:331-347);
 * layout 1: (q | k | v), each heads*ch wide (QKVAttention, unet_generator_attn_vid.py:334-363: the video UNet's
 * default, use_new_attention_order=True). */
int jg_attn_fwd(const void* qkv, int ldqkv, void* out, int ldo, float* lse, int N, int T, int heads, int ch,
                int layout, jg_stream_t stream);
/* ws: N*heads*T floats.  dqkv has the layout of qkv. */
int jg_attn_bwd(const void* qkv, int ldqkv, const void* out, int ldo, const void* d_out, int lddo, const float* lse,
                void* dqkv, int lddqkv, float* ws, int N, int T, int heads, int ch, int layout, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MotionModule of the video UNet (unet_generator_attn_vid.py:374-590), tokens = NHWC pixels of N = B*F frames
 * (frame index = n % F).  The Linear layers of the module are 1x1 convolutions (jg_conv2d_*).
 *   jg_layernorm_*      nn.LayerNorm(C) of TemporalTransformerBlock.norms / ff_norm (:555-563, :574-588), eps 1e-5;
 *                       pe fp32 [F][C] or NULL = PositionalEncoding.pe[0, :F] (:932-947) added after the affine,
 *                       i.e. VersatileAttention's pos_encoder on the "(b d) f c" view (:993-999);
 *                       stats fp32 [rows][2] = (mean, rstd) saved for the backward.
 *   jg_temporal_attn_*  VersatileAttention over the F <= 8 frames of each pixel (:978-1054, _attention :758-793):
 *                       qkv [B*F][HW][(q|k|v)] from to_q/to_k/to_v packed as one GEMM, scale dim_head^-1/2,
 *                       fp32 softmax; out [B*F][HW][heads*ch]; the backward recomputes the F x F probabilities.
 *   jg_geglu_*          GEGLU (:908-929): x = (a | gate) [rows][2*Cout] -> a * gelu(gate) (erf GELU).
 * ------------------------------------------------------------------------------------------- */
int jg_layernorm_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int C, float eps, const float* gamma,
                     const float* beta, const float* pe, int HW, int F, float* stats, jg_stream_t stream);
/* dgamma / dbeta [C] overwritten.  addend (optional, bf16 rows of stride ldadd): added to dx — the gradient of the
 * residual branch x also feeds (BasicTransformerBlock: norm(x) -> attn/ff -> + x, unet_generator_attn_vid.py:565-590).
 * dx_colsum (optional, [C], overwritten): sum over rows of dx = the bias gradient of the Linear that produced x. */
int jg_layernorm_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, const void* addend,
                     int ldadd, int64_t rows, int C, const float* gamma, const float* stats, float* dgamma,
                     float* dbeta, float* dx_colsum, jg_stream_t stream);
int jg_temporal_attn_fwd(const void* qkv, int ldqkv, void* out, int ldo, int B, int F, int HW, int heads, int ch,
                         jg_stream_t stream);
int jg_temporal_attn_bwd(const void* qkv, int ldqkv, const void* d_out, int lddo, void* dqkv, int lddqkv, int B, int F,
                         int HW, int heads, int ch, jg_stream_t stream);
int jg_geglu_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int Cout, jg_stream_t stream);
int jg_geglu_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int Cout,
                 jg_stream_t stream);

/* All emb_layers Linears of a UNet in ONE launch (unet_generator_attn.py:201-207, 247-258): n items
 * y_i[B][O_i] = act_in(x[B][I]) @ W_i^T + b_i on the same x.  Item i owns the contiguous [B][O_i] block of Y (and of
 * dY in the backward) at float offset B * off, off = the item's first output in the concatenation of all items.
 * Tiles: jg_linear_batched_tiles(O_i) per item, tile_start_dev = exclusive prefix sum.  Backward: dW [sum O][I] and
 * dB [sum O] (item i at rows off .. off + O_i) are overwritten; dx [B][I] (may be NULL) = the sum over items.
 * B <= 64, I <= 128. */
typedef struct {
  const float* w;  /* [O][I] */
  const float* b;  /* [O] or NULL */
  int O, off;
} jg_linear_item;
int jg_linear_batched_tiles(int O);
int jg_linear_batched_fwd(const float* x, const jg_linear_item* items_dev, const int* tile_start_dev, int n,
                          int total_tiles, float* Y, int B, int I, int act_in, jg_stream_t stream);
int jg_linear_batched_bwd(const float* x, const jg_linear_item* items_dev, const int* tile_start_dev, int n,
                          int total_tiles, const float* dY, float* dW, float* dB, float* dx, int B, int I, int act_in,
                          jg_stream_t stream);
/* ---------------------------------------------------------------------------------------------
 * Small fp32 Linear on [B, I] embeddings with optional SiLU on the input / output:
 *   ResBlock.emb_layers = SiLU -> Linear (unet_generator_attn.py:201-207),
 *   DiffusionGenerator.cond_embed = Linear -> SiLU -> Linear (diffusion_generator.py:72-76).
 * ------------------------------------------------------------------------------------------- */
int jg_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int I, int O, int act_in,
                  int act_out, jg_stream_t stream);
/* dx (=|+=) (may be NULL), dw [O][I] and db [O] overwritten (may be NULL). act_in as in the forward. */
int jg_linear_bwd(const float* x, const float* w, const float* dy, float* dx, int dx_accumulate, float* dw, float* db,
                  int B, int I, int O, int act_in, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * DiffusionGenerator.forward prologue (diffusion_generator.py:480-491): q_sample + mask blend +
 * cat([y_cond, y_noisy]) written as NHWC bf16 (channel stride ld, zero padded).  y0 / ycond / noise
 * fp32 NCHW [B,C,H,W]; mask [B,1,H,W] as fp32 or int64 (either may be NULL; clamp(mask,0,1) is
 * bit-exact); gammas fp32 [B] = sample_gammas.
 * ------------------------------------------------------------------------------------------- */
int jg_noise_pack_fwd(const float* y0, const float* ycond, const float* noise, const float* mask_f32,
                      const int64_t* mask_i64, const float* gammas, void* out, int B, int C, int H, int W, int ld,
                      jg_stream_t stream);

/* One reverse-diffusion step of the DDPM sampler: DiffusionGenerator.p_sample
 * (diffusion_generator.py:248-283) = predict_start_from_noise + clamp + q_posterior (diffusion_utils.py:122-137)
 * + noise, the mask blend of restoration_ddpm (:168-170), and the next step's cat([y_cond, y_t]) NHWC bf16 pack.
 * eps: UNet output NHWC bf16 (stride lde); y_t, y_cond, y_0, noise, y_next: fp32 NCHW [B,C,H,W]; noise NULL at t = 0;
 * coef fp32 [B][5] = (sqrt_recip_gammas, sqrt_recipm1_gammas, posterior_mean_coef1, posterior_mean_coef2,
 * exp(0.5 * posterior_log_variance_clipped)) gathered at t; x_next (may be NULL) NHWC bf16 [B,H,W,ld].
 * ddim != 0: the DDIM update of ddim_p_sample / ddim_p_mean_variance (:349-456) instead:
 * y = clamp(c1*y_t + c2*clamp(eps,-1,1), -1, 1) with coef = (sqrt(g_prev/g_t), coef_eps - sqrt(g_prev*(1-g_t)/g_t), ..). */
int jg_ddpm_step(const void* eps, int lde, const float* y_t, const float* y_cond, const float* y_0,
                 const float* mask_f32, const int64_t* mask_i64, const float* noise, const float* coef, float* y_next,
                 void* x_next, int B, int C, int H, int W, int ld, int ddim, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * PaletteModel.compute_palette_loss (palette_model.py:596-620):
 *   loss = lambda_G * mean_{b,c,h,w} (w_b * clamp(mask,0,1) * (noise - noise_hat))^2   (l1: |.|)
 * noise fp32 NCHW, noise_hat NHWC bf16 (stride ld); w_b fp32 [B] (min-SNR weight) or NULL.
 * loss: fp32 scalar on device (overwritten).  bwd writes d loss/d noise_hat * (*grad_out).
 * ------------------------------------------------------------------------------------------- */
int jg_palette_loss_fwd(const float* noise, const void* noise_hat, int ld, const float* mask_f32,
                        const int64_t* mask_i64, const float* w_b, int B, int C, int HW, float lambda_g, int l1,
                        float* loss, jg_stream_t stream);
int jg_palette_loss_bwd(const float* noise, const void* noise_hat, int ld, const float* mask_f32,
                        const int64_t* mask_i64, const float* w_b, int B, int C, int HW, float lambda_g, int l1,
                        const float* grad_out, void* d_noise_hat, int ldd, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Adam(W) + EMA over flat fp32 buffers: torch.optim.AdamW / Adam update (train.py:51-62,
 * base_model.py:1268-1274) followed by ema_step (base_model.py:1284-1297).  g is multiplied by
 * grad_scale first (1/world_size after a SUM all-reduce, 1/iter_size...).  ema may be NULL;
 * ema_init != 0 reproduces the first ema_step (deep copy of the updated net).  step is the 1-based
 * optimizer step on the host, or — when step_dev != NULL — a device counter that this call increments
 * and reads (CUDA-graph replay; ema_init is then derived as step == 1).
 * ------------------------------------------------------------------------------------------- */
int jg_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int adamw, int step, int* step_dev,
                      float grad_scale, float ema_beta, int ema_init, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GAN generator / discriminator helpers (NHWC bf16, HBM-bound).
 *   jg_pad2d_*    nn.ReflectionPad2d / ReplicationPad2d (resnet_generator.py:52-55, 207, 326-331); mode 0 reflect,
 *                 1 replicate
 *   jg_dilate2x   mode 0: zero insertion dst[2h][2w] = src[h][w] (dst is Hd x Wd = 2H x 2W) — prologue of
 *                 nn.ConvTranspose2d(k=3, s=2, p=1, output_padding=1) (resnet_generator.py:306-318) and of the
 *                 dgrad of stride-2 convolutions, both then run as stride-1 implicit GEMMs;
 *                 mode 1: the adjoint gather dst[h][w] = src[2h][2w] (src is 2Hd x 2Wd)
 *   jg_act_bwd    dx = dy * act'(x) from the OUTPUT y (tanh: 1 - y^2; (Leaky)ReLU: sign of y)
 *   jg_gan_loss_* GANLoss (loss.py:11-85) on PatchGAN logits [rows][C]: mode 0 lsgan mean((p-target)^2),
 *                 1 hinge mean(relu(1 - sign*p)), 2 linear mean(-sign*p)
 * ------------------------------------------------------------------------------------------- */
int jg_pad2d_fwd(const void* src, int lds, void* dst, int ldd, int N, int H, int W, int C, int pad, int mode,
                 jg_stream_t stream);
int jg_pad2d_bwd(const void* dpad, int ldp, void* dsrc, int lds, int N, int H, int W, int C, int pad, int mode,
                 jg_stream_t stream);
int jg_dilate2x(const void* src, int lds, void* dst, int ldd, int N, int Hd, int Wd, int C, int mode,
                jg_stream_t stream);
int jg_act_bwd(const void* y, int ldy, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int C, int act,
               jg_stream_t stream);
int jg_gan_loss_fwd(const void* pred, int ld, int64_t rows, int C, int mode, float target, float sign, float* loss,
                    jg_stream_t stream);
int jg_gan_loss_bwd(const void* pred, int ld, int64_t rows, int C, int mode, float target, float sign,
                    const float* grad_out, void* dpred, int ldd, jg_stream_t stream);

/* ---- CUT contrastive path (written in round 1, NOT yet verified on hardware) ----
 * PatchSampleF.forward (models/modules/cut_networks.py:38-73): dst[b*P + p] = src[b*HW + ids[p]] (bf16 rows, the same
 * ids for every image); the backward zeroes dsrc [B*HW rows] and scatters.  ids: int64 on the device, distinct. */
int jg_gather_rows(const void* src, int lds, const int64_t* ids, void* dst, int ldd, int B, int HW, int P, int C,
                   jg_stream_t stream);
int jg_gather_rows_bwd(const void* ddst, int ldd, const int64_t* ids, void* dsrc, int lds, int B, int HW, int P, int C,
                       jg_stream_t stream);
/* torch.nn.functional.normalize(x, eps) over D features (cut_networks.py:66): bf16 rows in, fp32 [rows][D] out,
 * norms [rows] kept for the backward (dx bf16). */
int jg_l2norm_fwd(const void* x, int ldx, float* y, float* norms, int64_t rows, int D, float eps, jg_stream_t stream);
int jg_l2norm_bwd(const float* y, const float* dy, const float* norms, void* dx, int lddx, int64_t rows, int D,
                  float eps, jg_stream_t stream);
/* PatchNCELoss (models/modules/NCE/base_NCE.py:17-77): q, k fp32 [G*P][D]; G groups of P patches (G = batch, or 1 with
 * P = batch * patches for --alg_cut_nce_includes_all_negatives_from_minibatch); loss / lse [G*P].  Backward:
 * grad_loss [G*P]; dq and/or dk [G*P][D] (k receives gradient through the negatives only, like the reference). */
int jg_patch_nce_fwd(const float* q, const float* k, int G, int P, int D, float T, float* loss, float* lse,
                     jg_stream_t stream);
int jg_patch_nce_bwd(const float* q, const float* k, const float* lse, const float* grad_loss, int G, int P, int D,
                     float T, float* dq, float* dk, jg_stream_t stream);

/* MoNCELoss (models/modules/NCE/monce.py:12-33 + sinkhorn.py; --alg_cut_nce_loss monce): PatchNCE whose negative logits
 * get T * log of the Sinkhorn optimal-transport weights (cost "hard", eps 1, `iters` scalings), differentiated through
 * the iterations w.r.t. q.  ws: jg_monce_ws_floats(G, P, iters, backward) floats, the SAME buffer for the forward and
 * the backward of one step (the forward leaves C, K and the scaling history in it).  P <= 1024 (one CTA per group).
 * Compiled, not yet run on hardware. */
size_t jg_monce_ws_floats(int G, int P, int iters, int backward);
int jg_monce_fwd(const float* q, const float* k, int G, int P, int D, float T, int num_patches_opt, int iters, float* ws,
                 float* loss, float* lse, jg_stream_t stream);
int jg_monce_bwd(const float* q, const float* k, const float* lse, const float* grad_loss, int G, int P, int D, float T,
                 int num_patches_opt, int iters, float* ws, float* dq, float* dk, jg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * b2b video backbone, JiTViD (models/modules/vit/vit_vid.py).  Tokens are rows of bf16
 * [rows][ld] tensors, rows = N * T (N = B * F frames, T tokens per frame); modulation vectors are fp32 slices
 * [N][.] (row stride ldm) of the adaLN Linear's output.  The Linears are jg_conv2d_* (1x1) on the same tensors.
 *   jg_rmsnorm_mod_*     RMSNorm (util/model_util.py:165-179, eps 1e-6, weight w) + modulate x*(1+scale)+shift
 *                        (vit_vid.py:47-48, JiTBlock :249-280, FinalLayer :283-308); shift = scale = NULL: plain RMSNorm.
 *                        rstd [rows] saved.  bwd: dx; dw [C] (overwritten); dshift / dscale [N][.] (row stride lddm).
 *   jg_qknorm_rope_*     Attention.forward :205-231: per-head RMSNorm of q and k (weights wq, wk [hd]) + rotary embedding
 *                        (cos / sin fp32 [T][hd], rotate_half on interleaved pairs, util/model_util.py:97-162) of the q and
 *                        k parts of qkv [rows][(q | k | v) each heads*hd]; out [rows][2*heads*hd]; rstd [rows][heads][2].
 *   jg_attn_small_*      softmax(q k^T / sqrt(hd)) v over the T <= 128 tokens of each frame, fp32; lse [N*heads][T].
 *   jg_swiglu_*          SwiGLUFFN :234-246: x = (x1 | x2) [rows][2H] -> silu(x1) * x2.
 *   jg_gated_residual_*  out = x + gate * y (:270-279); bwd: dy = gate * d, dgate [N][.] = sum_t d * y.
 * hd in {16, 32, 64}. */
int jg_rmsnorm_mod_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int C, int T, float eps, const float* w,
                       const float* shift, const float* scale, int ldm, float* rstd, jg_stream_t stream);
int jg_rmsnorm_mod_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int C, int T,
                       const float* w, const float* scale, int ldm, const float* rstd, float* dw, float* dshift,
                       float* dscale, int lddm, jg_stream_t stream);
int jg_qknorm_rope_fwd(const void* qkv, int ldq, void* out, int ldo, int64_t rows, int T, int heads, int hd, float eps,
                       const float* wq, const float* wk, const float* cosb, const float* sinb, float* rstd,
                       jg_stream_t stream);
int jg_qknorm_rope_bwd(const void* qkv, int ldq, const void* dout, int lddo, void* dqkv, int lddq, int64_t rows, int T,
                       int heads, int hd, const float* wq, const float* wk, const float* cosb, const float* sinb,
                       const float* rstd, float* dwq, float* dwk, jg_stream_t stream);
int jg_attn_small_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, float* lse,
                      int N, int T, int heads, int hd, jg_stream_t stream);
int jg_attn_small_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, int ldo,
                      const void* d_o, int lddo, const float* lse, void* dq, int lddq, void* dk, int lddk, void* dv,
                      int lddv, int N, int T, int heads, int hd, jg_stream_t stream);
int jg_swiglu_fwd(const void* x, int ldx, void* y, int ldy, int64_t rows, int H, jg_stream_t stream);
int jg_swiglu_bwd(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int64_t rows, int H,
                  jg_stream_t stream);
int jg_gated_residual_fwd(const void* x, int ldx, const void* y, int ldy, const float* gate, int ldm, void* out, int ldo,
                          int64_t rows, int C, int T, jg_stream_t stream);
int jg_gated_residual_bwd(const void* d, int ldd, const void* y, int ldy, const float* gate, int ldm, void* dy, int lddy,
                          float* dgate, int lddm, int64_t rows, int C, int T, jg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* JG_B200_H */
