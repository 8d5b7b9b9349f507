"""CPU oracle for the Palette diffusion-UNet training step (TEST INFRASTRUCTURE — see oracle/__init__.py).

A functional restatement, in plain fp32 PyTorch ops over a flat {state_dict key: tensor} mapping, of

  UNet / ResBlock / AttentionBlock / QKVAttentionLegacy
      /root/reference/models/modules/unet_generator_attn/unet_generator_attn.py:143-266, 269-347, 390-695
  GroupNorm wrapper (fp32 compute), InstanceNorm1d attention norm
      /root/reference/models/modules/unet_generator_attn/unet_attn_utils.py:42-48, 60-66, 94-117
  DiffusionGenerator.forward (t / gamma sampling, q_sample, mask blend, cond embed, min-SNR weight)
      /root/reference/models/modules/diffusion_generator.py:457-528
  gamma_embedding, linear beta schedule buffers
      /root/reference/models/modules/diffusion_utils.py:8-42, 45-119
  PaletteModel.compute_palette_loss (masked eps-MSE)
      /root/reference/models/palette_model.py:558-620
  BaseModel.optimize_parameters / compute_step / ema_step with torch.optim.AdamW / Adam
      /root/reference/models/base_model.py:1250-1377, /root/reference/train.py:51-62

The state_dict keys are exactly the reference's (`denoise_fn.model.input_blocks.1.0.in_layers.0.norm.weight`,
`cond_embed.0.weight`, ...), so weights move between the reference, this oracle and the B200 modules
without renaming.  Pinned against the real reference by oracle/gen_golden.py + tests/test_oracle_golden.py.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# configuration + structure (restates UNet.__init__, unet_generator_attn.py:419-658)
# --------------------------------------------------------------------------------------------
@dataclass
class UNetCfg:
    image_size: int = 256
    in_channel: int = 6
    inner_channel: int = 64
    out_channel: int = 3
    res_blocks: Tuple[int, ...] = (2, 2, 2, 2)
    attn_res: Tuple[int, ...] = (16,)
    channel_mults: Tuple[int, ...] = (1, 2, 4, 8)
    num_heads: int = 1
    num_head_channels: int = 32
    group_norm_size: int = 32
    cond_embed_dim: int = 32
    n_timestep_train: int = 2000
    n_timestep_test: int = 1000
    use_scale_shift_norm: bool = True
    efficient: bool = False
    # PaletteDenoiseFn conditioning (palette_denoise_fn.py:35-59): "" | "class" | "mask" | "class_mask"
    conditioning: str = ""
    nclasses: int = 2


@dataclass
class BlockSpec:
    kind: str  # "conv" | "res" | "attn"
    cin: int = 0
    cout: int = 0
    up: bool = False
    down: bool = False
    heads: int = 0


def unet_structure(cfg: UNetCfg):
    """Returns (input_blocks, middle_block, output_blocks): lists of lists of BlockSpec, in the
    order and with the channel bookkeeping of UNet.__init__ (unet_generator_attn.py:478-632)."""

    def heads_for(ch):
        if cfg.num_head_channels == -1:
            return cfg.num_heads
        assert ch % cfg.num_head_channels == 0
        return ch // cfg.num_head_channels

    ch = int(cfg.channel_mults[0] * cfg.inner_channel)
    input_blocks = [[BlockSpec("conv", cfg.in_channel, ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cfg.channel_mults):
        for _ in range(cfg.res_blocks[level]):
            out = int(mult * cfg.inner_channel)
            layers = [BlockSpec("res", ch, out)]
            ch = out
            if ds in cfg.attn_res:
                layers.append(BlockSpec("attn", ch, ch, heads=heads_for(ch)))
            input_blocks.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mults) - 1:
            input_blocks.append([BlockSpec("res", ch, ch, down=True)])
            chans.append(ch)
            ds *= 2
    middle = [BlockSpec("res", ch, ch), BlockSpec("attn", ch, ch, heads=heads_for(ch)), BlockSpec("res", ch, ch)]
    output_blocks = []
    for level, mult in list(enumerate(cfg.channel_mults))[::-1]:
        for i in range(cfg.res_blocks[level] + 1):
            ich = chans.pop()
            out = int(cfg.inner_channel * mult)
            layers = [BlockSpec("res", ch + ich, out)]
            ch = out
            if ds in cfg.attn_res:
                layers.append(BlockSpec("attn", ch, ch, heads=heads_for(ch)))
            if level and i == cfg.res_blocks[level]:
                layers.append(BlockSpec("res", ch, ch, up=True))
                ds //= 2
            output_blocks.append(layers)
    return input_blocks, middle, output_blocks


def param_shapes(cfg: UNetCfg, prefix="denoise_fn.model.") -> Dict[str, Tuple[int, ...]]:
    """Parameter name -> shape, in the reference's state_dict naming/order (parameters only)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    e = cfg.cond_embed_dim

    def conv(name, cin, cout, k):
        shapes[name + ".weight"] = (cout, cin, k, k)
        shapes[name + ".bias"] = (cout,)

    def res(name, b: BlockSpec):
        shapes[name + ".in_layers.0.norm.weight"] = (b.cin,)
        shapes[name + ".in_layers.0.norm.bias"] = (b.cin,)
        conv(name + ".in_layers.2", b.cin, b.cout, 3)
        n_emb = 2 * b.cout if cfg.use_scale_shift_norm else b.cout
        shapes[name + ".emb_layers.1.weight"] = (n_emb, e)
        shapes[name + ".emb_layers.1.bias"] = (n_emb,)
        shapes[name + ".out_layers.0.norm.weight"] = (b.cout,)
        shapes[name + ".out_layers.0.norm.bias"] = (b.cout,)
        conv(name + ".out_layers.3", b.cout, b.cout, 3)
        if b.cin != b.cout:
            conv(name + ".skip_connection", b.cin, b.cout, 1)

    def attn(name, b: BlockSpec):
        shapes[name + ".qkv.weight"] = (3 * b.cin, b.cin, 1)
        shapes[name + ".qkv.bias"] = (3 * b.cin,)
        shapes[name + ".proj_out.weight"] = (b.cin, b.cin, 1)
        shapes[name + ".proj_out.bias"] = (b.cin,)

    def block(name, layers):
        for j, b in enumerate(layers):
            if b.kind == "conv":
                conv("%s.%d" % (name, j), b.cin, b.cout, 3)
            elif b.kind == "res":
                res("%s.%d" % (name, j), b)
            else:
                attn("%s.%d" % (name, j), b)

    inp, mid, outb = unet_structure(cfg)
    for i, layers in enumerate(inp):
        block(prefix + "input_blocks.%d" % i, layers)
    block(prefix + "middle_block", mid)
    for i, layers in enumerate(outb):
        block(prefix + "output_blocks.%d" % i, layers)
    ch0 = int(cfg.channel_mults[0] * cfg.inner_channel)
    shapes[prefix + "out.0.norm.weight"] = (ch0,)
    shapes[prefix + "out.0.norm.bias"] = (ch0,)
    conv(prefix + "out.2", ch0, cfg.out_channel, 3)
    return shapes


def generator_param_shapes(cfg: UNetCfg) -> Dict[str, Tuple[int, ...]]:
    """DiffusionGenerator (palette, cond_embed "") parameters: UNet + cond_embed MLP
    (diffusion_generator.py:63-76), in state_dict order."""
    shapes = param_shapes(cfg, "denoise_fn.model.")
    e = cfg.cond_embed_dim
    # label embedders (palette_denoise_fn.py:46-59), then the gamma MLP whose width halves when a class / ref
    # embedding shares the embedding vector (diffusion_generator.py:63-76)
    if "class" in cfg.conditioning:
        shapes["denoise_fn.netl_embedder_class.embedding_table.weight"] = (cfg.nclasses, e // 2)
    if "mask" in cfg.conditioning:
        shapes["denoise_fn.netl_embedder_mask.embedding_table.weight"] = (cfg.nclasses, e)
    eg = e // 2 if "class" in cfg.conditioning else e
    shapes["cond_embed.0.weight"] = (eg, eg)
    shapes["cond_embed.0.bias"] = (eg,)
    shapes["cond_embed.2.weight"] = (eg, eg)
    shapes["cond_embed.2.bias"] = (eg,)
    return shapes


def init_params(cfg: UNetCfg, seed: int, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded, NON-zero parameters for every tensor (the reference zero-initialises the second
    conv of each ResBlock, attention proj_out and the final conv, which makes a freshly built net
    a vacuous parity test — SURVEY.md mismatch 8).  Fan-in scaled normal weights; norm weights
    ~ 1 + 0.1 N(0,1); biases 0.05 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in generator_param_shapes(cfg).items():
        if name.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            t = scale * torch.randn(shape, generator=g) / math.sqrt(fan_in)
        out[name] = t
    return out


# --------------------------------------------------------------------------------------------
# noise schedule buffers (diffusion_utils.py:45-119, UNet.beta_schedule unet_generator_attn.py:644-657)
# --------------------------------------------------------------------------------------------
def schedule_buffers(cfg: UNetCfg, phase="train") -> Dict[str, torch.Tensor]:
    if phase == "train":
        n, lo, hi = cfg.n_timestep_train, 1e-6, 0.01
    else:
        n, lo, hi = cfg.n_timestep_test, 1e-4, 0.09
    betas = np.linspace(lo, hi, n, dtype=np.float64)
    alphas = 1.0 - betas
    gammas = np.cumprod(alphas, axis=0)
    gammas_prev = np.append(1.0, gammas[:-1])
    posterior_variance = betas * (1.0 - gammas_prev) / (1.0 - gammas)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {  # diffusion_utils.set_new_noise_schedule, :79-119
        "gammas_" + phase: f32(gammas),
        "sqrt_recip_gammas_" + phase: f32(np.sqrt(1.0 / gammas)),
        "sqrt_recipm1_gammas_" + phase: f32(np.sqrt(1.0 / gammas - 1)),
        "posterior_log_variance_clipped_" + phase: f32(np.log(np.maximum(posterior_variance, 1e-20))),
        "posterior_mean_coef1_" + phase: f32(betas * np.sqrt(gammas_prev) / (1.0 - gammas)),
        "posterior_mean_coef2_" + phase: f32((1.0 - gammas_prev) * np.sqrt(alphas) / (1.0 - gammas)),
    }


def gamma_embedding(gammas: torch.Tensor, dim: int, max_period=10000) -> torch.Tensor:
    """diffusion_utils.py:8-42 for gammas of shape [B, 1]."""
    assert gammas.shape[1] == 1
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = gammas[:, 0:1].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# --------------------------------------------------------------------------------------------
# forward restatements
# --------------------------------------------------------------------------------------------
# Precision emulation (default OFF = the reference's fp32 arithmetic).  With EMULATE_BF16[0] = True the
# restatement rounds to bf16 exactly where the B200 path stores bf16 (every feature map written to HBM,
# conv weights), keeping fp32 statistics / accumulation — used by the GPU tests to separate "precision
# of bf16 storage" from "kernel bug": CUDA vs this emulation must agree far tighter than CUDA vs fp32.
EMULATE_BF16 = [False]


class _RoundBF16(torch.autograd.Function):
    """bf16 storage of a feature map: the value is rounded on the way forward and its gradient on the
    way back (the B200 path keeps both in bf16 NHWC buffers)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def _r(x):
    return _RoundBF16.apply(x) if EMULATE_BF16[0] else x


def _rw(w):  # conv weights are consumed as bf16 copies; their gradients stay fp32
    return w + (w.to(torch.bfloat16).float() - w).detach() if EMULATE_BF16[0] else w


def _conv2d(x, w, b, padding=0):
    return F.conv2d(x, _rw(w), b, padding=padding)


def group_norm(x, w, b, groups):
    # unet_attn_utils.py:42-48 — nn.GroupNorm(groups, C) computed in fp32, eps 1e-5
    return F.group_norm(x.float(), groups, w, b, eps=1e-5).type(x.dtype)


def res_block(sd, name, x, emb, b: BlockSpec, cfg: UNetCfg):
    """ResBlock._forward, unet_generator_attn.py:233-266."""
    h = _r(F.silu(group_norm(x, sd[name + ".in_layers.0.norm.weight"], sd[name + ".in_layers.0.norm.bias"],
                             cfg.group_norm_size)))
    if b.up and cfg.efficient:
        # (:239-242) efficient up-blocks convolve at the low resolution and upsample afterwards
        h = _r(_conv2d(h, sd[name + ".in_layers.2.weight"], sd[name + ".in_layers.2.bias"], padding=1))
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    else:
        if b.up:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        elif b.down:
            h = _r(F.avg_pool2d(h, 2, 2))
            x = _r(F.avg_pool2d(x, 2, 2))
        h = _r(_conv2d(h, sd[name + ".in_layers.2.weight"], sd[name + ".in_layers.2.bias"], padding=1))
    emb_out = F.linear(F.silu(emb), sd[name + ".emb_layers.1.weight"], sd[name + ".emb_layers.1.bias"])
    emb_out = emb_out[:, :, None, None]
    gn_w, gn_b = sd[name + ".out_layers.0.norm.weight"], sd[name + ".out_layers.0.norm.bias"]
    if cfg.use_scale_shift_norm:
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = group_norm(h, gn_w, gn_b, cfg.group_norm_size) * (1 + scale) + shift
        h = _r(F.silu(h))
    else:
        h = _r(h + emb_out)  # (identity in fp32; the B200 path stores the sum as bf16)
        h = _r(F.silu(group_norm(h, gn_w, gn_b, cfg.group_norm_size)))
    h = _conv2d(h, sd[name + ".out_layers.3.weight"], sd[name + ".out_layers.3.bias"], padding=1)
    if b.cin != b.cout:
        x = _r(_conv2d(x, sd[name + ".skip_connection.weight"], sd[name + ".skip_connection.bias"]))
    skipw = 1.0 / math.sqrt(2) if cfg.efficient else 1.0
    return _r(skipw * x + h)


def qkv_attention_legacy(qkv, n_heads):
    """QKVAttentionLegacy.forward, unet_generator_attn.py:331-347."""
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
    a = torch.einsum("bts,bcs->bct", weight, v)
    return a.reshape(bs, -1, length)


def attention_block(sd, name, x, b: BlockSpec):
    """AttentionBlock._forward, unet_generator_attn.py:310-319 (norm = InstanceNorm1d, no affine)."""
    bsz, c, hh, ww = x.shape
    xf = x.reshape(bsz, c, -1)
    xn = _r(F.instance_norm(xf.float(), eps=1e-5).type(xf.dtype))
    qkv = _r(F.conv1d(xn, _rw(sd[name + ".qkv.weight"]), sd[name + ".qkv.bias"]))
    h = _r(qkv_attention_legacy(qkv, b.heads))
    h = F.conv1d(h, _rw(sd[name + ".proj_out.weight"]), sd[name + ".proj_out.bias"])
    return _r(xf + h).reshape(bsz, c, hh, ww)


def _run_block(sd, name, layers, h, emb, cfg):
    for j, b in enumerate(layers):
        n = "%s.%d" % (name, j)
        if b.kind == "conv":
            h = _r(_conv2d(h, sd[n + ".weight"], sd[n + ".bias"], padding=1))
        elif b.kind == "res":
            h = res_block(sd, n, h, emb, b, cfg)
        else:
            h = attention_block(sd, n, h, b)
    return h


def unet_forward(sd, x, emb, cfg: UNetCfg, prefix="denoise_fn.model.", return_feats=False):
    """UNet.forward, unet_generator_attn.py:660-695."""
    inp, mid, outb = unet_structure(cfg)
    hs = []
    h = _r(x.float())
    for i, layers in enumerate(inp):
        h = _run_block(sd, prefix + "input_blocks.%d" % i, layers, h, emb, cfg)
        hs.append(h)
    h = _run_block(sd, prefix + "middle_block", mid, h, emb, cfg)
    feats = list(hs)
    for i, layers in enumerate(outb):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, prefix + "output_blocks.%d" % i, layers, h, emb, cfg)
    h = _r(F.silu(group_norm(h, sd[prefix + "out.0.norm.weight"], sd[prefix + "out.0.norm.bias"],
                             cfg.group_norm_size)))
    out = _r(_conv2d(h, sd[prefix + "out.2.weight"], sd[prefix + "out.2.bias"], padding=1))
    if return_feats:
        return out, feats
    return out


def sample_t_gamma(cfg: UNetCfg, batch: int, generator: Optional[torch.Generator] = None):
    """The three RNG draws of DiffusionGenerator.forward in reference order
    (diffusion_generator.py:467-480): t ~ randint(1, T), u ~ rand(b,1); noise is drawn by the caller."""
    t = torch.randint(1, cfg.n_timestep_train, (batch,), generator=generator).long()
    u = torch.rand((batch, 1), generator=generator)
    return t, u


def label_embed(table, idx):
    """LabelEmbedder.forward (palette_denoise_fn.py:14-31): nn.Embedding(max_norm=1.0, scale_grad_by_freq=True) — the
    looked-up rows are renormalised IN PLACE to norm <= 1 first, gradients are divided by the index frequency."""
    return F.embedding(idx, table, max_norm=1.0, scale_grad_by_freq=True)


def diffusion_forward(sd, y_0, y_cond, mask, noise, t, u, cfg: UNetCfg, unet=None, cls=None):
    """DiffusionGenerator.forward for 4-D inputs with explicit randomness (t, u, noise).
    Returns (noise, noise_hat, min_snr_loss_weight) like diffusion_generator.py:521.
    unet(sd, input, emb, cfg): the denoiser (default: the plain UNet; oracle.ref_oracle.denoiser(ref) for the
    reference-attention UNet, whose extra `ref` argument PaletteDenoiseFn forwards, palette_denoise_fn.py:111-112)."""
    unet = unet or unet_forward
    sched = schedule_buffers(cfg, "train")
    gammas = sched["gammas_train"]
    b = y_0.shape[0]
    gamma_t1 = gammas.gather(-1, t - 1).reshape(b, 1)
    gamma_t2 = gammas.gather(-1, t).reshape(b, 1)
    sample_gammas = (gamma_t2 - gamma_t1) * u + gamma_t1
    sample_gammas = sample_gammas.view(b, -1)
    g4 = sample_gammas.view(-1, 1, 1, 1)
    y_noisy = g4.sqrt() * y_0 + (1 - g4).sqrt() * noise
    eg = cfg.cond_embed_dim // 2 if "class" in cfg.conditioning else cfg.cond_embed_dim
    emb = gamma_embedding(sample_gammas, eg)
    emb = F.linear(emb, sd["cond_embed.0.weight"], sd["cond_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["cond_embed.2.weight"], sd["cond_embed.2.bias"])
    if mask is not None:
        temp_mask = torch.clamp(mask, min=0.0, max=1.0)
        y_noisy = y_noisy * temp_mask + (1.0 - temp_mask) * y_0
    inp = torch.cat([y_cond, y_noisy], dim=1)
    # PaletteDenoiseFn.forward / compute_cond (palette_denoise_fn.py:95-136)
    if "class" in cfg.conditioning:
        emb = torch.cat((emb, label_embed(sd["denoise_fn.netl_embedder_class.embedding_table.weight"], cls)), dim=1)
    if "mask" in cfg.conditioning:
        hw = mask.shape[-1]
        me = label_embed(sd["denoise_fn.netl_embedder_mask.embedding_table.weight"],
                         mask.to(torch.int32).squeeze(1).flatten(1))          # [b, h*w, e]
        inp = torch.cat([inp, me.reshape(b, -1, hw, me.shape[-1]).permute(0, 3, 1, 2)], dim=1)
    noise_hat = unet(sd, inp, emb, cfg)
    ksnr = 5.0
    snr1 = sched["sqrt_recip_gammas_train"].gather(-1, t)
    snr2 = sched["sqrt_recipm1_gammas_train"].gather(-1, t)
    snr = torch.pow(snr1 / snr2, 2)
    w = torch.stack([snr, ksnr * torch.ones_like(t)], dim=1).min(dim=1)[0] / snr
    return noise, noise_hat, w.view(-1, 1, 1, 1)


def _sampler_denoise(sd, y_cond, y_t, noise_level, mask, cls, cfg, unet):
    """p_mean_variance's denoiser call (diffusion_generator.py:213-246) through PaletteDenoiseFn.forward
    (palette_denoise_fn.py:95-108): noise-level embedding (+ class embedding), input (+ per-pixel mask embedding)."""
    b = y_cond.shape[0]
    eg = cfg.cond_embed_dim // 2 if "class" in cfg.conditioning else cfg.cond_embed_dim
    emb = gamma_embedding(noise_level, eg)
    emb = F.linear(emb, sd["cond_embed.0.weight"], sd["cond_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["cond_embed.2.weight"], sd["cond_embed.2.bias"])
    inp = torch.cat([y_cond, y_t], dim=1)
    if "class" in cfg.conditioning:
        emb = torch.cat((emb, label_embed(sd["denoise_fn.netl_embedder_class.embedding_table.weight"], cls)), dim=1)
    if "mask" in cfg.conditioning:
        hw = mask.shape[-1]
        me = label_embed(sd["denoise_fn.netl_embedder_mask.embedding_table.weight"],
                         mask.to(torch.int32).squeeze(1).flatten(1))
        inp = torch.cat([inp, me.reshape(b, -1, hw, me.shape[-1]).permute(0, 3, 1, 2)], dim=1)
    return unet(sd, inp, emb, cfg)


def restoration_ddpm(sd, y_cond, y_t, y_0, mask, noises, cfg: UNetCfg, sample_num=2, unet=None, cls=None):
    """DiffusionGenerator.restoration_ddpm (diffusion_generator.py:122-177) with p_sample / p_mean_variance
    (:192-283), predict_start_from_noise and q_posterior (diffusion_utils.py:122-137), class / mask conditioning
    included, no guidance.  `noises[i]` is the randn_like draw of step i (i > 0), in the reference's order.
    Returns (y_t, ret_arr)."""
    unet = unet or unet_forward
    sched = schedule_buffers(cfg, "test")
    T = int(sched["gammas_test"].shape[0])
    sample_inter = T // sample_num
    b = y_cond.shape[0]
    ret_arr = y_t
    for i in reversed(range(T)):
        t = torch.full((b,), i, dtype=torch.long)
        noise_level = sched["gammas_test"].gather(-1, t).reshape(b, 1)
        eps = _sampler_denoise(sd, y_cond, y_t, noise_level, mask, cls, cfg, unet)

        def ex(name):
            return sched[name + "_test"].gather(-1, t).reshape(b, 1, 1, 1)

        y0_hat = (ex("sqrt_recip_gammas") * y_t - ex("sqrt_recipm1_gammas") * eps).clamp(-1.0, 1.0)
        mean = ex("posterior_mean_coef1") * y0_hat + ex("posterior_mean_coef2") * y_t
        noise = noises[i] if i > 0 else torch.zeros_like(y_t)
        y_t = mean + noise * (0.5 * ex("posterior_log_variance_clipped")).exp()
        if mask is not None:
            m = torch.clamp(mask, min=0.0, max=1.0)
            y_t = y_0 * (1.0 - m) + m * y_t
        if i % sample_inter == 0:
            ret_arr = torch.cat([ret_arr, y_t], dim=0)
    return y_t, ret_arr


def restoration_ddim(sd, y_cond, y_t, y_0, mask, cfg: UNetCfg, sample_num=8, num_steps=10, eta=0.5, unet=None,
                     cls=None):
    """DiffusionGenerator.restoration_ddim / ddim_p_sample / ddim_p_mean_variance (diffusion_generator.py:286-456),
    conditioning "" and no guidance.  Deterministic given y_t (the reference's per-step noise draw is unused)."""
    unet = unet or unet_forward
    sched = schedule_buffers(cfg, "test")
    T = int(sched["gammas_test"].shape[0])
    sample_inter = T // sample_num
    b = y_cond.shape[0]
    ret_arr = y_t
    tseq = list(np.linspace(0, T - 1, num_steps).astype(int))
    gammas_prev = torch.cat([torch.ones(1), sched["gammas_test"][:-1]])
    for i in range(num_steps):
        t = torch.full((b,), int(tseq[-1 - i]), dtype=torch.long)
        prevt = torch.full((b,), int(tseq[-2 - i]) if i != num_steps - 1 else -1, dtype=torch.long)
        noise_level = sched["gammas_test"].gather(-1, t).reshape(b, 1)
        e = _sampler_denoise(sd, y_cond, y_t, noise_level, mask, cls, cfg, unet).clamp(-1.0, 1.0)
        g_t = sched["gammas_test"].gather(-1, t).reshape(b, 1, 1, 1)
        g_p = gammas_prev.gather(-1, prevt + 1).reshape(b, 1, 1, 1)
        sigma = eta * torch.sqrt((1 - g_p) / (1 - g_t) * (1 - g_t / g_p))
        coef_eps = 1 - g_p - sigma ** 2
        coef_eps[coef_eps < 0] = 0
        coef_eps = torch.sqrt(coef_eps)
        y_t = (torch.sqrt(g_p) * (y_t - torch.sqrt(1.0 - g_t) * e) / torch.sqrt(g_t) + coef_eps * e).clamp(-1.0, 1.0)
        if mask is not None:
            m = torch.clamp(mask, min=0.0, max=1.0)
            y_t = y_0 * (1.0 - m) + m * y_t
        if i % sample_inter == 0:
            ret_arr = torch.cat([ret_arr, y_t], dim=0)
    return y_t, ret_arr


def palette_loss(noise, noise_hat, mask, min_snr_w=None, lambda_G=1.0, use_minsnr=False, kind="MSE"):
    """PaletteModel.compute_palette_loss, palette_model.py:596-620."""
    w = min_snr_w if use_minsnr else 1.0
    if mask is not None:
        mb = torch.clamp(mask, min=0, max=1)
        a, b = w * mb * noise, w * mb * noise_hat
    else:
        a, b = w * noise, w * noise_hat
    loss = F.mse_loss(b, a) if kind == "MSE" else F.l1_loss(b, a)
    return lambda_G * loss


# --------------------------------------------------------------------------------------------
# training step
# --------------------------------------------------------------------------------------------
@dataclass
class OptimCfg:
    lr: float = 2e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 0.0
    kind: str = "adamw"  # "adamw" | "adam"  (train.py:51-62)
    ema_beta: float = 0.999
    iter_size: int = 1


@dataclass
class TrainState:
    params: Dict[str, torch.Tensor]
    exp_avg: Dict[str, torch.Tensor] = field(default_factory=dict)
    exp_avg_sq: Dict[str, torch.Tensor] = field(default_factory=dict)
    ema: Optional[Dict[str, torch.Tensor]] = None
    step: int = 0


def adam_update(state: TrainState, grads: Dict[str, torch.Tensor], oc: OptimCfg):
    """torch.optim.AdamW / Adam single-tensor update rule (what scaler.step(optimizer) applies,
    base_model.py:1268-1274), then ema_step (base_model.py:1284-1297)."""
    state.step += 1
    bc1 = 1 - oc.beta1 ** state.step
    bc2 = 1 - oc.beta2 ** state.step
    for k, p in state.params.items():
        g = grads[k]
        if oc.kind == "adam" and oc.weight_decay != 0:
            g = g + oc.weight_decay * p
        if oc.kind == "adamw":
            p.mul_(1 - oc.lr * oc.weight_decay)
        m = state.exp_avg.setdefault(k, torch.zeros_like(p))
        v = state.exp_avg_sq.setdefault(k, torch.zeros_like(p))
        m.mul_(oc.beta1).add_(g, alpha=1 - oc.beta1)
        v.mul_(oc.beta2).addcmul_(g, g, value=1 - oc.beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(oc.eps)
        p.addcdiv_(m, denom, value=-oc.lr / bc1)
    if state.ema is None:
        # first ema_step deep-copies the (already updated) network, then lerps it with itself
        state.ema = {k: p.clone() for k, p in state.params.items()}
    for k, p in state.params.items():
        pe = state.ema[k]
        pe.copy_(p + oc.ema_beta * (pe - p))  # p.lerp(p_ema, beta)


def train_step(state: TrainState, cfg: UNetCfg, oc: OptimCfg, y_0, y_cond, mask, noise, t, u,
               lambda_G=1.0, use_minsnr=False, forward=None):
    """One optimize_parameters() of the palette group: forward + loss + backward + Adam(W) + EMA.
    Returns (loss, noise_hat, grads).  forward(leaves) -> (noise, noise_hat, w): the generator forward for the other
    denoisers (reference-attention UNet: diffusion_forward(..., unet=ref_oracle.denoiser(ref)); video UNet:
    vid_oracle.diffusion_forward_vid); default = the plain UNet."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in state.params.items()}
    if forward is None:
        _, noise_hat, w = diffusion_forward(leaves, y_0, y_cond, mask, noise, t, u, cfg)
    else:
        _, noise_hat, w = forward(leaves)
    loss = palette_loss(noise, noise_hat, mask, w, lambda_G, use_minsnr)
    (loss / oc.iter_size).backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    with torch.no_grad():
        adam_update(state, grads, oc)
    return loss.detach(), noise_hat.detach(), grads


def synthetic_batch(batch, size, seed, device="cpu"):
    """BASELINE config-2 synthetic inputs: gt ~ N(0,0.5²) clamped, one random box mask
    per image covering 10–40 % of the area (int64 0/1), cond = gt·(1−m) + N(0,1)·m
    (= fill_mask_with_random, /root/reference/data/online_creation.py:1366-1376)."""
    g = torch.Generator().manual_seed(seed)
    gt = (0.5 * torch.randn(batch, 3, size, size, generator=g)).clamp(-1, 1)
    mask = torch.zeros(batch, 1, size, size, dtype=torch.int64)
    for i in range(batch):
        frac = 0.1 + 0.3 * float(torch.rand((), generator=g))
        side = max(1, int(round(size * math.sqrt(frac))))
        y0 = int(torch.randint(0, size - side + 1, (), generator=g))
        x0 = int(torch.randint(0, size - side + 1, (), generator=g))
        mask[i, 0, y0:y0 + side, x0:x0 + side] = 1
    rnd = torch.randn(batch, 3, size, size, generator=g)
    cond = gt * (1 - mask) + rnd * mask
    return {"gt": gt.to(device), "cond": cond.to(device), "mask": mask.to(device)}
