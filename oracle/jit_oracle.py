"""CPU oracle for the b2b video backbone (TEST INFRASTRUCTURE — see oracle/__init__.py): SURVEY.md section 8(f)
rank 2, `model_type b2b` + `G_netG vit_vid` (example_b2b_vid_mario.json, BASELINE.json config 5).

Functional fp32 restatement over {state_dict key: tensor} of the path the example configuration takes (no mask-size /
frame-step / global-context / object-reference conditioning, no register tokens, one MotionModule after the last
block — `motion_every == 0`):

  JiTViD.forward            /root/reference/models/modules/vit/vit_vid.py:1234-1358
    BottleneckPatchEmbed    :51-87     (conv p x p stride p, no bias -> conv 1x1) + fixed sin-cos pos_embed (a frozen
                                        Parameter of the state_dict)
    TimestepEmbedder / LabelEmbedder   :90-147
    JiTBlock                :249-280   adaLN (SiLU -> Linear -> 6 chunks), RMSNorm (util/model_util.py:165-179),
    Attention               :182-231   qkv Linear, per-head RMSNorm of q and k, 2-D rotary embedding
                                        (VisionRotaryEmbeddingFast, util/model_util.py:97-162: identity rotation for the
                                        in-context prefix tokens), softmax attention in fp32
    SwiGLUFFN               :234-246   hidden = int(4 * D * 2 / 3)
    in-context tokens       :1296-1316 in_context_len copies of the label embedding + a learned position table,
                                        prepended at block `in_context_start`, dropped after the last block
    MotionModule            models/modules/vit/vit_vid_per_layer_motion.py:281-466 on the patch grid
                                        (same arithmetic as the video UNet's: oracle.vid_oracle.motion_module)
    FinalLayer + unpatchify :283-308, 1062-1082
  B2BGenerator.b2b_forward / forward   /root/reference/models/modules/b2b_generator.py:238-348
      z = t x + (1 - t) e, known pixels kept under the mask, v = (x - z) / max(1 - t, t_eps); the model predicts x.
  B2BModel._masked_region_loss         /root/reference/models/b2b_model.py:1201-1217 (pseudo-Huber over the mask)

Randomness (t, e) is an input.  Pinned against the real reference by oracle/gen_golden_jit.py + tests/test_jit_oracle.py.
"""
import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict

import torch
import torch.nn.functional as F

from . import vid_oracle as V


@dataclass
class JitCfg:
    input_size: int = 128
    patch_size: int = 16
    in_channels: int = 6
    out_channels: int = 3
    hidden_size: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: float = 4.0
    num_classes: int = 1
    in_context_len: int = 32
    in_context_start: int = 4
    max_frames: int = 8
    motion_num_heads: int = 8
    motion_num_layers: int = 2
    t_eps: float = 0.05
    noise_scale: float = 1.0


# ---- tables ----------------------------------------------------------------------------------------------------------
def rope_tables(cfg: JitCfg, num_cls_token: int):
    """VisionRotaryEmbeddingFast(dim = head_dim / 2, pt_seq_len = grid): cos / sin [num_cls + grid^2, head_dim]."""
    dim = cfg.hidden_size // cfg.num_heads // 2
    n = cfg.input_size // cfg.patch_size
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(n) / n * n
    fr = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)                 # [n, dim]
    fr = torch.cat([fr[:, None, :].expand(n, n, dim), fr[None, :, :].expand(n, n, dim)], dim=-1).reshape(n * n, -1)
    cos, sin = fr.cos(), fr.sin()
    if num_cls_token > 0:
        cos = torch.cat([torch.ones(num_cls_token, cos.shape[1]), cos], dim=0)
        sin = torch.cat([torch.zeros(num_cls_token, sin.shape[1]), sin], dim=0)
    return cos, sin


def rotate_half(x):
    x1, x2 = x.reshape(*x.shape[:-1], -1, 2).unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(x.shape)


def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ---- blocks ----------------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps=1e-6):
    xf = x.float()
    return (w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(x.dtype)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def attention(sd, name, x, cos, sin, heads):
    b, n, c = x.shape
    qkv = _lin(sd, name + ".qkv", x).reshape(b, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rms_norm(q, sd[name + ".q_norm.weight"])
    k = rms_norm(k, sd[name + ".k_norm.weight"])
    q = q * cos + rotate_half(q) * sin
    k = k * cos + rotate_half(k) * sin
    w = torch.softmax(q.float() @ k.float().transpose(-2, -1) / math.sqrt(q.shape[-1]), dim=-1)
    o = (w @ v).transpose(1, 2).reshape(b, n, c)
    return _lin(sd, name + ".proj", o)


def swiglu(sd, name, x):
    x1, x2 = _lin(sd, name + ".w12", x).chunk(2, dim=-1)
    return _lin(sd, name + ".w3", F.silu(x1) * x2)


def jit_block(sd, name, x, c, cos, sin, heads):
    mod = _lin(sd, name + ".adaLN_modulation.1", F.silu(c))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=-1)
    x = x + gate_msa.unsqueeze(1) * attention(sd, name + ".attn", modulate(rms_norm(x, sd[name + ".norm1.weight"]),
                                                                           shift_msa, scale_msa), cos, sin, heads)
    x = x + gate_mlp.unsqueeze(1) * swiglu(sd, name + ".mlp", modulate(rms_norm(x, sd[name + ".norm2.weight"]),
                                                                       shift_mlp, scale_mlp))
    return x


def unpatchify(x, p, c):
    n, t, _ = x.shape
    h = w = int(t ** 0.5)
    x = x.reshape(n, h, w, p, p, c)
    return torch.einsum("nhwpqc->nchpwq", x).reshape(n, c, h * p, w * p)


def add_buffers(sd: Dict[str, torch.Tensor], cfg: JitCfg) -> Dict[str, torch.Tensor]:
    """The MotionModule's PositionalEncoding buffers (persistent in the reference's state_dict; derived here)."""
    sd = dict(sd)
    for k in list(sd.keys()):
        if k.endswith(".to_q.weight"):
            sd[k[: -len(".to_q.weight")] + ".pos_encoder.pe"] = V.positional_encoding(sd[k].shape[1], cfg.max_frames)
    return sd


def jit_vid_forward(sd, x5, t, y, cfg: JitCfg, prefix=""):
    """JiTViD.forward: x5 [B, F, C, H, W], t [B] or [B*F] in [0, 1], y [B] class labels -> [B, F, out_channels, H, W]."""
    b, f, c, hh, ww = x5.shape
    p = cfg.patch_size
    hp, wp = hh // p, ww // p
    x = x5.reshape(b * f, c, hh, ww)
    x = F.conv2d(x, sd[prefix + "x_embedder.proj1.weight"], None, stride=p)
    x = F.conv2d(x, sd[prefix + "x_embedder.proj2.weight"], sd[prefix + "x_embedder.proj2.bias"])
    x = x.flatten(2).transpose(1, 2) + sd[prefix + "pos_embed"]
    t = t.reshape(-1)
    t2 = t.repeat_interleave(f) if t.shape[0] == b else t
    y = y.reshape(-1)
    y2 = y.repeat_interleave(f) if y.shape[0] == b else y
    t_emb = _lin(sd, prefix + "t_embedder.mlp.2", F.silu(_lin(sd, prefix + "t_embedder.mlp.0", timestep_embedding(t2))))
    y_emb = sd[prefix + "y_embedder.embedding_table.weight"][y2]
    cvec = t_emb + y_emb
    cos0, sin0 = rope_tables(cfg, 0)
    cos1, sin1 = rope_tables(cfg, cfg.in_context_len)
    for i in range(cfg.depth):
        if i == cfg.in_context_start and cfg.in_context_len > 0:
            ctx = y_emb.unsqueeze(1).repeat(1, cfg.in_context_len, 1) + sd[prefix + "in_context_posemb"]
            x = torch.cat([ctx, x], dim=1)
        cos, sin = (cos0, sin0) if i < cfg.in_context_start else (cos1, sin1)
        x = jit_block(sd, prefix + "blocks.%d" % i, x, cvec, cos, sin, cfg.num_heads)
    if cfg.depth > cfg.in_context_start and cfg.in_context_len > 0:
        x = x[:, cfg.in_context_len:]
    # MotionModule on the patch grid: (b f) (h w) d -> (b f) d h w, frames attended per patch
    d = x.shape[-1]
    grid = x.reshape(b * f, hp, wp, d).permute(0, 3, 1, 2)
    mcfg = SimpleNamespace(num_transformer_blocks=cfg.motion_num_layers, num_attention_heads=cfg.motion_num_heads)
    grid = V.motion_module(sd, prefix + "motion_module", grid, f, mcfg)
    x = grid.permute(0, 2, 3, 1).reshape(b * f, hp * wp, d)
    # FinalLayer
    shift, scale = _lin(sd, prefix + "final_layer.adaLN_modulation.1", F.silu(cvec)).chunk(2, dim=1)
    x = modulate(rms_norm(x, sd[prefix + "final_layer.norm_final.weight"]), shift, scale)
    x = _lin(sd, prefix + "final_layer.linear", x)
    out = unpatchify(x, p, cfg.out_channels)
    return out.reshape(b, f, cfg.out_channels, hh, ww)


# ---- B2BGenerator ----------------------------------------------------------------------------------------------------
def b2b_forward(sd, x, mask, x_cond, label, t_base, e, cfg: JitCfg, prefix="b2b_model."):
    """B2BGenerator.forward for clips [B, F, C, H, W] with explicit randomness: t_base [B] in (0, 1) (one draw per clip,
    b2b_generator.py:257-259), e = randn_like(x) (already multiplied by nothing: noise_scale is applied here).
    Returns (v_pred, v, x_pred)."""
    b, f = x.shape[:2]
    t = t_base[:, None].repeat(1, f).view(b, f, 1, 1, 1)
    t_flat = t.reshape(b * f)
    if mask is not None:
        mask = torch.clamp(mask, min=0.0, max=1.0)
    z_t = t * x + (1.0 - t) * (e * cfg.noise_scale)
    z = z_t * mask + (1.0 - mask) * x if mask is not None else z_t
    z_model = z if x_cond is None else torch.cat([x_cond, z], dim=2)
    v = (x - z) / (1.0 - t).clamp_min(cfg.t_eps)
    x_pred = jit_vid_forward(sd, z_model, t_flat, label, cfg, prefix=prefix)
    if x_pred.shape[2] > x.shape[2]:
        x_pred = x_pred[:, :, -x.shape[2]:]
    if mask is not None:
        x_pred = x_pred * mask + (1 - mask) * x
    v_pred = (x_pred - z) / (1 - t).clamp_min(cfg.t_eps)
    return v_pred, v, x_pred


def masked_region_loss(pred, target, mask, kind="pseudo_huber", eps=1e-8):
    """B2BModel._masked_region_loss: the mean over the batch of (sum of the masked per-element loss / mask sum).  The
    mask is passed as the model holds it — ONE channel, broadcast over the image channels in the numerator but not in
    the denominator — so the value is `channels` times a per-element mean (pinned by b2b_plumbing.pt)."""
    if kind == "MSE":
        le = (pred - target) ** 2
    elif kind == "L1":
        le = (pred - target).abs()
    elif kind == "pseudo_huber":
        c = 0.00054 * math.sqrt(math.prod(pred.shape[1:]))
        le = torch.sqrt((pred - target) ** 2 + c ** 2) - c
    else:
        raise NotImplementedError(kind)
    dims = tuple(range(1, le.ndim))
    return ((le * mask).sum(dim=dims) / mask.sum(dim=dims).clamp_min(eps)).mean()


def b2b_loss(sd, x, mask, x_cond, label, t_base, e, cfg: JitCfg, kind="pseudo_huber", lambda_G=1.0,
             masked_region_only=True, prefix="b2b_model."):
    """B2BModel.compute_b2b_loss (b2b_model.py:1081-1168) without the perceptual terms (LPIPS / DISTS are third-party
    frozen networks, out of scope) and without min-SNR weighting."""
    v_pred, v, _ = b2b_forward(sd, x, mask, x_cond, label, t_base, e, cfg, prefix=prefix)
    mb = torch.clamp(mask, min=0, max=1)
    if masked_region_only:
        return lambda_G * masked_region_loss(v_pred, v, mb, kind)
    if kind != "pseudo_huber":
        raise NotImplementedError(kind)
    c = 0.00054 * math.sqrt(math.prod(v_pred.shape[1:]))
    return lambda_G * torch.mean(torch.sqrt((mb * v_pred - mb * v) ** 2 + c ** 2) - c)


def restoration(sd, y, y_cond, mask, labels, init_noise, cfg: JitCfg, steps: int, clip_denoised=False,
                disable_inference_clipping=True, prefix="b2b_model."):
    """B2BGenerator.restoration (b2b_generator.py:406-500) with cfg_scale 1 (guidance neutral): Heun steps on the
    linspace(0, 1, steps + 1) grid, a final Euler step, known pixels re-projected after every step, final clamp.
    y / y_cond [B, F, C, H, W]; init_noise = the randn_like(y) draw."""
    b, f = y.shape[:2]
    if mask is not None:
        mask = torch.clamp(mask, 0.0, 1.0)
        y_background = y * (1.0 - mask)
    else:
        y_background = y
    if labels is None:
        labels = torch.zeros(b, dtype=torch.long)
    x = y_background + init_noise * cfg.noise_scale
    if mask is not None:
        x = x * mask + y * (1.0 - mask)
    ts = torch.linspace(0.0, 1.0, steps + 1)

    def project(v):
        return v if mask is None else v * mask + y * (1.0 - mask)

    def velocity(xc, t):
        x_in = project(xc)
        inp = x_in if y_cond is None else torch.cat([y_cond, x_in], dim=2)
        xp = jit_vid_forward(sd, inp, torch.full((b * f,), float(t)), labels, cfg, prefix=prefix)
        xp = project(xp[:, :, -x_in.shape[2]:])
        den = 1.0 - t
        if not disable_inference_clipping:
            den = den.clamp_min(cfg.t_eps)
        return (xp - x_in) / den

    with torch.no_grad():
        for i in range(steps - 1):
            t, tn = ts[i], ts[i + 1]
            v_t = velocity(x, t)
            v_n = velocity(x + (tn - t) * v_t, tn)
            x = x + (tn - t) * 0.5 * (v_t + v_n)
            if clip_denoised:
                x = x.clamp(-1.0, 1.0)
            x = project(x)
        x = x + (ts[-1] - ts[-2]) * velocity(x, ts[-2])
        if clip_denoised:
            x = x.clamp(-1.0, 1.0)
        return project(x).clamp(-1.0, 1.0)
