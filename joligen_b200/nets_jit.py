"""B200 mirror of the b2b video backbone (BASELINE.json config 5 as written):
`JiTViD` (/root/reference/models/modules/vit/vit_vid.py:1234-1358) inside `B2BGenerator`
(/root/reference/models/modules/b2b_generator.py:238-348), for the path `example_b2b_vid_mario.json` takes: no
mask-size / frame-step / global-context / object-reference conditioning, no register tokens, one MotionModule after the
last block (`motion_every == 0`).

Same sub-module names => the reference's `state_dict` keys.  A clip `[B, F, C, H, W]` becomes ONE bf16 token tensor
`[N = B*F, T, 1, D]` (NHWC with H = T, W = 1): every `nn.Linear` / the patch embedding is a 1x1 tcgen05 convolution on
it, RMSNorm + adaLN modulate, per-head q/k RMSNorm + rotary, the per-frame attention over T <= 128 tokens, SwiGLU and the
gated residual are the kernels of csrc/jit.cu (ops_jit.py), the MotionModule on the patch grid is nets_vid's (the same
temporal kernels as the video UNet).  The O(B) conditioning vectors (timestep / label embeddings, adaLN outputs) are fp32.
"""
import math

import torch
import torch.nn as nn

from . import lib as L
from . import ops
from . import ops_jit as J
from .nets import ConvPack
from .nets_vid import MotionModule


def _lin(x, lin, pack, weight=None):
    """nn.Linear (or a 1x1 / patch convolution viewed as one) on bf16 tokens [N, T, 1, I] -> [N, T, 1, O]."""
    w = lin.weight if weight is None else weight
    while w.dim() < 4:
        w = w.unsqueeze(-1)
    return ops.conv2d(x, w, lin.bias, pack.get(), stride=1, pad=0)


def _lin_vec(v, lin, pack, silu=False):
    """nn.Linear on the O(B) fp32 conditioning vectors [N, I] -> fp32 [N, O] through the same 1x1 tensor-core path (one
    token per frame).  The thread-per-output fp32 kernels of ops.linear are meant for 32-wide embeddings: at 768 -> 4608
    their backward took 1.6 ms per adaLN Linear, 25 of the 47 ms of a step (profiles/r02_breakdown_cfg6.json)."""
    if silu:
        v = torch.nn.functional.silu(v)
    y = _lin(v.to(torch.bfloat16)[:, None, None, :].contiguous(), lin, pack)
    return y.float().reshape(v.shape[0], -1)


class _AsLinear:
    """A k x k stride-k patch convolution seen as the Linear it is on patchified tokens: weight [O, C*k*k]."""

    def __init__(self, conv):
        self.conv = conv

    @property
    def weight(self):
        return self.conv.weight.view(self.conv.weight.shape[0], -1)

    @property
    def bias(self):
        return self.conv.bias


def timestep_embedding(t, dim=256, max_period=10000):
    """TimestepEmbedder.timestep_embedding (vit_vid.py:107-122)"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rope_tables(head_dim, grid, num_prefix, device):
    """VisionRotaryEmbeddingFast(dim = head_dim / 2, pt_seq_len = grid) (util/model_util.py:97-162): cos / sin
    [num_prefix + grid^2, head_dim]; the prefix (in-context) tokens rotate by the identity."""
    dim = head_dim // 2
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid).float()
    fr = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)
    fr = torch.cat([fr[:, None, :].expand(grid, grid, dim), fr[None, :, :].expand(grid, grid, dim)], dim=-1)
    fr = fr.reshape(grid * grid, -1)
    cos, sin = fr.cos(), fr.sin()
    if num_prefix > 0:
        cos = torch.cat([torch.ones(num_prefix, cos.shape[1]), cos], dim=0)
        sin = torch.cat([torch.zeros(num_prefix, sin.shape[1]), sin], dim=0)
    return cos.contiguous().to(device), sin.contiguous().to(device)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class BottleneckPatchEmbed(nn.Module):
    """(:51-87) conv p x p stride p without bias -> conv 1x1."""

    def __init__(self, patch_size, in_chans, pca_dim, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.proj1 = nn.Conv2d(in_chans, pca_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.proj2 = nn.Conv2d(pca_dim, embed_dim, kernel_size=1, stride=1, bias=True)
        self._lin1 = _AsLinear(self.proj1)
        self._pack1 = ConvPack(self._lin1)
        self._pack2 = ConvPack(self.proj2)

    def forward_tokens(self, x):
        n, c, hh, ww = x.shape
        p = self.patch_size
        hp, wp = hh // p, ww // p
        tok = x.reshape(n, c, hp, p, wp, p).permute(0, 2, 4, 1, 3, 5).reshape(n, hp * wp, 1, c * p * p)
        tok = tok.to(torch.bfloat16).contiguous()
        return _lin(_lin(tok, self._lin1, self._pack1), self.proj2, self._pack2)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size
        self._pack0 = ConvPack(self.mlp[0])
        self._pack2 = ConvPack(self.mlp[2])

    def forward(self, t):
        h = _lin_vec(timestep_embedding(t, self.frequency_embedding_size), self.mlp[0], self._pack0)
        return _lin_vec(h, self.mlp[2], self._pack2, silu=True)


class LabelEmbedder(nn.Module):
    def __init__(self, num_classes, hidden_size):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + 1, hidden_size)

    def forward(self, labels):
        return self.embedding_table(labels)


class Attention(nn.Module):
    """(:182-231)"""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        hd = dim // num_heads
        self.q_norm = RMSNorm(hd)
        self.k_norm = RMSNorm(hd)
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        self._pack_qkv = ConvPack(self.qkv)
        self._pack_proj = ConvPack(self.proj)

    def forward_tokens(self, x, cos, sin):
        qkv = _lin(x, self.qkv, self._pack_qkv)
        qk = J.qknorm_rope(qkv, self.q_norm.weight, self.k_norm.weight, cos, sin, self.num_heads, self.q_norm.eps)
        return _lin(J.attn_small(qk, qkv, self.num_heads), self.proj, self._pack_proj)


class SwiGLUFFN(nn.Module):
    """(:234-246) hidden = int(hidden_dim * 2 / 3)"""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        hidden_dim = int(hidden_dim * 2 / 3)
        if hidden_dim % 8:
            raise NotImplementedError("B200 SwiGLUFFN: hidden width %d is not a multiple of 8 channels" % hidden_dim)
        self.w12 = nn.Linear(dim, 2 * hidden_dim, bias=True)
        self.w3 = nn.Linear(hidden_dim, dim, bias=True)
        self._pack12 = ConvPack(self.w12)
        self._pack3 = ConvPack(self.w3)

    def forward_tokens(self, x):
        return _lin(J.swiglu(_lin(x, self.w12, self._pack12)), self.w3, self._pack3)


class JiTBlock(nn.Module):
    """(:249-280)"""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = RMSNorm(hidden_size)
        self.attn = Attention(hidden_size, num_heads)
        self.norm2 = RMSNorm(hidden_size)
        self.mlp = SwiGLUFFN(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))
        self._pack_ada = ConvPack(self.adaLN_modulation[1])

    def forward_tokens(self, x, c, cos, sin):
        d = x.shape[-1]
        mod = _lin_vec(c, self.adaLN_modulation[1], self._pack_ada, silu=True)
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = [mod[:, i * d:(i + 1) * d] for i in range(6)]
        a = self.attn.forward_tokens(J.rmsnorm_mod(x, self.norm1.weight, shift_msa, scale_msa, self.norm1.eps), cos, sin)
        x = J.gated_residual(x, a, gate_msa)
        m = self.mlp.forward_tokens(J.rmsnorm_mod(x, self.norm2.weight, shift_mlp, scale_mlp, self.norm2.eps))
        return J.gated_residual(x, m, gate_mlp)


class FinalLayer(nn.Module):
    """(:283-308)"""

    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.norm_final = RMSNorm(hidden_size)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))
        self._pack = ConvPack(self.linear)
        self._pack_ada = ConvPack(self.adaLN_modulation[1])

    def forward_tokens(self, x, c):
        d = x.shape[-1]
        mod = _lin_vec(c, self.adaLN_modulation[1], self._pack_ada, silu=True)
        h = J.rmsnorm_mod(x, self.norm_final.weight, mod[:, :d], mod[:, d:], self.norm_final.eps)
        return _lin(h, self.linear, self._pack)


class JiTViD(nn.Module):
    def __init__(self, input_size=128, patch_size=16, in_channels=6, out_channels=3, hidden_size=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, num_classes=1, bottleneck_dim=None, in_context_len=32, in_context_start=4,
                 max_frames=8, motion_num_heads=8, motion_num_layers=2, motion_every=0):
        super().__init__()
        if motion_every != 0:
            raise NotImplementedError("B200 JiTViD: per-layer motion modules (motion_every > 0) are not supported")
        if (hidden_size // num_heads) not in (16, 32, 64):
            raise NotImplementedError("B200 JiTViD: head dim %d (16, 32, 64)" % (hidden_size // num_heads))
        self.input_size, self.patch_size, self.in_channels, self.out_channels = input_size, patch_size, in_channels, out_channels
        self.hidden_size, self.num_heads, self.depth = hidden_size, num_heads, depth
        self.in_context_len, self.in_context_start, self.max_frames = in_context_len, in_context_start, max_frames
        grid = input_size // patch_size
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.y_embedder = LabelEmbedder(num_classes, hidden_size)
        self.x_embedder = BottleneckPatchEmbed(patch_size, in_channels, bottleneck_dim or hidden_size, hidden_size)
        # "fixed" sin-cos table of the reference: a Parameter of its state_dict that b2b_model's set_requires_grad turns
        # trainable (oracle/jit_oracle.py pins that quirk): kept a Parameter, values come with the state_dict
        self.pos_embed = nn.Parameter(torch.zeros(1, grid * grid, hidden_size), requires_grad=False)
        if in_context_len > 0:
            self.in_context_posemb = nn.Parameter(torch.zeros(1, in_context_len, hidden_size))
        self.blocks = nn.ModuleList([JiTBlock(hidden_size, num_heads, mlp_ratio) for _ in range(depth)])
        self._clip = {"frames": 1}
        self.motion_module = MotionModule(in_channels=hidden_size, num_attention_heads=motion_num_heads,
                                          num_transformer_block=motion_num_layers,
                                          attention_block_types=("Temporal_Self", "Temporal_Self"),
                                          temporal_position_encoding=True,
                                          temporal_position_encoding_max_len=max_frames, clip=self._clip)
        self.final_layer = FinalLayer(hidden_size, patch_size, out_channels)
        self._rope = {}

    def _rope_for(self, prefix, device):
        key = (prefix, str(device))
        if key not in self._rope:
            self._rope[key] = rope_tables(self.hidden_size // self.num_heads, self.input_size // self.patch_size, prefix,
                                          device)
        return self._rope[key]

    def unpatchify(self, x, b, f):
        """(:1062-1082) tokens [N, T, p*p*C] -> [B, F, C, H, W]"""
        p, c = self.patch_size, self.out_channels
        n, t, _ = x.shape
        h = w = int(t ** 0.5)
        x = x.reshape(n, h, w, p, p, c)
        x = torch.einsum("nhwpqc->nchpwq", x).reshape(n, c, h * p, w * p)
        return x.reshape(b, f, c, h * p, w * p)

    def forward(self, x, t, y):
        """x [B, F, C, H, W] fp32, t [B] or [B*F] in [0, 1], y [B] labels -> [B, F, out_channels, H, W] fp32."""
        b, f, c, hh, ww = x.shape
        n = b * f
        hp, wp = hh // self.patch_size, ww // self.patch_size
        tok = self.x_embedder.forward_tokens(x.reshape(n, c, hh, ww).float())
        tok = (tok.float() + self.pos_embed[:, :, None, :]).to(torch.bfloat16)
        t = t.reshape(-1)
        t2 = t.repeat_interleave(f) if t.shape[0] == b else t
        y = y.reshape(-1)
        y2 = y.repeat_interleave(f) if y.shape[0] == b else y
        y_emb = self.y_embedder(y2)
        cvec = self.t_embedder(t2) + y_emb
        for i, blk in enumerate(self.blocks):
            if i == self.in_context_start and self.in_context_len > 0:
                ctx = y_emb.unsqueeze(1).repeat(1, self.in_context_len, 1) + self.in_context_posemb
                tok = torch.cat([ctx[:, :, None, :].to(torch.bfloat16), tok], dim=1).contiguous()
            prefix = self.in_context_len if (i >= self.in_context_start and self.in_context_len > 0) else 0
            cos, sin = self._rope_for(prefix, tok.device)
            tok = blk.forward_tokens(tok, cvec, cos, sin)
        if self.depth > self.in_context_start and self.in_context_len > 0:
            tok = tok[:, self.in_context_len:].contiguous()
        self._clip["frames"] = f
        d = tok.shape[-1]
        grid = self.motion_module.forward_nhwc(tok.reshape(n, hp, wp, d))
        tok = grid.reshape(n, hp * wp, 1, d)
        out = self.final_layer.forward_tokens(tok, cvec)
        return self.unpatchify(out.float().reshape(n, hp * wp, -1), b, f)


class B2BGenerator(nn.Module):
    """b2b_generator.B2BGenerator (:238-348) around `b2b_model` (JiTViD): flow-matching forward with explicit randomness
    (t_base [B], e = randn_like(x)) and the masked pseudo-Huber loss of B2BModel._masked_region_loss
    (/root/reference/models/b2b_model.py:1201-1217, incl. its one-channel-mask broadcast quirk)."""

    def __init__(self, b2b_model, t_eps=0.05, noise_scale=1.0, P_mean=-0.8, P_std=0.8, timestep_uniform_mix_prob=0.0,
                 label_drop_prob=0.0, num_classes=1, denoise_timesteps=50):
        super().__init__()
        self.b2b_model = b2b_model
        self.t_eps, self.noise_scale, self.P_mean, self.P_std = t_eps, noise_scale, P_mean, P_std
        self.timestep_uniform_mix_prob = float(timestep_uniform_mix_prob)   # --alg_b2b_timestep_uniform_mix_prob
        self.label_drop_prob = min(max(float(label_drop_prob), 0.0), 1.0)   # --alg_diffusion_dropout_prob (:81-84)
        self.num_classes = num_classes
        self.denoise_timesteps = denoise_timesteps
        # sampler options (b2b_generator.py:46-55); accelerate() copies the reference generator's values.  The velocity
        # denominator 1 - t is clamped at t_eps during sampling unless disable_inference_clipping (this mirror's
        # historical default; the two coincide whenever 1 / steps >= t_eps)
        self.clip_denoised_default = False
        self.disable_inference_clipping = True

    def sample_t(self, n, device):
        """(:192-210) logit-normal, optionally mixed with uniform draws; the draws come in the reference's order"""
        t = torch.sigmoid(torch.randn(n, device=device) * self.P_std + self.P_mean)
        if self.timestep_uniform_mix_prob <= 0.0:
            return t
        if self.timestep_uniform_mix_prob >= 1.0:
            return torch.rand_like(t)
        t_uniform = torch.rand_like(t)
        return torch.where(torch.rand_like(t) < self.timestep_uniform_mix_prob, t_uniform, t)

    def drop_labels(self, labels):
        """(:212-216) classifier-free label dropout: the dropped samples get the extra class `num_classes`"""
        if self.label_drop_prob <= 0.0:
            return labels
        drop = torch.rand(labels.shape, device=labels.device) < self.label_drop_prob
        return torch.where(drop, torch.full_like(labels, self.num_classes), labels)

    def forward_flow(self, x, mask, x_cond, label, t_base=None, e=None, use_gt=None, ref_idx=None):
        """b2b_forward + forward (:238-348) for clips [B, F, C, H, W] -> (v_pred, v, x_pred, raw x_pred).  Random draws
        in the reference's order: label dropout, t, then the noise."""
        b, f = x.shape[:2]
        if label is None:
            label = torch.zeros(b, dtype=torch.long, device=x.device)
        elif self.training:
            label = self.drop_labels(label)
        if t_base is None:
            t_base = self.sample_t(b, x.device)
        if e is None:
            e = torch.randn_like(x)
        t_cont = t_base[:, None].repeat(1, f)
        if use_gt is not None and ref_idx is not None and bool(use_gt.any()):
            b_idx = torch.arange(b, device=x.device)      # autoregressive training: the reference frame is clean (:266-268)
            t_cont[b_idx[use_gt], ref_idx[use_gt]] = 1.0
        t = t_cont.view(b, f, 1, 1, 1)
        if mask is not None:
            mask = torch.clamp(mask, min=0.0, max=1.0)
        z_t = t * x + (1.0 - t) * (e * self.noise_scale)
        z = z_t * mask + (1.0 - mask) * x if mask is not None else z_t
        z_model = z if x_cond is None else torch.cat([x_cond, z], dim=2)
        v = (x - z) / (1.0 - t).clamp_min(self.t_eps)
        raw = self.b2b_model(z_model, t_cont.reshape(b * f), label)
        if raw.shape[2] > x.shape[2]:
            raw = raw[:, :, -x.shape[2]:]
        x_pred = raw * mask + (1 - mask) * x if mask is not None else raw
        v_pred = (x_pred - z) / (1 - t).clamp_min(self.t_eps)
        return v_pred, v, x_pred, raw

    def forward(self, x, mask=None, x_cond=None, label=None, use_gt=None, ref_idx=None, temporal_frame_step=None,
                global_context=None, object_refs=None, return_x_pred=False, return_raw_x_pred=False, t_base=None,
                e=None):
        """The reference's call signature (b2b_model.py:1112-1122 calls it with keywords): (v_pred, v) [, x_pred
        [, raw x_pred]].  t_base / e: explicit random draws (tests)."""
        if temporal_frame_step is not None or global_context is not None or object_refs is not None:
            raise NotImplementedError("B200 B2BGenerator: frame-step / global-context / object-reference conditioning")
        if x.dim() != 5:
            raise NotImplementedError("B200 B2BGenerator: clips [B, F, C, H, W] only (the vit_vid backbone)")
        v_pred, v, x_pred, raw = self.forward_flow(x, mask, x_cond, label, t_base, e, use_gt, ref_idx)
        if return_raw_x_pred:
            return v_pred, v, x_pred, raw
        if return_x_pred:
            return v_pred, v, x_pred
        return v_pred, v

    @staticmethod
    def masked_region_loss(pred, target, mask, eps=1e-8):
        c = 0.00054 * math.sqrt(math.prod(pred.shape[1:]))
        le = torch.sqrt((pred - target) ** 2 + c ** 2) - c
        dims = tuple(range(1, le.ndim))
        return ((le * mask).sum(dim=dims) / mask.sum(dim=dims).clamp_min(eps)).mean()

    def forward_loss(self, x, mask, x_cond, label, t_base=None, e=None, lambda_G=1.0):
        v_pred, v, _, _ = self.forward_flow(x, mask, x_cond, label, t_base, e)
        return lambda_G * self.masked_region_loss(v_pred, v, torch.clamp(mask, min=0, max=1))

    @torch.no_grad()
    def restoration(self, y, y_cond=None, denoise_timesteps=None, mask=None, labels=None, clip_denoised=None,
                    use_gt=None, ref_idx=None, init_noise=None, temporal_frame_step=None, global_context=None,
                    object_refs=None, disable_inference_clipping=None):
        """b2b_generator.B2BGenerator.restoration (:406-500, same argument order: b2b_model.py:1312-1349 calls it
        positionally up to `labels` and by keyword after) with guidance neutral (cfg_scale 1): Heun steps on the
        linspace(0, 1, steps + 1) grid, a final Euler step, known pixels re-projected after every step, final clamp."""
        if temporal_frame_step is not None or global_context is not None or object_refs is not None:
            raise NotImplementedError("B200 B2BGenerator: frame-step / global-context / object-reference conditioning")
        if use_gt is not None and ref_idx is not None and bool(use_gt.any()):
            raise NotImplementedError("B200 B2BGenerator.restoration: autoregressive reference frames")
        b, f = y.shape[:2]
        if denoise_timesteps is None:
            denoise_timesteps = self.denoise_timesteps
        if isinstance(denoise_timesteps, (list, tuple)):
            denoise_timesteps = denoise_timesteps[0]
        if clip_denoised is None:
            clip_denoised = self.clip_denoised_default              # --alg_b2b_clip_denoised
        if disable_inference_clipping is None:
            disable_inference_clipping = self.disable_inference_clipping   # --alg_b2b_disable_inference_clipping
        steps = int(denoise_timesteps)
        if mask is not None:
            mask = torch.clamp(mask, 0.0, 1.0)
        if labels is None:
            labels = torch.zeros(b, dtype=torch.long, device=y.device)
        if init_noise is None:
            init_noise = torch.randn_like(y)

        def project(v):
            return v if mask is None else v * mask + y * (1.0 - mask)

        x = project((y * (1.0 - mask) if mask is not None else y) + init_noise * self.noise_scale)
        ts = torch.linspace(0.0, 1.0, steps + 1, device=y.device)

        def velocity(xc, t):
            x_in = project(xc)
            inp = x_in if y_cond is None else torch.cat([y_cond, x_in], dim=2)
            xp = self.b2b_model(inp, torch.full((b * f,), float(t), device=y.device), labels)
            xp = project(xp[:, :, -x_in.shape[2]:])
            den = 1.0 - t
            if not disable_inference_clipping:
                den = den.clamp_min(self.t_eps)
            return (xp - x_in) / den

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

