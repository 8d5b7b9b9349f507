"""TEST INFRASTRUCTURE (see oracle/__init__.py) — CPU restatement of the reference-image-conditioned UNet
(`UNetGeneratorRefAttn` / `AttentionBlockRef` / `EmbedSequentialRef`,
models/modules/unet_generator_attn/unet_generator_attn.py:1017-1645), row a-16 of SURVEY.md section 8.

Pinned by `tests/golden/refattn_small.pt` (oracle/gen_golden_ref.py, generated from the UNMODIFIED reference).
Functional style on a state_dict with the reference's key names; ResBlocks reuse oracle.palette_oracle.
"""
from dataclasses import dataclass
from typing import List

import torch
import torch.nn.functional as F

from . import palette_oracle as O
from .palette_oracle import _r, _rw


@dataclass
class RSpec:
    kind: str  # "conv" | "res" | "attn"
    cin: int = 0
    cout: int = 0
    up: bool = False
    down: bool = False
    heads: int = 0
    use_ref: bool = False
    terminal: bool = False


def ref_structure(cfg: O.UNetCfg):
    """UNetGeneratorRefAttn.__init__ (:1225-1540): (input_blocks, middle_block, output_blocks, input_blocks_ref,
    middle_block_ref, output_blocks_ref) as lists of lists of RSpec."""

    def heads_for(ch):
        return cfg.num_heads if cfg.num_head_channels == -1 else ch // cfg.num_head_channels

    def encoder(use_ref, ch):
        blocks = [[RSpec("conv", cfg.in_channel, ch)]]
        chans = [ch]
        ds = 1
        for level, mult in enumerate(cfg.channel_mults):
            for _ in range(cfg.res_blocks[level]):
                out = int(mult * cfg.inner_channel)
                layers = [RSpec("res", ch, out)]
                ch = out
                if ds in cfg.attn_res:
                    layers.append(RSpec("attn", ch, ch, heads=heads_for(ch), use_ref=use_ref))
                blocks.append(layers)
                chans.append(ch)
            if level != len(cfg.channel_mults) - 1:
                blocks.append([RSpec("res", ch, ch, down=True)])
                chans.append(ch)
                ds *= 2
        middle = [RSpec("res", ch, ch), RSpec("attn", ch, ch, heads=heads_for(ch), use_ref=use_ref),
                  RSpec("res", ch, ch)]
        return blocks, middle, chans, ch, ds

    inp, mid, chans, ch, ds = encoder(True, int(cfg.channel_mults[0] * cfg.inner_channel))
    # the reference builds the second encoder with `ch` carried over from the first one (:1330-1332): its first conv
    # has as many output channels as the bottleneck
    inp_r, mid_r, _, _, _ = encoder(False, ch)
    ch_ref, ds_ref, chans_ref = ch, ds, list(chans)
    out_r: List[List[RSpec]] = []
    is_terminal = False
    for level, mult in list(enumerate(cfg.channel_mults))[::-1]:
        for i in range(cfg.res_blocks[level] + 1):
            is_terminal = i == cfg.res_blocks[level] and ds_ref / 2 not in cfg.attn_res
            ich = chans_ref.pop()
            out = int(cfg.inner_channel * mult)
            layers = [RSpec("res", ch_ref + ich, out)]
            ch_ref = out
            if ds_ref in cfg.attn_res:
                layers.append(RSpec("attn", ch_ref, ch_ref, heads=heads_for(ch_ref), terminal=is_terminal))
            if level and i == cfg.res_blocks[level]:
                if not is_terminal:
                    layers.append(RSpec("res", ch_ref, ch_ref, up=True))
                ds_ref //= 2
            out_r.append(layers)
        if is_terminal:
            break
    outb = []
    for level, mult in list(enumerate(cfg.channel_mults))[::-1]:
        for i in range(cfg.res_blocks[level] + 1):
            ich = chans.pop()
            out = int(cfg.inner_channel * mult)
            layers = [RSpec("res", ch + ich, out)]
            ch = out
            if ds in cfg.attn_res:
                layers.append(RSpec("attn", ch, ch, heads=heads_for(ch), use_ref=True))
            if level and i == cfg.res_blocks[level]:
                layers.append(RSpec("res", ch, ch, up=True))
                ds //= 2
            outb.append(layers)
    return inp, mid, outb, inp_r, mid_r, out_r


def attention_block_ref(sd, name, x, b: RSpec, qkv_ref):
    """AttentionBlockRef._forward (:1098-1130), QKVAttentionLegacy order (the class default)."""
    bsz, c, hh, ww = x.shape
    xf = x.reshape(bsz, c, -1)
    xn = _r(F.instance_norm(xf.float(), eps=1e-5).type(xf.dtype))
    qkv = _r(F.conv1d(xn, _rw(sd[name + ".qkv.weight"]), sd[name + ".qkv.bias"]))
    if b.terminal:
        return None, qkv
    h = _r(O.qkv_attention_legacy(qkv, b.heads))
    if b.use_ref:
        q, _, _ = qkv.chunk(3, dim=1)
        _, k_ref, v_ref = qkv_ref.chunk(3, dim=1)
        h_ref = _r(O.qkv_attention_legacy(torch.cat([q, k_ref, v_ref], dim=1), b.heads))
        h = torch.cat([h, h_ref], dim=1)
    h = F.conv1d(h, _rw(sd[name + ".proj_out.weight"]), sd[name + ".proj_out.bias"])
    return _r(xf + h).reshape(bsz, c, hh, ww), qkv


def _run_block(sd, name, layers, h, emb, cfg, qkv_ref=None):
    """EmbedSequentialRef.forward (:1023-1038)."""
    qkv = []
    for j, b in enumerate(layers):
        n = "%s.%d" % (name, j)
        if b.kind == "conv":
            h = _r(O._conv2d(h, sd[n + ".weight"], sd[n + ".bias"], padding=1))
        elif b.kind == "res":
            h = O.res_block(sd, n, h, emb, b, cfg)
        else:
            cur = qkv_ref if (qkv_ref is None or type(qkv_ref) != list) else qkv_ref.pop(0)
            h, q = attention_block_ref(sd, n, h, b, cur)
            qkv.append(q)
    return h, qkv


def unet_ref_forward(sd, x, emb, ref, cfg: O.UNetCfg, prefix=""):
    """UNetGeneratorRefAttn.compute_feats + forward (:1565-1636)."""
    inp, mid, outb, inp_r, mid_r, out_r = ref_structure(cfg)
    qkv_list, hs_ref = [], []
    h = _r(torch.cat([ref, ref], dim=1).float())
    for i, layers in enumerate(inp_r):
        h, q = _run_block(sd, prefix + "input_blocks_ref.%d" % i, layers, h, emb, cfg)
        qkv_list.append(q)
        hs_ref.append(h)
    h_ref, q = _run_block(sd, prefix + "middle_block_ref", mid_r, h, emb, cfg)
    qkv_list.append(q)
    hs = []
    h = _r(x.float())
    for i, layers in enumerate(inp):
        h, _ = _run_block(sd, prefix + "input_blocks.%d" % i, layers, h, emb, cfg, qkv_ref=qkv_list.pop(0))
        hs.append(h)
    h, _ = _run_block(sd, prefix + "middle_block", mid, h, emb, cfg, qkv_ref=qkv_list.pop(0))
    qkv_list = []
    for i, layers in enumerate(out_r):
        h_ref = torch.cat([h_ref, hs_ref.pop()], dim=1)
        h_ref, q = _run_block(sd, prefix + "output_blocks_ref.%d" % i, layers, h_ref, emb, cfg)
        qkv_list.append(q)
    for i, layers in enumerate(outb):
        h = torch.cat([h, hs.pop()], dim=1)
        h, _ = _run_block(sd, prefix + "output_blocks.%d" % i, layers, h, emb, cfg,
                          qkv_ref=qkv_list.pop(0) if qkv_list else None)
    h = _r(F.silu(O.group_norm(h, sd[prefix + "out.0.norm.weight"], sd[prefix + "out.0.norm.bias"],
                               cfg.group_norm_size)))
    return _r(O._conv2d(h, sd[prefix + "out.2.weight"], sd[prefix + "out.2.bias"], padding=1))


def denoiser(ref, prefix="denoise_fn.model."):
    """The `unet` callable of oracle.palette_oracle.diffusion_forward / restoration_* for cfg 4: PaletteDenoiseFn hands
    the dataloader's reference image to the three-argument UNet (palette_denoise_fn.py:40-41, 111-112)."""
    return lambda sd, x, emb, cfg: unet_ref_forward(sd, x, emb, ref, cfg, prefix=prefix)
