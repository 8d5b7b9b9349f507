"""B200-backed mirrors of the reference's hot-path nn.Modules.

Same constructor arguments, same sub-module names and therefore the same `state_dict` keys as

    models/modules/unet_generator_attn/unet_generator_attn.py  (UNet, ResBlock, AttentionBlock, EmbedSequential)
    models/modules/unet_generator_attn/unet_attn_utils.py      (GroupNorm wrapper, normalization)
    models/modules/palette_denoise_fn.py                       (PaletteDenoiseFn)
    models/modules/diffusion_generator.py                      (DiffusionGenerator)

Parameters live in ordinary nn.Conv2d / nn.GroupNorm / nn.Linear containers (fp32, reference layout);
their `forward` is never used: the blocks call the sm_100a kernels through joligen_b200.ops on NHWC
bf16 activations.  `accelerate.accelerate()` swaps a reference-built tree for these classes while
SHARING the nn.Parameter objects.  There is no CPU path: calling forward without CUDA raises.
"""
import inspect
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K
from . import lib as L
from . import ops

# bumped by anything that rewrites parameters behind torch's back (fused optimizer, load_state_dict)
_PACK_EPOCH = [0]


def invalidate_packed_weights():
    _PACK_EPOCH[0] += 1


# The batched pack / staged wgrad tables hold raw DEVICE pointers: only CUDA weights take part.  (Cleared only by the
# kernel test double of tests/, which follows the same pointers in host memory.)
_DEVICE_WEIGHTS_ONLY = [True]


def on_device(w):
    return w.is_cuda or not _DEVICE_WEIGHTS_ONLY[0]


class ConvPack:
    """bf16 implicit-GEMM copies of one conv's fp32 master weight, refreshed when the weight changes.
    A trainer may take over (`managed`): it owns persistent packed buffers and refreshes ALL convolutions of the
    model with one batched launch after every optimizer step (WeightPackSet); get() then only hands them out."""

    def __init__(self, conv):
        self.conv = conv
        self.key = None
        self.packed = None
        self.managed_epoch = None  # _PACK_EPOCH value at which a trainer's batched pack last refreshed this entry
        self.managed_key = None    # (weight._version, data_ptr, bias._version) seen by that refresh

    def _weight4(self):
        w4 = self.conv.weight.detach()
        while w4.dim() < 4:  # Conv1d k=1 / nn.Linear
            w4 = w4.unsqueeze(-1)
        return w4

    def _bias_padded(self):
        b = self.conv.bias
        if b is None:
            return None
        cout = b.shape[0]
        cout8 = (cout + 7) // 8 * 8
        if cout8 != cout:
            bias_p = torch.zeros(cout8, dtype=torch.float32, device=b.device)
            bias_p[:cout] = b.detach()
            return bias_p
        return b.detach()  # aliases the parameter storage: always current

    def _torch_key(self):
        w, b = self.conv.weight, self.conv.bias
        return (w._version, w.data_ptr(), None if b is None else b._version)

    def get(self):
        # Managed entries are current as long as nothing wrote the master weights THROUGH torch since the trainer's
        # last batched pack: the fused optimizer writes them through raw pointers (no version bump) and re-packs right
        # after, whereas load_state_dict / copy_ / manual init bump `_version` and must trigger a re-pack here.
        if (self.managed_epoch is not None and self.managed_epoch == _PACK_EPOCH[0]
                and self.managed_key == self._torch_key()):
            return self.packed
        w = self.conv.weight
        b = self.conv.bias
        key = (w._version, _PACK_EPOCH[0], w.data_ptr(), None if b is None else (b._version, b.data_ptr()))
        if key != self.key:
            # a managed entry keeps its persistent buffers (the trainer's batched table points at them)
            out = (self.packed[0], self.packed[1]) if (self.managed_epoch is not None and self.packed) else None
            wf, wd = K.pack_conv_weight(self._weight4(), want_dgrad=True, out=out)
            bias_p = self._bias_padded()
            if out is not None and self.packed[2] is not None and bias_p is not None \
                    and self.packed[2].data_ptr() != bias_p.data_ptr():
                self.packed[2].copy_(bias_p)  # managed padded-bias copy: keep the buffer captured graphs point at
                bias_p = self.packed[2]
            self.packed = (wf, wd, bias_p)
            self.key = key
            if self.managed_epoch is not None:
                self.managed_epoch, self.managed_key = _PACK_EPOCH[0], self._torch_key()
        return self.packed


def conv_packs(module):
    """Every ConvPack reachable from `module`'s sub-modules (attributes holding a ConvPack or a list of them)."""
    out, seen = [], set()
    for m in module.modules():
        for v in vars(m).values():
            for c in (v if isinstance(v, (list, tuple)) else (v,)):
                if isinstance(c, ConvPack) and id(c) not in seen:
                    seen.add(id(c))
                    out.append(c)
    return out


class WeightPackSet:
    """All convolutions of a model packed by ONE launch (jg_pack_conv_weights_batched).  The packed buffers are
    persistent; call refresh() after the master weights changed (the trainer does, after every optimizer step)."""

    def __init__(self, module):
        self.packs = [p for p in conv_packs(module) if on_device(p.conv.weight)]
        items, dims = [], []
        for p in self.packs:
            w4 = p._weight4()
            cout, cin, r, s = w4.shape
            if not w4.is_contiguous() or w4.dtype != torch.float32:
                raise RuntimeError("WeightPackSet: master weights must be contiguous fp32")
            wf, wd = K.pack_conv_weight(w4, want_dgrad=True)  # allocates (zero-padded) and fills once
            p.packed = (wf, wd, p._bias_padded())
            cin8, cout8 = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
            items.append(L.PackItem(w4.data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, cin, r * s, cin8, cout8, 0))
            dims.append((cout, cin, r * s))
        self.table = K.WeightTable(items, dims, self.packs[0].conv.weight.device) if self.packs else None
        self._mark()

    def _mark(self):
        for p in self.packs:
            p.managed_epoch = _PACK_EPOCH[0]
            p.managed_key = p._torch_key()

    def refresh(self):
        if self.table is not None:
            K.pack_conv_weights_batched(self.table)
            for p in self.packs:  # padded biases are copies
                b = p.conv.bias
                if b is not None and p.packed[2] is not None and p.packed[2].data_ptr() != b.data_ptr():
                    p.packed[2][:b.shape[0]].copy_(b.detach())
        self._mark()


def _conv(x, conv, pack, residual=None, res_scale=1.0, out=None, want_stats=False):
    """want_stats: the output is read by a GroupNorm next (see ops.conv2d)."""
    w = conv.weight
    if w.dim() == 3:
        w = w.unsqueeze(-1)
    return ops.conv2d(x, w, conv.bias, pack.get(), stride=conv.stride[0], pad=conv.padding[0], residual=residual,
                      res_scale=res_scale, grad_sink=conv.weight, out=out, want_stats=want_stats)


class GroupNorm(nn.Module):
    """unet_attn_utils.GroupNorm: holds `self.norm = nn.GroupNorm(groups, channels)` (fp32 statistics)."""

    def __init__(self, group_size, channels):
        super().__init__()
        self.norm = nn.GroupNorm(group_size, channels)

    def forward_nhwc(self, x, film=None, act=L.ACT_NONE):
        return ops.group_norm(x, self.norm.weight, self.norm.bias, self.norm.num_groups, film=film, act=act)

    def forward_tap_nhwc(self, x, act=L.ACT_NONE):
        """-> (y, x_tap); every other consumer of x must read x_tap (gradient sum fused into the GN backward)."""
        return ops.group_norm_tap(x, self.norm.weight, self.norm.bias, self.norm.num_groups, film=None, act=act)

    def forward_tap2_nhwc(self, x, act=L.ACT_NONE):
        """-> (y, x_tap, x_tap2): a second hand-through for a consumer outside the block (the decoder's concat)."""
        return ops.group_norm_tap2(x, self.norm.weight, self.norm.bias, self.norm.num_groups, film=None, act=act)


def normalization(channels, norm="groupnorm32"):
    if "groupnorm" in norm:
        return GroupNorm(int(norm.split("groupnorm")[1]), channels)
    if norm == "instancenorm":
        return GroupNorm(channels, channels)
    if norm == "layernorm":
        return GroupNorm(1, channels)
    raise ValueError("%s is not implemented in the B200 UNet (groupnorm / instancenorm / layernorm only)" % norm)


class EmbedBlock(nn.Module):
    pass


class EmbedSequential(nn.Sequential, EmbedBlock):
    def forward_nhwc(self, x, emb, out=None, want_tap=False):
        """out: optional destination view for the LAST layer's output (see UNet.forward_nhwc).
        want_tap: also return a hand-through of the block INPUT made by the first layer's GroupNorm (a ResBlock):
        the encoder keeps it as the skip tensor, so that the concat's gradient is summed inside that GroupNorm's
        backward (-> (y, tap); tap is None when the first layer cannot provide one)."""
        last = len(self) - 1
        tap = None
        for i, layer in enumerate(self):
            kw = {"out": out} if (i == last and out is not None) else {}
            if i == 0 and want_tap and isinstance(layer, ResBlock):
                x, tap = layer.forward_nhwc(x, emb, want_tap=True, **kw)
            else:
                x = layer.forward_nhwc(x, emb, **kw) if isinstance(layer, EmbedBlock) else layer.forward_nhwc(x, **kw)
        return (x, tap) if want_tap else x

    def out_geometry(self, h, w):
        """(H, W, C) of the output for an (h, w) input, or None when the last layer cannot write into a view."""
        c = None
        for layer in self:
            if isinstance(layer, ResBlock):
                c = layer.out_channel
                if layer.updown:
                    h, w = (h * 2, w * 2) if layer.up else (h // 2, w // 2)
            elif isinstance(layer, AttentionBlock):
                c = layer.channels
            elif hasattr(layer, "temporal_transformer"):  # MotionModule (video UNet): shape-preserving
                pass
            else:
                return None
        return None if c is None else (h, w, (c + 7) // 8 * 8)


class _Haar(nn.Module):
    """freq_utils.HaarTransform / InverseHaarTransform (:21-59): holds the reference's four 2x2 filter buffers (they are
    part of its state_dict); the transform itself is jg_haar on fp32 NCHW tensors (csrc/prep.cu, bit-exact taps)."""

    def __init__(self, inverse):
        super().__init__()
        self.inverse = inverse
        s = 1 / (2 ** 0.5)
        low, high = s * torch.ones(1, 2), s * torch.ones(1, 2)
        high[0, 0] = -high[0, 0]
        sign = -1.0 if inverse else 1.0
        self.register_buffer("ll", low.T * low)
        self.register_buffer("lh", sign * (high.T * low))
        self.register_buffer("hl", sign * (low.T * high))
        self.register_buffer("hh", high.T * high)

    def forward(self, x):
        return ops.haar_iwt(x) if self.inverse else ops.haar_dwt(x)


class _Resample(nn.Module):
    """Upsample / Downsample with use_conv=False (what ResBlock(up=/down=) instantiates).  freq_space
    (--train_feat_wavelet, :69-96, 113-140): the tensor holds the four Haar bands of a 2x larger map with C/4 channels;
    resampling happens in pixel space: inverse transform -> resample -> transform."""

    def __init__(self, up, freq_space=False, channels=None):
        super().__init__()
        self.up = up
        self.freq_space = freq_space
        self.channels = channels
        if freq_space:
            if channels is None or channels % 4:
                raise ValueError("freq_space resampling needs a channel count that is a multiple of 4")
            self.iwt = _Haar(True)
            self.dwt = _Haar(False)

    def forward_nhwc(self, x):
        if not self.freq_space:
            return ops.upsample2x(x) if self.up else ops.avgpool2x(x)
        # (a composition of existing kernels through the fp32 NCHW layout the Haar kernel takes: correct, not tuned)
        pix = ops.to_nhwc(self.iwt(ops.to_nchw(x, self.channels)))
        pix = ops.upsample2x(pix) if self.up else ops.avgpool2x(pix)
        return ops.to_nhwc(self.dwt(ops.to_nchw(pix, self.channels // 4)))


class ConvIn(nn.Conv2d):
    """The first 3x3 conv of the UNet (input_blocks[0][0]), kept as an nn.Conv2d for its state_dict keys."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._pack = ConvPack(self)

    def forward_nhwc(self, x):
        return _conv(x, self, self._pack, want_stats=True)


class ResBlock(EmbedBlock):
    """unet_generator_attn.ResBlock (lines 143-266): GN->SiLU->[up/down]->conv3x3, FiLM GN->SiLU->conv3x3,
    skip 1x1 conv, residual add fused into the second conv's epilogue."""

    apply_skipw = True  # the video ResBlock computes skipw but never applies it (..._vid.py:272-275)

    def __init__(self, channels, emb_channels, dropout, norm, out_channel=None, use_conv=False,
                 use_scale_shift_norm=False, use_checkpoint=False, up=False, down=False, efficient=False,
                 freq_space=False):
        super().__init__()
        if use_conv or use_checkpoint:
            raise NotImplementedError("B200 ResBlock: use_conv / use_checkpoint are not supported")
        if dropout:
            raise NotImplementedError("B200 ResBlock: dropout > 0 is not supported")
        self.channels = channels
        self.emb_channels = emb_channels
        self.out_channel = out_channel or channels
        self.use_scale_shift_norm = use_scale_shift_norm
        self.up, self.down = up, down
        self.updown = up or down
        self.efficient = efficient
        self.in_layers = nn.Sequential(normalization(channels, norm), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channel, 3, padding=1))
        self.freq_space = freq_space
        if up or down:
            self.h_upd = _Resample(up, freq_space, channels)
            self.x_upd = _Resample(up, freq_space, channels)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(
            nn.SiLU(), nn.Linear(emb_channels, 2 * self.out_channel if use_scale_shift_norm else self.out_channel))
        self.out_layers = nn.Sequential(normalization(self.out_channel, norm), nn.SiLU(), nn.Dropout(p=dropout),
                                        nn.Conv2d(self.out_channel, self.out_channel, 3, padding=1))
        if self.out_channel == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channel, 1)
        self._pack_in = ConvPack(self.in_layers[2])
        self._pack_out = ConvPack(self.out_layers[3])
        self._pack_skip = ConvPack(self.skip_connection) if isinstance(self.skip_connection, nn.Conv2d) else None
        self._film_pre = None

    def forward_nhwc(self, x, emb, out=None, want_tap=False):
        tap2 = None
        if want_tap:
            h, x, tap2 = self.in_layers[0].forward_tap2_nhwc(x, act=L.ACT_SILU)
        else:
            h, x = self.in_layers[0].forward_tap_nhwc(x, act=L.ACT_SILU)
        if self.updown and self.efficient and self.up:
            # (:239-242) --G_unet_mha_vit_efficient: convolve at the low resolution, upsample afterwards (the
            # statistics of the upsampled tensor are taken by the GroupNorm's own pass)
            h = _conv(h, self.in_layers[2], self._pack_in)
            h = self.h_upd.forward_nhwc(h)
            x = self.x_upd.forward_nhwc(x)
        else:
            if self.updown:
                h = self.h_upd.forward_nhwc(h)
                x = self.x_upd.forward_nhwc(x)
            h = _conv(h, self.in_layers[2], self._pack_in, want_stats=True)
        lin = self.emb_layers[1]
        emb_out, self._film_pre = self._film_pre, None  # computed for all blocks at once by the UNet (EmbBank)
        if emb_out is None:
            emb_out = ops.linear(emb, lin.weight, lin.bias, act_in=L.ACT_SILU)  # [N, 2C] = (scale | shift)
        if emb_out.shape[0] != h.shape[0]:
            # video UNet: one embedding per clip, N = B*F frames (emb_out.repeat_interleave(f), ..._vid.py:260)
            emb_out = emb_out.repeat_interleave(h.shape[0] // emb_out.shape[0], dim=0)
        if self.use_scale_shift_norm:
            h = self.out_layers[0].forward_nhwc(h, film=emb_out, act=L.ACT_SILU)
        else:
            # h + emb_out, then GN -> SiLU (:259-261).  No option of the reference reaches this branch
            # (diffusion_networks.py:70/:231 pass use_scale_shift_norm=True), so the per-(n, c) add is a plain
            # broadcast add in front of the fused norm rather than another epilogue mode of the convolution.
            e = emb_out if emb_out.shape[1] == h.shape[-1] else F.pad(emb_out, (0, h.shape[-1] - emb_out.shape[1]))
            h = (h.float() + e[:, None, None, :].float()).to(h.dtype)
            h = self.out_layers[0].forward_nhwc(h, film=None, act=L.ACT_SILU)
        skipw = 1.0 / math.sqrt(2) if (self.efficient and self.apply_skipw) else 1.0
        if self._pack_skip is not None:
            x = _conv(x, self.skip_connection, self._pack_skip)
        # the block output is normalised next (the following block's in_layers / attention norm / the UNet's out head)
        y = _conv(h, self.out_layers[3], self._pack_out, residual=x, res_scale=skipw, out=out, want_stats=True)
        return (y, tap2) if want_tap else y

    def forward(self, x, emb):
        """Drop-in NCHW fp32 signature of the reference block."""
        y = self.forward_nhwc(ops.to_nhwc(x), emb)
        return ops.to_nchw(y, self.out_channel)


class EmbBank:
    """emb_layers of every ResBlock of a UNet evaluated in ONE launch: each ResBlock finds its
    (scale | shift) in `_film_pre` instead of running its own tiny Linear (and two tiny backward kernels)."""

    def __init__(self, root):
        self.blocks = [m for m in root.modules() if isinstance(m, ResBlock)]
        self.bank = None

    def distribute(self, emb):
        if not self.blocks:
            return
        lins = [b.emb_layers[1] for b in self.blocks]
        key = tuple((l.weight.data_ptr(), None if l.bias is None else l.bias.data_ptr()) for l in lins)
        if self.bank is None or self.bank.key != key:  # (parameters are re-pointed when a trainer flattens them)
            self.bank = K.LinearBank(lins, emb.device)
        outs = ops.linear_bank(emb, self.bank, lins, act_in=L.ACT_SILU)
        for blk, y in zip(self.blocks, outs):
            blk._film_pre = y


class _NoAffineInstanceNorm1d(nn.Module):
    """normalization1d(): InstanceNorm1d over T without affine (no parameters, unet_attn_utils.py:60-66)."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels


class AttentionBlock(nn.Module):
    """unet_generator_attn.AttentionBlock (lines 269-319) with QKVAttentionLegacy."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False, use_transformer=False):
        super().__init__()
        if use_transformer or use_checkpoint:
            raise NotImplementedError("B200 AttentionBlock: use_transformer / use_checkpoint are not supported")
        # QKVAttentionLegacy splits heads before q|k|v (layout 0); QKVAttention splits q|k|v first (layout 1)
        self.attention_layout = 1 if use_new_attention_order else 0
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0
            self.num_heads = channels // num_head_channels
        self.norm = _NoAffineInstanceNorm1d(channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = nn.Conv1d(channels, channels, 1)
        self._pack_qkv = ConvPack(self.qkv)
        self._pack_proj = ConvPack(self.proj_out)

    def forward_nhwc(self, x, out=None):
        c = self.channels
        xn, x = ops.group_norm_tap(x, None, None, c, film=None, act=L.ACT_NONE)  # per-(n, c) statistics over T
        qkv = _conv(xn, self.qkv, self._pack_qkv)
        a = ops.attention(qkv, self.num_heads, c // self.num_heads, self.attention_layout)
        return _conv(a, self.proj_out, self._pack_proj, residual=x, res_scale=1.0, out=out, want_stats=True)

    def forward(self, x):
        y = self.forward_nhwc(ops.to_nhwc(x))
        return ops.to_nchw(y, self.channels)


class _OutHead(nn.Sequential):
    pass


class UNet(nn.Module):
    """unet_generator_attn.UNet (lines 390-705), same constructor; forward(input NCHW fp32, embed_gammas)."""

    def __init__(self, image_size, in_channel, inner_channel, out_channel, res_blocks, attn_res, tanh,
                 n_timestep_train, n_timestep_test, norm, group_norm_size, cond_embed_dim, dropout=0,
                 channel_mults=(1, 2, 4, 8), conv_resample=True, use_checkpoint=False, use_fp16=False, num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True,
                 use_new_attention_order=False, efficient=False, freq_space=False):
        super().__init__()
        if tanh or not resblock_updown or use_fp16:
            raise NotImplementedError("B200 UNet: tanh / conv resampling / fp16 are not supported")
        self.freq_space = freq_space
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channel = in_channel
        self.inner_channel = inner_channel
        self.out_channel = out_channel
        self.res_blocks = res_blocks
        self.attn_res = attn_res
        self.channel_mults = channel_mults
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.cond_embed_dim = cond_embed_dim
        if freq_space:   # (:467-473) the layers work on the Haar bands of the input / output: 4x the channels at half
            self.iwt = _Haar(True)   # the resolution (the attributes above keep the pixel-space counts, like the reference)
            self.dwt = _Haar(False)
            in_channel, out_channel = in_channel * 4, out_channel * 4
        if norm == "groupnorm":
            norm = norm + str(group_norm_size)
        rb = dict(use_scale_shift_norm=use_scale_shift_norm, norm=norm, efficient=efficient, freq_space=freq_space)
        ch = input_ch = int(channel_mults[0] * inner_channel)
        self.input_blocks = nn.ModuleList([EmbedSequential(ConvIn(in_channel, ch, 3, padding=1))])
        input_block_chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mults):
            for _ in range(res_blocks[level]):
                layers = [ResBlock(ch, cond_embed_dim, 0.0, out_channel=int(mult * inner_channel), **rb)]
                ch = int(mult * inner_channel)
                if ds in attn_res:
                    layers.append(AttentionBlock(ch, num_heads=num_heads, num_head_channels=num_head_channels))
                self.input_blocks.append(EmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mults) - 1:
                self.input_blocks.append(
                    EmbedSequential(ResBlock(ch, cond_embed_dim, 0.0, out_channel=ch, down=True, **rb)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = EmbedSequential(
            ResBlock(ch, cond_embed_dim, dropout, **rb),
            AttentionBlock(ch, num_heads=num_heads, num_head_channels=num_head_channels),
            ResBlock(ch, cond_embed_dim, dropout, **rb))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mults))[::-1]:
            for i in range(res_blocks[level] + 1):
                ich = input_block_chans.pop()
                layers = [ResBlock(ch + ich, cond_embed_dim, 0.0, out_channel=int(inner_channel * mult), **rb)]
                ch = int(inner_channel * mult)
                if ds in attn_res:
                    layers.append(AttentionBlock(ch, num_heads=num_heads_upsample,
                                                 num_head_channels=num_head_channels))
                if level and i == res_blocks[level]:
                    layers.append(ResBlock(ch, cond_embed_dim, 0.0, out_channel=ch, up=True, **rb))
                    ds //= 2
                self.output_blocks.append(EmbedSequential(*layers))
        self.out = _OutHead(normalization(ch, norm), nn.SiLU(), nn.Conv2d(input_ch, out_channel, 3, padding=1))
        self._pack_outconv = ConvPack(self.out[2])
        self._emb_bank = None
        self.beta_schedule = {
            "train": {"schedule": "linear", "n_timestep": n_timestep_train, "linear_start": 1e-6,
                      "linear_end": 0.01},
            "test": {"schedule": "linear", "n_timestep": n_timestep_test, "linear_start": 1e-4, "linear_end": 0.09},
        }

    # -- NHWC bf16 fast path -------------------------------------------------------------------
    def forward_nhwc(self, x, emb):
        """x: NHWC bf16 [N,H,W,round_up(in_channel,8)] -> NHWC bf16 [N,H,W,round_up(out_channel,8)]."""
        if getattr(self, "freq_space", False):   # (:672, :692) Haar bands in, Haar bands out — through the fp32 NCHW layout of jg_haar
            bands = ops.to_nhwc(self.dwt(ops.to_nchw(x, self.in_channel)))
            out = self._forward_bands(bands, emb)
            return ops.to_nhwc(self.iwt(ops.to_nchw(out, 4 * self.out_channel)))
        return self._forward_bands(x, emb)

    def _forward_bands(self, x, emb):
        # Skip tensors: every encoder output h_k is consumed by the next block AND by the decoder's concat.  The
        # decoder reads a hand-through ("tap") of h_k made by the next block's first GroupNorm, so the two gradients
        # meet inside that GroupNorm's backward pass instead of in a separate (strided) add kernel.
        if emb.shape[0] <= 64 and emb.shape[1] <= 128:
            if getattr(self, "_emb_bank", None) is None:  # (subclasses build themselves without UNet.__init__)
                self._emb_bank = EmbBank(self)
            self._emb_bank.distribute(emb)
        hs = []
        h = x
        for module in self.input_blocks:
            h, tap = module.forward_nhwc(h, emb, want_tap=True)
            if tap is not None and hs:
                hs[-1] = tap
            hs.append(h)
        # torch.cat([h, hs.pop()], dim=1) (unet_generator_attn.py:687) without copying h: the block that produces h
        # writes it straight into the first channels of the next block's concat buffer; only the skip is copied.
        def concat_buffer(block, hin, win):
            geo = block.out_geometry(hin, win) if hs else None
            if geo is None or tuple(hs[-1].shape[1:3]) != geo[:2]:
                return None, None
            buf = torch.empty((h.shape[0], geo[0], geo[1], geo[2] + hs[-1].shape[-1]), dtype=torch.bfloat16,
                              device=h.device)
            return buf, buf[..., :geo[2]]

        buf, dst = concat_buffer(self.middle_block, h.shape[1], h.shape[2])
        h, tap = self.middle_block.forward_nhwc(h, emb, out=dst, want_tap=True)
        if tap is not None:
            hs[-1] = tap
        for module in self.output_blocks:
            skip = hs.pop()
            h = ops.cat_into(buf, h, skip) if buf is not None else ops.cat_channels(h, skip)
            buf, dst = concat_buffer(module, h.shape[1], h.shape[2])
            h = module.forward_nhwc(h, emb, out=dst)
        h = self.out[0].forward_nhwc(h, act=L.ACT_SILU)
        return _conv(h, self.out[2], self._pack_outconv)

    def forward(self, input, embed_gammas=None):
        if embed_gammas is None:
            embed_gammas = torch.ones((input.shape[0], self.cond_embed_dim), device=input.device)
        y = self.forward_nhwc(ops.to_nhwc(input), embed_gammas)
        return ops.to_nchw(y, self.out_channel)

    def compute_feats(self, input, embed_gammas):
        """UNet.compute_feats (:660-681): (bottleneck, [encoder block outputs], emb), NCHW fp32 like the reference;
        embed_gammas None = the GAN use (all-ones embedding)."""
        if embed_gammas is None:
            embed_gammas = torch.ones((input.shape[0], self.cond_embed_dim), device=input.device)
        hs = []
        h = ops.to_nhwc(self.dwt(input.float()) if getattr(self, "freq_space", False) else input)
        for module in self.input_blocks:
            h = module.forward_nhwc(h, embed_gammas)
            hs.append(h)
        h = self.middle_block.forward_nhwc(h, embed_gammas)
        return ops.to_nchw(h, h.shape[-1]), [ops.to_nchw(f, f.shape[-1]) for f in hs], embed_gammas

    def get_feats(self, input, extract_layer_ids):
        """UNet.get_feats (:697-705): the encoder features CUT's PatchNCE samples from."""
        _, hs, _ = self.compute_feats(input, embed_gammas=None)
        return [feat for i, feat in enumerate(hs) if i in extract_layer_ids]


def set_new_noise_schedule(model, phase):
    """diffusion_utils.set_new_noise_schedule (lines 79-119), linear schedule only: registers the same 7
    buffers per phase on `model` so that state_dict keys match the reference."""
    sched = model.beta_schedule[phase]
    assert sched["schedule"] == "linear"
    betas = np.linspace(sched["linear_start"], sched["linear_end"], sched["n_timestep"], dtype=np.float64)
    alphas = 1.0 - betas
    setattr(model, "num_timesteps_" + phase, int(betas.shape[0]))
    gammas = np.cumprod(alphas, axis=0)
    gammas_prev = np.append(1.0, gammas[:-1])
    param = next(model.parameters(), None)
    device = param.device if param is not None else torch.device("cpu")
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=device)
    model.register_buffer("gammas_" + phase, t(gammas))
    model.register_buffer("gammas_prev_" + phase, t(gammas_prev))
    model.register_buffer("sqrt_recip_gammas_" + phase, t(np.sqrt(1.0 / gammas)))
    model.register_buffer("sqrt_recipm1_gammas_" + phase, t(np.sqrt(1.0 / gammas - 1)))
    posterior_variance = betas * (1.0 - gammas_prev) / (1.0 - gammas)
    model.register_buffer("posterior_log_variance_clipped_" + phase, t(np.log(np.maximum(posterior_variance, 1e-20))))
    model.register_buffer("posterior_mean_coef1_" + phase, t(betas * np.sqrt(gammas_prev) / (1.0 - gammas)))
    model.register_buffer("posterior_mean_coef2_" + phase, t((1.0 - gammas_prev) * np.sqrt(alphas) / (1.0 - gammas)))


def gamma_embedding(gammas, dim, max_period=10000):
    """diffusion_utils.gamma_embedding for [B,1] gammas (host-side O(B) work, stays in torch)."""
    half = dim // 2
    # built on the target device: no host->device copy inside the step (CUDA-graph capturable)
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=gammas.device) / half)
    args = gammas[:, 0:1].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class LabelEmbedder(nn.Module):
    """palette_denoise_fn.LabelEmbedder (:14-31): nn.Embedding(num_classes, hidden, max_norm=1, scale_grad_by_freq)."""

    def __init__(self, num_classes, hidden_size):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes, hidden_size, max_norm=1.0, scale_grad_by_freq=True)
        self.num_classes = num_classes

    def forward(self, labels):
        return ops.label_embed(self.embedding_table.weight, labels)


class PaletteDenoiseFn(nn.Module):
    """palette_denoise_fn.PaletteDenoiseFn with conditioning "" / "class" / "mask" / "class_mask" (the "ref" image
    embedding needs a frozen CLIP / ImageBind backbone: third-party, out of scope).  A UNet whose forward takes three
    arguments (UNetGeneratorRefAttn) receives the dataloader's reference image (palette_denoise_fn.py:40-41, 111-114)."""

    def __init__(self, model, cond_embed_dim, ref_embed_net="", conditioning="", nclasses=2):
        super().__init__()
        if "ref" in conditioning:
            raise NotImplementedError("B200 PaletteDenoiseFn: conditioning %r needs the frozen %s backbone"
                                      % (conditioning, ref_embed_net or "clip"))
        self.model = model
        self.model_nargs = len(inspect.signature(model.forward).parameters)
        self.cond_embed_dim = cond_embed_dim
        self.conditioning = conditioning
        if "class" in conditioning:
            self.netl_embedder_class = LabelEmbedder(nclasses, cond_embed_dim // 2)
            nn.init.normal_(self.netl_embedder_class.embedding_table.weight, std=0.02)
        if "mask" in conditioning:
            self.netl_embedder_mask = LabelEmbedder(nclasses, cond_embed_dim)
            nn.init.normal_(self.netl_embedder_mask.embedding_table.weight, std=0.02)

    def mask_embed_channels(self):
        """extra input channels the mask embedding occupies (diffusion_networks.py:112-113)"""
        return self.cond_embed_dim if "mask" in self.conditioning else 0

    def embedding(self, embed_noise_level, cls):
        """(:96-100) the class embedding shares the embedding vector with the noise level"""
        if "class" in self.conditioning:
            if cls is None:
                raise RuntimeError('PaletteDenoiseFn: conditioning "class" needs the class labels (cls)')
            return torch.cat((embed_noise_level, self.netl_embedder_class(cls)), dim=1)
        return embed_noise_level

    def forward(self, input, embed_noise_level, cls=None, mask=None, ref=None):
        """Drop-in NCHW signature of the reference (:95-115)."""
        emb = self.embedding(embed_noise_level, cls)
        if "mask" in self.conditioning:
            e = self.cond_embed_dim
            c = input.shape[1]
            x = ops.to_nhwc(torch.cat([input, input.new_zeros((input.shape[0], e) + tuple(input.shape[2:]))], dim=1))
            x = ops.embed_rows_into(self.netl_embedder_mask.embedding_table.weight, mask.contiguous(), x, c)
            return ops.to_nchw(self.forward_nhwc(x, emb, self.pack_ref(ref) if self.model_nargs == 3 else None),
                               self.model.out_channel)
        if self.model_nargs == 3:
            return self.model(input, emb, ref)
        return self.model(input, emb)

    def pack_ref(self, ref):
        """ref NCHW fp32 -> what the UNet's NHWC path consumes (None for two-argument UNets)."""
        if self.model_nargs != 3:
            return None
        if ref is None:
            raise RuntimeError("B200 PaletteDenoiseFn: this UNet needs the reference image (ref)")
        return self.model.pack_ref(ref)

    def forward_nhwc(self, x, emb, ref_nhwc=None):
        if self.model_nargs == 3:
            return self.model.forward_nhwc(x, emb, ref_nhwc)
        return self.model.forward_nhwc(x, emb)


class DiffusionGenerator(nn.Module):
    """diffusion_generator.DiffusionGenerator: training forward (lines 457-528) on the B200 kernels."""

    def __init__(self, denoise_fn, sampling_method="ddpm", image_size=256, G_ngf=64,
                 loading_backward_compatibility=False):
        super().__init__()
        if loading_backward_compatibility:
            raise NotImplementedError("B200 DiffusionGenerator: backward-compatibility embedding is not supported")
        self.denoise_fn = denoise_fn
        self.sampling_method = sampling_method
        self.image_size = image_size
        set_new_noise_schedule(self.denoise_fn.model, "train")
        set_new_noise_schedule(self.denoise_fn.model, "test")
        e = self.denoise_fn.cond_embed_dim
        self.cond_embed_dim = e
        # the gamma embedding is half as wide when a class / ref embedding shares the vector (:63-76)
        eg = e // 2 if any(c in getattr(self.denoise_fn, "conditioning", "") for c in ("class", "ref")) else e
        self.cond_embed_gammas = eg
        self.cond_embed_gammas_in = eg
        self.cond_embed = nn.Sequential(nn.Linear(eg, eg), nn.SiLU(), nn.Linear(eg, eg))

    def set_new_sampling_method(self, sampling_method):
        """diffusion_generator.py:523-524 (PaletteModel.inference switches to --alg_palette_sampling_method_test)"""
        self.sampling_method = sampling_method

    def compute_gammas(self, gammas):
        emb = gamma_embedding(gammas, self.cond_embed_gammas_in)
        emb = ops.linear(emb, self.cond_embed[0].weight, self.cond_embed[0].bias)
        return ops.linear(emb, self.cond_embed[2].weight, self.cond_embed[2].bias, act_in=L.ACT_SILU)

    def sample_noise_level(self, b, device, t=None, u=None):
        """t ~ randint(1, T), gamma ~ U(gamma_{t-1}, gamma_t) (lines 467-478).  Index gathers are bit exact."""
        model = self.denoise_fn.model
        if t is None:
            t = torch.randint(1, model.num_timesteps_train, (b,), device=device).long()
        if u is None:
            u = torch.rand((b, 1), device=device)
        gammas = model.gammas_train
        g1 = gammas.gather(-1, t - 1).reshape(b, 1)
        g2 = gammas.gather(-1, t).reshape(b, 1)
        sample_gammas = ((g2 - g1) * u + g1).view(b, -1)
        snr1 = model.sqrt_recip_gammas_train.gather(-1, t)
        snr2 = model.sqrt_recipm1_gammas_train.gather(-1, t)
        snr = torch.pow(snr1 / snr2, 2)
        w = torch.stack([snr, 5.0 * torch.ones_like(t)], dim=1).min(dim=1)[0] / snr
        return t, sample_gammas, w

    def forward_nhwc(self, y_0, y_cond, mask, noise, t=None, u=None, ref=None, cls=None):
        """Returns (noise, noise_hat NHWC bf16 [N,H,W,8], min_snr_w [B]).  Video clips [B,F,C,H,W] (UNetVid,
        diffusion_generator.py:460-463, 497-500) are folded to N = B*F frames: one (t, gamma) draw per clip, the
        per-frame work is identical to the image path; noise is returned in the folded [N,C,H,W] layout."""
        b = y_0.shape[0]
        frames = 0
        # random draws in the reference's order (diffusion_generator.py:467-480): t, then u, then the noise — a run
        # seeded like the reference's sees the same (t, gamma, noise)
        _, sample_gammas, w = self.sample_noise_level(b, y_0.device, t, u)
        if y_0.dim() == 5:
            frames = y_0.shape[1]
            if noise is None:
                # the reference draws on the folded "b c (f h) w" tensor (rearrange_5dto4d_fh, :460-463)
                _, _, c, hh, ww = y_0.shape
                noise = torch.randn((b, c, frames * hh, ww), dtype=y_0.dtype, device=y_0.device)
                noise = noise.reshape(b, c, frames, hh, ww).permute(0, 2, 1, 3, 4)
            fold = lambda v: None if v is None else v.reshape((b * frames,) + tuple(v.shape[2:]))  # noqa: E731
            y_0, y_cond, mask, noise = fold(y_0), fold(y_cond), fold(mask), fold(noise)
            self.denoise_fn.model._clip["frames"] = frames
        if noise is None:
            noise = torch.randn_like(y_0)
        emb = self.compute_gammas(sample_gammas)
        g_per_image = sample_gammas.reshape(b)
        if frames:
            g_per_image = g_per_image.repeat_interleave(frames)
        dn = self.denoise_fn
        e_mask = dn.mask_embed_channels() if hasattr(dn, "mask_embed_channels") else 0
        x = K.noise_pack(y_0.contiguous().float(), y_cond.contiguous().float(), noise.contiguous().float(),
                         None if mask is None else mask.contiguous(), g_per_image.contiguous(),
                         ld=(2 * y_0.shape[1] + e_mask + 7) // 8 * 8)
        if e_mask:  # cat([input, mask_embed]) (palette_denoise_fn.py:104-108): the embedding lands behind the 2C images
            if frames:
                raise NotImplementedError("B200 DiffusionGenerator: mask conditioning with video clips")
            x = ops.embed_rows_into(dn.netl_embedder_mask.embedding_table.weight, mask.contiguous(), x,
                                    2 * y_0.shape[1])
        if hasattr(dn, "embedding"):
            emb = dn.embedding(emb, cls)
        noise_hat = dn.forward_nhwc(x, emb, dn.pack_ref(ref))
        return noise, noise_hat, w

    def forward(self, y_0, y_cond, mask, noise, cls=None, ref=None, dropout_prob=0.0, t=None, u=None):
        # (dropout_prob is accepted and unused, like the reference's: the conditioning dropout happens in
        # PaletteModel.compute_palette_loss, palette_model.py:565-584 -> PaletteTrainer)
        shape5 = tuple(y_0.shape) if y_0.dim() == 5 else None
        c = y_0.shape[2] if shape5 else y_0.shape[1]
        noise, noise_hat, w = self.forward_nhwc(y_0, y_cond, mask, noise, t, u, ref=ref, cls=cls)
        noise_hat = ops.to_nchw(noise_hat, c)
        if shape5:
            noise, noise_hat = noise.reshape(shape5), noise_hat.reshape(shape5)
        return noise, noise_hat, w.view(-1, 1, 1, 1)

    def forward_loss(self, y_0, y_cond, mask, noise=None, lambda_G=1.0, use_minsnr=False, l1=False, t=None, u=None,
                     ref=None, cls=None):
        """compute_palette_loss fused: the UNet output stays NHWC bf16 and feeds the eps-loss kernel directly."""
        if y_0.dim() == 5 and use_minsnr:
            raise NotImplementedError("B200 DiffusionGenerator: min-SNR weighting with video clips")
        b5 = y_0.shape[0] * y_0.shape[1] if y_0.dim() == 5 else None
        noise, noise_hat, w = self.forward_nhwc(y_0, y_cond, mask, noise, t, u, ref=ref, cls=cls)
        if b5 is not None and mask is not None:
            mask = mask.reshape((b5,) + tuple(mask.shape[2:]))
        return ops.palette_loss(noise_hat, noise.contiguous().float(), None if mask is None else mask.contiguous(),
                                w.contiguous() if use_minsnr else None, lambda_G, l1)

    def _sampling_conditioner(self, cls, mask):
        """What the samplers need of PaletteDenoiseFn's conditioning (palette_denoise_fn.py:95-108, called once per
        reverse step by p_mean_variance, diffusion_generator.py:213-246): the class embedding appended to the
        noise-level embedding, and the per-pixel mask embedding behind the (y_cond | y_t) channels of every step's
        UNet input (the step kernel writes the images; the embedding channels are refilled by `fill`)."""
        dn = self.denoise_fn
        conditioning = getattr(dn, "conditioning", "")
        if "class" in conditioning and cls is None:
            raise RuntimeError('restoration: conditioning "class" needs the class labels (cls)')
        if "mask" in conditioning and mask is None:
            raise RuntimeError('restoration: conditioning "mask" needs the mask')

        class _Cond:
            extra = dn.mask_embed_channels() if hasattr(dn, "mask_embed_channels") else 0

            @staticmethod
            def first_input(y_cond, y_t):
                parts = [y_cond, y_t]
                if _Cond.extra:
                    parts.append(y_cond.new_zeros((y_cond.shape[0], _Cond.extra) + tuple(y_cond.shape[2:])))
                _Cond.col0 = y_cond.shape[1] + y_t.shape[1]
                return ops.to_nhwc(torch.cat(parts, dim=1))

            @staticmethod
            def fill(x):
                if _Cond.extra:
                    x = ops.embed_rows_into(dn.netl_embedder_mask.embedding_table.weight, mask.contiguous(), x, _Cond.col0)
                return x

            @staticmethod
            def embedding(emb):
                return dn.embedding(emb, cls) if hasattr(dn, "embedding") else emb

        return _Cond

    @torch.no_grad()
    def restoration_ddpm(self, y_cond, y_t=None, y_0=None, mask=None, sample_num=2, cls=None, guidance_scale=0.0,
                         ref=None, noise_fn=None):
        """diffusion_generator.restoration_ddpm (:122-177), class / mask conditioning included, no guidance (`ref` is
        the reference image of a UNetGeneratorRefAttn denoiser):
        num_timesteps_test UNet forwards; per step ONE fused kernel does predict_start_from_noise, the clamp, the
        posterior mean, the noise injection, the mask blend and the next step's NHWC bf16 input pack.
        noise_fn(i, shape) -> fp32 NCHW noise for step i (default torch.randn on the device; the tests replay the
        reference's CPU draws).  Returns (y_t, ret_arr) like the reference."""
        if guidance_scale:
            # (the reference's guidance branch calls the denoiser with cls=None, mask=None, which its own
            # PaletteDenoiseFn.forward cannot evaluate for class / mask conditioning: palette_denoise_fn.py:95-108)
            raise NotImplementedError("B200 restoration_ddpm: classifier-free guidance")
        model = self.denoise_fn.model
        ref_p = self.denoise_fn.pack_ref(ref)
        cond_in = self._sampling_conditioner(cls, mask)
        T = model.num_timesteps_test
        assert T > sample_num, "num_timesteps must greater than sample_num"
        sample_inter = T // sample_num
        c = model.out_channel
        dev = y_cond.device
        if noise_fn is None:
            noise_fn = lambda i, shape: torch.randn(shape, device=dev)  # noqa: E731
        # clips [B, F, C, H, W] (UNetVid; the reference folds them inside p_mean_variance, :213-218): B*F frames through
        # the UNet and the step kernel, one noise level per clip; random draws keep the reference's 5-D shape
        clip = self._fold_clip(y_cond, cls, mask)
        bq = y_cond.shape[0]                      # rows of the noise-level embedding (clips, or images)
        if y_t is None:
            y_t = noise_fn(T, ((bq, clip, c) if clip else (bq, c)) + tuple(y_cond.shape[-2:]))
        shape_out = tuple(y_t.shape)
        y_cond, y_t, y_0, mask = (self._fold(v) for v in (y_cond, y_t, y_0, mask))
        b = y_cond.shape[0]
        y_cond = y_cond.contiguous().float()
        y_t = y_t.contiguous().float()
        if mask is not None:
            y_0 = y_0.contiguous().float()
            mask = mask.contiguous()
        ld = (y_cond.shape[1] + c + cond_in.extra + 7) // 8 * 8
        sigma = torch.exp(0.5 * model.posterior_log_variance_clipped_test)
        table = torch.stack([model.sqrt_recip_gammas_test, model.sqrt_recipm1_gammas_test,
                             model.posterior_mean_coef1_test, model.posterior_mean_coef2_test, sigma], dim=1)
        # first input: cat([y_cond, y_t]) (every later one comes out of the step kernel)
        x = cond_in.first_input(y_cond, y_t)
        ret_arr = [y_t]
        for i in reversed(range(T)):
            gam = model.gammas_test[i].reshape(1, 1).expand(bq, 1)
            eps = self.denoise_fn.forward_nhwc(cond_in.fill(x), cond_in.embedding(self.compute_gammas(gam)), ref_p)
            noise = self._fold(noise_fn(i, shape_out)).contiguous().float() if i > 0 else None
            coef = table[i].reshape(1, 5).expand(b, 5).contiguous()
            y_t, x = K.ddpm_step(eps, y_t, y_cond, y_0, mask, noise, coef, ld=ld, want_next_input=i > 0)
            if i % sample_inter == 0:
                ret_arr.append(y_t)
        return y_t.reshape(shape_out), torch.cat([r.reshape(shape_out) for r in ret_arr], dim=0)

    @staticmethod
    def _fold(v):
        """[B, F, ...] -> [B*F, ...] for 5-D tensors, anything else unchanged"""
        return v.reshape((v.shape[0] * v.shape[1],) + tuple(v.shape[2:])) if (v is not None and v.dim() == 5) else v

    def _fold_clip(self, y_cond, cls, mask):
        """frames per clip (0 for images); tells the video UNet how many frames its temporal layers see"""
        if y_cond.dim() != 5:
            return 0
        if getattr(self.denoise_fn, "conditioning", ""):
            raise NotImplementedError("B200 samplers: class / mask conditioning with video clips")
        frames = y_cond.shape[1]
        self.denoise_fn.model._clip["frames"] = frames
        return frames

    @torch.no_grad()
    def restoration_ddim(self, y_cond, y_t=None, y_0=None, mask=None, sample_num=8, cls=None, guidance_scale=0.0,
                         num_steps=10, eta=0.5, ref=None):
        """diffusion_generator.restoration_ddim (:286-347) with ddim_p_sample / ddim_p_mean_variance (:349-456):
        num_steps UNet forwards on the linear t sequence; the update is deterministic (the reference draws a noise
        tensor and does not use it), one fused kernel per step."""
        if guidance_scale:
            raise NotImplementedError("B200 restoration_ddim: classifier-free guidance")
        model = self.denoise_fn.model
        ref_p = self.denoise_fn.pack_ref(ref)
        cond_in = self._sampling_conditioner(cls, mask)
        T = model.num_timesteps_test
        assert T > sample_num, "num_timesteps must greater than sample_num"
        sample_inter = T // sample_num
        self._fold_clip(y_cond, cls, mask)
        bq = y_cond.shape[0]
        y_t = torch.randn_like(y_cond) if y_t is None else y_t
        shape_out = tuple(y_t.shape)
        y_cond, y_t, y_0, mask = (self._fold(v) for v in (y_cond, y_t, y_0, mask))
        b = y_cond.shape[0]
        y_cond = y_cond.contiguous().float()
        y_t = y_t.contiguous().float()
        c = y_t.shape[1]
        if mask is not None:
            y_0 = y_0.contiguous().float()
            mask = mask.contiguous()
        ld = (y_cond.shape[1] + c + cond_in.extra + 7) // 8 * 8
        tseq = list(np.linspace(0, T - 1, num_steps).astype(int))
        x = cond_in.first_input(y_cond, y_t)
        ret_arr = [y_t]
        for i in range(num_steps):
            t = int(tseq[-1 - i])
            prevt = int(tseq[-2 - i]) if i != num_steps - 1 else -1
            g_t = model.gammas_test[t]
            g_p = model.gammas_prev_test[prevt + 1]
            sigma = eta * torch.sqrt((1 - g_p) / (1 - g_t) * (1 - g_t / g_p))
            coef_eps = torch.sqrt(torch.clamp(1 - g_p - sigma ** 2, min=0))
            c1 = torch.sqrt(g_p) / torch.sqrt(g_t)
            c2 = coef_eps - torch.sqrt(g_p) * torch.sqrt(1.0 - g_t) / torch.sqrt(g_t)
            coef = torch.stack([c1, c2, c1 * 0, c1 * 0, c1 * 0]).reshape(1, 5).expand(b, 5).contiguous().float()
            eps = self.denoise_fn.forward_nhwc(cond_in.fill(x),
                                               cond_in.embedding(self.compute_gammas(g_t.reshape(1, 1).expand(bq, 1))), ref_p)
            y_t, x = K.ddpm_step(eps, y_t, y_cond, y_0, mask, None, coef, ld=ld, want_next_input=i != num_steps - 1,
                                 ddim=True)
            if i % sample_inter == 0:
                ret_arr.append(y_t)
        return y_t.reshape(shape_out), torch.cat([r.reshape(shape_out) for r in ret_arr], dim=0)

    def restoration(self, y_cond, y_t=None, y_0=None, mask=None, sample_num=8, cls=None, ref=None,
                    guidance_scale=0.0, ddim_num_steps=10, ddim_eta=0.5):
        """diffusion_generator.restoration (:83-118)."""
        if self.sampling_method == "ddpm":
            return self.restoration_ddpm(y_cond, y_t=y_t, y_0=y_0, mask=mask, sample_num=sample_num, cls=cls,
                                         guidance_scale=guidance_scale, ref=ref)
        return self.restoration_ddim(y_cond, y_t=y_t, y_0=y_0, mask=mask, sample_num=sample_num, cls=cls,
                                     guidance_scale=guidance_scale, num_steps=ddim_num_steps, eta=ddim_eta, ref=ref)


def build_palette_generator(image_size=256, in_channel=6, inner_channel=64, out_channel=3, res_blocks=(2, 2, 2, 2),
                            attn_res=(16,), channel_mults=(1, 2, 4, 8), num_heads=1, num_head_channels=32,
                            group_norm_size=32, cond_embed_dim=32, n_timestep_train=2000, n_timestep_test=1000,
                            efficient=False, conditioning="", nclasses=2, use_scale_shift_norm=True):
    """What diffusion_networks.define_G(model_type="palette", G_netG="unet_mha", ...) builds
    (models/diffusion_networks.py:114-139, 361-376), on the B200 modules."""
    if "mask" in conditioning:
        in_channel += cond_embed_dim  # diffusion_networks.py:112-113
    unet = UNet(image_size=image_size, in_channel=in_channel, inner_channel=inner_channel, out_channel=out_channel,
                res_blocks=list(res_blocks), attn_res=list(attn_res), tanh=False, n_timestep_train=n_timestep_train,
                n_timestep_test=n_timestep_test, norm="groupnorm", group_norm_size=group_norm_size,
                cond_embed_dim=cond_embed_dim, channel_mults=tuple(channel_mults), num_heads=num_heads,
                num_head_channels=num_head_channels, efficient=efficient, use_scale_shift_norm=use_scale_shift_norm)
    dn = PaletteDenoiseFn(model=unet, cond_embed_dim=cond_embed_dim, conditioning=conditioning, nclasses=nclasses)
    return DiffusionGenerator(denoise_fn=dn, sampling_method="ddpm", image_size=image_size, G_ngf=inner_channel)
