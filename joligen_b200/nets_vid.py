"""B200 mirror of the reference's video UNet (models/modules/unet_generator_attn/unet_generator_attn_vid.py):
`UNetVid` with a `MotionModule` (temporal transformer over the F frames of every pixel) after every ResBlock level.

Same constructor arguments, same sub-module names => identical `state_dict` keys as the reference classes, so
checkpoints and `load_state_dict` interchange.  Internally a clip `[B, F, C, H, W]` is ONE NHWC bf16 tensor
`[B*F, H, W, C]` (frame index = n % F): the spatial blocks are the image UNet's blocks on B*F images
(ResBlock / AttentionBlock / InflatedConv3d fold frames into the batch in the reference too, :239-240, :322-328),
and the MotionModule never rearranges anything — its tokens ARE the NHWC pixels, its Linear layers are 1x1
convolutions on that tensor, and the temporal attention kernel strides over frames.
"""
import math

import torch
import torch.nn as nn

from . import lib as L
from . import ops
from .nets import AttentionBlock, ConvIn, ConvPack, EmbedSequential, ResBlock, UNet, _OutHead, normalization


class VidResBlock(ResBlock):
    """unet_generator_attn_vid.ResBlock (:148-278): the image block on B*F frames; `emb` is per clip and repeated per
    frame; `skipw` is computed but NOT applied by the reference (:272-275)."""

    apply_skipw = False


def _linear_conv(x, lin, pack, residual=None, out=None):
    """nn.Linear on NHWC tokens = a 1x1 convolution ([O, I] weight viewed as [O, I, 1, 1])."""
    w = lin.weight.unsqueeze(-1).unsqueeze(-1)
    return ops.conv2d(x, w, lin.bias, pack.get(), stride=1, pad=0, residual=residual, res_scale=1.0,
                      grad_sink=lin.weight, out=out)


class PositionalEncoding(nn.Module):
    """unet_generator_attn_vid.PositionalEncoding (:932-947): sinusoidal table over the frame index."""

    def __init__(self, d_model, dropout=0.0, max_len=25):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class VersatileAttention(nn.Module):
    """Temporal self-attention (:950-1054 on CrossAttention :593-660): to_q / to_k / to_v (no bias), 8 heads,
    softmax over the F frames of each pixel, to_out[0] Linear(+bias)."""

    def __init__(self, query_dim, heads, dim_head, temporal_position_encoding=True,
                 temporal_position_encoding_max_len=25):
        super().__init__()
        inner = heads * dim_head
        if inner != query_dim:
            raise NotImplementedError("B200 VersatileAttention: inner_dim must equal query_dim")
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(query_dim, inner, bias=False)
        self.to_v = nn.Linear(query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.pos_encoder = (PositionalEncoding(query_dim, dropout=0.0, max_len=temporal_position_encoding_max_len)
                            if temporal_position_encoding else None)
        self.is_cross_attention = False
        self._packs = [ConvPack(m) for m in (self.to_q, self.to_k, self.to_v, self.to_out[0])]

    def forward_nhwc(self, normed, residual, frames):
        """normed = LayerNorm(h) + PE (the caller fuses both); returns to_out(attn(normed)) + residual."""
        n, hh, ww, c = normed.shape
        qkv = torch.empty((n, hh, ww, 3 * c), dtype=torch.bfloat16, device=normed.device)
        parts = [_linear_conv(normed, lin, pack, out=qkv[..., i * c:(i + 1) * c])
                 for i, (lin, pack) in enumerate(zip((self.to_q, self.to_k, self.to_v), self._packs[:3]))]
        a = ops.temporal_attention(ops.join_slices(qkv, *parts), frames, self.heads)
        return _linear_conv(a, self.to_out[0], self._packs[3], residual=residual)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """FeedForward (:862-905) with GEGLU: net = [GEGLU(dim, 4*dim), Dropout, Linear(4*dim, dim)]."""

    def __init__(self, dim, mult=4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])
        self._pack_in = ConvPack(self.net[0].proj)
        self._pack_out = ConvPack(self.net[2])

    def forward_nhwc(self, normed, residual):
        g = _linear_conv(normed, self.net[0].proj, self._pack_in)
        return _linear_conv(ops.geglu(g), self.net[2], self._pack_out, residual=residual)


class TemporalTransformerBlock(nn.Module):
    """(:516-590) 2 x [LayerNorm -> temporal self-attention + residual], LayerNorm -> GEGLU feed-forward + residual."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types,
                 temporal_position_encoding, temporal_position_encoding_max_len):
        super().__init__()
        for name in attention_block_types:
            if name.split("_")[0] != "Temporal" or name.endswith("_Cross"):
                raise NotImplementedError("B200 TemporalTransformerBlock: only temporal self-attention blocks")
        self.attention_blocks = nn.ModuleList([
            VersatileAttention(dim, num_attention_heads, attention_head_dim, temporal_position_encoding,
                               temporal_position_encoding_max_len) for _ in attention_block_types])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in attention_block_types])
        self.ff = FeedForward(dim)
        self.ff_norm = nn.LayerNorm(dim)

    def forward_nhwc(self, h, frames):
        for attn, norm in zip(self.attention_blocks, self.norms):
            pe = attn.pos_encoder.pe[0, :frames].contiguous() if attn.pos_encoder is not None else None
            # h feeds the norm and the residual add: the tap sums both gradients inside the LayerNorm backward
            normed, h = ops.layer_norm_tap(h, norm.weight, norm.bias, pe=pe, frames=frames, eps=norm.eps)
            h = attn.forward_nhwc(normed, h, frames)
        normed, h = ops.layer_norm_tap(h, self.ff_norm.weight, self.ff_norm.bias, eps=self.ff_norm.eps)
        return self.ff.forward_nhwc(normed, h)


class TemporalTransformer3DModel(nn.Module):
    """(:425-513) GroupNorm(32, C, eps 1e-6) -> proj_in -> transformer blocks -> proj_out, + residual."""

    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers, attention_block_types,
                 temporal_position_encoding, temporal_position_encoding_max_len, norm_num_groups=32):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            TemporalTransformerBlock(inner, num_attention_heads, attention_head_dim, attention_block_types,
                                     temporal_position_encoding, temporal_position_encoding_max_len)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels)
        self._pack_in = ConvPack(self.proj_in)
        self._pack_out = ConvPack(self.proj_out)

    def forward_nhwc(self, x, frames, out=None):
        # x feeds the norm and the final residual: the tap sums both gradients inside the GroupNorm backward
        h, x = ops.group_norm_tap(x, self.norm.weight, self.norm.bias, self.norm.num_groups, act=L.ACT_NONE,
                                  eps=self.norm.eps)
        h = _linear_conv(h, self.proj_in, self._pack_in)
        for blk in self.transformer_blocks:
            h = blk.forward_nhwc(h, frames)
        return _linear_conv(h, self.proj_out, self._pack_out, residual=x, out=out)


class MotionModule(nn.Module):
    """unet_generator_attn_vid.MotionModule (:374-422).  `clip` is the shared {"frames": F} set by UNetVid.forward."""

    def __init__(self, in_channels, num_attention_heads=8, num_transformer_block=2, cross_attention_dim=768,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), cross_frame_attention_mode=None,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=25,
                 temporal_attention_dim_div=1, zero_initialize=True, clip=None):
        super().__init__()
        if cross_frame_attention_mode is not None:
            raise NotImplementedError("B200 MotionModule: cross_frame_attention_mode is not supported")
        self.temporal_transformer = TemporalTransformer3DModel(
            in_channels=in_channels, num_attention_heads=num_attention_heads,
            attention_head_dim=in_channels // num_attention_heads // temporal_attention_dim_div,
            num_layers=num_transformer_block, attention_block_types=attention_block_types,
            temporal_position_encoding=temporal_position_encoding,
            temporal_position_encoding_max_len=temporal_position_encoding_max_len)
        if zero_initialize:
            for p in self.temporal_transformer.proj_out.parameters():
                p.detach().zero_()
        self._clip = clip if clip is not None else {"frames": 1}

    def forward_nhwc(self, x, out=None):
        return self.temporal_transformer.forward_nhwc(x, self._clip["frames"], out=out)


class UNetVid(UNet):
    """unet_generator_attn_vid.UNetVid (:1057-1407): forward(input [B, F, C, H, W], embed_gammas [B, E])."""

    def __init__(self, image_size, in_channel, inner_channel, out_channel, res_blocks, attn_res, tanh,
                 n_timestep_train, n_timestep_test, norm, group_norm_size, cond_embed_dim, dropout=0,
                 channel_mults=(1, 2, 4, 8), conv_resample=True, use_checkpoint=False, use_fp16=False, num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True,
                 use_new_attention_order=True, efficient=False, freq_space=False, max_sequence_length=25,
                 cross_attention_dim=768, num_attention_heads=8, num_transformer_blocks=2):
        nn.Module.__init__(self)
        if tanh or freq_space or not resblock_updown or use_fp16 or use_checkpoint:
            raise NotImplementedError("B200 UNetVid: tanh / freq_space / conv resampling / fp16 / checkpointing")
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
        self.max_sequence_length = max_sequence_length
        self._clip = {"frames": 1}
        if norm == "groupnorm":
            norm = norm + str(group_norm_size)
        rb = dict(use_scale_shift_norm=use_scale_shift_norm, norm=norm, efficient=efficient)
        at = dict(num_head_channels=num_head_channels, use_new_attention_order=use_new_attention_order)

        def motion(ch):
            return MotionModule(in_channels=ch, num_attention_heads=num_attention_heads,
                                num_transformer_block=num_transformer_blocks, cross_attention_dim=cross_attention_dim,
                                attention_block_types=("Temporal_self", "Temporal_Self"),
                                temporal_position_encoding=True,
                                temporal_position_encoding_max_len=max_sequence_length, clip=self._clip)

        ch = input_ch = int(channel_mults[0] * inner_channel)
        self.input_blocks = nn.ModuleList([EmbedSequential(ConvIn(in_channel, ch, 3, padding=1))])
        input_block_chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mults):
            for _ in range(res_blocks[level]):
                layers = [VidResBlock(ch, cond_embed_dim, 0.0, out_channel=int(mult * inner_channel), **rb)]
                ch = int(mult * inner_channel)
                if ds in attn_res:
                    layers.append(AttentionBlock(ch, num_heads=num_heads, **at))
                layers.append(motion(ch))
                self.input_blocks.append(EmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mults) - 1:
                self.input_blocks.append(
                    EmbedSequential(VidResBlock(ch, cond_embed_dim, 0.0, out_channel=ch, down=True, **rb)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = EmbedSequential(VidResBlock(ch, cond_embed_dim, dropout, **rb),
                                            AttentionBlock(ch, num_heads=num_heads, **at),
                                            VidResBlock(ch, cond_embed_dim, dropout, **rb))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mults))[::-1]:
            for i in range(res_blocks[level] + 1):
                ich = input_block_chans.pop()
                layers = [VidResBlock(ch + ich, cond_embed_dim, 0.0, out_channel=int(inner_channel * mult), **rb)]
                ch = int(inner_channel * mult)
                if ds in attn_res:
                    layers.append(AttentionBlock(ch, num_heads=num_heads_upsample, **at))
                layers.append(motion(ch))
                if level and i == res_blocks[level]:
                    layers.append(VidResBlock(ch, cond_embed_dim, 0.0, out_channel=ch, up=True, **rb))
                    ds //= 2
                self.output_blocks.append(EmbedSequential(*layers))
        self.out = _OutHead(normalization(ch, norm), nn.SiLU(), nn.Conv2d(input_ch, out_channel, 3, padding=1))
        self._pack_outconv = ConvPack(self.out[2])
        self.beta_schedule = {
            "train": {"schedule": "linear", "n_timestep": n_timestep_train, "linear_start": 1e-6,
                      "linear_end": 0.01},
            "test": {"schedule": "linear", "n_timestep": n_timestep_test, "linear_start": 1e-4, "linear_end": 0.09},
        }

    def forward(self, input, embed_gammas=None):
        b, f, c, hh, ww = input.shape
        if embed_gammas is None:
            embed_gammas = torch.ones((b, self.cond_embed_dim), device=input.device)
        self._clip["frames"] = f
        y = self.forward_nhwc(ops.to_nhwc(input.reshape(b * f, c, hh, ww)), embed_gammas)
        return ops.to_nchw(y, self.out_channel).reshape(b, f, -1, hh, ww)
