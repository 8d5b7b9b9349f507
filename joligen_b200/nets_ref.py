"""B200 mirror of the reference-image-conditioned UNet (`UNetGeneratorRefAttn`, `AttentionBlockRef`,
`EmbedSequentialRef`: models/modules/unet_generator_attn/unet_generator_attn.py:1017-1645).

Two UNets share the embedding: the *reference* UNet (`*_ref` blocks) runs on the reference image and hands the
qkv tensor of each of its attention blocks to the main UNet, whose attention blocks attend to their own keys/values
AND to the reference's (`use_ref=True`): h = proj_out(cat[attn(q,k,v), attn(q,k_ref,v_ref)]) + x.
Same constructor arguments and sub-module names as the reference => identical state_dict keys.  No new kernels:
the second attention is the same flash kernel on a qkv tensor whose first third of the channels is swapped in
(`ops.mix_qkv`), both outputs land in the two halves of one buffer, proj_out is a 1x1 conv with residual epilogue.
"""
import torch
import torch.nn as nn

from . import lib as L
from . import ops
from .nets import ConvIn, ConvPack, EmbedBlock, ResBlock, _OutHead, _conv, normalization


class AttentionBlockRef(nn.Module):
    """unet_generator_attn.AttentionBlockRef (:1041-1130)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False, use_transformer=False, use_ref=False, terminal=False):
        super().__init__()
        if use_transformer or use_checkpoint:
            raise NotImplementedError("B200 AttentionBlockRef: use_transformer / use_checkpoint are not supported")
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        self.attention_layout = 1 if use_new_attention_order else 0
        self.use_ref = use_ref
        self.terminal = terminal
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self._pack_qkv = ConvPack(self.qkv)
        if not terminal:
            self.proj_out = nn.Conv1d(channels * 2 if use_ref else channels, channels, 1)
            for p in self.proj_out.parameters():  # zero_module
                p.detach().zero_()
            self._pack_proj = ConvPack(self.proj_out)

    def forward_nhwc(self, x, qkv_ref=None):
        """-> (x + proj_out(...), qkv); a terminal block only produces its qkv (-> (None, qkv))."""
        c, heads = self.channels, self.num_heads
        xn, x = ops.group_norm_tap(x, None, None, c, film=None, act=L.ACT_NONE)  # InstanceNorm1d over T
        qkv = _conv(xn, self.qkv, self._pack_qkv)
        if self.terminal:
            return None, qkv
        if self.use_ref:
            assert qkv_ref is not None
            n, hh, ww, _ = qkv.shape
            buf = torch.empty((n, hh, ww, 2 * c), dtype=torch.bfloat16, device=qkv.device)
            a = ops.attention(qkv, heads, c // heads, self.attention_layout, out=buf[..., :c])
            a_ref = ops.attention(ops.mix_qkv(qkv, qkv_ref), heads, c // heads, self.attention_layout,
                                  out=buf[..., c:])
            a = ops.join_slices(buf, a, a_ref)
        else:
            a = ops.attention(qkv, heads, c // heads, self.attention_layout)
        return _conv(a, self.proj_out, self._pack_proj, residual=x, res_scale=1.0), qkv


class EmbedSequentialRef(nn.Sequential, EmbedBlock):
    """unet_generator_attn.EmbedSequentialRef (:1017-1038): returns (x, [qkv of every attention layer])."""

    def forward_nhwc(self, x, emb, qkv_ref=None):
        qkv = []
        for layer in self:
            if isinstance(layer, AttentionBlockRef):
                cur = qkv_ref if (qkv_ref is None or type(qkv_ref) != list) else qkv_ref.pop(0)
                x, q = layer.forward_nhwc(x, qkv_ref=cur)
                qkv.append(q)
            elif isinstance(layer, EmbedBlock):
                x = layer.forward_nhwc(x, emb)
            else:
                x = layer.forward_nhwc(x)
        return x, qkv


class UNetGeneratorRefAttn(nn.Module):
    """unet_generator_attn.UNetGeneratorRefAttn (:1136-1645): forward(input, embed_gammas, ref)."""

    def __init__(self, image_size, in_channel, inner_channel, out_channel, res_blocks, attn_res, tanh,
                 n_timestep_train, n_timestep_test, norm, group_norm_size, cond_embed_dim, dropout=0,
                 channel_mults=(1, 2, 4, 8), conv_resample=True, use_checkpoint=False, use_fp16=False, num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True,
                 use_new_attention_order=False, efficient=False, freq_space=False):
        super().__init__()
        if tanh or freq_space or not resblock_updown or use_fp16 or use_checkpoint or dropout:
            raise NotImplementedError("B200 UNetGeneratorRefAttn: tanh / freq_space / conv resampling / fp16 / "
                                      "checkpointing / dropout are not supported")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channel = in_channel
        self.inner_channel = inner_channel
        self.out_channel = out_channel
        self.res_blocks = res_blocks
        self.attn_res = attn_res
        self.channel_mults = channel_mults
        self.cond_embed_dim = cond_embed_dim
        if norm == "groupnorm":
            norm = norm + str(group_norm_size)
        rb = dict(use_scale_shift_norm=use_scale_shift_norm, norm=norm, efficient=efficient)
        at = dict(num_head_channels=num_head_channels, use_new_attention_order=use_new_attention_order)

        def encoder(use_ref, ch):
            blocks = nn.ModuleList([EmbedSequentialRef(ConvIn(in_channel, ch, 3, padding=1))])
            chans = [ch]
            ds = 1
            for level, mult in enumerate(channel_mults):
                for _ in range(res_blocks[level]):
                    layers = [ResBlock(ch, cond_embed_dim, 0.0, out_channel=int(mult * inner_channel), **rb)]
                    ch = int(mult * inner_channel)
                    if ds in attn_res:
                        layers.append(AttentionBlockRef(ch, num_heads=num_heads, use_ref=use_ref, **at))
                    blocks.append(EmbedSequentialRef(*layers))
                    chans.append(ch)
                if level != len(channel_mults) - 1:
                    blocks.append(EmbedSequentialRef(ResBlock(ch, cond_embed_dim, 0.0, out_channel=ch, down=True,
                                                              **rb)))
                    chans.append(ch)
                    ds *= 2
            middle = EmbedSequentialRef(ResBlock(ch, cond_embed_dim, 0.0, **rb),
                                        AttentionBlockRef(ch, num_heads=num_heads, use_ref=use_ref, **at),
                                        ResBlock(ch, cond_embed_dim, 0.0, **rb))
            return blocks, middle, chans, ch, ds

        input_ch = int(channel_mults[0] * inner_channel)
        self.input_blocks, self.middle_block, input_block_chans, ch, ds = encoder(True, input_ch)
        # the reference builds the second encoder with `ch` carried over from the first one (:1330-1332): its first
        # conv has as many output channels as the bottleneck
        self.input_blocks_ref, self.middle_block_ref, _, _, _ = encoder(False, ch)

        # reference decoder: stops at the last block that still feeds an attention layer of the main decoder
        ch_ref, ds_ref = ch, ds
        chans_ref = list(input_block_chans)
        self.output_blocks_ref = nn.ModuleList([])
        is_terminal = False
        for level, mult in list(enumerate(channel_mults))[::-1]:
            for i in range(res_blocks[level] + 1):
                is_terminal = i == res_blocks[level] and ds_ref / 2 not in attn_res
                ich = chans_ref.pop()
                layers = [ResBlock(ch_ref + ich, cond_embed_dim, 0.0, out_channel=int(inner_channel * mult), **rb)]
                ch_ref = int(inner_channel * mult)
                if ds_ref in attn_res:
                    layers.append(AttentionBlockRef(ch_ref, num_heads=num_heads_upsample, terminal=is_terminal, **at))
                if level and i == res_blocks[level]:
                    if not is_terminal:
                        layers.append(ResBlock(ch_ref, cond_embed_dim, 0.0, out_channel=ch_ref, up=True, **rb))
                    ds_ref //= 2
                self.output_blocks_ref.append(EmbedSequentialRef(*layers))
            if is_terminal:
                break

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mults))[::-1]:
            for i in range(res_blocks[level] + 1):
                ich = input_block_chans.pop()
                layers = [ResBlock(ch + ich, cond_embed_dim, 0.0, out_channel=int(inner_channel * mult), **rb)]
                ch = int(inner_channel * mult)
                if ds in attn_res:
                    layers.append(AttentionBlockRef(ch, num_heads=num_heads_upsample, use_ref=True, **at))
                if level and i == res_blocks[level]:
                    layers.append(ResBlock(ch, cond_embed_dim, 0.0, out_channel=ch, up=True, **rb))
                    ds //= 2
                self.output_blocks.append(EmbedSequentialRef(*layers))
        self.out = _OutHead(normalization(ch, norm), nn.SiLU(), nn.Conv2d(input_ch, out_channel, 3, padding=1))
        self._pack_outconv = ConvPack(self.out[2])
        self.beta_schedule = {
            "train": {"schedule": "linear", "n_timestep": n_timestep_train, "linear_start": 1e-6,
                      "linear_end": 0.01},
            "test": {"schedule": "linear", "n_timestep": n_timestep_test, "linear_start": 1e-4, "linear_end": 0.09},
        }

    def forward_nhwc(self, x, emb, ref):
        """x, ref: NHWC bf16 (ref already channel-doubled like the reference's cat([ref, ref], dim=1))."""
        qkv_list, hs_ref = [], []
        h = ref
        for module in self.input_blocks_ref:
            h, q = module.forward_nhwc(h, emb, qkv_ref=None)
            qkv_list.append(q)
            hs_ref.append(h)
        h_ref, q = self.middle_block_ref.forward_nhwc(h, emb, qkv_ref=None)
        qkv_list.append(q)
        hs = []
        h = x
        for module in self.input_blocks:
            h, _ = module.forward_nhwc(h, emb, qkv_ref=qkv_list.pop(0))
            hs.append(h)
        h, _ = self.middle_block.forward_nhwc(h, emb, qkv_ref=qkv_list.pop(0))
        qkv_list = []
        for module in self.output_blocks_ref:
            h_ref = ops.cat_channels(h_ref, hs_ref.pop())
            h_ref, q = module.forward_nhwc(h_ref, emb, qkv_ref=None)
            qkv_list.append(q)
        for module in self.output_blocks:
            h = ops.cat_channels(h, hs.pop())
            h, _ = module.forward_nhwc(h, emb, qkv_ref=qkv_list.pop(0) if qkv_list else None)
        h = self.out[0].forward_nhwc(h, act=L.ACT_SILU)
        return _conv(h, self.out[2], self._pack_outconv)

    @staticmethod
    def pack_ref(ref):
        """ref NCHW fp32 [N,3,H,W] -> NHWC bf16 of cat([ref, ref], dim=1) (:1577)."""
        return ops.to_nhwc(torch.cat([ref, ref], dim=1))

    def forward(self, input, embed_gammas=None, ref=None):
        if ref is None:
            raise NotImplementedError("B200 UNetGeneratorRefAttn: the reference image is required")
        if embed_gammas is None:
            embed_gammas = torch.ones((input.shape[0], self.cond_embed_dim), device=input.device)
        y = self.forward_nhwc(ops.to_nhwc(input), embed_gammas, self.pack_ref(ref))
        return ops.to_nchw(y, self.out_channel)
