"""B200-backed mirrors of the reference's GAN generator / discriminator (CUT, config 3):

    models/modules/resnet_architecture/resnet_generator.py   ResnetBlock :11-95, ResnetGenerator :98-164,
                                                              ResnetEncoder :167-271, ResnetDecoder :274-347
    models/modules/discriminators.py                          NLayerDiscriminator :10-117
    models/modules/loss.py                                    GANLoss :11-85

The modules hold the SAME nn.Sequential layouts as the reference (nn.ReflectionPad2d, nn.Conv2d,
nn.InstanceNorm2d(affine=False), nn.ReLU, nn.ConvTranspose2d, nn.Tanh, nn.LeakyReLU ...), hence the same
state_dict keys (`encoder.model.1.weight`, `encoder.model.10.conv_block.5.bias`, `decoder.model.0.weight`,
`model.11.weight`, ...); execution walks the sequence and dispatches fused NHWC bf16 kernels:

    ReflectionPad2d          -> jg_pad2d_fwd / _bwd
    Conv2d [+ LeakyReLU|Tanh] -> implicit-GEMM conv with the activation in the epilogue (stride 1 or 2)
    InstanceNorm2d + ReLU|LeakyReLU -> jg_groupnorm_fwd / _bwd with groups == C, no affine, fused activation
    ConvTranspose2d(3, s2, p1, op1) -> zero insertion + stride-1 implicit GEMM (dgrad form)
    ResnetBlock              -> x + conv_block(x)
"""
import functools

import torch.nn as nn

from . import kernels as K
from . import lib as L
from . import ops
from .nets import ConvPack


def _act_code(m):
    if isinstance(m, nn.ReLU):
        return L.ACT_RELU
    if isinstance(m, nn.LeakyReLU):
        if abs(m.negative_slope - 0.2) > 1e-12:
            raise NotImplementedError("LeakyReLU slope %g (0.2 is implemented)" % m.negative_slope)
        return L.ACT_LRELU02
    if isinstance(m, nn.Tanh):
        return L.ACT_TANH
    return None


class _SeqRunner:
    """Executes an nn.Sequential of the GAN building blocks on NHWC bf16 tensors."""

    def __init__(self):
        self.packs = {}

    def pack(self, conv):
        p = self.packs.get(id(conv))
        if p is None:
            p = self.packs[id(conv)] = ConvPack(conv)
        return p.get()

    def run(self, seq, x, collect=None, base_id=0):
        mods = list(seq)
        i = 0
        feats = []
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            consumed = 1
            if isinstance(m, nn.ReflectionPad2d):
                x = ops.reflection_pad(x, m.padding[0])
            elif isinstance(m, nn.ReplicationPad2d):
                x = ops.replication_pad(x, m.padding[0])
            elif isinstance(m, nn.Conv2d):
                act = _act_code(nxt) if isinstance(nxt, (nn.LeakyReLU, nn.Tanh)) else None
                if hasattr(m, "weight_orig"):
                    # torch's spectral_norm wrapper (--D_spectral): its pre-forward hook never runs here; one power
                    # iteration + the normalised weight are parameter preprocessing (nets_projd._sn_weight)
                    from .nets_projd import _sn_conv
                    x = _sn_conv(x, m, act=act if act is not None else L.ACT_NONE)
                else:
                    x = ops.conv_act(x, m.weight, m.bias, self.pack(m), stride=m.stride[0], pad=m.padding[0],
                                     act=act if act is not None else L.ACT_NONE)
                if act is not None:
                    consumed = 2
            elif isinstance(m, nn.ConvTranspose2d):
                if m.kernel_size != (3, 3) or m.stride != (2, 2) or m.padding != (1, 1) or m.output_padding != (1, 1):
                    raise NotImplementedError("ConvTranspose2d other than k3 s2 p1 op1")
                if hasattr(m, "weight_orig"):   # --G_spectral
                    from .nets_projd import _sn_conv_transpose
                    x = _sn_conv_transpose(x, m)
                else:
                    x = ops.conv_transpose2d(x, m.weight, m.bias, self.pack(m), pad=1)
            elif isinstance(m, nn.InstanceNorm2d):
                if m.affine or m.track_running_stats:
                    raise NotImplementedError("InstanceNorm2d with affine / running stats")
                act = _act_code(nxt) if isinstance(nxt, (nn.ReLU, nn.LeakyReLU)) else None
                x = ops.group_norm(x, None, None, x.shape[-1], film=None, act=act if act is not None else L.ACT_NONE)
                if act is not None:
                    consumed = 2
            elif isinstance(m, ResnetBlock):
                x = ops.add(x, self.run(m.conv_block, x))
            elif isinstance(m, (nn.Identity, nn.Dropout)):
                if isinstance(m, nn.Dropout) and m.p > 0 and m.training:
                    raise NotImplementedError("Dropout in the B200 GAN path")
            else:
                raise NotImplementedError("B200 GAN path: unsupported layer %s" % type(m).__name__)
            if collect is not None:
                # a fused (layer, in-place activation) pair: the reference's compute_feats records the layer's
                # output tensor, which the in-place activation then overwrites -> both ids see the activated map
                for k in range(consumed):
                    if base_id + i + k in collect:
                        if consumed == 2 and k == 0 and not getattr(nxt, "inplace", False):
                            raise NotImplementedError("feature requested before a non-in-place activation")
                        feats.append((base_id + i + k, x))
            i += consumed
        return (x, feats) if collect is not None else x


def get_norm_layer(norm_type="instance"):
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    raise NotImplementedError("B200 GAN path: norm %r (instance is implemented)" % norm_type)


def _pad_layers(padding_type, n):
    """--G_padding_type (resnet_generator.py:42-50, 328-336) -> ([pad layer], conv padding)"""
    if padding_type == "reflect":
        return [nn.ReflectionPad2d(n)], 0
    if padding_type == "replicate":
        return [nn.ReplicationPad2d(n)], 0
    if padding_type == "zeros":
        return [], n
    raise NotImplementedError("padding [%s] is not implemented" % padding_type)


def _sn(use_spectral):
    """models/modules/utils.spectral_norm(module, mode) (:201-204)"""
    return nn.utils.spectral_norm if use_spectral else (lambda m: m)


def _uses_bias(norm_layer):
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return f == nn.InstanceNorm2d


class ResnetBlock(nn.Module):
    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias, use_spectral=False, conv=nn.Conv2d):
        super().__init__()
        if use_dropout:
            raise NotImplementedError("B200 ResnetBlock: dropout")
        sn = _sn(use_spectral)
        pad, p = _pad_layers(padding_type, 1)
        pad2, _ = _pad_layers(padding_type, 1)
        self.conv_block = nn.Sequential(
            *pad, sn(nn.Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias)), norm_layer(dim), nn.ReLU(True),
            *pad2, sn(nn.Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias)), norm_layer(dim))


class ResnetEncoder(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=None, use_dropout=False, n_blocks=6,
                 padding_type="reflect", use_spectral=False, conv=nn.Conv2d):
        super().__init__()
        norm_layer = norm_layer or get_norm_layer("instance")
        use_bias = _uses_bias(norm_layer)
        sn = _sn(use_spectral)
        model = [nn.ReflectionPad2d(3), sn(nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias)),
                 norm_layer(ngf), nn.ReLU(True)]
        for i in range(2):
            mult = 2 ** i
            model += [sn(nn.Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1, bias=use_bias)),
                      norm_layer(ngf * mult * 2), nn.ReLU(True)]
        for _ in range(n_blocks):
            # (the reference's encoder does NOT hand use_spectral to its ResnetBlocks, resnet_generator.py:237-246: with
            # --G_spectral the stem, the down- and the up-sampling layers are wrapped, the blocks are not)
            model += [ResnetBlock(ngf * 4, padding_type, norm_layer, use_dropout, use_bias)]
        self.model = nn.Sequential(*model)


class ResnetDecoder(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=None, use_dropout=False, n_blocks=6,
                 padding_type="reflect", use_spectral=False):
        super().__init__()
        norm_layer = norm_layer or get_norm_layer("instance")
        use_bias = _uses_bias(norm_layer)
        model = []
        for i in range(2):
            mult = 2 ** (2 - i)
            model += [_sn(use_spectral)(nn.ConvTranspose2d(ngf * mult, ngf * mult // 2, kernel_size=3, stride=2,
                                                            padding=1, output_padding=1, bias=use_bias)),
                      norm_layer(ngf * mult // 2), nn.ReLU(True)]
        pad, p = _pad_layers(padding_type, 3)
        model += pad + [nn.Conv2d(ngf, output_nc, kernel_size=7, padding=p), nn.Tanh()]
        self.model = nn.Sequential(*model)


class ResnetGenerator(nn.Module):
    """resnet_generator.ResnetGenerator(input_nc, output_nc, ngf, norm_layer, use_dropout, n_blocks, padding_type)."""

    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=None, use_dropout=False, n_blocks=6,
                 padding_type="reflect", use_spectral=False, mobile=False):
        super().__init__()
        if mobile:
            raise NotImplementedError("B200 ResnetGenerator: the mobile (separable convolution) variant")
        self.output_nc = output_nc
        self.encoder = ResnetEncoder(input_nc, output_nc, ngf, norm_layer, use_dropout, n_blocks, padding_type,
                                     use_spectral)
        self.decoder = ResnetDecoder(input_nc, output_nc, ngf, norm_layer, use_dropout, n_blocks, padding_type,
                                     use_spectral)
        self._runner = _SeqRunner()

    def forward_nhwc(self, x):
        h = self._runner.run(self.encoder.model, x)
        return self._runner.run(self.decoder.model, h)

    def forward(self, input):
        y = self.forward_nhwc(ops.to_nhwc(input))
        return ops.to_nchw(y, self.output_nc)

    def get_feats(self, input, extract_layer_ids=[]):
        """ResnetEncoder.compute_feats (:238-256): activations after the listed encoder layers, NCHW fp32."""
        _, feats = self._runner.run(self.encoder.model, ops.to_nhwc(input), collect=set(extract_layer_ids))
        out = []
        for lid, f in feats:
            m = self.encoder.model[lid]
            c = f.shape[-1]
            if isinstance(m, nn.ReflectionPad2d) and lid == 0:
                c = input.shape[1]
            out.append(ops.to_nchw(f, c))
        return out


class NLayerDiscriminator(nn.Module):
    """discriminators.NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer): PatchGAN."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_dropout=False, use_spectral=False,
                 freq_space=False):
        super().__init__()
        if use_dropout:
            raise NotImplementedError("B200 NLayerDiscriminator: dropout")
        sn = nn.utils.spectral_norm if use_spectral else (lambda conv: conv)   # (discriminators.py:51-108)
        self.freq_space = freq_space
        self.input_nc = input_nc
        if freq_space:   # (discriminators.py:40-46) the PatchGAN sees the four Haar bands of the image
            from .nets import _Haar
            self.iwt = _Haar(True)
            self.dwt = _Haar(False)
            input_nc *= 4
        norm_layer = norm_layer or get_norm_layer("instance")
        use_bias = _uses_bias(norm_layer)
        kw, padw = 4, 1
        seq = [sn(nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw)), nn.LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_prev, nf_mult = nf_mult, min(2 ** n, 8)
            seq += [sn(nn.Conv2d(ndf * nf_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias)),
                    norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        nf_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        seq += [sn(nn.Conv2d(ndf * nf_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias)),
                norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        seq += [sn(nn.Conv2d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw))]
        self.model = nn.Sequential(*seq)
        self._runner = _SeqRunner()

    def forward_nhwc(self, x):
        """-> logits NHWC bf16 [N, h, w, 8] (channel 0 is the prediction, 1..7 are zero padding)."""
        if self.freq_space:
            x = ops.to_nhwc(self.dwt(ops.to_nchw(x, self.input_nc)))
        return self._runner.run(self.model, x)

    def forward(self, input):
        if self.freq_space:
            return ops.to_nchw(self._runner.run(self.model, ops.to_nhwc(self.dwt(input.float()))), 1)
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(input)), 1)


class GANLoss(nn.Module):
    """loss.GANLoss for lsgan / wgangp / projected, evaluated by one fused kernel on the NHWC logits."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        if gan_mode not in ("lsgan", "wgangp", "projected"):
            raise NotImplementedError("gan mode %s not implemented on the B200 path" % gan_mode)
        self.gan_mode = gan_mode
        self.real_label, self.fake_label = target_real_label, target_fake_label

    def forward_nhwc(self, pred, target_is_real, relu=True, c_real=1):
        if self.gan_mode == "lsgan":
            return ops.gan_loss(pred, c_real, K.GAN_LSGAN, self.real_label if target_is_real else self.fake_label, 1.0)
        sign = 1.0 if target_is_real else -1.0
        if self.gan_mode == "projected" and relu:
            return ops.gan_loss(pred, c_real, K.GAN_HINGE, 0.0, sign)
        if self.gan_mode == "projected":
            return ops.gan_loss(pred, c_real, K.GAN_LINEAR, 0.0, 1.0)  # (-prediction).mean()
        return ops.gan_loss(pred, c_real, K.GAN_LINEAR, 0.0, sign)

    def forward(self, prediction, target_is_real, relu=True):
        """prediction: NCHW fp32 [N, 1, h, w] (the reference signature)."""
        return self.forward_nhwc(ops.to_nhwc(prediction), target_is_real, relu, c_real=prediction.shape[1])
