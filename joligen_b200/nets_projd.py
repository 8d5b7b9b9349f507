"""The projected discriminator's trainable part on the B200 kernels: mirrors of

    SingleDisc, MultiScaleD   /root/reference/models/modules/projected_d/discriminator.py:13-77, 166-230
    DownBlock, conv2d (= spectral_norm(nn.Conv2d)), NormLayer   /root/reference/models/modules/projected_d/blocks.py

with the reference's constructor arguments and state_dict keys (`...weight_orig`, `weight_u`, `weight_v`: the layers ARE
torch's spectral_norm-wrapped nn.Conv2d, so checkpoints load unchanged).  Every layer is an existing kernel of the GAN
path: 4x4 stride-2 implicit GEMM (the NLayerDiscriminator's), GroupNorm(c/2 groups) with the LeakyReLU(0.2) fused,
4x4 valid conv with one output channel.  The spectral normalisation itself — one power iteration on a [Cout, Cin*16]
matrix per step — is parameter preprocessing (two matrix-vector products per layer), done with torch ops on the device.
The frozen timm feature network in front of MultiScaleD is third-party and stays the reference's.

Parity on the B200 (tests/test_gpu_widen_cut.py, no markers left): logits, hinge loss, parameter gradients and the
power-iteration state vs the reference's golden vectors, feature gradients at the measured bf16 floor.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from . import kernels as K
from . import lib as L
from . import ops

CHANNEL_DICT = {4: 512, 8: 512, 16: 256, 32: 128, 64: 64, 128: 64, 256: 32, 512: 16, 1024: 8}


def sn_conv2d(*args, **kwargs):
    return spectral_norm(nn.Conv2d(*args, **kwargs))


def _sn_weight(conv, eps=1e-12, dim=0):
    """What torch's SpectralNorm.compute_weight does in its forward pre-hook (which never runs here, because the conv's
    own forward is not called): in training mode one power iteration updates the u / v buffers in place, then
    W = W_orig / (u^T W_mat v) with u, v constants for autograd.  dim: the weight dimension that forms the matrix rows
    (0 for nn.Conv2d, 1 for nn.ConvTranspose2d — torch.nn.utils.spectral_norm's own default per module type)."""
    w = conv.weight_orig
    u, v = conv.weight_u, conv.weight_v
    wm = w if dim == 0 else w.permute(dim, *[d for d in range(w.dim()) if d != dim])
    wm = wm.reshape(wm.shape[0], -1)
    if conv.training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
    sigma = torch.dot(u.detach().clone(), torch.mv(wm, v.detach().clone()))
    return w / sigma


def _sn_conv(x, conv, act=L.ACT_NONE):
    """x NHWC bf16 -> conv with the spectrally normalised weight; the bf16 packed copies are rebuilt from it (the
    normalised weight changes with every power iteration)."""
    w = _sn_weight(conv)
    wf, wd = K.pack_conv_weight(w.detach())
    cout8 = (w.shape[0] + 7) // 8 * 8
    bias_p = None
    if conv.bias is not None:
        bias_p = torch.zeros(cout8, dtype=torch.float32, device=w.device)
        bias_p[: w.shape[0]] = conv.bias.detach()
    return ops.conv_act(x, w, conv.bias, (wf, wd, bias_p), stride=conv.stride[0], pad=conv.padding[0], act=act)


def _sn_conv_transpose(x, convt):
    """spectral_norm(nn.ConvTranspose2d(k=3, s=2, p=1, op=1)) (--G_spectral, resnet_generator.py:306-318)"""
    w = _sn_weight(convt, dim=1)                      # [Cin, Cout, 3, 3]
    wf, wd = K.pack_conv_weight(w.detach())
    cout = w.shape[1]
    bias_p = None
    if convt.bias is not None:
        bias_p = torch.zeros((cout + 7) // 8 * 8, dtype=torch.float32, device=w.device)
        bias_p[:cout] = convt.bias.detach()
    return ops.conv_transpose2d(x, w, convt.bias, (wf, wd, bias_p), pad=convt.padding[0])


class DownBlock(nn.Module):
    def __init__(self, in_planes, out_planes, separable=False):
        super().__init__()
        if separable:
            raise NotImplementedError("B200 DownBlock: separable convolutions")
        self.main = nn.Sequential(sn_conv2d(in_planes, out_planes, 4, 2, 1), nn.GroupNorm(out_planes // 2, out_planes),
                                  nn.LeakyReLU(0.2, inplace=True))

    def forward_nhwc(self, x):
        x = _sn_conv(x, self.main[0])
        gn = self.main[1]
        return ops.group_norm(x, gn.weight, gn.bias, gn.num_groups, film=None, act=L.ACT_LRELU02, eps=gn.eps)


class SingleDisc(nn.Module):
    def __init__(self, nc=None, ndf=None, start_sz=256, end_sz=8, head=None, separable=False, patch=False):
        super().__init__()
        if head or patch:
            raise NotImplementedError("B200 SingleDisc: head / patch variants")
        if start_sz not in CHANNEL_DICT:
            start_sz = min(CHANNEL_DICT.keys(), key=lambda s: abs(s - start_sz))
        self.start_sz = start_sz
        nfc = dict(CHANNEL_DICT) if ndf is None else {k: ndf for k in CHANNEL_DICT}
        if nc is not None:
            nfc[start_sz] = nc
        layers = []
        while start_sz > end_sz:
            layers.append(DownBlock(nfc[start_sz], nfc[start_sz // 2], separable))
            start_sz //= 2
        layers.append(sn_conv2d(nfc[end_sz], 1, 4, 1, 0, bias=False))
        self.main = nn.Sequential(*layers)

    def forward_nhwc(self, x):
        for layer in list(self.main)[:-1]:
            x = layer.forward_nhwc(x)
        return _sn_conv(x, self.main[-1])

    def forward(self, x):
        """x NCHW fp32 feature map -> logits NCHW fp32 [N, 1, h, w]."""
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)), 1)


class MultiScaleD(nn.Module):
    def __init__(self, channels, resolutions, conv, feats, num_discs=4, proj_type=2, cond=0, separable=False,
                 patch=False, **kwargs):
        super().__init__()
        assert num_discs in [1, 2, 3, 4]
        if cond or not conv or patch:
            raise NotImplementedError("B200 MultiScaleD: conditional / MLP (ViT backbone) / patch mini-discriminators")
        self.disc_in_channels = channels[:num_discs]
        self.disc_in_res = resolutions[:num_discs]
        self.mini_discs = nn.ModuleDict({
            str(i): SingleDisc(nc=cin, start_sz=res, end_sz=8, separable=separable, patch=patch)
            for i, (cin, res) in enumerate(zip(self.disc_in_channels, self.disc_in_res))})

    def forward(self, features):
        """features: {"0": NCHW fp32 map, ...} (what the frozen projector returns) -> [N, sum of logits]."""
        all_logits = []
        for k, disc in self.mini_discs.items():
            all_logits.append(disc(features[k]).reshape(features[k].size(0), -1))
        return torch.cat(all_logits, dim=1)
