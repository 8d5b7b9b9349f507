"""CPU oracle for the GAN generator / discriminator operators (TEST INFRASTRUCTURE — see oracle/__init__.py).

Functional fp32 restatement over {state_dict key: tensor} of

  ResnetBlock / ResnetEncoder / ResnetDecoder / ResnetGenerator
      /root/reference/models/modules/resnet_architecture/resnet_generator.py:11-95, 98-164, 167-271, 274-347
      (G_netG "resnet": reflect padding, InstanceNorm2d(affine=False), use_bias=True)
  NLayerDiscriminator   /root/reference/models/modules/discriminators.py:10-117
  GANLoss (lsgan / wgangp / projected)   /root/reference/models/modules/loss.py:11-85

Pinned against the real reference by oracle/gen_golden_gan.py + tests/test_oracle_golden.py.  The bf16-storage
emulation switch of oracle.palette_oracle (EMULATE_BF16) applies here too.
"""
import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .palette_oracle import _r, _rw


def resnet_param_shapes(input_nc=3, output_nc=3, ngf=64, n_blocks=9) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(name, cin, cout, k):
        s[name + ".weight"] = (cout, cin, k, k)
        s[name + ".bias"] = (cout,)

    conv("encoder.model.1", input_nc, ngf, 7)
    conv("encoder.model.4", ngf, ngf * 2, 3)
    conv("encoder.model.7", ngf * 2, ngf * 4, 3)
    for i in range(n_blocks):
        conv("encoder.model.%d.conv_block.1" % (10 + i), ngf * 4, ngf * 4, 3)
        conv("encoder.model.%d.conv_block.5" % (10 + i), ngf * 4, ngf * 4, 3)
    s["decoder.model.0.weight"] = (ngf * 4, ngf * 2, 3, 3)  # ConvTranspose2d: [Cin, Cout, k, k]
    s["decoder.model.0.bias"] = (ngf * 2,)
    s["decoder.model.3.weight"] = (ngf * 2, ngf, 3, 3)
    s["decoder.model.3.bias"] = (ngf,)
    conv("decoder.model.7", ngf, output_nc, 7)
    return s


def nlayer_d_param_shapes(input_nc=3, ndf=64, n_layers=3) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    idx = 0

    def conv(cin, cout):
        s["model.%d.weight" % idx] = (cout, cin, 4, 4)
        s["model.%d.bias" % idx] = (cout,)

    conv(input_nc, ndf)
    idx = 2
    mult = 1
    for n in range(1, n_layers):
        prev, mult = mult, min(2 ** n, 8)
        conv(ndf * prev, ndf * mult)
        idx += 3
    prev, mult = mult, min(2 ** n_layers, 8)
    conv(ndf * prev, ndf * mult)
    idx += 3
    conv(ndf * mult, 1)
    return s


def init_from_shapes(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in shapes.items():
        if name.endswith(".bias"):
            out[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            out[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
    return out


def _in(x):
    return F.instance_norm(x, eps=1e-5)


def _conv(x, sd, name, stride=1, padding=0):
    return F.conv2d(x, _rw(sd[name + ".weight"]), sd[name + ".bias"], stride=stride, padding=padding)


def resnet_encoder(sd, x, n_blocks=9, prefix="encoder.model.", feats=None):
    """ResnetEncoder.compute_feats (resnet_generator.py:238-256); feats: optional dict layer_id -> tensor."""
    def rec(i, t):
        if feats is not None:
            feats[i] = t
        return t

    def in_relu(i, t):
        # layer i = InstanceNorm2d, i + 1 = nn.ReLU(True): the ReLU is IN-PLACE, so the tensor that
        # compute_feats recorded for layer i (:249-250) is overwritten — features at i and i + 1 are both the
        # activated map (checked against the reference's get_feats golden).
        out = _r(F.relu(_in(t)))
        rec(i, out)
        return rec(i + 1, out)

    h = rec(0, _r(F.pad(_r(x), (3, 3, 3, 3), mode="reflect")))
    h = in_relu(2, rec(1, _r(_conv(h, sd, prefix + "1"))))
    h = in_relu(5, rec(4, _r(_conv(h, sd, prefix + "4", stride=2, padding=1))))
    h = in_relu(8, rec(7, _r(_conv(h, sd, prefix + "7", stride=2, padding=1))))
    for i in range(n_blocks):
        p = prefix + "%d.conv_block." % (10 + i)
        t = _r(F.pad(h, (1, 1, 1, 1), mode="reflect"))
        t = _r(_conv(t, sd, p + "1"))
        t = _r(F.relu(_in(t)))
        t = _r(F.pad(t, (1, 1, 1, 1), mode="reflect"))
        t = _r(_conv(t, sd, p + "5"))
        t = _r(_in(t))
        h = rec(10 + i, _r(h + t))  # ResnetBlock.forward (:92-95)
    return h


def resnet_decoder(sd, h, prefix="decoder.model."):
    h = _r(F.conv_transpose2d(h, _rw(sd[prefix + "0.weight"]), sd[prefix + "0.bias"], stride=2, padding=1,
                              output_padding=1))
    h = _r(F.relu(_in(h)))
    h = _r(F.conv_transpose2d(h, _rw(sd[prefix + "3.weight"]), sd[prefix + "3.bias"], stride=2, padding=1,
                              output_padding=1))
    h = _r(F.relu(_in(h)))
    h = _r(F.pad(h, (3, 3, 3, 3), mode="reflect"))
    return _r(torch.tanh(_conv(h, sd, prefix + "7")))


def resnet_generator(sd, x, n_blocks=9):
    return resnet_decoder(sd, resnet_encoder(sd, x, n_blocks))


def nlayer_discriminator(sd, x, n_layers=3, prefix="model."):
    h = _r(F.leaky_relu(_conv(_r(x), sd, prefix + "0", stride=2, padding=1), 0.2))
    idx = 2
    for n in range(1, n_layers):
        h = _r(_conv(h, sd, prefix + str(idx), stride=2, padding=1))
        h = _r(F.leaky_relu(_in(h), 0.2))
        idx += 3
    h = _r(_conv(h, sd, prefix + str(idx), stride=1, padding=1))
    h = _r(F.leaky_relu(_in(h), 0.2))
    idx += 3
    return _r(_conv(h, sd, prefix + str(idx), stride=1, padding=1))


def gan_loss(pred, target_is_real, mode="lsgan", relu=True):
    """GANLoss.__call__ (loss.py:57-85)."""
    if mode == "lsgan":
        return F.mse_loss(pred, torch.full_like(pred, 1.0 if target_is_real else 0.0))
    if mode == "wgangp":
        return -pred.mean() if target_is_real else pred.mean()
    if mode == "projected":
        if relu:
            return F.relu(1 - pred).mean() if target_is_real else F.relu(1 + pred).mean()
        return (-pred).mean()
    raise NotImplementedError(mode)


def gan_train_step(state_G, state_D, oc_G, oc_D, real_A, real_B, n_blocks=9, n_layers=3, lambda_gan=1.0, mode="lsgan"):
    """One optimize_parameters() of the (G), (D) groups restricted to the GAN terms
    (cut_model.py:406-437, base_gan_model.py:382-419,457-503, loss.py:288-313).  Returns (loss_G, loss_D)."""
    from .palette_oracle import adam_update
    gl = {k: v.detach().clone().requires_grad_(True) for k, v in state_G.params.items()}
    dl = {k: v.detach().clone() for k, v in state_D.params.items()}
    fake = resnet_generator(gl, real_A, n_blocks)
    loss_G = lambda_gan * gan_loss(nlayer_discriminator(dl, fake, n_layers), True, mode, relu=False)
    loss_G.backward()
    with torch.no_grad():
        adam_update(state_G, {k: v.grad for k, v in gl.items()}, oc_G)
    dl = {k: v.detach().clone().requires_grad_(True) for k, v in state_D.params.items()}
    loss_D = 0.5 * (gan_loss(nlayer_discriminator(dl, real_B, n_layers), True, mode)
                    + gan_loss(nlayer_discriminator(dl, fake.detach(), n_layers), False, mode))
    loss_D.backward()
    with torch.no_grad():
        adam_update(state_D, {k: v.grad for k, v in dl.items()}, oc_D)
    return loss_G.detach(), loss_D.detach()
