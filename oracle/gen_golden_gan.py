"""Golden vectors for the GAN operators from the UNMODIFIED reference (TEST INFRASTRUCTURE).

    python -m oracle.gen_golden_gan      # writes tests/golden/gan_*.pt

  gan_resnet.pt   ResnetGenerator(3, 3, ngf=16, instance norm, 2 blocks, reflect) 32x32 b=2: output, encoder
                  features at layers (0,4,8,12... as CUT samples), all parameter gradients for a random dy
  gan_nlayerd.pt  NLayerDiscriminator(3, ndf=16, n_layers=3, instance) 64x64 b=2: logits, lsgan losses
                  (real / fake), all parameter gradients of the real-label loss
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gan_oracle as G  # noqa: E402
from oracle import ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ref_stubs.install()
    from models.modules.discriminators import NLayerDiscriminator
    from models.modules.loss import GANLoss
    from models.modules.resnet_architecture.resnet_generator import ResnetGenerator
    from models.modules.utils import get_norm_layer

    torch.set_num_threads(8)
    norm = get_norm_layer("instance")
    # ---- generator
    ngf, nb = 16, 2
    net = ResnetGenerator(3, 3, ngf, norm_layer=norm, use_dropout=False, n_blocks=nb, padding_type="reflect")
    shapes = G.resnet_param_shapes(3, 3, ngf, nb)
    assert [(k, tuple(v.shape)) for k, v in net.named_parameters()] == list(shapes.items())
    params = G.init_from_shapes(shapes, 31)
    net.load_state_dict(params)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    y = net(x)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    feats = net.get_feats(x, [0, 4, 8, 11])
    torch.save({"ngf": ngf, "n_blocks": nb, "wseed": 31, "x": x, "dy": dy, "y": y.detach(),
                "feat_ids": [0, 4, 8, 11], "feats": [f.detach() for f in feats],
                "grads": {k: p.grad.detach().clone() for k, p in net.named_parameters()},
                "torch_version": str(torch.__version__)}, os.path.join(GOLDEN, "gan_resnet.pt"))
    print("gan_resnet.pt", float(y.abs().max()))
    # ---- discriminator + lsgan
    ndf = 16
    d = NLayerDiscriminator(3, ndf, n_layers=3, norm_layer=norm)
    shapes = G.nlayer_d_param_shapes(3, ndf, 3)
    assert [(k, tuple(v.shape)) for k, v in d.named_parameters()] == list(shapes.items())
    params = G.init_from_shapes(shapes, 32)
    d.load_state_dict(params)
    x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    pred = d(x)
    crit = GANLoss("lsgan")
    loss_real = crit(pred, True)
    loss_fake = crit(pred, False)
    loss_real.backward()
    torch.save({"ndf": ndf, "wseed": 32, "x": x, "pred": pred.detach(), "loss_real": float(loss_real),
                "loss_fake": float(loss_fake),
                "hinge_real": float(GANLoss("projected")(pred, True)), "hinge_fake": float(GANLoss("projected")(pred, False)),
                "grads": {k: p.grad.detach().clone() for k, p in d.named_parameters()},
                "torch_version": str(torch.__version__)}, os.path.join(GOLDEN, "gan_nlayerd.pt"))
    print("gan_nlayerd.pt", tuple(pred.shape), float(loss_real))


if __name__ == "__main__":
    main()
