"""Generate the reference-attention UNet golden vectors from the UNMODIFIED reference (/root/reference) imported on
CPU in the build container (TEST INFRASTRUCTURE — see oracle/__init__.py).

    python -m oracle.gen_golden_ref        # writes tests/golden/refattn_small.pt

Fixture: UNetGeneratorRefAttn(64 ch, mults (1,2,4), res (1,1,1), attention at ds=2 and 4, head channels 32) on
x [2, 6, 32, 32], ref [2, 3, 32, 32]; seeded de-zeroed weights (regenerated from the stored (key, shape) list by
oracle.vid_oracle.init_params_from_shapes); y, loss sum(y*g), per-parameter gradient (sum, L2, first 16 values).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import palette_oracle as O  # noqa: E402
from oracle import ref_oracle as R  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CFG = dict(image_size=32, in_channel=6, inner_channel=64, out_channel=3, res_blocks=(1, 1, 1), attn_res=(2, 4),
           channel_mults=(1, 2, 4), num_head_channels=32)


def build_reference(cfg: O.UNetCfg):
    from models.modules.unet_generator_attn.unet_generator_attn import UNetGeneratorRefAttn
    return UNetGeneratorRefAttn(image_size=cfg.image_size, in_channel=cfg.in_channel,
                                inner_channel=cfg.inner_channel, out_channel=cfg.out_channel,
                                res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res), tanh=False,
                                n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test,
                                norm="groupnorm", group_norm_size=cfg.group_norm_size,
                                cond_embed_dim=cfg.cond_embed_dim, channel_mults=cfg.channel_mults,
                                num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                                efficient=cfg.efficient)


def inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cfg.in_channel, cfg.image_size, cfg.image_size, generator=g)
    ref = torch.randn(batch, cfg.in_channel // 2, cfg.image_size, cfg.image_size, generator=g)
    emb = torch.randn(batch, cfg.cond_embed_dim, generator=g)
    gy = torch.randn(batch, cfg.out_channel, cfg.image_size, cfg.image_size, generator=g)
    return x, ref, emb, gy


def main():
    ref_stubs.install()
    cfg = O.UNetCfg(**CFG)
    net = build_reference(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    params = init_params_from_shapes(shapes, seed=7)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    batch = 2
    x, ref, emb, gy = inputs(cfg, batch, seed=11)
    y = net(x, emb, ref)
    loss = (y * gy).sum()
    loss.backward()
    grads = {}
    for k, p in net.named_parameters():
        g = p.grad.detach() if p.grad is not None else torch.zeros_like(p)
        grads[k] = {"sum": float(g.double().sum()), "l2": float(g.double().norm()),
                    "head": g.flatten()[:16].clone(), "full": g.clone() if g.numel() <= 4096 else None,
                    "none": p.grad is None}
    out = {"cfg": CFG, "batch": batch, "wseed": 7, "dseed": 11, "torch_version": str(torch.__version__),
           "shapes": shapes, "y": y.detach().clone(), "loss": float(loss.detach()), "grads": grads}
    torch.save(out, os.path.join(GOLDEN, "refattn_small.pt"))
    print("refattn_small.pt: %d parameter tensors, loss %.6f, |y|max %.4f, params without grad: %d" % (
        len(shapes), out["loss"], float(y.abs().max()), sum(g["none"] for g in grads.values())))
    params_r = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    yo = R.unet_ref_forward(params_r, x, emb, ref, cfg)
    (yo * gy).sum().backward()
    err = float((yo - y).abs().max() / y.abs().max())
    gerr = max(float(((params_r[k].grad if params_r[k].grad is not None else torch.zeros_like(p)) -
                      (p.grad if p.grad is not None else torch.zeros_like(p))).norm() /
                     ((p.grad.norm() if p.grad is not None else 0) + 1e-9)) for k, p in net.named_parameters())
    print("oracle vs reference: y rel max err %.2e, worst grad rel L2 err %.2e" % (err, gerr))


if __name__ == "__main__":
    main()
