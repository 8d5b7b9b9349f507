"""Generate the video-UNet golden vectors from the UNMODIFIED reference (/root/reference) imported on CPU in the
build container (TEST INFRASTRUCTURE — see oracle/__init__.py).

    python -m oracle.gen_golden_vid        # writes tests/golden/vid_small.pt

Fixture: UNetVid(64 ch, mults (1,2), res (1,1), attention at ds=2, head channels 32, 8 temporal heads,
2 transformer blocks) on a clip [b=2, f=4, 6, 16, 16]; seeded de-zeroed weights (regenerated from the stored
(key, shape) list by oracle.vid_oracle.init_params_from_shapes); stores y, the loss sum(y*g) and for every parameter
gradient (sum, L2, first 16 values) — full tensors for the small ones.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_stubs  # noqa: E402
from oracle import vid_oracle as V  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CFG = dict(image_size=16, in_channel=6, inner_channel=64, out_channel=3, res_blocks=(1, 1), attn_res=(2,),
           channel_mults=(1, 2), num_head_channels=32, max_sequence_length=25, num_attention_heads=8,
           num_transformer_blocks=2)


def build_reference(cfg: V.VidCfg):
    from models.modules.unet_generator_attn.unet_generator_attn_vid import UNetVid
    return UNetVid(image_size=cfg.image_size, in_channel=cfg.in_channel, inner_channel=cfg.inner_channel,
                   out_channel=cfg.out_channel, res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res),
                   tanh=False, n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test,
                   norm="groupnorm", group_norm_size=cfg.group_norm_size, cond_embed_dim=cfg.cond_embed_dim,
                   channel_mults=cfg.channel_mults, num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                   efficient=cfg.efficient, max_sequence_length=cfg.max_sequence_length,
                   num_attention_heads=cfg.num_attention_heads, num_transformer_blocks=cfg.num_transformer_blocks)


def inputs(cfg, batch, frames, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, frames, cfg.in_channel, cfg.image_size, cfg.image_size, generator=g)
    emb = torch.randn(batch, cfg.cond_embed_dim, generator=g)
    gy = torch.randn(batch, frames, cfg.out_channel, cfg.image_size, cfg.image_size, generator=g)
    return x, emb, gy


def main():
    ref_stubs.install()
    cfg = V.VidCfg(**CFG)
    net = build_reference(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    params = V.init_params_from_shapes(shapes, seed=5)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all(m.endswith("pos_encoder.pe") for m in missing), (missing, unexpected)
    # the oracle's positional-encoding buffers must be the reference's
    sd = V.add_buffers(params, cfg)
    for k, v in net.state_dict().items():
        if k.endswith("pos_encoder.pe"):
            assert torch.equal(v, sd[k]), k
    batch, frames = 2, 4
    x, emb, gy = inputs(cfg, batch, frames, seed=9)
    y = net(x, emb)
    loss = (y * gy).sum()
    loss.backward()
    grads = {}
    for k, p in net.named_parameters():
        g = p.grad.detach()
        grads[k] = {"sum": float(g.double().sum()), "l2": float(g.double().norm()),
                    "head": g.flatten()[:16].clone(), "full": g.clone() if g.numel() <= 4096 else None}
    out = {"cfg": CFG, "batch": batch, "frames": frames, "wseed": 5, "dseed": 9,
           "torch_version": str(torch.__version__), "shapes": shapes, "y": y.detach().clone(), "loss": float(loss),
           "grads": grads}
    torch.save(out, os.path.join(GOLDEN, "vid_small.pt"))
    print("vid_small.pt: %d parameters tensors, loss %.6f, |y|max %.4f" % (len(shapes), out["loss"],
                                                                          float(y.abs().max())))
    # the restatement against the reference, right here
    params_r = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    yo = V.unet_vid_forward(V.add_buffers(params_r, cfg), x, emb, cfg)
    (yo * gy).sum().backward()
    err = float((yo - y).abs().max() / y.abs().max())
    gerr = max(float((params_r[k].grad - p.grad).norm() / (p.grad.norm() + 1e-12)) for k, p in net.named_parameters())
    print("oracle vs reference: y rel max err %.2e, worst grad rel L2 err %.2e" % (err, gerr))


def generator_inputs(cfg, batch, frames, seed):
    g = torch.Generator().manual_seed(seed)
    shp = (batch, frames, 3, cfg.image_size, cfg.image_size)
    gt = (0.5 * torch.randn(shp, generator=g)).clamp(-1, 1)
    mask = (torch.rand((batch, frames, 1, cfg.image_size, cfg.image_size), generator=g) > 0.6).long()
    cond = gt * (1 - mask) + torch.randn(shp, generator=g) * mask
    noise = torch.randn(shp, generator=g)
    return gt, cond, mask, noise


def generator_golden():
    """cfg 5 end to end: DiffusionGenerator(PaletteDenoiseFn(UNetVid)) forward on a clip + the Palette loss + backward."""
    from einops import rearrange
    from models.modules.diffusion_generator import DiffusionGenerator
    from models.modules.palette_denoise_fn import PaletteDenoiseFn
    from oracle import palette_oracle as O
    cfg = V.VidCfg(**CFG)
    dn = PaletteDenoiseFn(model=build_reference(cfg), cond_embed_dim=cfg.cond_embed_dim, ref_embed_net="",
                          conditioning="", nclasses=2)
    net = DiffusionGenerator(denoise_fn=dn, sampling_method="ddpm", image_size=cfg.image_size, G_ngf=cfg.inner_channel,
                             loading_backward_compatibility=False)
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    params = V.init_params_from_shapes(shapes, seed=6)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all(("gammas" in m or "posterior" in m or m.endswith("pos_encoder.pe")) for m in missing), \
        (missing, unexpected)
    batch, frames, rseed = 2, 4, 77
    gt, cond, mask, noise = generator_inputs(cfg, batch, frames, seed=10)
    torch.manual_seed(rseed)
    noise_fh = rearrange(noise, "b f c h w -> b c (f h) w")
    n_out, noise_hat, _ = net(y_0=gt, y_cond=cond, mask=mask, noise=noise_fh, cls=None, ref=None)
    assert torch.equal(n_out, noise)
    mb = torch.clamp(mask, min=0, max=1)
    loss = torch.nn.MSELoss()(mb * n_out, mb * noise_hat)
    loss.backward()
    torch.manual_seed(rseed)
    t, u = O.sample_t_gamma(cfg, batch)
    grads = {k: {"sum": float(p.grad.double().sum()), "l2": float(p.grad.double().norm()),
                 "head": p.grad.flatten()[:16].clone()} for k, p in net.named_parameters()}
    torch.save({"cfg": CFG, "batch": batch, "frames": frames, "wseed": 6, "dseed": 10, "rseed": rseed, "t": t, "u": u,
                "torch_version": str(torch.__version__), "shapes": shapes, "noise_hat": noise_hat.detach().clone(),
                "loss": float(loss.detach()), "grads": grads}, os.path.join(GOLDEN, "vid_generator.pt"))
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    sd = V.add_buffers(leaves, cfg)
    _, nh = V.diffusion_forward_vid(sd, gt, cond, mask, noise, t, u, cfg)
    lo = torch.nn.MSELoss()(mb * noise, mb * nh)
    lo.backward()
    gerr = max(float((leaves[k].grad - p.grad).norm() / (p.grad.norm() + 1e-12)) for k, p in net.named_parameters())
    print("vid_generator.pt: loss %.6f (oracle %.6f), noise_hat rel max err %.2e, worst grad rel L2 err %.2e" % (
        float(loss), float(lo), float((nh - noise_hat).abs().max() / noise_hat.abs().max()), gerr))


if __name__ == "__main__":
    ref_stubs.install()
    main()
    generator_golden()
