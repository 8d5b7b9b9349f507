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


def generator_inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    shp = (batch, 3, cfg.image_size, cfg.image_size)
    gt = (0.5 * torch.randn(shp, generator=g)).clamp(-1, 1)
    mask = (torch.rand((batch, 1, cfg.image_size, cfg.image_size), generator=g) > 0.6).long()
    cond = gt * (1 - mask) + torch.randn(shp, generator=g) * mask
    noise = torch.randn(shp, generator=g)
    ref = 0.5 * torch.randn(shp, generator=g)
    return gt, cond, mask, noise, ref


def generator_golden():
    """cfg 4 end to end: DiffusionGenerator(PaletteDenoiseFn(UNetGeneratorRefAttn)) — training forward + Palette loss
    + backward with the reference image, then DDPM (n_timestep_test steps) and DDIM sampling with it."""
    from models.modules.diffusion_generator import DiffusionGenerator
    from models.modules.palette_denoise_fn import PaletteDenoiseFn
    cfgd = dict(CFG, n_timestep_test=8)
    cfg = O.UNetCfg(**cfgd)
    dn = PaletteDenoiseFn(model=build_reference(cfg), cond_embed_dim=cfg.cond_embed_dim, ref_embed_net="",
                          conditioning="", nclasses=2)
    assert dn.model_nargs == 3
    net = DiffusionGenerator(denoise_fn=dn, sampling_method="ddpm", image_size=cfg.image_size, G_ngf=cfg.inner_channel,
                             loading_backward_compatibility=False)
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    params = init_params_from_shapes(shapes, seed=9)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all(("gammas" in m or "posterior" in m) for m in missing), (missing, unexpected)
    batch, rseed = 2, 55
    gt, cond, mask, noise, ref = generator_inputs(cfg, batch, seed=13)
    torch.manual_seed(rseed)
    n_out, noise_hat, _ = net(y_0=gt, y_cond=cond, mask=mask, noise=noise, cls=None, ref=ref)
    mb = torch.clamp(mask, min=0, max=1)
    loss = torch.nn.MSELoss()(mb * n_out, mb * noise_hat)
    loss.backward()
    torch.manual_seed(rseed)
    t, u = O.sample_t_gamma(cfg, batch)
    grads = {}
    for k, p in net.named_parameters():
        g = p.grad.detach() if p.grad is not None else torch.zeros_like(p)
        grads[k] = {"sum": float(g.double().sum()), "l2": float(g.double().norm()), "head": g.flatten()[:16].clone(),
                    "none": p.grad is None}
    # sampling with the reference image
    sample_num = 2
    torch.manual_seed(rseed + 1)
    with torch.no_grad():
        y, ret = net.restoration(cond, y_t=None, y_0=gt, mask=mask, sample_num=sample_num, ref=ref)
    torch.manual_seed(rseed + 1)
    y_t0 = torch.randn_like(gt)
    noises = {i: torch.randn_like(gt) for i in reversed(range(1, cfg.n_timestep_test))}
    net.sampling_method = "ddim"
    with torch.no_grad():
        yd, retd = net.restoration(cond, y_t=y_t0.clone(), y_0=gt, mask=mask, sample_num=sample_num, ref=ref,
                                   ddim_num_steps=4, ddim_eta=0.5)
    torch.save({"cfg": cfgd, "batch": batch, "wseed": 9, "dseed": 13, "rseed": rseed, "t": t, "u": u,
                "torch_version": str(torch.__version__), "shapes": shapes, "noise_hat": noise_hat.detach().clone(),
                "loss": float(loss.detach()), "grads": grads, "sample_num": sample_num, "y": y.clone(),
                "ret_arr": ret.clone(), "ddim_steps": 4, "ddim_eta": 0.5, "y_ddim": yd.clone(),
                "ret_arr_ddim": retd.clone()}, os.path.join(GOLDEN, "refattn_generator.pt"))
    # the restatement against the reference, right here
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    _, nh, _ = O.diffusion_forward(leaves, gt, cond, mask, noise, t, u, cfg, unet=R.denoiser(ref))
    lo = torch.nn.MSELoss()(mb * noise, mb * nh)
    lo.backward()
    zero = lambda g, p: g if g is not None else torch.zeros_like(p)  # noqa: E731
    gerr = max(float((zero(leaves[k].grad, p) - zero(p.grad, p)).norm() / (zero(p.grad, p).norm() + 1e-9))
               for k, p in net.named_parameters())
    print("refattn_generator.pt: loss %.6f (oracle %.6f), noise_hat rel max err %.2e, worst grad rel L2 err %.2e" % (
        float(loss), float(lo), float((nh - noise_hat).abs().max() / noise_hat.abs().max()), gerr))
    with torch.no_grad():
        yo, reto = O.restoration_ddpm(params, cond, y_t0, gt, mask, noises, cfg, sample_num, unet=R.denoiser(ref))
        ydo, retdo = O.restoration_ddim(params, cond, y_t0.clone(), gt, mask, cfg, sample_num, num_steps=4, eta=0.5,
                                        unet=R.denoiser(ref))
    print("sampling oracle vs reference: DDPM y %.2e ret %.2e, DDIM y %.2e ret %.2e" % (
        float((yo - y).abs().max() / y.abs().max()), float((reto - ret).abs().max() / ret.abs().max()),
        float((ydo - yd).abs().max() / yd.abs().max()), float((retdo - retd).abs().max() / retd.abs().max())))


if __name__ == "__main__":
    main()
    generator_golden()
