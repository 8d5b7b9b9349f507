"""Golden vectors for the conditioned Palette generator (alg_diffusion_cond_embed "class" / "mask" / "class_mask",
example_ddpm_mario.json ships "class") from the UNMODIFIED reference:

    python -m oracle.gen_golden_cond       -> tests/golden/palette_cond_{class,mask,class_mask}.pt

DiffusionGenerator(PaletteDenoiseFn(UNet, conditioning=...)) forward + loss + every parameter gradient, with the label
tables scaled so that some rows exceed the embedding's max_norm (the in-place renormalisation is part of the forward).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import palette_oracle as O  # noqa: E402
from oracle import ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
BASE = dict(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,), num_head_channels=16)


def cond_cfg(conditioning, nclasses=4):
    e = 32
    return O.UNetCfg(in_channel=6 + (e if "mask" in conditioning else 0), conditioning=conditioning, nclasses=nclasses,
                     **BASE)


def cond_params(cfg, seed):
    """oracle init; label tables: N(0,1)/sqrt(dim) rows, the first one scaled to norm ~2 (> max_norm = 1)."""
    params = O.init_params(cfg, seed)
    for k in params:
        if "embedding_table" in k:
            params[k] = params[k].clone()
            params[k][0] *= 2.0 / params[k][0].norm()
    return params


def cond_batch(cfg, batch, seed):
    data = O.synthetic_batch(batch, cfg.image_size, seed)
    g = torch.Generator().manual_seed(seed + 1)
    # semantic masks with several classes (0 = background) instead of the binary box, class labels per image
    cls_map = torch.randint(1, cfg.nclasses, (batch, 1, 1, 1), generator=g)
    data["mask"] = data["mask"] * cls_map
    data["cls"] = torch.randint(0, cfg.nclasses, (batch,), generator=g)
    data["cls"][0] = 0  # the over-long table row is looked up
    return data


def main():
    ref_stubs.install()
    from models.modules.diffusion_generator import DiffusionGenerator
    from models.modules.palette_denoise_fn import PaletteDenoiseFn
    from models.modules.unet_generator_attn.unet_generator_attn import UNet

    for conditioning in ("class", "mask", "class_mask"):
        cfg = cond_cfg(conditioning)
        unet = UNet(image_size=cfg.image_size, in_channel=cfg.in_channel, inner_channel=cfg.inner_channel,
                    out_channel=cfg.out_channel, res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res),
                    tanh=False, n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test,
                    norm="groupnorm", group_norm_size=cfg.group_norm_size, cond_embed_dim=cfg.cond_embed_dim,
                    channel_mults=cfg.channel_mults, num_heads=cfg.num_heads,
                    num_head_channels=cfg.num_head_channels, efficient=cfg.efficient)
        dn = PaletteDenoiseFn(model=unet, cond_embed_dim=cfg.cond_embed_dim, ref_embed_net="",
                              conditioning=conditioning, nclasses=cfg.nclasses)
        net = DiffusionGenerator(denoise_fn=dn, sampling_method="ddpm", image_size=cfg.image_size,
                                 G_ngf=cfg.inner_channel, loading_backward_compatibility=False)
        wseed, dseed, rseed, batch = 31, 41, 51, 3
        params = cond_params(cfg, wseed)
        ref_shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
        assert ref_shapes == list(O.generator_param_shapes(cfg).items()), "oracle parameter list != reference"
        missing, unexpected = net.load_state_dict(params, strict=False)
        assert not unexpected and all("gammas" in m or "posterior" in m for m in missing), (missing, unexpected)
        data = cond_batch(cfg, batch, dseed)
        torch.manual_seed(rseed)
        noise, noise_hat, w = net(y_0=data["gt"], y_cond=data["cond"], mask=data["mask"], noise=None,
                                  cls=data["cls"] if "class" in conditioning else None, ref=None)
        mask_binary = torch.clamp(data["mask"], min=0, max=1)
        loss = torch.nn.MSELoss()(mask_binary * noise, mask_binary * noise_hat)
        loss.backward()
        out = {"conditioning": conditioning, "nclasses": cfg.nclasses, "batch": batch, "wseed": wseed, "dseed": dseed,
               "rseed": rseed, "torch_version": str(torch.__version__), "noise_hat": noise_hat.detach().clone(),
               "loss": float(loss),
               "grad_stats": {k: (float(p.grad.double().sum()), float(p.grad.double().norm()))
                              for k, p in net.named_parameters()},
               "grads": {k: p.grad.detach().clone() for k, p in net.named_parameters()
                         if any(t in k for t in ("embedding_table", "cond_embed", "input_blocks.0.0", "out.2"))},
               "tables_after": {k: v.detach().clone() for k, v in net.state_dict().items() if "embedding_table" in k}}
        if conditioning == "class_mask":
            # sampling with both conditionings (row (f)-1): DDPM over the 8 test steps and DDIM (4 steps) of the
            # unmodified reference, the random draws recorded by replaying the seed (as oracle/gen_golden.py)
            scfg = O.UNetCfg(in_channel=cfg.in_channel, conditioning=conditioning, nclasses=cfg.nclasses,
                             n_timestep_test=8, **BASE)
            sunet = UNet(image_size=scfg.image_size, in_channel=scfg.in_channel, inner_channel=scfg.inner_channel,
                         out_channel=scfg.out_channel, res_blocks=list(scfg.res_blocks), attn_res=list(scfg.attn_res),
                         tanh=False, n_timestep_train=scfg.n_timestep_train, n_timestep_test=scfg.n_timestep_test,
                         norm="groupnorm", group_norm_size=scfg.group_norm_size, cond_embed_dim=scfg.cond_embed_dim,
                         channel_mults=scfg.channel_mults, num_heads=scfg.num_heads,
                         num_head_channels=scfg.num_head_channels, efficient=scfg.efficient)
            sdn = PaletteDenoiseFn(model=sunet, cond_embed_dim=scfg.cond_embed_dim, ref_embed_net="",
                                   conditioning=conditioning, nclasses=scfg.nclasses)
            snet = DiffusionGenerator(denoise_fn=sdn, sampling_method="ddpm", image_size=scfg.image_size,
                                      G_ngf=scfg.inner_channel, loading_backward_compatibility=False)
            snet.load_state_dict(params, strict=False)
            srseed, sample_num = 61, 2
            torch.manual_seed(srseed)
            with torch.no_grad():
                y, ret = snet.restoration(data["cond"], y_t=None, y_0=data["gt"], mask=data["mask"],
                                          sample_num=sample_num, cls=data["cls"])
            torch.manual_seed(srseed)
            y_t0 = torch.randn_like(data["gt"])
            noises = {i: torch.randn_like(data["gt"]) for i in reversed(range(1, scfg.n_timestep_test))}
            sparams = {k: v.clone() for k, v in snet.state_dict().items()}   # (tables renormalised by the lookups)
            with torch.no_grad():
                yo, reto = O.restoration_ddpm(sparams, data["cond"], y_t0, data["gt"], data["mask"], noises, scfg,
                                              sample_num, cls=data["cls"])
            print("conditioned DDPM oracle vs reference: y %.2e ret %.2e" % (
                float((yo - y).abs().max() / y.abs().max()), float((reto - ret).abs().max() / ret.abs().max())))
            snet.sampling_method = "ddim"
            with torch.no_grad():
                yd, retd = snet.restoration(data["cond"], y_t=y_t0.clone(), y_0=data["gt"], mask=data["mask"],
                                            sample_num=sample_num, cls=data["cls"], ddim_num_steps=4, ddim_eta=0.5)
                ydo, _ = O.restoration_ddim(sparams, data["cond"], y_t0.clone(), data["gt"], data["mask"], scfg,
                                            sample_num, num_steps=4, eta=0.5, cls=data["cls"])
            print("conditioned DDIM oracle vs reference: y %.2e" % float((ydo - yd).abs().max() / yd.abs().max()))
            out["sampling"] = {"n_timestep_test": 8, "rseed": srseed, "sample_num": sample_num, "y": y.clone(),
                               "ret_arr": ret.clone(), "ddim_steps": 4, "ddim_eta": 0.5, "y_ddim": yd.clone(),
                               "ret_arr_ddim": retd.clone()}
        torch.save(out, os.path.join(GOLDEN, "palette_cond_%s.pt" % conditioning))
        print(conditioning, "loss", out["loss"], "noise_hat absmax", float(noise_hat.abs().max()))


if __name__ == "__main__":
    sys.exit(main())
