"""Generate golden vectors from the UNMODIFIED reference (/root/reference) imported on CPU in the
build container (TEST INFRASTRUCTURE — see oracle/__init__.py).

    python -m oracle.gen_golden            # writes tests/golden/*.pt

The reference cannot travel to the GPU box, so the fixtures are committed.  Weights are NOT stored:
they are regenerated from a seed by oracle.palette_oracle.init_params (this script checks that the
reference's named_parameters() are exactly that key/shape list and loads them with load_state_dict).

Fixtures
  palette_small.pt   DiffusionGenerator(UNet 32ch, mults (1,2), attn at ds=2) 32x32 b=2:
                     t, noise_hat, loss, every parameter gradient (full tensors, fp16-free fp32)
  palette_mid.pt     4-level UNet (32ch, mults (1,2,4,8), res (1,1,1,1), attn at ds=8) 64x64 b=2:
                     noise_hat + loss + per-parameter grad (sum, L2)
  palette_plumbing.pt  full joliGEN plumbing: TrainOptions.parse_json -> create_model -> setup ->
                     set_input -> optimize_parameters() x2 on CPU (gpu_ids=-1): loss per step,
                     per-parameter (sum, L2) of weights and EMA weights after 2 steps
"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import palette_oracle as O  # noqa: E402
from oracle import ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SMALL = dict(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,),
             num_head_channels=16)
MID = dict(image_size=64, inner_channel=32, channel_mults=(1, 2, 4, 8), res_blocks=(1, 1, 1, 1), attn_res=(8,),
           num_head_channels=32)


def build_reference_generator(cfg: O.UNetCfg):
    from models.modules.diffusion_generator import DiffusionGenerator
    from models.modules.palette_denoise_fn import PaletteDenoiseFn
    from models.modules.unet_generator_attn.unet_generator_attn import UNet

    unet = UNet(image_size=cfg.image_size, in_channel=cfg.in_channel, inner_channel=cfg.inner_channel,
                out_channel=cfg.out_channel, res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res),
                tanh=False, n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test,
                norm="groupnorm", group_norm_size=cfg.group_norm_size, cond_embed_dim=cfg.cond_embed_dim,
                channel_mults=cfg.channel_mults, num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                efficient=cfg.efficient)
    dn = PaletteDenoiseFn(model=unet, cond_embed_dim=cfg.cond_embed_dim, ref_embed_net="", conditioning="",
                          nclasses=2)
    return DiffusionGenerator(denoise_fn=dn, sampling_method="ddpm", image_size=cfg.image_size,
                              G_ngf=cfg.inner_channel, loading_backward_compatibility=False)


def module_golden(name, cfgd, batch, wseed, dseed, rseed, full_grads):
    cfg = O.UNetCfg(**cfgd)
    net = build_reference_generator(cfg)
    params = O.init_params(cfg, wseed)
    ref_shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    assert ref_shapes == list(O.generator_param_shapes(cfg).items()), "oracle parameter list != reference"
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all("gammas" in m or "posterior" in m for m in missing), (missing, unexpected)
    data = O.synthetic_batch(batch, cfg.image_size, dseed)
    torch.manual_seed(rseed)
    noise, noise_hat, w = net(y_0=data["gt"], y_cond=data["cond"], mask=data["mask"], noise=None, cls=None, ref=None)
    mask_binary = torch.clamp(data["mask"], min=0, max=1)
    loss = torch.nn.MSELoss()(mask_binary * noise, mask_binary * noise_hat)
    loss.backward()
    # the same three draws, in the reference's order, for the record
    torch.manual_seed(rseed)
    t, u = O.sample_t_gamma(cfg, batch)
    out = {
        "cfg": cfgd, "batch": batch, "wseed": wseed, "dseed": dseed, "rseed": rseed,
        "torch_version": str(torch.__version__),
        "t": t, "u": u,
        "noise_sum": float(noise.double().sum()),
        "noise_hat": noise_hat.detach().clone(),
        "min_snr_w": w.detach().clone(),
        "loss": float(loss),
        "grad_stats": {k: (float(p.grad.double().sum()), float(p.grad.double().norm())) for k, p in
                       net.named_parameters()},
    }
    if full_grads:
        out["grads"] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    torch.save(out, os.path.join(GOLDEN, name))
    print(name, "loss", out["loss"], "noise_hat absmax", float(noise_hat.abs().max()))


def create_reference_model(size=32, batch=2, extra=None):
    """options -> create_model -> setup of the reference's PaletteModel on CPU (train.py:183-281) -> (model, opt)"""
    import train as ref_train
    from models import create_model
    from options.train_options import TrainOptions

    with open(os.path.join(ref_stubs.REFERENCE_ROOT, "examples", "example_ddpm_mario.json")) as f:
        nested = json.load(f)

    def flatten(d, prefix=""):
        flat = {}
        for k, v in d.items():
            if isinstance(v, dict):
                flat.update(flatten(v, prefix + k + "_"))
            else:
                flat[prefix + k] = v
        return flat

    flat = flatten(nested)
    tmp = tempfile.mkdtemp()
    flat.update({
        "gpu_ids": "-1", "data_crop_size": size, "data_load_size": size, "train_batch_size": batch,
        "dataroot": tmp, "checkpoints_dir": tmp, "name": "golden",
        "alg_diffusion_cond_embed": "", "G_ngf": 32, "G_unet_mha_channel_mults": [1, 2],
        "G_unet_mha_res_blocks": [1, 1], "G_unet_mha_attn_res": [2], "G_unet_mha_num_head_channels": 16,
        "train_optim": "adamw", "train_G_lr": 1e-3, "train_G_ema": True, "train_G_ema_beta": 0.9,
        "train_optim_weight_decay": 0.01, "output_no_html": True,
    })
    flat.update(extra or {})
    opt = TrainOptions().parse_json(flat, save_config=False)
    opt.use_cuda = False
    opt.optim = ref_train.optim
    opt.jg_dir = ref_stubs.REFERENCE_ROOT
    opt.total_iters = 0
    opt.num_test_images = 0
    torch.manual_seed(5)
    model = create_model(opt, 0)
    model.setup(opt)
    model.use_temporal = False
    return model, opt


def plumbing_golden(name):
    """The reference's own control path: options -> create_model -> optimize_parameters (train.py:183-281)."""
    size, batch = 32, 2
    model, opt = create_reference_model(size, batch)
    cfg = O.UNetCfg(**SMALL)
    wseed = 21
    params = O.init_params(cfg, wseed)
    assert [(k, tuple(v.shape)) for k, v in model.netG_A.named_parameters()] == list(
        O.generator_param_shapes(cfg).items())
    model.netG_A.load_state_dict(params, strict=False)
    losses = []
    for step in range(2):
        data = O.synthetic_batch(batch, size, 100 + step)
        model.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"],
                         "B_label_cls": torch.zeros(batch, dtype=torch.long), "A_img_paths": ["a"] * batch})
        torch.manual_seed(1000 + step)
        model.optimize_parameters()
        losses.append(float(model.loss_G_tot))
    stat = lambda net: {k: (float(p.double().sum()), float(p.double().norm())) for k, p in net.named_parameters()}
    out = {
        "cfg": SMALL, "batch": batch, "size": size, "wseed": wseed, "data_seeds": [100, 101],
        "rng_seeds": [1000, 1001], "torch_version": str(torch.__version__),
        "optim": dict(lr=1e-3, beta1=opt.train_beta1, beta2=opt.train_beta2, eps=opt.train_optim_eps,
                      weight_decay=0.01, kind="adamw", ema_beta=0.9, iter_size=opt.train_iter_size),
        "lambda_G": opt.alg_diffusion_lambda_G, "minsnr": bool(opt.alg_palette_minsnr),
        "losses": losses,
        "param_stats": stat(model.netG_A),
        "ema_stats": stat(model.netG_A_ema),
        "sample_param": model.netG_A.state_dict()["denoise_fn.model.middle_block.1.qkv.weight"].detach().clone(),
    }
    torch.save(out, os.path.join(GOLDEN, name))
    print(name, "losses", losses)


def sampling_golden(name, cfgd, batch, wseed, dseed, rseed, sample_num):
    """DiffusionGenerator.restoration (DDPM, n_timestep_test steps) of the unmodified reference on CPU; the random
    draws (initial y_t, one randn_like per step with t > 0) are recorded by replaying the seed."""
    cfg = O.UNetCfg(**cfgd)
    net = build_reference_generator(cfg)
    params = O.init_params(cfg, wseed)
    net.load_state_dict(params, strict=False)
    data = O.synthetic_batch(batch, cfg.image_size, dseed)
    torch.manual_seed(rseed)
    with torch.no_grad():
        y, ret = net.restoration(data["cond"], y_t=None, y_0=data["gt"], mask=data["mask"], sample_num=sample_num)
    # the same draws, in the reference's order
    torch.manual_seed(rseed)
    y_t0 = torch.randn_like(data["gt"])
    noises = {i: torch.randn_like(data["gt"]) for i in reversed(range(1, cfg.n_timestep_test))}
    with torch.no_grad():
        yo, reto = O.restoration_ddpm(params, data["cond"], y_t0, data["gt"], data["mask"], noises, cfg, sample_num)
    print(name, "oracle vs reference: y rel max err %.2e, ret_arr %.2e" % (
        float((yo - y).abs().max() / y.abs().max()), float((reto - ret).abs().max() / ret.abs().max())))
    # DDIM: 5 steps, eta 0.5, from the same initial y_t (the reference's y_t default has y_cond's shape)
    net.sampling_method = "ddim"
    torch.manual_seed(rseed + 1)
    with torch.no_grad():
        yd, retd = net.restoration(data["cond"], y_t=y_t0.clone(), y_0=data["gt"], mask=data["mask"],
                                   sample_num=sample_num, ddim_num_steps=5, ddim_eta=0.5)
        ydo, retdo = O.restoration_ddim(params, data["cond"], y_t0.clone(), data["gt"], data["mask"], cfg, sample_num,
                                        num_steps=5, eta=0.5)
    print(name, "DDIM oracle vs reference: y rel max err %.2e, ret_arr %.2e" % (
        float((ydo - yd).abs().max() / yd.abs().max()), float((retdo - retd).abs().max() / retd.abs().max())))
    torch.save({"cfg": cfgd, "batch": batch, "wseed": wseed, "dseed": dseed, "rseed": rseed, "sample_num": sample_num,
                "torch_version": str(torch.__version__), "y": y.clone(), "ret_arr": ret.clone(),
                "ddim_steps": 5, "ddim_eta": 0.5, "y_ddim": yd.clone(), "ret_arr_ddim": retd.clone()},
               os.path.join(GOLDEN, name))


def main():
    ref_stubs.install()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    module_golden("palette_small.pt", SMALL, batch=2, wseed=7, dseed=11, rseed=123, full_grads=True)
    module_golden("palette_mid.pt", MID, batch=2, wseed=8, dseed=12, rseed=124, full_grads=False)
    plumbing_golden("palette_plumbing.pt")
    sampling_golden("palette_sampling.pt", dict(SMALL, n_timestep_test=12), batch=2, wseed=7, dseed=11, rseed=321,
                    sample_num=3)


if __name__ == "__main__":
    main()
