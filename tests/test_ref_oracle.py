"""Reference-image-conditioned UNet, CPU side: the oracle restatement against the
golden vectors of the unmodified reference `UNetGeneratorRefAttn` (oracle/gen_golden_ref.py), and the B200 module
tree against the reference's parameter list."""
import os

import torch

from oracle import palette_oracle as O
from oracle import ref_oracle as R
from oracle.vid_oracle import init_params_from_shapes


def _load(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "refattn_small.pt"))
    return gold, O.UNetCfg(**gold["cfg"]), init_params_from_shapes(gold["shapes"], gold["wseed"])


def build_b200(cfg):
    from joligen_b200 import nets_ref
    return nets_ref.UNetGeneratorRefAttn(
        image_size=cfg.image_size, in_channel=cfg.in_channel, inner_channel=cfg.inner_channel,
        out_channel=cfg.out_channel, res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res), tanh=False,
        n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test, norm="groupnorm",
        group_norm_size=cfg.group_norm_size, cond_embed_dim=cfg.cond_embed_dim, channel_mults=cfg.channel_mults,
        num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels)


def test_ref_oracle_matches_reference_forward_backward(golden_dir):
    from oracle.gen_golden_ref import inputs
    gold, cfg, params = _load(golden_dir)
    x, ref, emb, gy = inputs(cfg, gold["batch"], gold["dseed"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y = R.unet_ref_forward(leaves, x, emb, ref, cfg)
    assert float((y - gold["y"]).abs().max()) < 1e-4 * float(gold["y"].abs().max())
    (y * gy).sum().backward()
    scale = max(g["l2"] for g in gold["grads"].values())
    for k, g in gold["grads"].items():
        mine = leaves[k].grad
        assert abs(float(mine.double().norm()) - g["l2"]) < 1e-4 * max(g["l2"], 1e-3 * scale), k
        assert float((mine.flatten()[:16] - g["head"]).abs().max()) < 1e-4 * max(g["l2"], 1e-3 * scale), k


def test_b200_refattn_unet_has_reference_parameter_list(golden_dir):
    gold, cfg, _ = _load(golden_dir)
    net = build_b200(cfg)
    assert [(k, tuple(v.shape)) for k, v in net.named_parameters()] == [(k, tuple(s)) for k, s in gold["shapes"]]


def test_ref_generator_oracle_matches_reference_loss_grads_and_samplers(golden_dir):
    """cfg 4 end to end on CPU: DiffusionGenerator forward + Palette loss + backward with the reference image, and the
    DDPM / DDIM samplers with it, against the unmodified reference (oracle/gen_golden_ref.py: generator_golden)."""
    from oracle.gen_golden_ref import generator_inputs
    gold = torch.load(os.path.join(golden_dir, "refattn_generator.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    gt, cond, mask, noise, ref = generator_inputs(cfg, gold["batch"], gold["dseed"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    _, nh, _ = O.diffusion_forward(leaves, gt, cond, mask, noise, gold["t"], gold["u"], cfg, unet=R.denoiser(ref))
    assert float((nh - gold["noise_hat"]).abs().max()) < 1e-4 * float(gold["noise_hat"].abs().max())
    mb = torch.clamp(mask, min=0, max=1)
    loss = torch.nn.MSELoss()(mb * noise, mb * nh)
    assert abs(float(loss) - gold["loss"]) < 1e-5 * gold["loss"]
    loss.backward()
    scale = max(g["l2"] for g in gold["grads"].values())
    for k, g in gold["grads"].items():
        mine = leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])
        assert abs(float(mine.double().norm()) - g["l2"]) < 1e-4 * max(g["l2"], 1e-3 * scale), k
    torch.manual_seed(gold["rseed"] + 1)
    y_t0 = torch.randn_like(gt)
    noises = {i: torch.randn_like(gt) for i in reversed(range(1, cfg.n_timestep_test))}
    with torch.no_grad():
        y, ret = O.restoration_ddpm(params, cond, y_t0, gt, mask, noises, cfg, gold["sample_num"], unet=R.denoiser(ref))
        yd, retd = O.restoration_ddim(params, cond, y_t0.clone(), gt, mask, cfg, gold["sample_num"],
                                      num_steps=gold["ddim_steps"], eta=gold["ddim_eta"], unet=R.denoiser(ref))
    for mine, key in ((y, "y"), (ret, "ret_arr"), (yd, "y_ddim"), (retd, "ret_arr_ddim")):
        assert float((mine - gold[key]).abs().max()) < 1e-4 * float(gold[key].abs().max()), key
