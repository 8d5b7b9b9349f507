"""The CPU oracle (oracle/palette_oracle.py) against golden vectors produced by the unmodified
reference (oracle/gen_golden.py).  fp32 on both sides: tolerance 1e-4 relative to the tensor scale
(different summation order between nn.Module and functional paths only)."""
import os

import pytest
import torch

from oracle import palette_oracle as O


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)


def _forward_backward(gold):
    cfg = O.UNetCfg(**gold["cfg"])
    params = O.init_params(cfg, gold["wseed"])
    data = O.synthetic_batch(gold["batch"], cfg.image_size, gold["dseed"])
    torch.manual_seed(gold["rseed"])
    t, u = O.sample_t_gamma(cfg, gold["batch"])
    noise = torch.randn_like(data["gt"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    _, noise_hat, w = O.diffusion_forward(leaves, data["gt"], data["cond"], data["mask"], noise, t, u, cfg)
    loss = O.palette_loss(noise, noise_hat, data["mask"])
    loss.backward()
    return cfg, t, u, noise, noise_hat, w, loss, leaves


@pytest.mark.parametrize("name", ["palette_small.pt", "palette_mid.pt"])
def test_oracle_matches_reference_forward_backward(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, name))
    cfg, t, u, noise, noise_hat, w, loss, leaves = _forward_backward(gold)
    assert torch.equal(t, gold["t"])  # index draws: bit exact
    assert torch.equal(u, gold["u"])
    assert abs(float(noise.double().sum()) - gold["noise_sum"]) < 1e-6 * noise.numel()
    assert _rel(noise_hat.detach(), gold["noise_hat"]) < 1e-4
    assert _rel(w, gold["min_snr_w"]) < 1e-6
    assert abs(float(loss) - gold["loss"]) < 1e-5 * abs(gold["loss"])
    for k, (gsum, gnorm) in gold["grad_stats"].items():
        g = leaves[k].grad
        assert abs(float(g.double().norm()) - gnorm) <= 2e-4 * gnorm + 1e-9, k
    if "grads" in gold:
        for k, gref in gold["grads"].items():
            assert _rel(leaves[k].grad, gref) < 2e-4, k


def test_oracle_matches_reference_train_steps(golden_dir):
    """Two optimize_parameters() of the full joliGEN plumbing (AdamW + weight decay + EMA)."""
    gold = torch.load(os.path.join(golden_dir, "palette_plumbing.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    oc = O.OptimCfg(**gold["optim"])
    state = O.TrainState(params=O.init_params(cfg, gold["wseed"]))
    for step in range(2):
        data = O.synthetic_batch(gold["batch"], gold["size"], gold["data_seeds"][step])
        torch.manual_seed(gold["rng_seeds"][step])
        t, u = O.sample_t_gamma(cfg, gold["batch"])
        noise = torch.randn_like(data["gt"])
        loss, _, _ = O.train_step(state, cfg, oc, data["gt"], data["cond"], data["mask"], noise, t, u,
                                  lambda_G=gold["lambda_G"], use_minsnr=gold["minsnr"])
        assert abs(float(loss) - gold["losses"][step]) < 2e-5 * abs(gold["losses"][step]), step
    for k, (s, n) in gold["param_stats"].items():
        assert abs(float(state.params[k].double().norm()) - n) <= 1e-5 * n + 1e-9, k
    for k, (s, n) in gold["ema_stats"].items():
        assert abs(float(state.ema[k].double().norm()) - n) <= 1e-5 * n + 1e-9, k
    key = "denoise_fn.model.middle_block.1.qkv.weight"
    assert _rel(state.params[key], gold["sample_param"]) < 1e-4


def test_param_count_of_headline_config():
    """the config-2 Palette UNet has 59.35 M parameters in 338 state entries."""
    cfg = O.UNetCfg()
    shapes = O.generator_param_shapes(cfg)
    n = sum(int(torch.Size(s).numel()) for s in shapes.values())
    assert abs(n - 59.35e6) < 0.05e6
    assert len(shapes) + 14 == 338


def test_gan_oracle_matches_reference(golden_dir):
    """ResnetGenerator / NLayerDiscriminator / GANLoss restatements vs the unmodified reference."""
    from oracle import gan_oracle as G
    gold = torch.load(os.path.join(golden_dir, "gan_resnet.pt"))
    shapes = G.resnet_param_shapes(3, 3, gold["ngf"], gold["n_blocks"])
    leaves = {k: v.requires_grad_(True) for k, v in G.init_from_shapes(shapes, gold["wseed"]).items()}
    feats = {}
    y = G.resnet_decoder(leaves, G.resnet_encoder(leaves, gold["x"], gold["n_blocks"], feats=feats))
    assert _rel(y.detach(), gold["y"]) < 1e-4
    for lid, f in zip(gold["feat_ids"], gold["feats"]):
        assert _rel(feats[lid].detach(), f) < 1e-4, lid
    y.backward(gold["dy"])
    for k, gref in gold["grads"].items():
        assert _rel(leaves[k].grad, gref) < 2e-4, k
    gold = torch.load(os.path.join(golden_dir, "gan_nlayerd.pt"))
    shapes = G.nlayer_d_param_shapes(3, gold["ndf"], 3)
    leaves = {k: v.requires_grad_(True) for k, v in G.init_from_shapes(shapes, gold["wseed"]).items()}
    pred = G.nlayer_discriminator(leaves, gold["x"])
    assert _rel(pred.detach(), gold["pred"]) < 1e-4
    assert abs(float(G.gan_loss(pred, True)) - gold["loss_real"]) < 1e-5
    assert abs(float(G.gan_loss(pred, False)) - gold["loss_fake"]) < 1e-5
    assert abs(float(G.gan_loss(pred, True, "projected")) - gold["hinge_real"]) < 1e-5
    assert abs(float(G.gan_loss(pred, False, "projected")) - gold["hinge_fake"]) < 1e-5
    G.gan_loss(pred, True).backward()
    for k, gref in gold["grads"].items():
        assert _rel(leaves[k].grad, gref) < 2e-4, k


def _sampling_draws(gold, cfg):
    data = O.synthetic_batch(gold["batch"], cfg.image_size, gold["dseed"])
    torch.manual_seed(gold["rseed"])
    y_t0 = torch.randn_like(data["gt"])
    noises = {i: torch.randn_like(data["gt"]) for i in reversed(range(1, cfg.n_timestep_test))}
    return data, y_t0, noises


def test_oracle_ddpm_sampler_matches_reference(golden_dir):
    """DiffusionGenerator.restoration (DDPM) of the unmodified reference."""
    gold = torch.load(os.path.join(golden_dir, "palette_sampling.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    params = O.init_params(cfg, gold["wseed"])
    data, y_t0, noises = _sampling_draws(gold, cfg)
    with torch.no_grad():
        y, ret = O.restoration_ddpm(params, data["cond"], y_t0, data["gt"], data["mask"], noises, cfg,
                                    gold["sample_num"])
    assert ret.shape == gold["ret_arr"].shape
    assert _rel(y, gold["y"]) < 1e-4 and _rel(ret, gold["ret_arr"]) < 1e-4


def test_oracle_ddim_sampler_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "palette_sampling.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    params = O.init_params(cfg, gold["wseed"])
    data, y_t0, _ = _sampling_draws(gold, cfg)
    with torch.no_grad():
        y, ret = O.restoration_ddim(params, data["cond"], y_t0, data["gt"], data["mask"], cfg, gold["sample_num"],
                                    num_steps=gold["ddim_steps"], eta=gold["ddim_eta"])
    assert _rel(y, gold["y_ddim"]) < 1e-4 and _rel(ret, gold["ret_arr_ddim"]) < 1e-4
