"""oracle/palette_oracle.py with PaletteDenoiseFn conditioning ("class", "mask", "class_mask") against the unmodified
reference's vectors (tests/golden/palette_cond_*.pt, oracle/gen_golden_cond.py).  CPU only."""
import os

import pytest
import torch


@pytest.mark.parametrize("conditioning", ["class", "mask", "class_mask"])
def test_conditioned_generator_oracle_matches_reference(golden_dir, conditioning):
    from oracle import palette_oracle as O
    from oracle.gen_golden_cond import cond_batch, cond_cfg, cond_params
    g = torch.load(os.path.join(golden_dir, "palette_cond_%s.pt" % conditioning))
    cfg = cond_cfg(conditioning, g["nclasses"])
    params = cond_params(cfg, g["wseed"])
    data = cond_batch(cfg, g["batch"], g["dseed"])
    torch.manual_seed(g["rseed"])
    t, u = O.sample_t_gamma(cfg, g["batch"])
    noise = torch.randn_like(data["gt"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    _, noise_hat, _ = O.diffusion_forward(leaves, data["gt"], data["cond"], data["mask"], noise, t, u, cfg,
                                          cls=data["cls"] if "class" in conditioning else None)
    assert torch.allclose(noise_hat, g["noise_hat"], atol=2e-5)
    loss = O.palette_loss(noise, noise_hat, data["mask"])
    assert abs(float(loss) - g["loss"]) < 1e-5 * abs(g["loss"])
    loss.backward()
    for k, (s, n) in g["grad_stats"].items():
        gr = leaves[k].grad
        assert abs(float(gr.double().norm()) - n) <= 1e-4 * n + 1e-7, k
    for k, ref in g["grads"].items():
        assert torch.allclose(leaves[k].grad, ref, atol=1e-5 * float(ref.abs().max()) + 1e-8), k
    # the looked-up over-long rows were renormalised in place (nn.Embedding max_norm)
    for k, ref in g["tables_after"].items():
        assert torch.allclose(leaves[k].detach(), ref, atol=1e-6), k
        assert float(ref[0].norm()) <= 1.0 + 1e-5 < float(params[k][0].norm())


def test_conditioned_samplers_match_reference(golden_dir):
    """restoration_ddpm / restoration_ddim of the unmodified reference with class + mask conditioning, random draws
    replayed (tests/golden/palette_cond_class_mask.pt["sampling"], oracle/gen_golden_cond.py)."""
    import os
    import torch
    from oracle import palette_oracle as O
    from oracle.gen_golden_cond import BASE, cond_batch, cond_cfg, cond_params
    g = torch.load(os.path.join(golden_dir, "palette_cond_class_mask.pt"))
    s = g["sampling"]
    cfg0 = cond_cfg("class_mask", g["nclasses"])
    cfg = O.UNetCfg(in_channel=cfg0.in_channel, conditioning="class_mask", nclasses=g["nclasses"],
                    n_timestep_test=s["n_timestep_test"], **BASE)
    params = cond_params(cfg, g["wseed"])
    for k, v in g["tables_after"].items():   # the tables as the training forward left them (renormalised rows)
        params[k] = v.clone()
    data = cond_batch(cfg, g["batch"], g["dseed"])
    torch.manual_seed(s["rseed"])
    y_t0 = torch.randn_like(data["gt"])
    noises = {i: torch.randn_like(data["gt"]) for i in reversed(range(1, cfg.n_timestep_test))}
    with torch.no_grad():
        y, ret = O.restoration_ddpm(params, data["cond"], y_t0, data["gt"], data["mask"], noises, cfg, s["sample_num"],
                                    cls=data["cls"])
        yd, _ = O.restoration_ddim(params, data["cond"], y_t0.clone(), data["gt"], data["mask"], cfg, s["sample_num"],
                                   num_steps=s["ddim_steps"], eta=s["ddim_eta"], cls=data["cls"])
    assert torch.allclose(y, s["y"], atol=1e-5) and torch.allclose(ret, s["ret_arr"], atol=1e-5)
    assert torch.allclose(yd, s["y_ddim"], atol=1e-5)
