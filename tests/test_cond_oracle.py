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
