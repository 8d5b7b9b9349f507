"""b2b video backbone (JiTViD + B2BGenerator flow-matching forward + masked pseudo-Huber
loss), CPU side: the oracle restatement against the golden vectors of the unmodified reference
(oracle/gen_golden_jit.py)."""
import os

import torch

from oracle import jit_oracle as J
from oracle.vid_oracle import init_params_from_shapes


import pytest


@pytest.mark.parametrize("name", ["jit_small.pt", "jit_b200.pt"])
def test_jit_b2b_oracle_matches_reference(golden_dir, name):
    from oracle.gen_golden_jit import inputs
    gold = torch.load(os.path.join(golden_dir, name))
    cfg = J.JitCfg(**gold["cfg"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    gt, cond, mask, label = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    torch.manual_seed(gold["rseed"])
    t_base = torch.sigmoid(torch.randn(gold["batch"]) * 0.8 - 0.8)     # sample_t, b2b_generator.py:192-199
    e = torch.randn_like(gt)
    assert torch.equal(t_base, gold["t_base"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    sd = J.add_buffers({**leaves, **gold["frozen"]}, cfg)
    v_pred, v, x_pred = J.b2b_forward(sd, gt, mask, cond, label, t_base, e, cfg)
    assert float((x_pred - gold["x_pred"]).abs().max()) < 1e-5 * float(gold["x_pred"].abs().max())
    m = mask.bool().expand_as(gt)
    assert torch.equal(x_pred[~m], gt[~m])              # known pixels are kept exactly
    loss = J.masked_region_loss(v_pred, v, torch.clamp(mask, 0, 1))
    assert abs(float(loss) - gold["loss"]) < 1e-5 * gold["loss"]
    loss.backward()
    scale = max(g["l2"] for g in gold["grads"].values())
    for k, g in gold["grads"].items():
        mine = leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])
        assert abs(float(mine.double().norm()) - g["l2"]) < 1e-4 * max(g["l2"], 1e-3 * scale), k
        assert float((mine.flatten()[:16] - g["head"]).abs().max()) < 1e-4 * max(g["l2"], 1e-3 * scale), k


def test_b2b_restoration_oracle_matches_reference(golden_dir):
    """B2BGenerator.restoration: Heun steps + the final Euler step on the flow ODE, known pixels re-projected."""
    from oracle.gen_golden_jit import inputs
    gold = torch.load(os.path.join(golden_dir, "jit_small.pt"))
    cfg = J.JitCfg(**gold["cfg"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    gt, cond, mask, label = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    torch.manual_seed(gold["rseed"] + 1)
    init_noise = torch.randn_like(gt)
    sd = J.add_buffers({**params, **gold["frozen"]}, cfg)
    out = J.restoration(sd, gt, cond, mask, label, init_noise, cfg, steps=gold["denoise_timesteps"])
    assert float((out - gold["restored"]).abs().max()) < 1e-4 * float(gold["restored"].abs().max())
    m = mask.bool().expand_as(gt)
    assert torch.equal(out[~m], gt[~m].clamp(-1, 1))


def test_rope_tables_and_unpatchify():
    cfg = J.JitCfg(input_size=32, patch_size=8, hidden_size=96, num_heads=6, in_context_len=4)
    cos, sin = J.rope_tables(cfg, 4)
    assert cos.shape == (4 + 16, 16) and sin.shape == cos.shape
    assert torch.equal(cos[:4], torch.ones(4, 16)) and torch.equal(sin[:4], torch.zeros(4, 16))  # prefix: identity
    x = torch.randn(2, 20, 16)
    rot = x * cos + J.rotate_half(x) * sin
    assert torch.allclose(rot.norm(dim=-1), x.norm(dim=-1), atol=1e-5)      # rotations keep the norm
    img = torch.arange(2 * 3 * 16 * 16, dtype=torch.float32).reshape(2, 3, 16, 16)
    patches = img.reshape(2, 3, 2, 8, 2, 8).permute(0, 2, 4, 3, 5, 1).reshape(2, 4, 8 * 8 * 3)
    assert torch.equal(J.unpatchify(patches, 8, 3), img)


def test_b2b_train_steps_match_reference_plumbing(golden_dir):
    """BASELINE.json config 5 as written (b2b_model + vit_vid, JiTVid-B/16, 156 M parameters) through the reference's
    own control path: two optimize_parameters() with AdamW(0.9, 0.95) + EMA.  The oracle reproduces both losses and the
    parameters / EMA afterwards — including the reference's quirks that the one-channel mask is broadcast only in the
    numerator of the masked loss and that the "fixed" pos_embed is trained (set_requires_grad turns it on)."""
    from oracle import palette_oracle as O
    from oracle.gen_golden_b2b_plumbing import batch, draws
    gold = torch.load(os.path.join(golden_dir, "b2b_plumbing.pt"))
    gen = gold["gen"]
    cfg = J.JitCfg(**gold["cfg"], t_eps=gen["t_eps"], noise_scale=gen["noise_scale"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    state = O.TrainState(params={**params, **{k: v.clone() for k, v in gold["frozen"].items()}})
    label = torch.zeros(gold["batch"], dtype=torch.long)
    for step in range(2):
        data = batch(gold["data_seeds"][step])
        t_base, e = draws(gold["rng_seeds"][step], gen["P_mean"], gen["P_std"], gen["mix"])
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in state.params.items()}
        loss = J.b2b_loss(J.add_buffers(leaves, cfg), data["B"], data["B_label_mask"].float(), None, label, t_base, e,
                          cfg, kind=gold["loss_kind"], lambda_G=gold["lambda_G"],
                          masked_region_only=gold["masked_region_only"])
        assert abs(float(loss) - gold["losses"][step]) < 2e-5 * gold["losses"][step], (step, float(loss))
        loss.backward()
        with torch.no_grad():
            O.adam_update(state, {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()},
                          O.OptimCfg(**gold["optim"]))
    for k, (s, n) in gold["param_stats"].items():
        assert abs(float(state.params[k].double().norm()) - n) <= 2e-4 * n + 1e-9, k
    for k, (s, n) in gold["ema_stats"].items():
        assert abs(float(state.ema[k].double().norm()) - n) <= 2e-4 * n + 1e-9, k
