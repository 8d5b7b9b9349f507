"""End-to-end GPU parity of the Palette generator (forward, loss, every parameter gradient, two
optimisation steps) against the CPU oracle and against the golden vectors produced by the
unmodified reference (tests/golden/*.pt, oracle/gen_golden.py).

bf16 activations vs the fp32 reference: tolerances are relative L2 / max-normalised 1e-2-class
numbers (north star: "within 1e-2 bf16"); deep-net gradients accumulate bf16 rounding over ~40
layers, so per-parameter gradients are checked at 5e-2 relative L2 with the aggregate at 2e-2.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    return nets, O


def build(nets, O, cfg, params):
    g = nets.build_palette_generator(image_size=cfg.image_size, inner_channel=cfg.inner_channel,
                                     res_blocks=cfg.res_blocks, attn_res=cfg.attn_res,
                                     channel_mults=cfg.channel_mults, num_head_channels=cfg.num_head_channels)
    missing, unexpected = g.load_state_dict(params, strict=False)
    assert not unexpected and all("gammas" in m or "posterior" in m for m in missing)
    return g.cuda()


def rel_l2(a, b):
    a, b = a.float().cpu().double(), b.float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", ["palette_small.pt", "palette_mid.pt"])
def test_generator_matches_reference_golden(env, golden_dir, name):
    nets, O = env
    gold = torch.load(os.path.join(golden_dir, name))
    cfg = O.UNetCfg(**gold["cfg"])
    params = O.init_params(cfg, gold["wseed"])
    net = build(nets, O, cfg, params)
    data = O.synthetic_batch(gold["batch"], cfg.image_size, gold["dseed"])
    torch.manual_seed(gold["rseed"])
    t, u = O.sample_t_gamma(cfg, gold["batch"])
    noise = torch.randn_like(data["gt"])
    assert torch.equal(t, gold["t"])
    noise_d, noise_hat, w = net(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise.cuda(),
                                t=t.cuda(), u=u.cuda())
    assert rel_l2(noise_hat.detach(), gold["noise_hat"]) < 1e-2
    assert float((noise_hat.detach().cpu() - gold["noise_hat"]).abs().max()) < 3e-2 * float(gold["noise_hat"].abs().max())
    assert rel_l2(w.detach(), gold["min_snr_w"]) < 1e-6
    loss = net.forward_loss(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise=noise.cuda(),
                            t=t.cuda(), u=u.cuda())
    assert abs(float(loss) - gold["loss"]) < 1e-2 * abs(gold["loss"])
    loss.backward()
    tot_err, tot_ref = 0.0, 0.0
    for k, p in net.named_parameters():
        gsum, gnorm = gold["grad_stats"][k]
        assert p.grad is not None, k
        n = float(p.grad.double().norm())
        assert abs(n - gnorm) <= 5e-2 * gnorm + 1e-7, (k, n, gnorm)
        if "grads" in gold:
            gref = gold["grads"][k]
            e = float((p.grad.cpu().double() - gref.double()).norm())
            assert e <= 5e-2 * gnorm + 1e-7, (k, e, gnorm)
            tot_err += e * e
            tot_ref += gnorm * gnorm
    if "grads" in gold:
        assert (tot_err / tot_ref) ** 0.5 < 2e-2


def test_train_steps_match_reference_plumbing(env, golden_dir):
    """Two optimize_parameters() (AdamW + weight decay + EMA) vs the reference's own control path."""
    nets, O = env
    from joligen_b200.trainer import PaletteTrainer
    gold = torch.load(os.path.join(golden_dir, "palette_plumbing.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    net = build(nets, O, cfg, O.init_params(cfg, gold["wseed"]))
    oc = gold["optim"]
    tr = PaletteTrainer(net, lr=oc["lr"], beta1=oc["beta1"], beta2=oc["beta2"], eps=oc["eps"],
                        weight_decay=oc["weight_decay"], optim=oc["kind"], ema=True, ema_beta=oc["ema_beta"],
                        iter_size=oc["iter_size"], lambda_G=gold["lambda_G"], use_minsnr=gold["minsnr"])
    for step in range(2):
        data = O.synthetic_batch(gold["batch"], gold["size"], gold["data_seeds"][step])
        torch.manual_seed(gold["rng_seeds"][step])
        t, u = O.sample_t_gamma(cfg, gold["batch"])
        noise = torch.randn_like(data["gt"])
        tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"]})
        loss = tr.optimize_parameters(noise=noise.cuda(), t=t.cuda(), u=u.cuda())
        assert abs(float(loss) - gold["losses"][step]) < 2e-2 * abs(gold["losses"][step]), step
    sd = net.state_dict()
    key = "denoise_fn.model.middle_block.1.qkv.weight"
    # Adam's first steps move every weight by ~lr regardless of gradient scale: compare the UPDATE
    p0 = O.init_params(cfg, gold["wseed"])[key]
    upd_ref = gold["sample_param"] - p0
    upd = sd[key].cpu() - p0
    assert rel_l2(upd, upd_ref) < 0.15
    for k, (s, n) in gold["param_stats"].items():
        assert abs(float(sd[k].double().norm()) - n) <= 2e-3 * n + 1e-6, k
    ema = tr.ema_state_dict()
    for k, (s, n) in gold["ema_stats"].items():
        assert abs(float(ema[k].double().norm()) - n) <= 2e-3 * n + 1e-6, k


def test_state_dict_roundtrip_and_modulewise_dropin(env):
    """state_dict keys equal the reference's; a single ResBlock / AttentionBlock called through the
    reference's NCHW fp32 signature matches the oracle's restatement."""
    nets, O = env
    cfg = O.UNetCfg(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,),
                    num_head_channels=16)
    params = O.init_params(cfg, 3)
    net = build(nets, O, cfg, params)
    assert [k for k, _ in net.named_parameters()] == list(O.generator_param_shapes(cfg).keys())
    sd = net.state_dict()
    for k, v in params.items():
        assert torch.equal(sd[k].cpu(), v)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 32, 16, 16, generator=g)
    emb = torch.randn(2, 32, generator=g)
    inp, mid, outb = O.unet_structure(cfg)
    name = "denoise_fn.model.input_blocks.1.0"
    ref = O.res_block(params, name, x, emb, inp[1][0], cfg)
    got = net.denoise_fn.model.input_blocks[1][0](x.cuda(), emb.cuda())
    assert rel_l2(got, ref) < 1e-2
    x2 = torch.randn(2, 64, 16, 16, generator=g)
    name = "denoise_fn.model.middle_block.1"
    ref = O.attention_block(params, name, x2, mid[1])
    got = net.denoise_fn.model.middle_block[1](x2.cuda())
    assert rel_l2(got, ref) < 1e-2
