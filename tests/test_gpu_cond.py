"""Class / mask conditioning of the Palette generator on the GPU (alg_diffusion_cond_embed "class" — what
example_ddpm_mario.json ships, BASELINE config 1 — "mask" and "class_mask") against the unmodified reference's vectors
(tests/golden/palette_cond_*.pt).  The label lookups are index operations: bit exact."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_embed_rows_lookup_bit_exact_and_scatter_gradient():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import kernels as K
    g = torch.Generator().manual_seed(0)
    table = torch.randn(5, 32, generator=g)
    for mask in (torch.randint(0, 5, (3, 1, 24, 16), generator=g), torch.randint(0, 5, (3, 1, 24, 16), generator=g).float()):
        x = torch.full((3, 24, 16, 40), 3.0, dtype=torch.bfloat16, device="cuda")
        K.embed_rows(table.cuda(), mask.cuda(), x, 6)
        want = table[mask.long().squeeze(1)].to(torch.bfloat16)          # [3, 24, 16, 32]
        assert torch.equal(x[..., 6:38].cpu(), want)                      # the gather is exact
        assert float((x[..., :6].float() - 3).abs().max()) == 0 and float((x[..., 38:].float() - 3).abs().max()) == 0
        d = torch.randn(3, 24, 16, 40, generator=g).to(torch.bfloat16)
        dtable, counts = K.embed_rows_bwd(d.cuda(), mask.cuda(), 6, 5, 32)
        ref = torch.zeros(5, 32).index_add_(0, mask.long().flatten(), d[..., 6:38].float().reshape(-1, 32))
        assert torch.allclose(dtable.cpu(), ref, rtol=1e-5, atol=1e-4)
        assert torch.equal(counts.cpu(), torch.bincount(mask.long().flatten(), minlength=5).float())


@pytest.mark.parametrize("conditioning", ["class", "mask", "class_mask"])
def test_conditioned_generator_vs_reference_golden(golden_dir, conditioning):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    from oracle.gen_golden_cond import BASE, cond_batch, cond_cfg, cond_params
    g = torch.load(os.path.join(golden_dir, "palette_cond_%s.pt" % conditioning))
    cfg = cond_cfg(conditioning, g["nclasses"])
    params = cond_params(cfg, g["wseed"])
    net = nets.build_palette_generator(conditioning=conditioning, nclasses=g["nclasses"], **BASE)
    assert [(k, tuple(v.shape)) for k, v in net.named_parameters()] == list(O.generator_param_shapes(cfg).items())
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all("gammas" in m or "posterior" in m for m in missing)
    net = net.cuda()
    data = cond_batch(cfg, g["batch"], g["dseed"])
    torch.manual_seed(g["rseed"])
    t, u = O.sample_t_gamma(cfg, g["batch"])
    noise = torch.randn_like(data["gt"])
    cls = data["cls"].cuda() if "class" in conditioning else None
    _, noise_hat, _ = net(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise.cuda(), cls=cls,
                          t=t.cuda(), u=u.cuda())
    # bf16 storage vs the fp32 golden: the floor of these tiny random nets is ~1.5e-2 .. 3e-2 (tests/test_gpu_palette.py)
    assert rel_l2(noise_hat, g["noise_hat"]) < 4e-2
    # the looked-up over-long rows were renormalised in place, exactly like nn.Embedding(max_norm=1)
    sd = net.state_dict()
    for k, ref in g["tables_after"].items():
        assert torch.allclose(sd[k].cpu(), ref, atol=1e-6), k
    net.load_state_dict(params, strict=False)
    loss = net.forward_loss(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise=noise.cuda(), cls=cls,
                            t=t.cuda(), u=u.cuda())
    assert abs(float(loss) - g["loss"]) < 1e-2 * abs(g["loss"])
    loss.backward()
    # bf16-storage floor of this tiny random net: the oracle with the CUDA path's rounding points vs fp32 (8 - 11 % on
    # the first conv / the label tables / cond_embed: early-layer gradients amplify the rounding of the whole chain)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.EMULATE_BF16[0] = True
    try:
        _, nh_e, _ = O.diffusion_forward(leaves, data["gt"], data["cond"], data["mask"], noise, t, u, cfg,
                                         cls=data["cls"] if "class" in conditioning else None)
        O.palette_loss(noise, nh_e, data["mask"]).backward()
    finally:
        O.EMULATE_BF16[0] = False
    floor = 1e-3 * max(n for _, n in g["grad_stats"].values())
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        _, gnorm = g["grad_stats"][k]
        emu = abs(float(leaves[k].grad.double().norm()) - gnorm) / (gnorm + 1e-30)
        assert abs(float(p.grad.double().norm()) - gnorm) <= max(6e-2, 2.5 * emu) * gnorm + floor, k
    for k, ref in g["grads"].items():
        got = dict(net.named_parameters())[k].grad
        rn = float(ref.double().norm())
        emu = float((leaves[k].grad.double() - ref.double()).norm()) / rn
        assert float((got.cpu().double() - ref.double()).norm()) <= max(6e-2, 2.5 * emu) * rn + floor, (k, emu)


def test_trainer_with_class_conditioning_and_dropout(golden_dir):
    """PaletteTrainer with cond_embed "class": B_label_cls travels through set_input; conditioning dropout
    (palette_model.py:565-584) replaces the dropped samples' class by num_classes - 1 — checked against the same step
    with the classes edited by hand."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from joligen_b200.trainer import PaletteTrainer
    from oracle import palette_oracle as O
    from oracle.gen_golden_cond import BASE, cond_batch, cond_cfg, cond_params
    cfg = cond_cfg("class", 4)
    params = cond_params(cfg, 3)
    data = cond_batch(cfg, 4, 9)
    torch.manual_seed(1)
    t, u = O.sample_t_gamma(cfg, 4)
    noise = torch.randn_like(data["gt"])
    drop_u = torch.tensor([0.01, 0.9, 0.05, 0.5])
    losses = []
    for mode in ("dropout", "by_hand"):
        net = nets.build_palette_generator(conditioning="class", nclasses=4, **BASE)
        net.load_state_dict(params, strict=False)
        cls, mask = data["cls"].clone(), data["mask"].clone()
        if mode == "by_hand":   # the reference fills BOTH the class and the mask of a dropped sample (:571-583)
            cls[drop_u < 0.1] = 3
            mask[drop_u < 0.1] = 3
        tr = PaletteTrainer(net, lr=1e-3, device="cuda", dropout_prob=0.1 if mode == "dropout" else 0.0, num_classes=4)
        tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": mask, "B_label_cls": cls})
        losses.append(float(tr.compute_palette_loss(noise=noise.cuda(), t=t.cuda(), u=u.cuda(),
                                                    drop_u=drop_u.cuda() if mode == "dropout" else None)))
    # the same arithmetic on the same inputs; fp32 atomics (split-K, fused statistics) reorder sums between two runs
    assert abs(losses[0] - losses[1]) < 1e-3 * abs(losses[1])
    loss = tr.optimize_parameters(noise=noise.cuda(), t=t.cuda(), u=u.cuda())
    assert torch.isfinite(loss)


def test_conditioned_samplers_vs_reference_golden(golden_dir):
    """Row (f)-1 with conditioning: restoration_ddpm (8 reverse steps) and restoration_ddim (4 steps) of a class + mask
    conditioned generator against the unmodified reference's samplers with replayed draws.  Outside the mask y_0 is
    copied exactly; inside, the bf16 UNet error is fed back every step: bounded by the bf16-emulating oracle's own
    distance to the fp32 golden."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    from oracle.gen_golden_cond import BASE, cond_batch, cond_cfg, cond_params
    g = torch.load(os.path.join(golden_dir, "palette_cond_class_mask.pt"))
    s = g["sampling"]
    cfg0 = cond_cfg("class_mask", g["nclasses"])
    cfg = O.UNetCfg(in_channel=cfg0.in_channel, conditioning="class_mask", nclasses=g["nclasses"],
                    n_timestep_test=s["n_timestep_test"], **BASE)
    params = cond_params(cfg, g["wseed"])
    for k, v in g["tables_after"].items():
        params[k] = v.clone()
    data = cond_batch(cfg, g["batch"], g["dseed"])
    torch.manual_seed(s["rseed"])
    y_t0 = torch.randn_like(data["gt"])
    noises = {i: torch.randn_like(data["gt"]) for i in reversed(range(1, cfg.n_timestep_test))}
    net = nets.build_palette_generator(conditioning="class_mask", nclasses=g["nclasses"],
                                       n_timestep_test=s["n_timestep_test"], **BASE)
    net.load_state_dict(params, strict=False)
    net = net.cuda()
    y, ret = net.restoration_ddpm(data["cond"].cuda(), y_t=y_t0.cuda(), y_0=data["gt"].cuda(), mask=data["mask"].cuda(),
                                  sample_num=s["sample_num"], cls=data["cls"].cuda(),
                                  noise_fn=lambda i, shape: noises[i].cuda())
    assert ret.shape == s["ret_arr"].shape
    m = data["mask"].clamp(0, 1).bool().expand_as(data["gt"])
    assert torch.equal(y.cpu()[~m], data["gt"][~m])
    O.EMULATE_BF16[0] = True
    try:
        with torch.no_grad():
            yo, _ = O.restoration_ddpm(params, data["cond"], y_t0, data["gt"], data["mask"], noises, cfg,
                                       s["sample_num"], cls=data["cls"])
            ydo, _ = O.restoration_ddim(params, data["cond"], y_t0.clone(), data["gt"], data["mask"], cfg,
                                        s["sample_num"], num_steps=s["ddim_steps"], eta=s["ddim_eta"], cls=data["cls"])
    finally:
        O.EMULATE_BF16[0] = False
    assert rel_l2(y, s["y"]) < max(3e-2, 2.5 * rel_l2(yo, s["y"])), (rel_l2(y, s["y"]), rel_l2(yo, s["y"]))
    net.sampling_method = "ddim"
    yd, retd = net.restoration(data["cond"].cuda(), y_t=y_t0.cuda(), y_0=data["gt"].cuda(), mask=data["mask"].cuda(),
                               sample_num=s["sample_num"], cls=data["cls"].cuda(), ddim_num_steps=s["ddim_steps"],
                               ddim_eta=s["ddim_eta"])
    assert retd.shape == s["ret_arr_ddim"].shape
    assert torch.equal(yd.cpu()[~m], data["gt"][~m])
    assert rel_l2(yd, s["y_ddim"]) < max(3e-2, 2.5 * rel_l2(ydo, s["y_ddim"]))
    # an unconditioned call of a conditioned net is an error, like the reference's (cls / mask embeddings missing)
    with pytest.raises(RuntimeError):
        net.restoration(data["cond"].cuda(), y_t=y_t0.cuda(), y_0=data["gt"].cuda(), mask=data["mask"].cuda())
