"""BASELINE.json config 2 AT ITS REAL SIZE (Palette UNet ngf 64, mults 1-2-4-8, 2 ResBlocks / level, attention at
32x32, 256x256) against the oracle: the net the benchmark times, not a toy.

  (A) every block of the full-width net at its TRUE resolution (first conv, 8 encoder ResBlocks + 3 down blocks,
      middle ResBlock - attention - ResBlock, 12 decoder ResBlocks on 192 .. 1024-channel concatenations + 3 up
      blocks), batch 1, forward + backward against the bf16-storage-emulating oracle on identical inputs:
      activations / input gradients 3e-3, parameter gradients 6e-3 relative L2 (same bounds as the toy-size
      test_gpu_palette.py (A)).  These are the kernel variants and the module compositions of the timed step.
  (B) one full training step of that net (batch 2): loss against the emulating oracle and against the fp32 oracle at
      2e-2; UNet output at the bf16 chaos floor (see test_gpu_palette.py docstring); aggregate parameter gradient.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    cfg = O.UNetCfg(image_size=256)
    assert (cfg.inner_channel, tuple(cfg.channel_mults), tuple(cfg.res_blocks), tuple(cfg.attn_res)) == \
        (64, (1, 2, 4, 8), (2, 2, 2, 2), (16,))
    params = O.init_params(cfg, 2024)
    g = nets.build_palette_generator(image_size=256)
    missing, unexpected = g.load_state_dict(params, strict=False)
    assert not unexpected and all("gammas" in m or "posterior" in m for m in missing)
    return nets, O, cfg, params, g.cuda()


def rel_l2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _blocks_with_resolution(O, cfg, unet):
    inp, mid, outb = O.unet_structure(cfg)
    res = cfg.image_size
    for i, layers in enumerate(inp):
        for j, b in enumerate(layers):
            yield "denoise_fn.model.input_blocks.%d.%d" % (i, j), b, unet.input_blocks[i][j], res
            if b.kind == "res" and b.down:
                res //= 2
    for j, b in enumerate(mid):
        yield "denoise_fn.model.middle_block.%d" % j, b, unet.middle_block[j], res
    for i, layers in enumerate(outb):
        for j, b in enumerate(layers):
            yield "denoise_fn.model.output_blocks.%d.%d" % (i, j), b, unet.output_blocks[i][j], res
            if b.kind == "res" and b.up:
                res *= 2


def test_every_config2_block_at_true_resolution(env):
    nets, O, cfg, params, net = env
    from joligen_b200 import ops
    unet = net.denoise_fn.model
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(1, cfg.cond_embed_dim, generator=g)
    seen = 0
    for name, b, mod, hw in _blocks_with_resolution(O, cfg, unet):
        x = torch.randn(1, b.cin, hw, hw, generator=g).to(torch.bfloat16).float()
        ho = hw * 2 if getattr(b, "up", False) else (hw // 2 if getattr(b, "down", False) else hw)
        dy = torch.randn(1, b.cout, ho, ho, generator=g).to(torch.bfloat16).float()
        keys = [k for k in params if k.startswith(name + ".")]
        leaves = dict(params)
        for k in keys:
            leaves[k] = params[k].clone().requires_grad_(True)
        xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
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
        for p in mod.parameters():
            p.grad = None
        xd, ed = x.cuda().requires_grad_(True), emb.cuda().requires_grad_(True)
        xin = ops.to_nhwc(xd)
        y = mod.forward_nhwc(xin, ed) if b.kind == "res" else mod.forward_nhwc(xin)
        yn = ops.to_nchw(y, b.cout)
        yn.backward(dy.cuda())
        errs = {"out": rel_l2(yn, ref), "dx": rel_l2(xd.grad, xr.grad)}
        if b.kind == "res":
            errs["demb"] = rel_l2(ed.grad, er.grad)
        local = dict(mod.named_parameters())
        gscale = max(float(leaves[k].grad.double().norm()) for k in keys)
        for k in keys:
            short = k[len(name) + 1:]
            ga, gb = local[short].grad.detach().cpu().double(), leaves[k].grad.double()
            errs["d" + short] = float((ga - gb).norm()) / max(float(gb.norm()), 5e-3 * gscale)
        bad = {k: v for k, v in errs.items() if v > (3e-3 if k in ("out", "dx", "demb") else 6e-3)}
        assert not bad, (name, b.kind, hw, bad)
        seen += 1
    assert seen == 1 + 8 + 3 + 3 + 12 + 3  # conv, encoder res, down, middle, decoder res, up


def test_full_config2_train_step_vs_oracle(env):
    nets, O, cfg, params, _ = env
    from joligen_b200.trainer import PaletteTrainer
    batch = 2
    net = nets.build_palette_generator(image_size=256)
    net.load_state_dict(params, strict=False)
    tr = PaletteTrainer(net, lr=1e-4, optim="adamw", ema=True, ema_beta=0.999, device="cuda")
    data = O.synthetic_batch(batch, 256, 31)
    torch.manual_seed(77)
    t, u = O.sample_t_gamma(cfg, batch)
    noise = torch.randn_like(data["gt"])
    # UNet output and loss before the step
    with torch.no_grad():
        _, nh, _ = tr.netG_A(data["gt"].cuda(), data["cond"].cuda(), data["mask"].cuda(), noise.cuda(),
                             t=t.cuda(), u=u.cuda())
    tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"]})
    loss = tr.optimize_parameters(noise=noise.cuda(), t=t.cuda(), u=u.cuda())
    torch.cuda.synchronize()
    p0 = {k: v.clone() for k, v in params.items()}
    oc = O.OptimCfg(lr=1e-4, ema_beta=0.999)
    fp = O.TrainState(params={k: v.clone() for k, v in p0.items()})
    fp_loss, fp_hat, _ = O.train_step(fp, cfg, oc, data["gt"], data["cond"], data["mask"], noise, t, u)
    emu = O.TrainState(params={k: v.clone() for k, v in p0.items()})
    O.EMULATE_BF16[0] = True
    try:
        emu_loss, emu_hat, _ = O.train_step(emu, cfg, oc, data["gt"], data["cond"], data["mask"], noise, t, u)
    finally:
        O.EMULATE_BF16[0] = False
    assert abs(float(loss) - float(emu_loss)) < 2e-2 * abs(float(emu_loss)), (float(loss), float(emu_loss))
    assert abs(float(loss) - float(fp_loss)) < 2e-2 * abs(float(fp_loss)), (float(loss), float(fp_loss))
    # the emulating oracle's own distance to fp32 is the storage-precision floor of the whole net
    floor = rel_l2(emu_hat, fp_hat)
    assert rel_l2(nh, fp_hat) < max(3e-2, 2.5 * floor), (rel_l2(nh, fp_hat), floor)
    # one AdamW step moves every weight by ~lr: compare the aggregate update with the emulating oracle's
    sd = tr.netG_A.state_dict()
    num = den = 0.0
    for k in p0:
        upd = sd[k].cpu().double() - p0[k].double()
        upd_ref = emu.params[k].double() - p0[k].double()
        num += float((upd - upd_ref).norm()) ** 2
        den += float(upd_ref.norm()) ** 2
    assert (num / den) ** 0.5 < 0.35
