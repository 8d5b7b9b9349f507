"""GPU parity of the b2b video backbone: the kernels of csrc/jit.cu against the oracle's
functions on identical bf16-rounded inputs, and B2BGenerator(JiTViD) forward + loss + every parameter gradient against
the golden vectors of the unmodified reference (tests/golden/jit_b200.pt, oracle/gen_golden_jit.py)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def r16(x):
    return x.to(torch.bfloat16).float()


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import ops_jit
    from oracle import jit_oracle as J
    return ops_jit, J


def tok(x):   # [N, T, C] fp32 cpu -> bf16 cuda [N, T, 1, C]
    return x.to(torch.bfloat16).cuda()[:, :, None, :].contiguous()


def test_rmsnorm_modulate_fwd_bwd(env):
    OJ, J = env
    g = torch.Generator().manual_seed(0)
    n, t, c = 3, 20, 128
    x = r16(torch.randn(n, t, c, generator=g))
    w = 1 + 0.2 * torch.randn(c, generator=g)
    mod = 0.3 * torch.randn(n, 6 * c, generator=g)
    dy = r16(torch.randn(n, t, c, generator=g))
    xr, wr, mr = x.clone().requires_grad_(True), w.clone().requires_grad_(True), mod.clone().requires_grad_(True)
    ref = J.modulate(J.rms_norm(xr, wr), mr[:, c:2 * c], mr[:, 3 * c:4 * c])
    ref.backward(dy)
    xd, wd, md = tok(x).requires_grad_(True), w.cuda().requires_grad_(True), mod.cuda().requires_grad_(True)
    y = OJ.rmsnorm_mod(xd, wd, md[:, c:2 * c], md[:, 3 * c:4 * c])
    y.backward(tok(dy))
    assert rel(y[:, :, 0], ref) < 6e-3
    assert rel(xd.grad[:, :, 0], xr.grad) < 1e-2 and rel(wd.grad, wr.grad) < 1e-2 and rel(md.grad, mr.grad) < 1e-2
    # plain RMSNorm (no modulation)
    xr2, wr2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref2 = J.rms_norm(xr2, wr2)
    ref2.backward(dy)
    xd2, wd2 = tok(x).requires_grad_(True), w.cuda().requires_grad_(True)
    y2 = OJ.rmsnorm_mod(xd2, wd2)
    y2.backward(tok(dy))
    assert rel(y2[:, :, 0], ref2) < 6e-3 and rel(xd2.grad[:, :, 0], xr2.grad) < 1e-2 and rel(wd2.grad, wr2.grad) < 1e-2


@pytest.mark.parametrize("hd", [16, 32, 64])
def test_qknorm_rope_attention_fwd_bwd(env, hd):
    """Attention.forward (vit_vid.py:205-231) between the qkv and proj Linears: q/k RMSNorm + rotary + softmax attention."""
    OJ, J = env
    from joligen_b200.nets_jit import rope_tables
    g = torch.Generator().manual_seed(hd)
    n, heads, grid, prefix = 2, 3, 4, 4
    t, d = prefix + grid * grid, heads * hd
    qkv = r16(torch.randn(n, t, 3 * d, generator=g))
    wq = 1 + 0.2 * torch.randn(hd, generator=g)
    wk = 1 + 0.2 * torch.randn(hd, generator=g)
    do = r16(torch.randn(n, t, d, generator=g))
    cos, sin = rope_tables(hd, grid, prefix, "cpu")
    cfg = J.JitCfg(input_size=grid * 8, patch_size=8, hidden_size=d, num_heads=heads, in_context_len=prefix)
    c0, s0 = J.rope_tables(cfg, prefix)
    assert torch.allclose(cos, c0) and torch.allclose(sin, s0)
    qr, wqr, wkr = qkv.clone().requires_grad_(True), wq.clone().requires_grad_(True), wk.clone().requires_grad_(True)
    q3 = qr.reshape(n, t, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q = J.rms_norm(q3[0], wqr)
    k = J.rms_norm(q3[1], wkr)
    q = q * cos + J.rotate_half(q) * sin
    k = k * cos + J.rotate_half(k) * sin
    w = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(hd), dim=-1)
    ref = (w @ q3[2]).transpose(1, 2).reshape(n, t, d)
    ref.backward(do)
    qd, wqd, wkd = tok(qkv).requires_grad_(True), wq.cuda().requires_grad_(True), wk.cuda().requires_grad_(True)
    qk = OJ.qknorm_rope(qd, wqd, wkd, cos.cuda(), sin.cuda(), heads)
    o = OJ.attn_small(qk, qd, heads)
    o.backward(tok(do))
    assert rel(o[:, :, 0], ref) < 1e-2
    assert rel(qd.grad[:, :, 0], qr.grad) < 2e-2
    assert rel(wqd.grad, wqr.grad) < 2e-2 and rel(wkd.grad, wkr.grad) < 2e-2


def test_swiglu_and_gated_residual(env):
    OJ, J = env
    g = torch.Generator().manual_seed(3)
    n, t, h = 2, 20, 96
    x = r16(torch.randn(n, t, 2 * h, generator=g))
    dy = r16(torch.randn(n, t, h, generator=g))
    xr = x.clone().requires_grad_(True)
    a, b = xr.chunk(2, dim=-1)
    ref = torch.nn.functional.silu(a) * b
    ref.backward(dy)
    xd = tok(x).requires_grad_(True)
    y = OJ.swiglu(xd)
    y.backward(tok(dy))
    assert rel(y[:, :, 0], ref) < 6e-3 and rel(xd.grad[:, :, 0], xr.grad) < 1e-2
    c = 128
    x = r16(torch.randn(n, t, c, generator=g))
    br = r16(torch.randn(n, t, c, generator=g))
    gate = 0.5 * torch.randn(n, 3 * c, generator=g)
    d = r16(torch.randn(n, t, c, generator=g))
    xr, brr, gr = x.clone().requires_grad_(True), br.clone().requires_grad_(True), gate.clone().requires_grad_(True)
    ref = xr + gr[:, c:2 * c].unsqueeze(1) * brr
    ref.backward(d)
    xd, bd, gd = tok(x).requires_grad_(True), tok(br).requires_grad_(True), gate.cuda().requires_grad_(True)
    out = OJ.gated_residual(xd, bd, gd[:, c:2 * c])
    out.backward(tok(d))
    assert rel(out[:, :, 0], ref) < 6e-3
    assert rel(xd.grad[:, :, 0], xr.grad) < 1e-6 and rel(bd.grad[:, :, 0], brr.grad) < 6e-3 and rel(gd.grad, gr.grad) < 1e-2


def test_b2b_generator_vs_reference_golden(env, golden_dir):
    """B2BGenerator(JiTViD) flow-matching forward, masked pseudo-Huber loss and all parameter gradients vs the unmodified
    reference (fp32) — bf16 storage: x_pred / loss at 3e-2, gradient norms at 6e-2 of the golden's."""
    OJ, J = env
    from joligen_b200 import nets_jit
    from oracle.gen_golden_jit import inputs
    from oracle.vid_oracle import init_params_from_shapes
    gold = torch.load(os.path.join(golden_dir, "jit_b200.pt"))
    cfg = J.JitCfg(**gold["cfg"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    net = nets_jit.B2BGenerator(nets_jit.JiTViD(
        input_size=cfg.input_size, patch_size=cfg.patch_size, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
        hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads, num_classes=cfg.num_classes,
        in_context_len=cfg.in_context_len, in_context_start=cfg.in_context_start, max_frames=cfg.max_frames,
        motion_num_heads=cfg.motion_num_heads, motion_num_layers=cfg.motion_num_layers), t_eps=cfg.t_eps,
        noise_scale=cfg.noise_scale)
    mine = {k: tuple(v.shape) for k, v in net.named_parameters() if v.requires_grad}
    assert mine == dict(gold["shapes"]), set(mine) ^ set(dict(gold["shapes"]))
    missing, unexpected = net.load_state_dict({**params, **gold["frozen"]}, strict=False)
    assert not unexpected and all(m.endswith("pos_encoder.pe") for m in missing), (missing, unexpected)
    net = net.cuda()
    gt, cond, mask, label = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    torch.manual_seed(gold["rseed"])
    t_base = torch.sigmoid(torch.randn(gold["batch"]) * 0.8 - 0.8)
    e = torch.randn_like(gt)
    v_pred, v, x_pred = net(gt.cuda(), mask.cuda(), cond.cuda(), label.cuda(), return_x_pred=True, t_base=t_base.cuda(),
                            e=e.cuda())
    assert rel(x_pred, gold["x_pred"]) < 3e-2
    m = mask.bool().expand_as(gt)
    assert torch.equal(x_pred.cpu()[~m], gt[~m])
    loss = net.masked_region_loss(v_pred, v, torch.clamp(mask.cuda(), 0, 1))
    assert abs(float(loss) - gold["loss"]) < 3e-2 * gold["loss"]
    loss.backward()
    scale = max(g["l2"] for g in gold["grads"].values())
    named = dict(net.named_parameters())
    bad = {}
    for k, g in gold["grads"].items():
        got = named[k].grad
        assert got is not None, k
        tol = 6e-2 * max(g["l2"], 2e-2 * scale)
        if abs(float(got.double().norm()) - g["l2"]) > tol or \
                float((got.flatten()[:16].cpu() - g["head"]).abs().max()) > tol:
            bad[k] = (float(got.double().norm()), g["l2"])
    assert not bad, bad


def test_b2b_trainer_two_steps_vs_reference_plumbing(env, golden_dir):
    """BASELINE.json config 5 as written (b2b_model + vit_vid, JiTVid-B/16, 156 M parameters): two optimize_parameters()
    of the unmodified reference (AdamW(0.9, 0.95) + EMA) with its random draws replayed — losses, parameter and EMA norms
    after the two steps (tests/golden/b2b_plumbing.pt, oracle/gen_golden_b2b_plumbing.py)."""
    OJ, J = env
    from joligen_b200 import nets_jit
    from joligen_b200.trainer_b2b import B2BTrainer
    from oracle.gen_golden_b2b_plumbing import batch, draws
    from oracle.vid_oracle import init_params_from_shapes
    gold = torch.load(os.path.join(golden_dir, "b2b_plumbing.pt"))
    gen = gold["gen"]
    c = gold["cfg"]
    net = nets_jit.B2BGenerator(nets_jit.JiTViD(**c), t_eps=gen["t_eps"], noise_scale=gen["noise_scale"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    missing, unexpected = net.load_state_dict({**params, **gold["frozen"]}, strict=False)
    assert not unexpected and all(m.endswith("pos_encoder.pe") for m in missing), (missing, unexpected)
    o = gold["optim"]
    tr = B2BTrainer(net, lr=o["lr"], beta1=o["beta1"], beta2=o["beta2"], eps=o["eps"], weight_decay=o["weight_decay"],
                    optim=o["kind"], ema=True, ema_beta=o["ema_beta"], lambda_G=gold["lambda_G"])
    for step in range(2):
        tr.set_input(batch(gold["data_seeds"][step]))
        t_base, e = draws(gold["rng_seeds"][step], gen["P_mean"], gen["P_std"], gen["mix"])
        loss = tr.optimize_parameters(t_base=t_base.cuda(), e=e.cuda())
        assert abs(float(loss) - gold["losses"][step]) < 3e-2 * gold["losses"][step], (step, float(loss))
    torch.cuda.synchronize()
    got, ema = tr.params(), tr.ema_state_dict()
    for k, (s, nrm) in gold["param_stats"].items():
        assert abs(float(got[k].double().norm()) - nrm) <= 2e-3 * nrm + 1e-6, k
    for k, (s, nrm) in gold["ema_stats"].items():
        assert abs(float(ema[k].double().norm()) - nrm) <= 2e-3 * nrm + 1e-6, k


def test_b2b_restoration_vs_reference_golden(env, golden_dir):
    """B2BGenerator.restoration (2 Heun steps + the final Euler step on the flow ODE) vs the unmodified reference; known
    pixels are re-projected exactly."""
    OJ, J = env
    from joligen_b200 import nets_jit
    from oracle.gen_golden_jit import inputs
    from oracle.vid_oracle import init_params_from_shapes
    gold = torch.load(os.path.join(golden_dir, "jit_b200.pt"))
    cfg = J.JitCfg(**gold["cfg"])
    net = nets_jit.B2BGenerator(nets_jit.JiTViD(
        input_size=cfg.input_size, patch_size=cfg.patch_size, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
        hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads, num_classes=cfg.num_classes,
        in_context_len=cfg.in_context_len, in_context_start=cfg.in_context_start, max_frames=cfg.max_frames,
        motion_num_heads=cfg.motion_num_heads, motion_num_layers=cfg.motion_num_layers), t_eps=cfg.t_eps,
        noise_scale=cfg.noise_scale)
    net.load_state_dict({**init_params_from_shapes(gold["shapes"], gold["wseed"]), **gold["frozen"]}, strict=False)
    net = net.cuda()
    gt, cond, mask, label = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    torch.manual_seed(gold["rseed"] + 1)
    init_noise = torch.randn_like(gt)
    out = net.restoration(gt.cuda(), cond.cuda(), gold["denoise_timesteps"], mask=mask.cuda(), labels=label.cuda(),
                          init_noise=init_noise.cuda())
    m = mask.bool().expand_as(gt)
    assert torch.equal(out.cpu()[~m], gt[~m].clamp(-1, 1))
    assert rel(out, gold["restored"]) < 3e-2
