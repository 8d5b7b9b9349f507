"""GPU parity of the video UNet path: MotionModule kernels against plain
PyTorch fp32 references, and `nets_vid.UNetVid` against the reference's golden vectors / the bf16-emulating oracle."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import kernels
    return kernels


def nhwc(K, t):  # [N, C, H, W] fp32 cpu -> NHWC bf16 cuda
    return K.nchw_to_nhwc(t.cuda())


@pytest.mark.parametrize("case", [(6, 5, 64, 3, True), (4, 8, 512, 2, True), (3, 7, 1024, 1, False), (8, 4, 96, 4, True),
                                  (5, 3, 128, 5, True), (2, 4, 256, 2, False), (8, 32, 64, 8, True)])
def test_layernorm_pe_fwd_bwd(K, case):
    n, hw, c, frames, use_pe = case
    g = torch.Generator().manual_seed(c + n)
    x = bf16_round(torch.randn(n, c, hw, hw, generator=g) * 1.3 + 0.2)
    gamma = 1 + 0.2 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    pe = torch.randn(frames, c, generator=g) if use_pe else None
    dy = bf16_round(torch.randn(n, c, hw, hw, generator=g))
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr.permute(0, 2, 3, 1), (c,), gr, br, eps=1e-5)
    if use_pe:
        ref = ref + pe[torch.arange(n) % frames][:, None, None, :]
    ref.backward(dy.permute(0, 2, 3, 1))
    xd = nhwc(K, x)
    y, stats = K.layernorm_fwd(xd, gamma.cuda(), beta.cuda(), pe=None if pe is None else pe.cuda(), frames=frames)
    assert rel(y.float().cpu(), ref.detach()) < 6e-3
    dx, dg, db = K.layernorm_bwd(xd, nhwc(K, dy), gamma.cuda(), stats)
    assert rel(dx.float().cpu(), xr.grad.permute(0, 2, 3, 1)) < 1e-2
    assert rel(dg, gr.grad) < 3e-3 and rel(db, br.grad) < 3e-3
    # residual hand-through: dx + addend in the same pass, column sums of the result on the side
    add = bf16_round(torch.randn(n, c, hw, hw, generator=g))
    cs = torch.empty(c, device="cuda")
    dx2, dg2, db2 = K.layernorm_bwd(xd, nhwc(K, dy), gamma.cuda(), stats, addend=nhwc(K, add), colsum=cs)
    want = dx.float() + nhwc(K, add).float()
    assert rel(dx2.float().cpu(), want.cpu()) < 6e-3
    assert rel(cs, want.reshape(-1, c).sum(0)) < 3e-3
    assert rel(dg2, gr.grad) < 3e-3 and rel(db2, br.grad) < 3e-3


@pytest.mark.parametrize("case", [(2, 8, 6, 64, 8), (1, 4, 5, 128, 8), (3, 3, 4, 512, 8), (2, 1, 3, 64, 4)])
def test_temporal_attention_fwd_bwd(K, case):
    b, frames, hw, c, heads = case
    ch = c // heads
    g = torch.Generator().manual_seed(c + frames)
    qkv = bf16_round(torch.randn(b * frames, 3 * c, hw, hw, generator=g))
    do = bf16_round(torch.randn(b * frames, c, hw, hw, generator=g))
    qr = qkv.clone().requires_grad_(True)
    # reference: VersatileAttention on the "(b d) f c" view
    t = qr.permute(0, 2, 3, 1).reshape(b, frames, hw * hw, 3, heads, ch)  # b f d (q|k|v) h c
    q, k, v = [t[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3)]    # b d h f c
    p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * ch ** -0.5, dim=-1)
    o = torch.matmul(p, v).permute(0, 3, 1, 2, 4).reshape(b * frames, hw, hw, c)
    o.backward(do.permute(0, 2, 3, 1))
    qd = nhwc(K, qkv)
    out = K.temporal_attn_fwd(qd, frames, heads)
    assert rel(out.float().cpu(), o.detach()) < 6e-3
    dqkv = K.temporal_attn_bwd(qd, nhwc(K, do), frames, heads)
    assert rel(dqkv.float().cpu(), qr.grad.permute(0, 2, 3, 1)) < 1e-2


def test_geglu_fwd_bwd(K):
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(3, 2 * 256, 6, 5, generator=g) * 1.5)
    dy = bf16_round(torch.randn(3, 256, 6, 5, generator=g))
    xr = x.clone().requires_grad_(True)
    a, gate = xr.permute(0, 2, 3, 1).chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    ref.backward(dy.permute(0, 2, 3, 1))
    xd = nhwc(K, x)
    assert rel(K.geglu_fwd(xd).float().cpu(), ref.detach()) < 6e-3
    assert rel(K.geglu_bwd(xd, nhwc(K, dy)).float().cpu(), xr.grad.permute(0, 2, 3, 1)) < 6e-3


@pytest.mark.parametrize("case", [(2, 64, 4, 32), (1, 256, 2, 64)])
def test_attention_new_order_layout(K, case):
    """QKVAttention (q | k | v chunks): the layout of the video UNet's spatial attention blocks."""
    from oracle.vid_oracle import qkv_attention_new
    n, t, heads, ch = case
    c = heads * ch
    side = int(math.isqrt(t))
    g = torch.Generator().manual_seed(t + ch)
    qkv = bf16_round(torch.randn(n, 3 * c, t, generator=g))
    qr = qkv.clone().requires_grad_(True)
    ref = qkv_attention_new(qr, heads)
    do = bf16_round(torch.randn(ref.shape, generator=g))
    ref.backward(do)
    qd = K.nchw_to_nhwc(qkv.reshape(n, 3 * c, side, side).cuda())
    out, lse = K.attn_fwd(qd, heads, ch, layout=1)
    assert rel(K.nhwc_to_nchw(out).reshape(n, c, t), ref.detach()) < 1e-2
    dqkv = K.attn_bwd(qd, out, K.nchw_to_nhwc(do.reshape(n, c, side, side).cuda()), lse, heads, ch, layout=1)
    assert rel(K.nhwc_to_nchw(dqkv).reshape(n, 3 * c, t), qr.grad) < 1.5e-2


def _build(cfg, params):
    from joligen_b200 import nets_vid
    net = nets_vid.UNetVid(image_size=cfg.image_size, in_channel=cfg.in_channel, inner_channel=cfg.inner_channel,
                           out_channel=cfg.out_channel, res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res),
                           tanh=False, n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test,
                           norm="groupnorm", group_norm_size=cfg.group_norm_size, cond_embed_dim=cfg.cond_embed_dim,
                           channel_mults=cfg.channel_mults, num_heads=cfg.num_heads,
                           num_head_channels=cfg.num_head_channels, max_sequence_length=cfg.max_sequence_length,
                           num_attention_heads=cfg.num_attention_heads,
                           num_transformer_blocks=cfg.num_transformer_blocks)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all(m.endswith("pos_encoder.pe") for m in missing)
    return net.cuda()


def test_unetvid_forward_backward_vs_reference_golden(golden_dir):
    """End to end against the unmodified reference's fp32 vectors (bf16-storage floor, see test_gpu_palette.py) and,
    tighter, against the oracle in bf16-storage emulation."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import palette_oracle as O
    from oracle import vid_oracle as V
    from oracle.gen_golden_vid import inputs
    gold = torch.load(os.path.join(golden_dir, "vid_small.pt"))
    cfg = V.VidCfg(**gold["cfg"])
    params = V.init_params_from_shapes(gold["shapes"], gold["wseed"])
    x, emb, gy = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    net = _build(cfg, params)
    y = net(x.cuda(), emb.cuda())
    (y * gy.cuda()).sum().backward()
    # bf16-emulating oracle with the same inputs
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.EMULATE_BF16[0] = True
    try:
        yo = V.unet_vid_forward(V.add_buffers(leaves, cfg), x, emb, cfg)
        (yo * gy).sum().backward()
    finally:
        O.EMULATE_BF16[0] = False
    e_gold, e_emul = rel(y, gold["y"]), rel(y, yo)
    emul_floor = rel(yo, gold["y"])
    assert e_gold < max(3e-2, 2 * emul_floor), (e_gold, emul_floor)
    assert e_emul < max(3e-2, 2 * emul_floor), (e_emul, emul_floor)
    named = dict(net.named_parameters())
    scale = max(g["l2"] for g in gold["grads"].values())
    worst = 0.0
    for k, g in gold["grads"].items():
        mine = named[k].grad.detach().cpu().double()
        ref = leaves[k].grad.double()
        floor = max(float(ref.norm()), 2e-2 * scale)
        err = float((mine - ref).norm()) / floor
        emu = abs(float(ref.norm()) - g["l2"]) / max(g["l2"], 2e-2 * scale)  # emulation vs fp32 reference
        worst = max(worst, err)
        assert err < max(8e-2, 3 * emu), (k, err, emu)
    assert worst < 0.2


def test_refattn_unet_forward_backward_vs_reference_golden(golden_dir):
    """Row a-16: UNetGeneratorRefAttn (two UNets, attention on own and reference keys/values) end to end against the
    unmodified reference's fp32 vectors and the oracle in bf16-storage emulation."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import palette_oracle as O
    from oracle import ref_oracle as R
    from oracle.gen_golden_ref import inputs
    from oracle.vid_oracle import init_params_from_shapes
    from test_ref_oracle import build_b200  # tests/ is on sys.path (pytest prepend import mode)
    gold = torch.load(os.path.join(golden_dir, "refattn_small.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    x, ref, emb, gy = inputs(cfg, gold["batch"], gold["dseed"])
    net = build_b200(cfg)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not missing and not unexpected
    net = net.cuda()
    y = net(x.cuda(), emb.cuda(), ref.cuda())
    (y * gy.cuda()).sum().backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.EMULATE_BF16[0] = True
    try:
        yo = R.unet_ref_forward(leaves, x, emb, ref, cfg)
        (yo * gy).sum().backward()
    finally:
        O.EMULATE_BF16[0] = False
    emul_floor = rel(yo, gold["y"])
    assert rel(y, gold["y"]) < max(3e-2, 2 * emul_floor), (rel(y, gold["y"]), emul_floor)
    assert rel(y, yo) < max(3e-2, 2 * emul_floor), (rel(y, yo), emul_floor)
    named = dict(net.named_parameters())
    scale = max(g["l2"] for g in gold["grads"].values())
    for k, g in gold["grads"].items():
        mine = named[k].grad.detach().cpu().double()
        refg = leaves[k].grad.double()
        err = float((mine - refg).norm()) / max(float(refg.norm()), 2e-2 * scale)
        emu = abs(float(refg.norm()) - g["l2"]) / max(g["l2"], 2e-2 * scale)
        assert err < max(8e-2, 3 * emu), (k, err, emu)


def test_refattn_generator_trainer_and_samplers_vs_reference_golden(golden_dir):
    """cfg 4 end to end: DiffusionGenerator(PaletteDenoiseFn(UNetGeneratorRefAttn)) with the dataloader's reference
    image — loss and gradients against the unmodified reference / the bf16-emulating oracle, the same batch through
    PaletteTrainer (data["ref_A"]), and the DDPM / DDIM samplers with the reference image."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from joligen_b200.trainer import PaletteTrainer
    from oracle import palette_oracle as O
    from oracle import ref_oracle as R
    from oracle.gen_golden_ref import generator_inputs
    from oracle.vid_oracle import init_params_from_shapes
    from test_ref_oracle import build_b200
    gold = torch.load(os.path.join(golden_dir, "refattn_generator.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    gt, cond, mask, noise, ref = generator_inputs(cfg, gold["batch"], gold["dseed"])

    def build():
        g = nets.DiffusionGenerator(nets.PaletteDenoiseFn(build_b200(cfg), cfg.cond_embed_dim),
                                    image_size=cfg.image_size, G_ngf=cfg.inner_channel)
        missing, unexpected = g.load_state_dict(params, strict=False)
        assert not unexpected and not [m for m in missing if "gammas" not in m and "posterior" not in m]
        return g.cuda()

    g = build()
    assert g.denoise_fn.model_nargs == 3
    t, u = gold["t"].cuda(), gold["u"].cuda()
    n_out, nh, _ = g(gt.cuda(), cond.cuda(), mask.cuda(), noise.cuda(), cls=None, ref=ref.cuda(), t=t, u=u)
    assert torch.equal(n_out.cpu(), noise)
    loss = g.forward_loss(gt.cuda(), cond.cuda(), mask.cuda(), noise=noise.cuda(), t=t, u=u, ref=ref.cuda())
    loss.backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    mb = torch.clamp(mask, min=0, max=1)
    O.EMULATE_BF16[0] = True
    try:
        _, nho, _ = O.diffusion_forward(leaves, gt, cond, mask, noise, gold["t"], gold["u"], cfg, unet=R.denoiser(ref))
        lo = torch.nn.MSELoss()(mb * noise, mb * nho)
        lo.backward()
    finally:
        O.EMULATE_BF16[0] = False
    floor = rel(nho, gold["noise_hat"])
    assert rel(nh, gold["noise_hat"]) < max(3e-2, 2 * floor), (rel(nh, gold["noise_hat"]), floor)
    assert abs(float(loss) - gold["loss"]) < 2e-2 * gold["loss"], (float(loss), gold["loss"])
    assert abs(float(loss) - float(lo)) < 2e-2 * gold["loss"]
    named = dict(g.named_parameters())
    scale = max(x["l2"] for x in gold["grads"].values())
    for k, x in gold["grads"].items():
        mine = named[k].grad.detach().cpu().double()
        refg = leaves[k].grad.double()
        err = float((mine - refg).norm()) / max(float(refg.norm()), 2e-2 * scale)
        emu = abs(float(refg.norm()) - x["l2"]) / max(x["l2"], 2e-2 * scale)
        assert err < max(8e-2, 3 * emu), (k, err, emu)
    # trainer: first loss = the generator's; a batch without ref_A is refused
    net2 = build()
    tr = PaletteTrainer(net2, lr=1e-4, optim="adamw", ema=True, ema_beta=0.9, device="cuda")
    with pytest.raises(RuntimeError):
        tr.set_input({"A": cond, "B": gt, "B_label_mask": mask})
    tr.set_input({"A": cond, "B": gt, "B_label_mask": mask, "ref_A": ref})
    l1 = float(tr.optimize_parameters(noise=noise.cuda(), t=t, u=u))
    assert abs(l1 - float(loss)) < 1e-3 * abs(float(loss)), (l1, float(loss))
    changed = sum(1 for k, p in net2.named_parameters() if not torch.equal(p.detach().cpu(), params[k]))
    assert changed == len(gold["shapes"]), (changed, len(gold["shapes"]))
    # samplers with the reference image
    torch.manual_seed(gold["rseed"] + 1)
    y_t0 = torch.randn_like(gt)
    noises = {i: torch.randn_like(gt) for i in reversed(range(1, cfg.n_timestep_test))}
    y, ret = g.restoration_ddpm(cond.cuda(), y_t=y_t0.cuda(), y_0=gt.cuda(), mask=mask.cuda(), ref=ref.cuda(),
                                sample_num=gold["sample_num"], noise_fn=lambda i, shape: noises[i].cuda())
    g.sampling_method = "ddim"
    yd, retd = g.restoration(cond.cuda(), y_t=y_t0.cuda(), y_0=gt.cuda(), mask=mask.cuda(), ref=ref.cuda(),
                             sample_num=gold["sample_num"], ddim_num_steps=gold["ddim_steps"], ddim_eta=gold["ddim_eta"])
    O.EMULATE_BF16[0] = True
    try:
        with torch.no_grad():
            yo, reto = O.restoration_ddpm(params, cond, y_t0, gt, mask, noises, cfg, gold["sample_num"],
                                          unet=R.denoiser(ref))
            ydo, retdo = O.restoration_ddim(params, cond, y_t0.clone(), gt, mask, cfg, gold["sample_num"],
                                            num_steps=gold["ddim_steps"], eta=gold["ddim_eta"], unet=R.denoiser(ref))
    finally:
        O.EMULATE_BF16[0] = False
    m = mask.clamp(0, 1).bool().expand_as(gt)
    assert torch.equal(y.cpu()[~m], gt[~m]) and torch.equal(yd.cpu()[~m], gt[~m])
    l2 = lambda a, b: float((a.float().cpu() - b.float().cpu()).norm() / b.float().norm())  # noqa: E731
    assert ret.shape == gold["ret_arr"].shape and retd.shape == gold["ret_arr_ddim"].shape
    assert l2(y, gold["y"]) < max(3e-2, 2.5 * l2(yo, gold["y"])), (l2(y, gold["y"]), l2(yo, gold["y"]))
    assert l2(yd, gold["y_ddim"]) < max(3e-2, 2.5 * l2(ydo, gold["y_ddim"])), (l2(yd, gold["y_ddim"]),
                                                                                l2(ydo, gold["y_ddim"]))


def test_video_generator_and_trainer_vs_reference_golden(golden_dir):
    """cfg 5 end to end: DiffusionGenerator(PaletteDenoiseFn(UNetVid)) on a clip — loss and all gradients against the
    unmodified reference's vectors / the bf16-emulating oracle; then the same clip through PaletteTrainer (flat
    buffers, batched pack / unpack, fused AdamW+EMA): the first step's loss is the generator's and every parameter
    moves by the first Adam step -lr * g / (|g| + eps) computed from the gradients checked above."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import nets
    from joligen_b200.trainer import PaletteTrainer
    from oracle import palette_oracle as O
    from oracle import vid_oracle as V
    from oracle.gen_golden_vid import generator_inputs
    gold = torch.load(os.path.join(golden_dir, "vid_generator.pt"))
    cfg = V.VidCfg(**gold["cfg"])
    params = V.init_params_from_shapes(gold["shapes"], gold["wseed"])
    gt, cond, mask, noise = generator_inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])

    def build():
        unet = _build(cfg, {k[len("denoise_fn.model."):]: v for k, v in params.items()
                            if k.startswith("denoise_fn.model.")}).cpu()
        g = nets.DiffusionGenerator(nets.PaletteDenoiseFn(unet, cfg.cond_embed_dim), image_size=cfg.image_size,
                                    G_ngf=cfg.inner_channel)
        missing, unexpected = g.load_state_dict(params, strict=False)
        assert not unexpected
        return g.cuda()

    g = build()
    t, u = gold["t"].cuda(), gold["u"].cuda()
    loss = g.forward_loss(gt.cuda(), cond.cuda(), mask.cuda(), noise=noise.cuda(), t=t, u=u)
    loss.backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.EMULATE_BF16[0] = True
    try:
        _, nh = V.diffusion_forward_vid(V.add_buffers(leaves, cfg), gt, cond, mask, noise, gold["t"], gold["u"], cfg)
        mb = torch.clamp(mask, min=0, max=1)
        lo = torch.nn.MSELoss()(mb * noise, mb * nh)
        lo.backward()
    finally:
        O.EMULATE_BF16[0] = False
    assert abs(float(loss) - gold["loss"]) < 2e-2 * gold["loss"], (float(loss), gold["loss"])
    assert abs(float(loss) - float(lo)) < 2e-2 * gold["loss"]
    named = dict(g.named_parameters())
    scale = max(x["l2"] for x in gold["grads"].values())
    for k, x in gold["grads"].items():
        mine = named[k].grad.detach().cpu().double()
        refg = leaves[k].grad.double()
        err = float((mine - refg).norm()) / max(float(refg.norm()), 2e-2 * scale)
        emu = abs(float(refg.norm()) - x["l2"]) / max(x["l2"], 2e-2 * scale)
        assert err < max(8e-2, 3 * emu), (k, err, emu)
    # trainer on the clip
    lr = 1e-3
    net2 = build()
    before = {k: v.detach().clone() for k, v in net2.named_parameters()}
    tr = PaletteTrainer(net2, lr=lr, optim="adam", ema=True, ema_beta=0.9, device="cuda")
    tr.set_input({"A": cond, "B": gt, "B_label_mask": mask})
    l1 = float(tr.optimize_parameters(noise=noise.cuda(), t=t, u=u))
    assert abs(l1 - float(loss)) < 1e-3 * abs(float(loss)), (l1, float(loss))
    moved = 0
    for k, pnew in net2.named_parameters():
        gk = named[k].grad
        step = pnew.detach() - before[k]
        if gk is None:
            assert float(step.abs().max()) == 0.0, k
            continue
        # away from zero the first Adam step is -lr * sign(g); small elements take the sign of the bf16 / summation
        # order noise (two runs of the same backward differ by ~1.5% of each tensor's norm)
        big = (gk.abs() > 0.25 * gk.abs().max()) & (gk.abs() > 1e-5)
        if not bool(big.any()):
            continue
        err = float((step[big] + lr * torch.sign(gk[big])).abs().max()) / lr
        assert err < 0.05, (k, err)
        moved += 1
    assert moved > 100, moved
