"""GPU parity of each sm_100a kernel against the CPU oracle's arithmetic (plain fp32 torch ops on the
same bf16-rounded operands).  Tolerances: bf16 storage => 1e-2 relative to the tensor's max (north star:
"1e-2 bf16"); fp32 paths (weight gradients, small linears, loss) 1e-3; index / mask handling bit exact.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import kernels
    from joligen_b200 import lib
    assert lib.load().jg_check_device() == 0, lib.load().jg_last_error()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return kernels


def rel(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max()) / (float(b.abs().max()) + 1e-12)


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def to_dev_nhwc(K, x_nchw_cpu):
    return K.nchw_to_nhwc(x_nchw_cpu.cuda())


CONV_CASES = [
    # n, h, w, cin, cout, k, pad
    (2, 16, 16, 64, 64, 3, 1),
    (2, 32, 32, 128, 256, 3, 1),
    (1, 32, 32, 192, 64, 3, 1),
    (2, 16, 16, 384, 128, 1, 0),
    (2, 32, 32, 6, 64, 3, 1),     # first UNet conv: 6 input channels, zero-padded to 8
    (2, 32, 32, 64, 3, 3, 1),     # last UNet conv: 3 output channels, padded to 8
    (3, 8, 8, 512, 512, 3, 1),    # several images per M tile
    (1, 30, 30, 64, 128, 4, 1),   # ragged tiles (NLayerDiscriminator stride-1 layers)
    (2, 8, 8, 512, 4096, 1, 0),   # widest layer: the video UNet's GEGLU projection (Linear as a 1x1 conv)
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(K, case):
    n, h, w, cin, cout, k, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = 0.1 * torch.randn(cout, generator=g)
    wb = bf16_round(wt)
    xr = x.clone().requires_grad_(True)
    wr = wb.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=pad)
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    ref.backward(dy)

    cout8 = (cout + 7) // 8 * 8
    x_d = to_dev_nhwc(K, x)
    wf, wd = K.pack_conv_weight(wt.cuda())
    bias_p = torch.zeros(cout8, device="cuda")
    bias_p[:cout] = b.cuda()
    y = K.conv2d_fwd(x_d, wf, bias_p, cout8, k, k, pad=pad)
    assert rel(K.nhwc_to_nchw(y, cout), ref.detach()) < 1e-2
    if cout8 != cout:
        assert float(y[..., cout:].float().abs().max()) == 0.0  # padding channels are exactly zero
    dy_d = to_dev_nhwc(K, dy)
    dx = K.conv2d_fwd(dy_d, wd, None, x_d.shape[-1], k, k, pad=k - 1 - pad)
    assert rel(K.nhwc_to_nchw(dx, cin), xr.grad) < 1e-2
    dw = K.conv2d_wgrad(x_d, dy_d, cout8, k, k, pad=pad)[:cout, :cin]
    assert rel(dw, wr.grad) < 1e-3
    db = K.bias_grad(dy_d)[:cout]
    assert rel(db, br.grad) < 1e-3


def test_conv_residual_epilogue_and_stride2(K):
    g = torch.Generator().manual_seed(5)
    x = bf16_round(torch.randn(2, 64, 32, 32, generator=g))
    wt = torch.randn(128, 64, 3, 3, generator=g) / 24.0
    r = bf16_round(torch.randn(2, 128, 32, 32, generator=g))
    wf, _ = K.pack_conv_weight(wt.cuda())
    y = K.conv2d_fwd(to_dev_nhwc(K, x), wf, None, 128, 3, 3, pad=1, residual=to_dev_nhwc(K, r),
                     res_scale=1.0 / math.sqrt(2))
    ref = F.conv2d(x, bf16_round(wt), padding=1) + r / math.sqrt(2)
    assert rel(K.nhwc_to_nchw(y), ref) < 1e-2
    y2 = K.conv2d_fwd(to_dev_nhwc(K, x), wf, None, 128, 3, 3, stride=2, pad=1)
    assert rel(K.nhwc_to_nchw(y2), F.conv2d(x, bf16_round(wt), stride=2, padding=1)) < 1e-2


STAGED_CASES = [
    # n, h, w, cin, cout, k, stride  (both accumulator layouts, the N=128/256 second pass, 1x1 and strided generics)
    (8, 16, 16, 64, 64, 3, 1), (8, 16, 16, 64, 128, 3, 1), (8, 8, 8, 128, 64, 1, 1), (8, 16, 16, 64, 64, 3, 2),
    (2, 32, 32, 256, 256, 3, 1), (8, 16, 16, 192, 64, 3, 1), (2, 16, 16, 512, 1024, 1, 1),
]


def test_staged_wgrad_and_batched_pack_match_immediate(K):
    """The trainer's path — raw split-K accumulation into persistent slots (jg_conv2d_wgrad_acc), ONE batched unpack
    that adds into the OIHW gradients and re-zeroes the slots, ONE batched bf16 weight pack — against the immediate
    per-convolution entry points, two accumulation rounds (the second checks the slots were left at zero)."""
    from joligen_b200 import lib as L
    g = torch.Generator().manual_seed(11)
    work, unpack_items, pack_items, dims = [], [], [], []
    for (n, h, w, cin, cout, k, stride) in STAGED_CASES:
        pad = (k - 1) // 2
        ho = (h + 2 * pad - k) // stride + 1
        x = torch.randn(n, h, w, cin, generator=g).bfloat16().cuda()
        dy = torch.randn(n, ho, ho, cout, generator=g).bfloat16().cuda()
        wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).cuda()
        ref = K.conv2d_wgrad(x, dy, cout, k, k, stride=stride, pad=pad)
        acc = torch.zeros(cout * cin * k * k, device="cuda")
        grad = torch.zeros(cout, cin, k, k, device="cuda")
        wf_ref, wd_ref = K.pack_conv_weight(wt)
        wf, wd = torch.zeros_like(wf_ref), torch.zeros_like(wd_ref)
        work.append((x, dy, wt, ref, acc, grad, wf_ref, wd_ref, wf, wd, (cout, k, stride, pad)))
        pack_items.append(L.PackItem(wt.data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, cin, k * k, cin, cout, 0))
        dims.append((cout, cin, k * k))
    K.pack_conv_weights_batched(K.WeightTable(pack_items, dims, torch.device("cuda")))
    for rnd in range(2):
        unpack_items = []
        for (x, dy, wt, ref, acc, grad, _, _, _, _, (cout, k, stride, pad)) in work:
            lay = K.conv2d_wgrad_acc(x, dy, cout, k, k, acc, stride=stride, pad=pad)
            unpack_items.append(L.UnpackItem(acc.data_ptr(), grad.data_ptr(), cout, x.shape[-1], k * k, lay))
        K.wgrad_unpack_batched(K.WeightTable(unpack_items, dims, torch.device("cuda")))
    for (x, dy, wt, ref, acc, grad, wf_ref, wd_ref, wf, wd, _) in work:
        assert rel(grad / 2, ref) < 1e-5
        assert float(acc.abs().max()) == 0.0
        assert torch.equal(wf, wf_ref) and torch.equal(wd, wd_ref)


GN_CASES = [
    # n, hw, c, groups, film, silu
    (2, 16, 64, 32, False, True),
    (2, 16, 192, 32, True, True),    # 6 channels per group: groups straddle 8-channel vectors
    (3, 8, 512, 32, True, True),
    (2, 16, 128, 128, False, False),  # attention InstanceNorm1d: one group per channel, no affine
    (1, 32, 1024, 32, True, True),
    (3, 256, 64, 32, True, True),
]


@pytest.mark.parametrize("case", GN_CASES)
def test_groupnorm_film_silu_fwd_bwd(K, case):
    n, hw, c, groups, use_film, silu = case
    from joligen_b200 import lib as L
    g = torch.Generator().manual_seed(c + groups)
    x = bf16_round(torch.randn(n, c, hw, hw, generator=g) * 1.5 + 0.3)
    affine = groups != c
    gamma = (1 + 0.2 * torch.randn(c, generator=g)) if affine else None
    beta = (0.1 * torch.randn(c, generator=g)) if affine else None
    film = 0.3 * torch.randn(n, 2 * c, generator=g) if use_film else None
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True) if affine else None
    br = beta.clone().requires_grad_(True) if affine else None
    fr = film.clone().requires_grad_(True) if use_film else None
    h = F.group_norm(xr, groups, gr, br, eps=1e-5)
    if use_film:
        scale, shift = torch.chunk(fr[:, :, None, None], 2, dim=1)
        h = h * (1 + scale) + shift
    ref = F.silu(h) if silu else h
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    ref.backward(dy)

    act = L.ACT_SILU if silu else L.ACT_NONE
    dev = lambda t: None if t is None else t.cuda()
    x_d = to_dev_nhwc(K, x)
    y, stats, ab = K.groupnorm_fwd(x_d, dev(gamma), dev(beta), groups, film=dev(film), act=act)
    assert rel(K.nhwc_to_nchw(y), ref.detach()) < 1e-2
    dx, dgamma, dbeta, dfilm = K.groupnorm_bwd(x_d, to_dev_nhwc(K, dy), dev(gamma), dev(beta), groups, dev(film), act,
                                               stats, ab, need_param_grads=affine, need_film_grad=use_film)
    assert rel(K.nhwc_to_nchw(dx), xr.grad) < 1e-2
    if affine:
        assert rel(dgamma, gr.grad) < 2e-3
        assert rel(dbeta, br.grad) < 2e-3
    if use_film:
        assert rel(dfilm, fr.grad) < 2e-3


def test_groupnorm_bwd_addend_colsum(K):
    """dx = GN'(dy) + addend, plus the per-channel sums of dx produced on the side (the producer conv's bias
    gradient); 96 channels: 12 vectors per row, so 4 threads of each block are idle."""
    from joligen_b200 import lib as L
    n, hw, c, groups = 5, 24, 96, 32
    g = torch.Generator().manual_seed(11)
    x = bf16_round(torch.randn(n, c, hw, hw, generator=g) * 1.2 - 0.2)
    gamma = 1 + 0.2 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    film = 0.3 * torch.randn(n, 2 * c, generator=g)
    xr = x.clone().requires_grad_(True)
    h = F.group_norm(xr, groups, gamma, beta, eps=1e-5)
    scale, shift = torch.chunk(film[:, :, None, None], 2, dim=1)
    ref = F.silu(h * (1 + scale) + shift)
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    add = bf16_round(torch.randn(ref.shape, generator=g))
    ref.backward(dy)
    want = xr.grad + add
    x_d = to_dev_nhwc(K, x)
    y, stats, ab = K.groupnorm_fwd(x_d, gamma.cuda(), beta.cuda(), groups, film=film.cuda(), act=L.ACT_SILU)
    assert rel(K.nhwc_to_nchw(y), ref.detach()) < 1e-2
    colsum = torch.full((c,), 123.0, device="cuda")
    dx, dgamma, dbeta, dfilm = K.groupnorm_bwd(x_d, to_dev_nhwc(K, dy), gamma.cuda(), beta.cuda(), groups, film.cuda(),
                                               L.ACT_SILU, stats, ab, need_film_grad=True,
                                               addend=to_dev_nhwc(K, add), colsum=colsum)
    assert rel(K.nhwc_to_nchw(dx), want) < 1e-2
    assert rel(colsum.cpu(), want.sum(dim=(0, 2, 3))) < 5e-3


@pytest.mark.parametrize("case", [(2, 256, 4, 16), (2, 64, 16, 32), (1, 1024, 2, 32), (1, 128, 2, 64)])
def test_attention_fwd_bwd(K, case):
    from oracle.palette_oracle import qkv_attention_legacy
    n, t, heads, ch = case
    c = heads * ch
    g = torch.Generator().manual_seed(t + ch)
    side = int(math.isqrt(t))
    hh, ww = (side, side) if side * side == t else (t // 8, 8)
    qkv = bf16_round(torch.randn(n, 3 * c, t, generator=g))
    qr = qkv.clone().requires_grad_(True)
    ref = qkv_attention_legacy(qr, heads)  # [n, c, t]
    do = bf16_round(torch.randn(ref.shape, generator=g))
    ref.backward(do)
    qkv_d = K.nchw_to_nhwc(qkv.reshape(n, 3 * c, hh, ww).cuda())
    out, lse = K.attn_fwd(qkv_d, heads, ch)
    assert rel(K.nhwc_to_nchw(out).reshape(n, c, t), ref.detach()) < 1e-2
    do_d = K.nchw_to_nhwc(do.reshape(n, c, hh, ww).cuda())
    dqkv = K.attn_bwd(qkv_d, out, do_d, lse, heads, ch)
    assert rel(K.nhwc_to_nchw(dqkv).reshape(n, 3 * c, t), qr.grad) < 1.5e-2


def test_attention_sliced_operands_tc_path(K):
    """The tcgen05 path with out / d_out as the SECOND channel half of a wider buffer that ends exactly at its
    allocation (AttentionBlockRef writes both attentions into the halves of one tensor, unet_generator_attn.py:1098-1130):
    the TMA boxes of the last head must end in zero fill, never past the slice (regression: illegal memory access in the
    cfg 4 step)."""
    from oracle.palette_oracle import qkv_attention_legacy
    n, t, heads, ch = 2, 256, 4, 32
    c = heads * ch
    g = torch.Generator().manual_seed(5)
    qkv = bf16_round(torch.randn(n, 3 * c, t, generator=g))
    qr = qkv.clone().requires_grad_(True)
    ref = qkv_attention_legacy(qr, heads)
    do = bf16_round(torch.randn(ref.shape, generator=g))
    ref.backward(do)
    qkv_d = K.nchw_to_nhwc(qkv.reshape(n, 3 * c, 16, 16).cuda())
    wide = torch.zeros(n, 16, 16, 2 * c, dtype=torch.bfloat16, device="cuda")
    out, lse = K.attn_fwd(qkv_d, heads, ch, out=wide[..., c:])
    assert rel(K.nhwc_to_nchw(out.contiguous()).reshape(n, c, t), ref.detach()) < 1e-2
    assert float(wide[..., :c].abs().max()) == 0.0
    dwide = torch.zeros(n, 16, 16, 2 * c, dtype=torch.bfloat16, device="cuda")
    dwide[..., c:] = K.nchw_to_nhwc(do.reshape(n, c, 16, 16).cuda())
    dqkv = K.attn_bwd(qkv_d, out, dwide[..., c:], lse, heads, ch)
    torch.cuda.synchronize()
    assert rel(K.nhwc_to_nchw(dqkv).reshape(n, 3 * c, t), qr.grad) < 1.5e-2


def test_linear_as_conv_wider_than_the_bias_staging(K):
    """nn.Linear on tokens [N, T, 1, I] with 4608 outputs (> 4096: the b2b backbone's 768 -> 6144 GEGLU projection takes
    this path): forward in output-channel chunks, one dgrad, one wgrad — vs fp32 torch on the bf16-rounded operands."""
    from joligen_b200 import nets, ops
    g = torch.Generator().manual_seed(2)
    n, t, i, o = 6, 36, 128, 4608
    lin = torch.nn.Linear(i, o).cuda()
    with torch.no_grad():
        lin.weight.copy_(bf16_round(torch.randn(o, i, generator=g) / 8).cuda())
        lin.bias.copy_(torch.randn(o, generator=g).cuda())
    x = bf16_round(torch.randn(n, t, i, generator=g))
    dy = bf16_round(torch.randn(n, t, o, generator=g))
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.linear(xr, lin.weight.detach().cpu(), lin.bias.detach().cpu())
    ref.backward(dy)
    wref = torch.einsum("nto,nti->oi", dy, x)
    pack = nets.ConvPack(lin)
    xd = x.to(torch.bfloat16).cuda()[:, :, None, :].contiguous().requires_grad_(True)
    y = ops.conv2d(xd, lin.weight[:, :, None, None], lin.bias, pack.get(), stride=1, pad=0)
    assert tuple(y.shape) == (n, t, 1, o)
    y.backward(dy.to(torch.bfloat16).cuda()[:, :, None, :].contiguous())
    assert rel(y[:, :, 0].float().cpu(), ref.detach()) < 6e-3
    assert rel(xd.grad[:, :, 0].float().cpu(), xr.grad) < 1e-2
    assert rel(lin.weight.grad.cpu(), wref) < 6e-3
    assert rel(lin.bias.grad.cpu(), dy.sum(dim=(0, 1))) < 6e-3


def test_layout_resample_concat_bit_exact(K):
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.randn(2, 24, 16, 16, generator=g))
    x_d = to_dev_nhwc(K, x)
    assert torch.equal(K.nhwc_to_nchw(x_d).cpu(), x)
    up = K.nhwc_to_nchw(K.resample2x(x_d, 0)).cpu()
    assert torch.equal(up, F.interpolate(x, scale_factor=2, mode="nearest"))
    pool = K.nhwc_to_nchw(K.resample2x(x_d, 1)).cpu()
    assert rel(pool, F.avg_pool2d(x, 2, 2)) < 4e-3
    y = bf16_round(torch.randn(2, 40, 16, 16, generator=g))
    from joligen_b200 import ops
    cat = ops.cat_channels(x_d, to_dev_nhwc(K, y))
    assert torch.equal(K.nhwc_to_nchw(cat).cpu(), torch.cat([x, y], dim=1))


def test_linear_noise_pack_loss_adam(K):
    from joligen_b200 import lib as L
    from oracle import palette_oracle as O
    g = torch.Generator().manual_seed(3)
    # emb_layers: SiLU -> Linear
    x = torch.randn(4, 32, generator=g)
    w = torch.randn(256, 32, generator=g) / 6
    b = torch.randn(256, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.linear(F.silu(xr), wr, br)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    y = K.linear_fwd(x.cuda(), w.cuda(), b.cuda(), act_in=L.ACT_SILU)
    assert rel(y, ref.detach()) < 1e-5
    dx, dw, db = K.linear_bwd(x.cuda(), w.cuda(), dy.cuda(), act_in=L.ACT_SILU)
    assert rel(dx, xr.grad) < 1e-5 and rel(dw, wr.grad) < 1e-5 and rel(db, br.grad) < 1e-5

    # prologue: q_sample + mask blend + cat, mask clamp bit exact (int64 masks with values outside {0,1})
    data = O.synthetic_batch(2, 32, 9)
    mask = data["mask"].clone()
    mask[0, 0, :2] = 3
    mask[1, 0, :2] = -1
    noise = torch.randn(2, 3, 32, 32, generator=g)
    gam = torch.tensor([0.3, 0.9])
    packed = K.noise_pack(data["gt"].cuda(), data["cond"].cuda(), noise.cuda(), mask.cuda(), gam.cuda())
    g4 = gam.view(-1, 1, 1, 1)
    yn = g4.sqrt() * data["gt"] + (1 - g4).sqrt() * noise
    m = torch.clamp(mask, min=0.0, max=1.0)
    yn = yn * m + (1.0 - m) * data["gt"]
    ref_in = torch.cat([data["cond"], yn], dim=1)
    got = K.nhwc_to_nchw(packed, 6).cpu()
    assert torch.equal(got, bf16_round(ref_in))  # elementwise fp32 math + one bf16 rounding: bit exact
    assert float(packed[..., 6:].float().abs().max()) == 0.0

    # eps-loss forward / backward
    nh = bf16_round(torch.randn(2, 3, 32, 32, generator=g))
    nhr = nh.clone().requires_grad_(True)
    wsnr = torch.tensor([0.7, 1.0])
    ref_loss = O.palette_loss(noise, nhr, mask, wsnr.view(-1, 1, 1, 1), lambda_G=2.0, use_minsnr=True)
    ref_loss.backward()
    nh_d = K.nchw_to_nhwc(nh.cuda())
    loss = K.palette_loss_fwd(noise.cuda(), nh_d, mask.cuda(), wsnr.cuda(), lambda_g=2.0)
    assert abs(float(loss) - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    gout = torch.tensor(0.5, device="cuda")
    dnh = K.palette_loss_bwd(noise.cuda(), nh_d, mask.cuda(), wsnr.cuda(), gout, lambda_g=2.0)
    assert rel(K.nhwc_to_nchw(dnh, 3), 0.5 * nhr.grad) < 1e-2

    # fused AdamW + EMA against the oracle's restatement of torch.optim.AdamW + ema_step, 3 steps
    p0 = torch.randn(1000, generator=g)
    st = O.TrainState(params={"p": p0.clone()})
    oc = O.OptimCfg(lr=1e-2, weight_decay=0.05, ema_beta=0.9)
    p = p0.clone().cuda()
    m_, v_, e_ = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(1000, generator=g)
        O.adam_update(st, {"p": gr}, oc)
        K.adamw_ema_step(p, (2 * gr).cuda(), m_, v_, e_, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.05,
                         adamw=True, step=step, grad_scale=0.5, ema_beta=0.9, ema_init=(step == 1))
    assert rel(p, st.params["p"]) < 1e-5
    assert rel(e_, st.ema["p"]) < 1e-5
