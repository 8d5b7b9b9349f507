"""Parity of the kernel VARIANTS the benchmark actually runs: every distinct convolution / GroupNorm / attention
shape of BASELINE.json config 2 (Palette UNet ngf 64, mults 1-2-4-8, 256x256; SURVEY.md section 8 a-1 / a-2 / a-3 /
a-9).  The launchers pick a template variant from C / H / W (conv_halo.cu launch_conv_halo, launch_wgrad_halo:
BLOCK_N 32..256, resident weights, N = 64 / 128 / 256 wgrad passes), so the toy shapes of test_gpu_ops.py do not
cover them.  Batch 2 by default (the variant does not depend on N), plus a few batch-32 cases for the persistent
multi-tile loop.

Reference = plain PyTorch fp32 on the same bf16-rounded operands (computed on the GPU with TF32 off).
Tolerances: the north star's 1e-2 (bf16 storage) relative to the tensor's max; fp32 weight gradients 1e-3.
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
    return float((a.float() - b.float()).abs().max()) / (float(b.float().abs().max()) + 1e-12)


# (cin, cout, size, k): the 22 distinct 3x3 and 15 distinct 1x1 convolutions of one forward pass
CONV3 = [(512, 512, 32), (64, 64, 256), (128, 128, 128), (256, 256, 64), (512, 512, 64), (256, 256, 128),
         (128, 128, 256), (1024, 512, 32), (128, 64, 256), (768, 512, 32), (768, 256, 64), (512, 256, 64),
         (384, 256, 64), (384, 128, 128), (256, 128, 128), (192, 128, 128), (192, 64, 256), (64, 128, 128),
         (128, 256, 64), (256, 512, 32), (6, 64, 256), (64, 3, 256)]
CONV1 = [(64, 128, 128), (128, 256, 64), (256, 512, 32), (1024, 512, 32), (768, 512, 32), (768, 256, 64),
         (512, 256, 64), (384, 256, 64), (384, 128, 128), (256, 128, 128), (192, 128, 128), (192, 64, 256),
         (128, 64, 256), (512, 1536, 32), (512, 512, 32)]  # + the attention block's qkv / proj_out
CONV_CASES = [(2, ci, co, s, 3) for ci, co, s in CONV3] + [(2, ci, co, s, 1) for ci, co, s in CONV1] + [
    (32, 64, 64, 256, 3), (32, 512, 512, 32, 3), (32, 128, 64, 256, 3), (32, 192, 64, 256, 1), (16, 256, 256, 64, 3)]


def _nhwc(t_nchw):
    return t_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(t_nhwc, c=None):
    t = t_nhwc.float().permute(0, 3, 1, 2)
    return t if c is None else t[:, :c]


def _pad_channels(t_nhwc, c8):
    n, h, w, c = t_nhwc.shape
    if c == c8:
        return t_nhwc
    out = torch.zeros((n, h, w, c8), dtype=t_nhwc.dtype, device=t_nhwc.device)
    out[..., :c] = t_nhwc
    return out


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "n%d_%dto%d_%d_k%d" % c)
def test_conv_production_shape(K, case):
    """fwd (+bias, +residual, written into a channel slice of a wider buffer as the decoder does), dgrad from a
    channel-sliced gradient (the concat backward hands out views), wgrad, bias gradient."""
    n, cin, cout, size, k = case
    pad = (k - 1) // 2
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout * 3 + size + k + n)
    rnd = lambda *shape: torch.randn(*shape, generator=g, device="cuda")  # noqa: E731
    x = rnd(n, cin, size, size).bfloat16().float()
    wt = rnd(cout, cin, k, k) / math.sqrt(cin * k * k)
    wb = wt.bfloat16().float()
    b = 0.1 * rnd(cout)
    res = rnd(n, cout, size, size).bfloat16().float()
    xr, wr, br = x.clone().requires_grad_(True), wb.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_plain = F.conv2d(xr, wr, br, padding=pad)
    dy = rnd(*y_plain.shape).bfloat16().float()
    y_plain.backward(dy)
    ref_res = (y_plain.detach() + res / math.sqrt(2))

    cin8, cout8 = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    x_d = _pad_channels(_nhwc(x), cin8)
    wf, wd = K.pack_conv_weight(wt)
    bias_p = torch.zeros(cout8, device="cuda")
    bias_p[:cout] = b
    # forward into a channel slice (ldy = cout8 + 64), residual read from a channel slice too
    buf = torch.full((n, size, size, cout8 + 64), 7.0, dtype=torch.bfloat16, device="cuda")
    resbuf = torch.zeros((n, size, size, cout8 + 8), dtype=torch.bfloat16, device="cuda")
    resbuf[..., 8:8 + cout] = _nhwc(res)
    y = K.conv2d_fwd(x_d, wf, bias_p, cout8, k, k, pad=pad, out=buf[..., :cout8])
    assert rel(_nchw(y, cout), y_plain.detach()) < 1e-2
    assert float((buf[..., cout8:].float() - 7.0).abs().max()) == 0.0  # nothing written past the slice
    if cout8 != cout:
        assert float(y[..., cout:].float().abs().max()) == 0.0
    y2 = K.conv2d_fwd(x_d, wf, bias_p, cout8, k, k, pad=pad, residual=resbuf[..., 8:8 + cout8],
                      res_scale=1.0 / math.sqrt(2))
    assert rel(_nchw(y2, cout), ref_res) < 1e-2
    # dgrad: the incoming gradient is a channel slice of a wider tensor
    dybuf = torch.zeros((n, size, size, cout8 + 16), dtype=torch.bfloat16, device="cuda")
    dybuf[..., 16:16 + cout] = _nhwc(dy)
    dy_d = dybuf[..., 16:]
    dx = K.conv2d_fwd(dy_d, wd, None, cin8, k, k, pad=k - 1 - pad)
    assert rel(_nchw(dx, cin), xr.grad) < 1e-2
    dw = K.conv2d_wgrad(x_d, dy_d, cout8, k, k, pad=pad)[:cout, :cin]
    assert rel(dw, wr.grad) < 1e-3
    assert rel(K.bias_grad(dy_d)[:cout], br.grad) < 1e-3
    # GroupNorm work fused into the epilogues (jg_conv_epilogue): the statistics of the stored output ...
    stats = torch.zeros(n, cout8, 2, device="cuda")
    y3 = K.conv2d_fwd(x_d, wf, bias_p, cout8, k, k, pad=pad, residual=resbuf[..., 8:8 + cout8],
                      res_scale=1.0 / math.sqrt(2), stats=stats)
    assert torch.equal(y3, y2)  # the fused reduction does not disturb the output
    yf = y3.float()
    want = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1)
    assert rel(stats, want) < 1e-4
    assert rel(K.chan_stats(y3), want) < 1e-4  # the stand-alone pass computes the same thing
    # ... and the GroupNorm-backward sums in the dgrad epilogue: du = dx * silu'(a * x_gn + b), (sum du, sum du * x_gn)
    from joligen_b200 import lib as L
    x_gn = _pad_channels(_nhwc((rnd(n, cin, size, size) * 1.5).bfloat16().float()), cin8)
    ab = torch.stack([1 + 0.3 * rnd(n, cin8), 0.2 * rnd(n, cin8)], dim=-1).contiguous()
    sums = torch.zeros(n, cin8, 2, device="cuda")
    dx2 = K.conv2d_fwd(dy_d, wd, None, cin8, k, k, pad=k - 1 - pad, gn=(x_gn, ab, L.ACT_SILU, sums))
    assert torch.equal(dx2, dx)
    u = x_gn.float() * ab[:, None, None, :, 0] + ab[:, None, None, :, 1]
    sg = torch.sigmoid(u)
    du = dx.float() * (sg * (1 + u * (1 - sg)))
    want = torch.stack([du.sum(dim=(1, 2)), (du * x_gn.float()).sum(dim=(1, 2))], dim=-1)
    assert rel(sums, want) < 2e-3
    # the trainer's path: raw accumulation into a persistent slot + batched unpack
    if cin8 == cin and cout8 == cout:
        from joligen_b200 import lib as L
        acc = torch.zeros(cout * cin * k * k, device="cuda")
        grad = torch.zeros(cout, cin, k, k, device="cuda")
        lay = K.conv2d_wgrad_acc(x_d, dy_d, cout, k, k, acc, pad=pad)
        K.wgrad_unpack_batched(K.WeightTable([L.UnpackItem(acc.data_ptr(), grad.data_ptr(), cout, cin, k * k, lay)],
                                             [(cout, cin, k * k)], torch.device("cuda")))
        assert rel(grad, wr.grad) < 1e-3


# (C, size): every GroupNorm(32) of one forward pass (ResBlock in / out norms incl. the concatenated decoder inputs)
GN_SHAPES = [(64, 256), (128, 256), (192, 256), (128, 128), (256, 128), (384, 128), (192, 128), (256, 64), (512, 64),
             (768, 64), (384, 64), (512, 32), (1024, 32), (768, 32)]


@pytest.mark.parametrize("case", [(2,) + s for s in GN_SHAPES] + [(32, 64, 256), (32, 1024, 32)],
                         ids=lambda c: "n%d_c%d_%d" % c)
def test_groupnorm_production_shape(K, case):
    """GN(32) + FiLM + SiLU forward; backward with two extra gradients of x summed in (skip path + decoder concat)
    and the column sums (the producer conv's bias gradient), as the UNet uses it."""
    from joligen_b200 import lib as L
    n, c, size = case
    g = torch.Generator(device="cuda").manual_seed(c + size + n)
    rnd = lambda *shape: torch.randn(*shape, generator=g, device="cuda")  # noqa: E731
    x = (rnd(n, c, size, size) * 1.5 + 0.3).bfloat16().float()
    gamma, beta = 1 + 0.2 * rnd(c), 0.1 * rnd(c)
    film = 0.3 * rnd(n, 2 * c)
    xr, gr, br, fr = (t.clone().requires_grad_(True) for t in (x, gamma, beta, film))
    h = F.group_norm(xr, 32, gr, br, eps=1e-5)
    scale, shift = torch.chunk(fr[:, :, None, None], 2, dim=1)
    ref = F.silu(h * (1 + scale) + shift)
    dy = rnd(*ref.shape).bfloat16().float()
    a1, a2 = rnd(*ref.shape).bfloat16().float(), rnd(*ref.shape).bfloat16().float()
    ref.backward(dy)
    want_dx = xr.grad + a1 + a2
    x_d = _nhwc(x)
    y, stats, ab = K.groupnorm_fwd(x_d, gamma, beta, 32, film=film, act=L.ACT_SILU)
    assert rel(_nchw(y), ref.detach()) < 1e-2
    colsum = torch.empty(c, device="cuda")
    dx, dgamma, dbeta, dfilm = K.groupnorm_bwd(x_d, _nhwc(dy), gamma, beta, 32, film, L.ACT_SILU, stats, ab,
                                               need_param_grads=True, need_film_grad=True, addend=_nhwc(a1),
                                               addend2=_nhwc(a2), colsum=colsum)
    assert rel(_nchw(dx), want_dx) < 1e-2
    assert rel(dgamma, gr.grad) < 2e-3 and rel(dbeta, br.grad) < 2e-3 and rel(dfilm, fr.grad) < 2e-3
    assert rel(colsum, want_dx.sum(dim=(0, 2, 3))) < 5e-3


@pytest.mark.parametrize("n", [2, 32])
def test_attention_production_shape(K, n):
    """The middle block's attention: T = 32 x 32 = 1024 tokens, 16 heads x 32 channels, legacy head interleave."""
    from oracle.palette_oracle import qkv_attention_legacy
    t, heads, ch = 1024, 16, 32
    c = heads * ch
    g = torch.Generator(device="cuda").manual_seed(n)
    qkv = torch.randn(n, 3 * c, t, generator=g, device="cuda").bfloat16().float()
    do = torch.randn(n, c, t, generator=g, device="cuda").bfloat16().float()
    # reference in chunks of 2 images (the materialised T x T logits of 32 images x 16 heads would be 2 GB)
    refs, grads = [], []
    for i in range(0, n, 2):
        qr = qkv[i:i + 2].clone().requires_grad_(True)
        r = qkv_attention_legacy(qr, heads)
        r.backward(do[i:i + 2])
        refs.append(r.detach())
        grads.append(qr.grad)
    ref, gref = torch.cat(refs), torch.cat(grads)
    qkv_d = _nhwc(qkv.reshape(n, 3 * c, 32, 32))
    out, lse = K.attn_fwd(qkv_d, heads, ch)
    assert rel(_nchw(out).reshape(n, c, t), ref) < 1e-2
    dqkv = K.attn_bwd(qkv_d, out, _nhwc(do.reshape(n, c, 32, 32)), lse, heads, ch)
    assert rel(_nchw(dqkv).reshape(n, 3 * c, t), gref) < 1.5e-2
