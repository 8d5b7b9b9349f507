"""TEST DOUBLE for `joligen_b200.kernels` — test infrastructure, never shipped, never measured.

The product has no CPU path (every op raises without the CUDA library).  That leaves the HOST logic above the C ABI —
`ops.py` (autograd plumbing: which tensor goes into which kernel argument, stamps, taps, concat slices), `nets.py`
(module mirrors), `accelerate.py` (the swap under the reference's `BaseModel`) — untestable where there is no GPU.
`install()` replaces the allocation-only wrappers of `joligen_b200/kernels.py` (the ONE place that crosses the C ABI)
by torch-CPU restatements of each kernel's documented contract (`include/jg_b200.h`): same arguments, same layouts
(NHWC bf16 channel slices, packed bf16 weights `[Cout8][R*S][Cin8]` / flipped `[Cin8][R*S][Cout8]`, fp32 `stats` /
`ab` / `chan_stats`), bf16 rounding where the kernels store bf16, fp32 arithmetic in between.  With it the whole
host stack runs on the CPU and is compared with the reference's golden vectors (tests/test_host_double.py) — and the
UNMODIFIED reference's `PaletteModel.optimize_parameters()` runs with `accelerate(netG_A)` swapped in.

Only `tests/` may import this module (it is a checker, like `oracle/`).  It says nothing about the CUDA kernels: those
are held to the oracle by the `-m gpu` tests.
"""
import contextlib

import torch
import torch.nn.functional as F

from joligen_b200 import kernels as K
from joligen_b200 import lib as L

BF = torch.bfloat16


def _act(u, act):
    if act == L.ACT_NONE:
        return u
    if act == L.ACT_RELU:
        return torch.relu(u)
    if act == L.ACT_LRELU02:
        return F.leaky_relu(u, 0.2)
    if act == L.ACT_TANH:
        return torch.tanh(u)
    if act == L.ACT_SILU:
        return F.silu(u)
    raise ValueError(act)


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)



def _raw(dst):
    """The kernels write their outputs through raw pointers: torch's version counters never see it (a producer writes
    a channel slice of a buffer other views of which autograd already holds).  Every write of the double into a
    caller-provided tensor happens under this guard, so that the host logic sees exactly that behaviour."""
    return torch.autograd._unsafe_preserve_version_counter(dst)


def _check_rows(t):
    K._ld(t)  # the real wrappers assert that every operand is an NHWC channel slice
    assert t.dtype == BF


# ---------------------------------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------------------------------
def pack_conv_weight(w_oihw, want_dgrad=True, out=None):
    cout, cin, r, s = w_oihw.shape
    w = w_oihw.detach().float()
    cin8, cout8 = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    wf = torch.zeros((cout8, r * s, cin8), dtype=BF)
    wf[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, r * s, cin).to(BF)
    wd = None
    if want_dgrad:
        wd = torch.zeros((cin8, r * s, cout8), dtype=BF)
        wd[:cin, :, :cout] = w.permute(1, 2, 3, 0).reshape(cin, r * s, cout).flip(1).to(BF)
    if out is not None:
        with _raw(out[0]):
            out[0].copy_(wf)
        if wd is not None and out[1] is not None:
            with _raw(out[1]):
                out[1].copy_(wd)
        return out
    return wf, wd


def _weight_oihw(w_packed, cout, r, s):
    rows, rs, cin8 = w_packed.shape
    assert rs == r * s and rows >= cout, (w_packed.shape, cout, r, s)
    return w_packed[:cout].float().reshape(cout, r, s, cin8).permute(0, 3, 1, 2)


def conv2d_fwd(x, w_packed, bias, cout, r, s, stride=1, pad=None, act=L.ACT_NONE, residual=None, res_scale=1.0,
               out=None, stats=None, gn=None):
    _check_rows(x)
    n, h, w, cin = x.shape
    if pad is None:
        pad = (r - 1) // 2
    assert cin == w_packed.shape[2], "x channels %d vs packed Cin8 %d" % (cin, w_packed.shape[2])
    assert cin % 8 == 0 and w_packed.dtype == BF
    ho, wo = K.conv_out_size(h, w, r, s, stride, pad)
    y = F.conv2d(_nchw(x), _weight_oihw(w_packed, cout, r, s), None, stride=stride, padding=pad)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= cout
        y = y + bias[:cout].float().view(1, -1, 1, 1)
    if residual is not None:
        _check_rows(residual)
        assert gn is None, "gn_sums is not combinable with a residual operand"
        assert tuple(residual.shape) == (n, ho, wo, cout), (residual.shape, (n, ho, wo, cout))
        y = y + float(res_scale) * _nchw(residual)
    y = _nhwc(_act(y, act)).to(BF)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=BF)
    _check_rows(out)
    assert tuple(out.shape) == (n, ho, wo, cout), (out.shape, (n, ho, wo, cout))
    with _raw(out):
        out.copy_(y)
    if stats is not None:
        assert stats.is_contiguous() and stats.dtype == torch.float32 and stats.numel() == n * cout * 2
        yf = out.float().reshape(n, -1, cout)
        with _raw(stats):
            stats.view(n, cout, 2).add_(torch.stack([yf.sum(1), (yf * yf).sum(1)], dim=-1))
    if gn is not None:
        x_gn, ab, gn_act, sums = gn
        assert tuple(x_gn.shape) == (n, ho, wo, cout) and ab.numel() == n * cout * 2 and sums.numel() == n * cout * 2
        a = ab.view(n, 1, 1, cout, 2)[..., 0]
        b = ab.view(n, 1, 1, cout, 2)[..., 1]
        u = (a * x_gn.float() + b).requires_grad_(True)
        with torch.enable_grad():
            v = _act(u, gn_act)
        (du,) = torch.autograd.grad(v, u, out.float())
        with _raw(sums):
            sums.view(n, cout, 2).add_(torch.stack([du.reshape(n, -1, cout).sum(1),
                                                    (du * x_gn.float()).reshape(n, -1, cout).sum(1)], dim=-1))
    return out


def chan_stats(x):
    _check_rows(x)
    n, h, w, c = x.shape
    xf = x.float().reshape(n, -1, c)
    return torch.stack([xf.sum(1), (xf * xf).sum(1)], dim=-1).contiguous()


def conv2d_cropped(x, w_packed, bias, cout, r, s, pad, out_hw, act=L.ACT_NONE):
    _check_rows(x)
    y = F.conv2d(_nchw(x), _weight_oihw(w_packed, cout, r, s), None, stride=1, padding=pad)
    if bias is not None:
        y = y + bias[:cout].float().view(1, -1, 1, 1)
    ho, wo = out_hw
    assert y.shape[2] >= ho and y.shape[3] >= wo
    return _nhwc(_act(y[:, :, :ho, :wo], act)).to(BF).contiguous()


def conv2d_wgrad(x, dy, cout, r, s, stride=1, pad=None, out=None, beta=0.0):
    _check_rows(x)
    _check_rows(dy)
    if pad is None:
        pad = (r - 1) // 2
    cin = x.shape[-1]
    assert dy.shape[-1] >= cout
    dw = torch.nn.grad.conv2d_weight(_nchw(x).contiguous(), (cout, cin, r, s), _nchw(dy[..., :cout]).contiguous(),
                                     stride=stride, padding=pad)
    if out is None:
        return dw.contiguous()
    with _raw(out):
        out.mul_(beta).add_(dw)
    return out


def conv2d_wgrad_acc(x, dy, cout, r, s, acc, stride=1, pad=None):
    """layout 1 ([Cout][R*S][Cin]) accumulator, like the halo wgrad kernels report."""
    dw = conv2d_wgrad(x, dy, cout, r, s, stride=stride, pad=pad)
    with _raw(acc):
        acc.view(cout, r * s, x.shape[-1]).add_(dw.permute(0, 2, 3, 1).reshape(cout, r * s, x.shape[-1]))
    return 1


def bias_grad(dy):
    _check_rows(dy)
    return dy.float().reshape(-1, dy.shape[-1]).sum(0)


# ---------------------------------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------------------------------
def nchw_to_nhwc(x, ld=None):
    n, c, h, w = x.shape
    if ld is None:
        ld = (c + 7) // 8 * 8
    out = torch.zeros((n, h, w, ld), dtype=BF)
    out[..., :c] = x.float().permute(0, 2, 3, 1).to(BF)
    return out


def nhwc_to_nchw(x, c=None):
    _check_rows(x)
    if c is None:
        c = x.shape[-1]
    return x[..., :c].float().permute(0, 3, 1, 2).contiguous()


def copy_channels(src, dst, accumulate=False):
    _check_rows(src)
    _check_rows(dst)
    assert dst.shape == src.shape
    with _raw(dst):
        if accumulate:
            dst.copy_((dst.float() + src.float()).to(BF))
        else:
            dst.copy_(src)
    return dst


def resample2x(x, mode):
    _check_rows(x)
    xf = _nchw(x)
    if mode == 0:      # F.interpolate(scale_factor=2, mode="nearest")
        y = F.interpolate(xf, scale_factor=2, mode="nearest")
    elif mode == 1:    # nn.AvgPool2d(2, 2)
        y = F.avg_pool2d(xf, 2, 2)
    elif mode == 2:    # backward of mode 0: sum over each 2x2 block
        y = F.avg_pool2d(xf, 2, 2) * 4.0
    elif mode == 3:    # backward of mode 1: 0.25 * nearest upsample
        y = F.interpolate(xf, scale_factor=2, mode="nearest") * 0.25
    else:
        raise ValueError(mode)
    return _nhwc(y).to(BF).contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# GroupNorm (+FiLM)(+act)
# ---------------------------------------------------------------------------------------------------------------------
_EPS = 1e-5


def _gn_forward(x32, gamma, beta, groups, film, act, mean=None, rstd=None, eps=_EPS):
    """fp32 [N,H,W,C] -> act(((x - mean) * rstd * gamma + beta) * (1 + scale) + shift); differentiable.
    eps: a float, or a [N,1,G,1] tensor (the backward recovers it from the saved rstd)."""
    n, h, w, c = x32.shape
    cg = c // groups
    xg = x32.reshape(n, h * w, groups, cg)
    if mean is None:
        mean = xg.mean(dim=(1, 3), keepdim=True)
        var = ((xg - mean) ** 2).mean(dim=(1, 3), keepdim=True)
        rstd = (var + eps).rsqrt()
    xh = ((xg - mean) * rstd).reshape(n, h, w, c)
    if gamma is not None:
        xh = xh * gamma.view(1, 1, 1, c) + beta.view(1, 1, 1, c)
    if film is not None:
        xh = xh * (1.0 + film[:, :c].reshape(n, 1, 1, c)) + film[:, c:].reshape(n, 1, 1, c)
    return _act(xh, act), mean.reshape(n, groups), rstd.reshape(n, groups)


def groupnorm_fwd(x, gamma, beta, groups, film=None, act=L.ACT_NONE, eps=1e-5, out=None, chan_stats=None):
    _check_rows(x)
    n, h, w, c = x.shape
    assert c % groups == 0
    x32 = x.float()
    mean = rstd = None
    if chan_stats is not None:
        # the statistics pass is skipped: (mean, rstd) come from the producer's per-(image, channel) sums — which must
        # therefore describe THIS tensor (a stale or mis-routed stamp shows up as a parity failure here)
        assert chan_stats.is_contiguous() and chan_stats.dtype == torch.float32 and chan_stats.numel() == n * c * 2
        cs = chan_stats.view(n, groups, c // groups, 2).sum(2)
        cnt = float(h * w * (c // groups))
        m = cs[..., 0] / cnt
        var = (cs[..., 1] / cnt - m * m).clamp_min(0.0)
        mean, rstd = m.view(n, 1, groups, 1), (var + eps).rsqrt().view(n, 1, groups, 1)
    if film is not None:
        assert film.dtype == torch.float32 and tuple(film.shape) == (n, 2 * c), (film.shape, (n, 2 * c))
    y, mean, rstd = _gn_forward(x32, None if gamma is None else gamma.detach().float(),
                                None if beta is None else beta.detach().float(), groups,
                                None if film is None else film.detach(), act, mean, rstd, eps)
    y = y.to(BF)
    if out is None:
        out = y.contiguous()
    else:
        _check_rows(out)
        with _raw(out):
            out.copy_(y)
    stats = torch.stack([mean, rstd], dim=-1).contiguous()
    cg = c // groups
    g = torch.ones(c) if gamma is None else gamma.detach().float()
    bt = torch.zeros(c) if beta is None else beta.detach().float()
    a = rstd.repeat_interleave(cg, dim=1) * g
    b = bt - mean.repeat_interleave(cg, dim=1) * a
    if film is not None:
        sc, sh = film[:, :c], film[:, c:]
        a, b = a * (1 + sc), b * (1 + sc) + sh
    return out, stats, torch.stack([a, b], dim=-1).contiguous()


def groupnorm_bwd(x, dy, gamma, beta, groups, film, act, stats, ab, need_param_grads=True, need_film_grad=False,
                  dx=None, addend=None, colsum=None, addend2=None, sums_pre=None, dfilm_out=None):
    _check_rows(x)
    _check_rows(dy)
    n, h, w, c = x.shape
    assert tuple(dy.shape) == tuple(x.shape), (dy.shape, x.shape)
    assert tuple(stats.shape) == (n, groups, 2) and tuple(ab.shape) == (n, c, 2)
    x32 = x.float().requires_grad_(True)
    leaves = [x32]
    g32 = b32 = f32 = None
    if gamma is not None:
        g32, b32 = gamma.detach().float().requires_grad_(True), beta.detach().float().requires_grad_(True)
        leaves += [g32, b32]
    if film is not None:
        f32 = film.detach().float().requires_grad_(True)
        leaves.append(f32)
    # eps is not an argument of the backward: it is whatever the saved rstd was made with
    with torch.no_grad():
        xg = x.float().reshape(n, h * w, groups, c // groups)
        var = ((xg - xg.mean(dim=(1, 3), keepdim=True)) ** 2).mean(dim=(1, 3), keepdim=True)
        eps = (stats[..., 1].view(n, 1, groups, 1) ** -2 - var).clamp_min(0.0)
    with torch.enable_grad():
        y, mean, rstd = _gn_forward(x32, g32, b32, groups, f32, act, eps=eps)
    # the saved statistics must be the ones of this x (they are what the real kernel differentiates with)
    assert torch.allclose(mean, stats[..., 0], rtol=1e-3, atol=1e-3), "groupnorm_bwd: stats do not belong to x"
    assert torch.allclose(rstd, stats[..., 1], rtol=2e-2), "groupnorm_bwd: stats do not belong to x"
    if sums_pre is not None:
        a = ab[..., 0].view(n, 1, 1, c)
        b = ab[..., 1].view(n, 1, 1, c)
        u = (a * x.float() + b).requires_grad_(True)
        with torch.enable_grad():
            v = _act(u, act)
        (du,) = torch.autograd.grad(v, u, dy.float())
        ref = torch.stack([du.reshape(n, -1, c).sum(1), (du * x.float()).reshape(n, -1, c).sum(1)], dim=-1)
        assert torch.allclose(sums_pre.view(n, c, 2), ref, rtol=2e-2, atol=1e-2 * float(ref.abs().max())), \
            "groupnorm_bwd: sums_pre do not belong to (x, dy)"
    grads = torch.autograd.grad(y, leaves, dy.float(), allow_unused=True)
    d = grads[0]
    if addend is not None:
        _check_rows(addend)
        d = d + addend.float()
    if addend2 is not None:
        assert addend is not None, "addend2 requires addend"
        _check_rows(addend2)
        d = d + addend2.float()
    dxo = d.to(BF)
    if dx is None:
        dx = dxo.contiguous()
    else:
        with _raw(dx):
            dx.copy_(dxo)
    if colsum is not None:
        with _raw(colsum):
            colsum.copy_(d.reshape(-1, c).sum(0))
    dgamma = dbeta = dfilm = None
    i = 1
    if gamma is not None:
        if need_param_grads:
            dgamma, dbeta = grads[1].contiguous(), grads[2].contiguous()
        i = 3
    if film is not None and need_film_grad:
        df = grads[i]
        ok = dfilm_out is not None and dfilm_out.is_contiguous() and tuple(dfilm_out.shape) == (n, 2 * c)
        if ok:
            with _raw(dfilm_out):
                dfilm_out.copy_(df)
            dfilm = dfilm_out
        else:
            dfilm = df.contiguous()
    return dx, dgamma, dbeta, dfilm


# ---------------------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------------------
def _attn(qkv32, heads, ch, layout):
    n, h, w, c3 = qkv32.shape
    t = h * w
    assert c3 == 3 * heads * ch
    z = qkv32.reshape(n, t, c3)
    if layout == 0:   # QKVAttentionLegacy: per head (q | k | v)
        z = z.reshape(n, t, heads, 3, ch)
        q, k, v = z[:, :, :, 0], z[:, :, :, 1], z[:, :, :, 2]
    else:             # QKVAttention: (q | k | v), each heads * ch wide
        z = z.reshape(n, t, 3, heads, ch)
        q, k, v = z[:, :, 0], z[:, :, 1], z[:, :, 2]
    scale = ch ** -0.25
    s = torch.einsum("nthc,nshc->nhts", q * scale, k * scale)
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("nhts,nshc->nthc", p, v)
    return o.reshape(n, h, w, heads * ch), torch.logsumexp(s, dim=-1).reshape(n * heads, t)


def attn_fwd(qkv, heads, ch, layout=0, out=None):
    _check_rows(qkv)
    o, lse = _attn(qkv.float(), heads, ch, layout)
    o = o.to(BF)
    if out is None:
        out = o.contiguous()
    else:
        _check_rows(out)
        with _raw(out):
            out.copy_(o)
    return out, lse.contiguous()


def attn_bwd(qkv, out, d_out, lse, heads, ch, layout=0):
    _check_rows(qkv)
    _check_rows(d_out)
    z = qkv.float().requires_grad_(True)
    with torch.enable_grad():
        o, _ = _attn(z, heads, ch, layout)
    (d,) = torch.autograd.grad(o, z, d_out.float())
    return d.to(BF).contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# small fp32 Linears
# ---------------------------------------------------------------------------------------------------------------------
def linear_fwd(x, w, b, act_in=L.ACT_NONE, act_out=L.ACT_NONE):
    assert x.dtype == torch.float32 and x.is_contiguous()
    return _act(F.linear(_act(x, act_in), w.detach(), None if b is None else b.detach()), act_out)


def linear_bwd(x, w, dy, act_in=L.ACT_NONE, need_dx=True):
    x32 = x.detach().requires_grad_(True)
    w32 = w.detach().requires_grad_(True)
    with torch.enable_grad():
        y = F.linear(_act(x32, act_in), w32)
    dx, dw = torch.autograd.grad(y, [x32, w32], dy)
    return (dx if need_dx else None), dw, dy.sum(0)


class LinearBank:
    """kernels.LinearBank without the device table: keeps the Linear modules themselves."""

    def __init__(self, linears, device):
        self.linears = list(linears)
        self.widths = [lin.weight.shape[0] for lin in linears]
        self.in_features = linears[0].weight.shape[1]
        self.offsets, off = [], 0
        for lin in linears:
            assert lin.weight.dtype == torch.float32 and lin.weight.is_contiguous()
            self.offsets.append(off)
            off += lin.weight.shape[0]
        self.total_out, self.n = off, len(linears)
        self.key = tuple((lin.weight.data_ptr(), 0 if lin.bias is None else lin.bias.data_ptr()) for lin in linears)


def linear_batched_fwd(x, bank, act_in=L.ACT_NONE):
    bsz = x.shape[0]
    y = torch.empty((bsz * bank.total_out,), dtype=torch.float32)
    for lin, off, o in zip(bank.linears, bank.offsets, bank.widths):
        y[bsz * off:bsz * (off + o)] = linear_fwd(x, lin.weight, lin.bias, act_in).reshape(-1)
    return y


def linear_batched_bwd(x, bank, dy, act_in=L.ACT_NONE, need_dx=True):
    bsz, i = x.shape
    dw = torch.empty((bank.total_out, i), dtype=torch.float32)
    db = torch.empty((bank.total_out,), dtype=torch.float32)
    dx = torch.zeros_like(x) if need_dx else None
    for lin, off, o in zip(bank.linears, bank.offsets, bank.widths):
        d = dy[bsz * off:bsz * (off + o)].view(bsz, o)
        dxi, dwi, dbi = linear_bwd(x, lin.weight, d, act_in, need_dx)
        dw[off:off + o], db[off:off + o] = dwi, dbi
        if need_dx:
            dx += dxi
    return dx, dw, db


# ---------------------------------------------------------------------------------------------------------------------
# prologue / loss / optimizer
# ---------------------------------------------------------------------------------------------------------------------
def _mask01(mask):
    if mask is None:
        return None
    assert mask.is_contiguous() and mask.dtype in (torch.int64, torch.float32), "mask must be int64 or float32"
    return mask.float().clamp(0.0, 1.0)


def noise_pack(y0, ycond, noise, mask, gammas, ld=8):
    b, c, h, w = y0.shape
    g = gammas.reshape(b, 1, 1, 1).float()
    yn = g.sqrt() * y0 + (1 - g).sqrt() * noise
    m = _mask01(mask)
    if m is not None:
        m = m.reshape(b, 1, h, w)
        yn = yn * m + (1 - m) * y0
    out = torch.zeros((b, h, w, ld), dtype=BF)
    out[..., :c] = ycond.permute(0, 2, 3, 1).to(BF)
    out[..., c:2 * c] = yn.permute(0, 2, 3, 1).to(BF)
    return out


def _loss_terms(noise, noise_hat, mask, w_b):
    b, c, h, w = noise.shape
    diff = noise - noise_hat[..., :c].float().permute(0, 3, 1, 2)
    m = _mask01(mask)
    scale = torch.ones((b, 1, 1, 1))
    if m is not None:
        scale = scale * m.reshape(b, 1, h, w)
    if w_b is not None:
        scale = scale * w_b.reshape(b, 1, 1, 1)
    return diff, scale


def palette_loss_fwd(noise, noise_hat, mask, w_b, lambda_g=1.0, l1=False):
    _check_rows(noise_hat)
    diff, scale = _loss_terms(noise, noise_hat, mask, w_b)
    e = scale * diff
    return float(lambda_g) * (e.abs().mean() if l1 else (e * e).mean())


def palette_loss_bwd(noise, noise_hat, mask, w_b, grad_out, lambda_g=1.0, l1=False):
    diff, scale = _loss_terms(noise, noise_hat, mask, w_b)
    e = scale * diff
    coef = float(lambda_g) / e.numel() * grad_out.reshape(())
    de = torch.sign(e) if l1 else 2.0 * e
    dnh = -(coef * de * scale)  # d/d noise_hat
    d = torch.zeros(noise_hat.shape, dtype=BF)
    d[..., :noise.shape[1]] = dnh.permute(0, 2, 3, 1).to(BF)
    return d


def adamw_ema_step(p, g, m, v, ema, lr, beta1, beta2, eps, weight_decay, adamw, step, grad_scale=1.0, ema_beta=0.999,
                   ema_init=False, step_dev=None):
    import contextlib as _cl
    with _cl.ExitStack() as st:
        for t in (p, m, v, ema, step_dev):
            if t is not None:
                st.enter_context(_raw(t))
        _adamw(p, g, m, v, ema, lr, beta1, beta2, eps, weight_decay, adamw, step, grad_scale, ema_beta, ema_init,
               step_dev)


def _adamw(p, g, m, v, ema, lr, beta1, beta2, eps, weight_decay, adamw, step, grad_scale, ema_beta, ema_init, step_dev):
    if step_dev is not None:
        step_dev.add_(1)
        step = int(step_dev.reshape(-1)[0])
        ema_init = step == 1
    gr = g * grad_scale
    if adamw:
        p.mul_(1.0 - lr * weight_decay)
    elif weight_decay:
        gr = gr + weight_decay * p
    m.mul_(beta1).add_(gr, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)
    if ema is not None:
        if ema_init:
            ema.copy_(p)
        else:
            ema.mul_(ema_beta).add_(p, alpha=1 - ema_beta)



# ---------------------------------------------------------------------------------------------------------------------
# GAN generator / discriminator helpers
# ---------------------------------------------------------------------------------------------------------------------
def pad2d(x, pad, mode=0):
    _check_rows(x)
    return _nhwc(F.pad(_nchw(x), (pad,) * 4, mode="reflect" if mode == 0 else "replicate")).to(BF).contiguous()


def pad2d_bwd(dpad, pad, mode=0):
    _check_rows(dpad)
    n, hp, wp, c = dpad.shape
    x = torch.zeros((n, c, hp - 2 * pad, wp - 2 * pad), requires_grad=True)
    with torch.enable_grad():
        y = F.pad(x, (pad,) * 4, mode="reflect" if mode == 0 else "replicate")
    (d,) = torch.autograd.grad(y, x, _nchw(dpad))
    return _nhwc(d).to(BF).contiguous()


def dilate2x(x):
    _check_rows(x)
    n, h, w, c = x.shape
    out = torch.zeros((n, 2 * h, 2 * w, c), dtype=BF)
    out[:, ::2, ::2] = x
    return out


def undilate2x(x):
    _check_rows(x)
    return x[:, ::2, ::2].contiguous()


def act_bwd(y, dy, act):
    _check_rows(y)
    _check_rows(dy)
    yf = y.float()
    if act == L.ACT_TANH:
        g = 1.0 - yf * yf
    elif act == L.ACT_LRELU02:
        g = torch.where(yf > 0, 1.0, 0.2)
    elif act == L.ACT_RELU:
        g = (yf > 0).float()
    else:
        raise ValueError(act)
    return (dy.float() * g).to(BF).contiguous()


def _gan_terms(pred, c_real, mode, target, sign):
    p = pred[..., :c_real].float().requires_grad_(True)
    with torch.enable_grad():
        if mode == K.GAN_LSGAN:
            loss = ((p - target) ** 2).mean()
        elif mode == K.GAN_HINGE:
            loss = torch.relu(1.0 - sign * p).mean()
        else:
            loss = (-sign * p).mean()
    return p, loss


def gan_loss_fwd(pred, c_real, mode, target, sign):
    _check_rows(pred)
    return _gan_terms(pred, c_real, mode, target, sign)[1].detach()


def gan_loss_bwd(pred, c_real, mode, target, sign, grad_out):
    p, loss = _gan_terms(pred, c_real, mode, target, sign)
    (d,) = torch.autograd.grad(loss, p, grad_out.reshape(()).float())
    out = torch.zeros(pred.shape, dtype=BF)
    out[..., :c_real] = d.to(BF)
    return out



# ---------------------------------------------------------------------------------------------------------------------
# MotionModule (video UNet): LayerNorm (+ frame positional encoding), temporal attention, GEGLU
# ---------------------------------------------------------------------------------------------------------------------
def _ln(x32, gamma, beta, pe, frames, eps=1e-5):
    n, h, w, c = x32.shape
    mean = x32.mean(-1, keepdim=True)
    var = ((x32 - mean) ** 2).mean(-1, keepdim=True)
    rstd = (var + eps).rsqrt()
    y = (x32 - mean) * rstd * gamma.view(1, 1, 1, c) + beta.view(1, 1, 1, c)
    if pe is not None:
        assert tuple(pe.shape) == (frames, c) and n % frames == 0
        y = y + pe.float()[torch.arange(n) % frames].view(n, 1, 1, c)  # frame index = n % F
    return y, mean, rstd


def layernorm_fwd(x, gamma, beta, eps=1e-5, pe=None, frames=1):
    _check_rows(x)
    y, mean, rstd = _ln(x.float(), gamma.detach().float(), beta.detach().float(), pe, frames, eps)
    return y.to(BF).contiguous(), torch.cat([mean, rstd], dim=-1).reshape(-1, 2).contiguous()


def layernorm_bwd(x, dy, gamma, stats, addend=None, colsum=None):
    _check_rows(x)
    _check_rows(dy)
    n, h, w, c = x.shape
    x32 = x.float().requires_grad_(True)
    g32 = gamma.detach().float().requires_grad_(True)
    b32 = torch.zeros(c, requires_grad=True)
    with torch.enable_grad():
        y, mean, _ = _ln(x32, g32, b32, None, 1)
    assert torch.allclose(mean.reshape(-1), stats[:, 0], rtol=1e-3, atol=1e-3), "layernorm_bwd: stats do not belong to x"
    dx, dg, db = torch.autograd.grad(y, [x32, g32, b32], dy.float())
    if addend is not None:
        _check_rows(addend)
        dx = dx + addend.float()
    if colsum is not None:
        with _raw(colsum):
            colsum.copy_(dx.reshape(-1, c).sum(0))
    return dx.to(BF).contiguous(), dg.contiguous(), db.contiguous()


def _tattn(qkv32, frames, heads):
    n, h, w, c3 = qkv32.shape
    c = c3 // 3
    ch = c // heads
    b = n // frames
    z = qkv32.reshape(b, frames, h * w, 3, heads, ch)
    q, k, v = z[:, :, :, 0], z[:, :, :, 1], z[:, :, :, 2]                 # [b, f, p, heads, ch]
    s = torch.einsum("bfphc,bgphc->bphfg", q, k) * (ch ** -0.5)
    o = torch.einsum("bphfg,bgphc->bfphc", torch.softmax(s, dim=-1), v)
    return o.reshape(n, h, w, c)


def temporal_attn_fwd(qkv, frames, heads):
    _check_rows(qkv)
    return _tattn(qkv.float(), frames, heads).to(BF).contiguous()


def temporal_attn_bwd(qkv, d_out, frames, heads):
    _check_rows(qkv)
    _check_rows(d_out)
    z = qkv.float().requires_grad_(True)
    with torch.enable_grad():
        o = _tattn(z, frames, heads)
    (d,) = torch.autograd.grad(o, z, d_out.float())
    return d.to(BF).contiguous()


def _geglu(x32):
    c = x32.shape[-1] // 2
    return x32[..., :c] * F.gelu(x32[..., c:])


def geglu_fwd(x):
    _check_rows(x)
    return _geglu(x.float()).to(BF).contiguous()


def geglu_bwd(x, dy):
    _check_rows(x)
    _check_rows(dy)
    z = x.float().requires_grad_(True)
    with torch.enable_grad():
        y = _geglu(z)
    (d,) = torch.autograd.grad(y, z, dy.float())
    return d.to(BF).contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# conditioning embeddings, reverse-diffusion step
# ---------------------------------------------------------------------------------------------------------------------
def _labels(idx, k):
    assert idx.dtype in (torch.int64, torch.float32)
    return idx.reshape(-1).to(torch.int64).clamp(0, k - 1)  # (.to(torch.int32) of the reference truncates)


def embed_rows(table, idx, out, col0):
    _check_rows(out)
    k, e = table.shape
    n, h, w, ld = out.shape
    assert e % 2 == 0 and col0 % 2 == 0 and col0 + e <= ld and idx.numel() == n * h * w
    with _raw(out):
        out[..., col0:col0 + e] = table.detach().float()[_labels(idx, k)].reshape(n, h, w, e).to(BF)
    return out


def embed_rows_bwd(d, idx, col0, k, e):
    _check_rows(d)
    lab = _labels(idx, k)
    rows = d[..., col0:col0 + e].float().reshape(-1, e)
    dtable = torch.zeros((k, e)).index_add_(0, lab, rows)
    counts = torch.zeros((k,)).index_add_(0, lab, torch.ones(lab.numel()))
    return dtable, counts


def ddpm_step(eps, y_t, y_cond, y_0, mask, noise, coef, ld=8, want_next_input=True, ddim=False):
    """include/jg_b200.h jg_ddpm_step: coef [B][5] = (sqrt_recip_gammas, sqrt_recipm1_gammas, posterior_mean_coef1,
    posterior_mean_coef2, exp(0.5 * posterior_log_variance_clipped)) gathered at t."""
    _check_rows(eps)
    b, c, h, w = y_t.shape
    e = eps[..., :c].float().permute(0, 3, 1, 2)
    cf = coef.reshape(b, 5, 1, 1, 1).float()
    if ddim:
        y = (cf[:, 0] * y_t + cf[:, 1] * e.clamp(-1, 1)).clamp(-1, 1)
    else:
        y0_hat = (cf[:, 0] * y_t - cf[:, 1] * e).clamp(-1, 1)
        y = cf[:, 2] * y0_hat + cf[:, 3] * y_t
        if noise is not None:
            y = y + noise * cf[:, 4]
    m = _mask01(mask)
    if m is not None:
        m = m.reshape(b, 1, h, w)
        y = y_0 * (1.0 - m) + m * y
    x_next = None
    if want_next_input:
        x_next = torch.zeros((b, h, w, ld), dtype=BF)
        x_next[..., :c] = y_cond.permute(0, 2, 3, 1).to(BF)
        x_next[..., c:2 * c] = y.permute(0, 2, 3, 1).to(BF)
    return y.contiguous(), x_next



# ---------------------------------------------------------------------------------------------------------------------
# trainer-side batched weight kernels: the item tables hold RAW POINTERS (here: host addresses) — follow them
# ---------------------------------------------------------------------------------------------------------------------
def _at(ptr, shape, dtype):
    """the tensor living at a raw address (what a kernel sees of a jg_pack_item / jg_unpack_item field)"""
    import ctypes
    n = 1
    for v in shape:
        n *= v
    buf = (ctypes.c_char * (n * torch.empty((), dtype=dtype).element_size())).from_address(ptr)
    return torch.frombuffer(buf, dtype=dtype).view(shape)


class WeightTable:
    """kernels.WeightTable without the device copy: keeps the ctypes items (jg_pack_item / jg_unpack_item)."""

    def __init__(self, items, dims, device):
        self.structs = list(items)
        self.n = len(items)
        self.total_tiles = sum(L.load().jg_weight_tiles(cout, cin, rs) for cout, cin, rs in dims)


def pack_conv_weights_batched(table):
    for it in table.structs:
        w = _at(it.w, (it.Cout, it.Cin, it.RS), torch.float32)
        wf = _at(it.wf, (it.Cout8, it.RS, it.Cin8), BF)
        wf[:it.Cout, :, :it.Cin] = w.permute(0, 2, 1).to(BF)
        if it.wd:
            wd = _at(it.wd, (it.Cin8, it.RS, it.Cout8), BF)
            wd[:it.Cin, :, :it.Cout] = w.permute(1, 2, 0).flip(1).to(BF)


def wgrad_unpack_batched(table):
    """dw_oihw += acc (permuted from its layout), then acc = 0"""
    for it in table.structs:
        dw = _at(it.dw, (it.Cout, it.Cin, it.RS), torch.float32)
        if it.layout == 1:
            acc = _at(it.acc, (it.Cout, it.RS, it.Cin), torch.float32)
            dw += acc.permute(0, 2, 1)
        else:
            acc = _at(it.acc, (it.RS, it.Cin, it.Cout), torch.float32)
            dw += acc.permute(2, 1, 0)
        acc.zero_()



# ---------------------------------------------------------------------------------------------------------------------
# b2b video backbone (csrc/jit.cu): ops_jit.py calls the C ABI directly, so these doubles stand in for its public
# functions (differentiable torch expressions, bf16 where the kernels store bf16); nets_jit / trainer_b2b are what runs.
# ---------------------------------------------------------------------------------------------------------------------
def _j_rmsnorm_mod(x, w, shift=None, scale=None, eps=1e-6):
    xf = x.float()
    y = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))
    if scale is not None:
        n, c = x.shape[0], x.shape[-1]
        assert tuple(scale.shape) == (n, c) and tuple(shift.shape) == (n, c) and scale.dtype == torch.float32
        y = y * (1.0 + scale[:, None, None, :]) + shift[:, None, None, :]
    return y.to(BF)


def _rotate_half(x):
    x1, x2 = x.reshape(*x.shape[:-1], -1, 2).unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(x.shape)


def _j_qknorm_rope(qkv, wq, wk, cos, sin, heads, eps=1e-6):
    n, t, _, c3 = qkv.shape
    d = c3 // 3
    hd = d // heads
    assert tuple(cos.shape) == (t, hd) and tuple(sin.shape) == (t, hd), (cos.shape, (t, hd))
    out = []
    for part, w in ((qkv[..., :d], wq), (qkv[..., d:2 * d], wk)):
        z = part.float().reshape(n, t, heads, hd)
        z = w * (z * torch.rsqrt(z.pow(2).mean(-1, keepdim=True) + eps))
        z = z * cos[None, :, None, :] + _rotate_half(z) * sin[None, :, None, :]
        out.append(z.reshape(n, t, 1, d))
    return torch.cat(out, dim=-1).to(BF)


def _j_attn_small(qk, qkv, heads):
    n, t, _, d2 = qk.shape
    d = d2 // 2
    hd = d // heads
    q = qk[..., :d].float().reshape(n, t, heads, hd)
    k = qk[..., d:].float().reshape(n, t, heads, hd)
    v = qkv[..., 2 * d:].float().reshape(n, t, heads, hd)
    p = torch.softmax(torch.einsum("nthc,nshc->nhts", q, k) / (hd ** 0.5), dim=-1)
    return torch.einsum("nhts,nshc->nthc", p, v).reshape(n, t, 1, d).to(BF)


def _j_swiglu(x):
    h = x.shape[-1] // 2
    return (F.silu(x[..., :h].float()) * x[..., h:].float()).to(BF)


def _j_gated_residual(x, y, gate):
    n, c = x.shape[0], x.shape[-1]
    assert tuple(gate.shape) == (n, c) and gate.dtype == torch.float32
    return (x.float() + gate[:, None, None, :] * y.float()).to(BF)


# ---------------------------------------------------------------------------------------------------------------------
# CUT contrastive path (csrc/nce.cu)
# ---------------------------------------------------------------------------------------------------------------------
def gather_rows(feat, ids):
    _check_rows(feat)
    b, h, w, c = feat.shape
    assert ids.dtype == torch.int64 and ids.dim() == 1
    return feat.reshape(b, h * w, c)[:, ids].reshape(b * ids.numel(), c).contiguous()


def gather_rows_bwd(d_out, ids, shape):
    b, h, w, c = shape
    d = torch.zeros((b, h * w, c), dtype=BF)
    d[:, ids] = d_out.reshape(b, ids.numel(), c).to(BF)   # the positions are distinct
    return d.reshape(b, h, w, c)


def l2norm_fwd(x, eps=1e-7):
    assert x.dtype == BF and x.dim() == 2
    xf = x.float()
    norms = xf.norm(dim=1)
    return xf / norms.clamp_min(eps)[:, None], norms


def l2norm_bwd(y, dy, norms, eps=1e-7):
    return ((dy - y * (y * dy).sum(1, keepdim=True)) / norms.clamp_min(eps)[:, None]).to(BF)


def _nce(q, k, groups, temperature, kind, num_patches_opt=256):
    """oracle/cut_oracle.patch_nce_loss restates PatchNCELoss / MoNCELoss (pinned on the reference's cut_plumbing*.pt)"""
    from oracle import cut_oracle as C
    return C.patch_nce_loss(q, k, groups, T=temperature, all_negatives_from_minibatch=False, kind=kind,
                            num_patches_opt=num_patches_opt)


def patch_nce_fwd(q, k, groups, temperature):
    loss = _nce(q, k, groups, temperature, "patchnce")
    return loss, torch.zeros_like(loss)


def _nce_bwd(q, k, grad_loss, groups, temperature, kind, num_patches_opt, need_dq, need_dk):
    q, k = q.detach().requires_grad_(True), k.detach().requires_grad_(True)
    with torch.enable_grad():
        loss = _nce(q, k, groups, temperature, kind, num_patches_opt)
    dq, dk = torch.autograd.grad(loss, [q, k], grad_loss)
    return (dq if need_dq else None), (dk if need_dk else None)


def patch_nce_bwd(q, k, lse, grad_loss, groups, temperature, need_dq=True, need_dk=True):
    return _nce_bwd(q, k, grad_loss, groups, temperature, "patchnce", 256, need_dq, need_dk)


def monce_fwd(q, k, groups, temperature, num_patches_opt, iters=50):
    assert iters == 50
    loss = _nce(q, k, groups, temperature, "monce", num_patches_opt)
    return loss, torch.zeros_like(loss), torch.zeros(1)


def monce_bwd(q, k, lse, grad_loss, ws, groups, temperature, num_patches_opt, iters=50, need_dk=True):
    return _nce_bwd(q, k, grad_loss, groups, temperature, "monce", num_patches_opt, True, need_dk)



def mask_class_dropout(mask, drop_u, prob, fill):
    """palette_model.py:565-584: samples with drop_u[n] < prob get the unconditioned class `fill` everywhere"""
    mask = mask.contiguous()
    drop = (drop_u.float() < prob).view(-1, *([1] * (mask.dim() - 1)))
    return torch.where(drop, torch.full_like(mask, fill), mask)



def haar(x, mode):
    """mode 0: DWT forward, 1: its backward, 2: IWT forward, 3: its backward (oracle/prep_oracle.py restates
    freq_utils.HaarTransform / InverseHaarTransform, pinned on the reference's prep_small.pt)"""
    from oracle import prep_oracle as P
    x = x.contiguous().float()
    if mode == 0:
        return P.haar_dwt(x)
    if mode == 2:
        return P.haar_iwt(x)
    n, c, h, w = x.shape
    src = torch.zeros((n, c // 4, 2 * h, 2 * w) if mode == 1 else (n, 4 * c, h // 2, w // 2), requires_grad=True)
    with torch.enable_grad():
        y = P.haar_dwt(src) if mode == 1 else P.haar_iwt(src)
    (d,) = torch.autograd.grad(y, src, x)
    return d


_JIT_DOUBLES = dict(rmsnorm_mod=_j_rmsnorm_mod, qknorm_rope=_j_qknorm_rope, attn_small=_j_attn_small, swiglu=_j_swiglu,
                    gated_residual=_j_gated_residual)


_DOUBLES = dict(pack_conv_weight=pack_conv_weight, conv2d_fwd=conv2d_fwd, chan_stats=chan_stats,
                conv2d_cropped=conv2d_cropped, conv2d_wgrad=conv2d_wgrad, conv2d_wgrad_acc=conv2d_wgrad_acc,
                bias_grad=bias_grad, nchw_to_nhwc=nchw_to_nhwc, nhwc_to_nchw=nhwc_to_nchw, copy_channels=copy_channels,
                resample2x=resample2x, groupnorm_fwd=groupnorm_fwd, groupnorm_bwd=groupnorm_bwd, attn_fwd=attn_fwd,
                attn_bwd=attn_bwd, linear_fwd=linear_fwd, linear_bwd=linear_bwd, LinearBank=LinearBank,
                linear_batched_fwd=linear_batched_fwd, linear_batched_bwd=linear_batched_bwd, noise_pack=noise_pack,
                palette_loss_fwd=palette_loss_fwd, palette_loss_bwd=palette_loss_bwd, adamw_ema_step=adamw_ema_step,
                pad2d=pad2d, pad2d_bwd=pad2d_bwd, dilate2x=dilate2x, undilate2x=undilate2x, act_bwd=act_bwd,
                gan_loss_fwd=gan_loss_fwd, gan_loss_bwd=gan_loss_bwd, layernorm_fwd=layernorm_fwd,
                layernorm_bwd=layernorm_bwd, temporal_attn_fwd=temporal_attn_fwd, temporal_attn_bwd=temporal_attn_bwd,
                geglu_fwd=geglu_fwd, geglu_bwd=geglu_bwd, embed_rows=embed_rows, embed_rows_bwd=embed_rows_bwd,
                ddpm_step=ddpm_step, WeightTable=WeightTable, pack_conv_weights_batched=pack_conv_weights_batched,
                wgrad_unpack_batched=wgrad_unpack_batched, gather_rows=gather_rows, gather_rows_bwd=gather_rows_bwd,
                l2norm_fwd=l2norm_fwd, l2norm_bwd=l2norm_bwd, patch_nce_fwd=patch_nce_fwd, patch_nce_bwd=patch_nce_bwd,
                monce_fwd=monce_fwd, monce_bwd=monce_bwd, mask_class_dropout=mask_class_dropout, haar=haar)


def _refuse(name):
    def f(*a, **k):
        raise AssertionError("kernel double: kernels.%s has no CPU restatement (host test reached an unexpected op)"
                             % name)
    return f


@contextlib.contextmanager
def installed():
    """Swap every public callable of joligen_b200.kernels: the restated ones by their doubles, the rest by a refusal
    (so that nothing can reach the CUDA library by accident)."""
    from joligen_b200 import nets
    saved = {}
    device_only, nets._DEVICE_WEIGHTS_ONLY[0] = nets._DEVICE_WEIGHTS_ONLY[0], False
    keep = {"conv_out_size", "make_conv_desc", "_ld", "_device_table", "_mask_ptrs"}
    for name, obj in list(vars(K).items()):
        if name.startswith("__") or name in keep or not (callable(obj)) or getattr(obj, "__module__", "") != K.__name__:
            continue
        saved[name] = obj
        setattr(K, name, _DOUBLES.get(name, _refuse(name)))
    from joligen_b200 import ops_jit
    saved_jit = {name: getattr(ops_jit, name) for name in _JIT_DOUBLES}
    for name, fn in _JIT_DOUBLES.items():
        setattr(ops_jit, name, fn)
    try:
        yield
    finally:
        nets._DEVICE_WEIGHTS_ONLY[0] = device_only
        for name, obj in saved.items():
            setattr(K, name, obj)
        for name, obj in saved_jit.items():
            setattr(ops_jit, name, obj)
