"""Thin, allocation-only Python wrappers over the C ABI (no autograd here; see ops.py).

Activations are NHWC bf16 torch tensors (shape [N, H, W, C], possibly channel slices of a wider
buffer: stride(-1) == 1, stride(-2) == ld).
"""
import ctypes

import torch

from . import lib as L


def _ld(t):
    assert t.stride(-1) == 1, "channel dim must be contiguous"
    ld = t.stride(-2)
    n, h, w, _ = t.shape
    assert t.stride(1) == w * ld and (n == 1 or t.stride(0) == h * w * ld), "not an NHWC channel slice"
    return ld


def conv_out_size(h, w, r, s, stride, pad):
    return (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1


def make_conv_desc(x, cout, r, s, stride=1, pad=None, ldy=None, act=L.ACT_NONE, ldres=0, res_scale=1.0):
    n, h, w, cin = x.shape
    if pad is None:
        pad = (r - 1) // 2
    ho, wo = conv_out_size(h, w, r, s, stride, pad)
    d = L.ConvDesc()
    d.N, d.H, d.W, d.Cin, d.ldx = n, h, w, cin, _ld(x)
    d.Ho, d.Wo, d.Cout, d.ldy = ho, wo, cout, (ldy if ldy else cout)
    d.R, d.S, d.stride, d.pad, d.up2x, d.act = r, s, stride, pad, 0, act
    d.ldres, d.res_scale = ldres, res_scale
    return d


def pack_conv_weight(w_oihw, want_dgrad=True):
    """fp32 OIHW -> (bf16 [Cout][RS][Cin8], bf16 [Cin][RS][Cout8] or None)."""
    cout, cin, r, s = w_oihw.shape
    w = w_oihw.detach().contiguous().float()
    cin8, cout8 = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    wf = torch.empty((cout, r * s, cin8), dtype=torch.bfloat16, device=w.device)
    wd = torch.empty((cin, r * s, cout8), dtype=torch.bfloat16, device=w.device) if want_dgrad else None
    L.call("jg_pack_conv_weight", L.ptr(w), L.ptr(wf), L.ptr(wd), cout, cin, r, s, L.stream())
    return wf, wd


def conv2d_fwd(x, w_packed, bias, cout, r, s, stride=1, pad=None, act=L.ACT_NONE, residual=None, res_scale=1.0,
               out=None):
    """x: NHWC bf16 (C multiple of 8). Returns NHWC bf16 [N,Ho,Wo,cout]; `out` may be a channel slice."""
    n, h, w, _ = x.shape
    if pad is None:
        pad = (r - 1) // 2
    ho, wo = conv_out_size(h, w, r, s, stride, pad)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    d = make_conv_desc(x, cout, r, s, stride, pad, ldy=_ld(out), act=act,
                       ldres=(_ld(residual) if residual is not None else 0), res_scale=res_scale)
    L.call("jg_conv2d_fwd", ctypes.byref(d), L.ptr(x), L.ptr(w_packed), L.ptr(bias), L.ptr(residual), L.ptr(out),
           L.stream())
    return out


def conv2d_wgrad(x, dy, cout, r, s, stride=1, pad=None):
    """Returns fp32 OIHW weight gradient [cout, Cin, r, s] (Cin = x channels)."""
    cin = x.shape[-1]
    d = make_conv_desc(x, cout, r, s, stride, pad)
    acc = torch.zeros((cout, r * s, cin), dtype=torch.float32, device=x.device)
    L.call("jg_conv2d_wgrad", ctypes.byref(d), L.ptr(x), L.ptr(dy), _ld(dy), L.ptr(acc), L.stream())
    out = torch.empty((cout, cin, r, s), dtype=torch.float32, device=x.device)
    L.call("jg_unpack_conv_wgrad", L.ptr(acc), L.ptr(out), cout, cin, r, s, 0.0, L.stream())
    return out


def bias_grad(dy):
    c = dy.shape[-1]
    rows = dy.numel() // c if dy.is_contiguous() else dy.shape[0] * dy.shape[1] * dy.shape[2]
    db = torch.empty((c,), dtype=torch.float32, device=dy.device)
    L.call("jg_bias_grad", L.ptr(dy), rows, c, _ld(dy), L.ptr(db), L.stream())
    return db


def nchw_to_nhwc(x, ld=None):
    """fp32 NCHW -> bf16 NHWC with channels padded (zeros) to `ld` (default: round up to 8)."""
    n, c, h, w = x.shape
    if ld is None:
        ld = (c + 7) // 8 * 8
    x = x.contiguous().float()
    out = torch.empty((n, h, w, ld), dtype=torch.bfloat16, device=x.device)
    L.call("jg_nchw_f32_to_nhwc_bf16", L.ptr(x), L.ptr(out), n, c, h, w, ld, L.stream())
    return out


def nhwc_to_nchw(x, c=None):
    """bf16 NHWC (channel stride ld) -> fp32 NCHW with the first c channels."""
    n, h, w, cc = x.shape
    if c is None:
        c = cc
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    L.call("jg_nhwc_bf16_to_nchw_f32", L.ptr(x), L.ptr(out), n, c, h, w, _ld(x), L.stream())
    return out


def copy_channels(src, dst, accumulate=False):
    n, h, w, c = src.shape
    assert dst.shape == src.shape
    L.call("jg_copy_channels", L.ptr(src), _ld(src), L.ptr(dst), _ld(dst), n * h * w, c, int(accumulate), L.stream())
    return dst


def resample2x(x, mode):
    n, h, w, c = x.shape
    up = mode in (0, 3)
    ho, wo = (h * 2, w * 2) if up else (h // 2, w // 2)
    out = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.device)
    L.call("jg_resample2x", L.ptr(x), _ld(x), L.ptr(out), c, n, h, w, c, mode, L.stream())
    return out
