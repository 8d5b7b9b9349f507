"""Autograd wrappers of the b2b video backbone's memory-bound kernels (csrc/jit.cu; the reference classes they replace
are cited in include/jg_b200.h).  Token tensors are bf16 [N, T, 1, C] (NHWC with H = T, W = 1: the layout the 1x1
tcgen05 convolutions take), modulation vectors fp32 [N, C] views of the adaLN Linear's output."""
import torch

from . import lib as L


def _ld(t):
    return t.stride(-2) if t.dim() == 4 else t.stride(0)


def _ldm(v):
    return v.stride(0)


def _tok(t):
    n, tt = t.shape[0], t.shape[1]
    return n, tt, n * tt


class RmsNormModFn(torch.autograd.Function):
    """y = w * x / rms(x) [* (1 + scale) + shift]  (util/model_util.py:165-179 + vit_vid.py:47-48)"""

    @staticmethod
    def forward(ctx, x, w, shift, scale, eps):
        n, t, rows = _tok(x)
        c = x.shape[-1]
        y = torch.empty_like(x)
        rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
        L.call("jg_rmsnorm_mod_fwd", L.ptr(x), _ld(x), L.ptr(y), _ld(y), rows, c, t, float(eps), L.ptr(w), L.ptr(shift),
               L.ptr(scale), _ldm(scale) if scale is not None else 0, L.ptr(rstd), L.stream())
        ctx.save_for_backward(x, w, scale, rstd)
        ctx.has_mod = scale is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, scale, rstd = ctx.saved_tensors
        n, t, rows = _tok(x)
        c = x.shape[-1]
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        dshift = dscale = None
        if ctx.has_mod:
            dshift = torch.empty((n, c), dtype=torch.float32, device=x.device)
            dscale = torch.empty((n, c), dtype=torch.float32, device=x.device)
        L.call("jg_rmsnorm_mod_bwd", L.ptr(x), _ld(x), L.ptr(dy), _ld(dy), L.ptr(dx), _ld(dx), rows, c, t, L.ptr(w),
               L.ptr(scale), _ldm(scale) if ctx.has_mod else 0, L.ptr(rstd), L.ptr(dw), L.ptr(dshift), L.ptr(dscale), c,
               L.stream())
        return dx, dw, dshift, dscale, None


def rmsnorm_mod(x, w, shift=None, scale=None, eps=1e-6):
    return RmsNormModFn.apply(x, w, shift, scale, eps)


class QkNormRopeFn(torch.autograd.Function):
    """(q | k) parts of qkv -> per-head RMSNorm + rotary (vit_vid.py:205-231); returns [N, T, 1, 2D]."""

    @staticmethod
    def forward(ctx, qkv, wq, wk, cos, sin, heads, eps):
        n, t, rows = _tok(qkv)
        d = qkv.shape[-1] // 3
        hd = d // heads
        out = torch.empty((n, t, 1, 2 * d), dtype=torch.bfloat16, device=qkv.device)
        rstd = torch.empty((rows * heads * 2,), dtype=torch.float32, device=qkv.device)
        L.call("jg_qknorm_rope_fwd", L.ptr(qkv), _ld(qkv), L.ptr(out), _ld(out), rows, t, heads, hd, float(eps), L.ptr(wq),
               L.ptr(wk), L.ptr(cos), L.ptr(sin), L.ptr(rstd), L.stream())
        ctx.save_for_backward(qkv, wq, wk, cos, sin, rstd)
        ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, wq, wk, cos, sin, rstd = ctx.saved_tensors
        n, t, rows = _tok(qkv)
        d = qkv.shape[-1] // 3
        hd = d // ctx.heads
        dout = dout.contiguous()
        # gradient w.r.t. the whole qkv tensor: the v third is zero here (the attention op returns its own dv)
        dqkv = torch.zeros_like(qkv)
        dwq = torch.empty_like(wq)
        dwk = torch.empty_like(wk)
        L.call("jg_qknorm_rope_bwd", L.ptr(qkv), _ld(qkv), L.ptr(dout), _ld(dout), L.ptr(dqkv), _ld(dqkv), rows, t, ctx.heads,
               hd, L.ptr(wq), L.ptr(wk), L.ptr(cos), L.ptr(sin), L.ptr(rstd), L.ptr(dwq), L.ptr(dwk), L.stream())
        return dqkv, dwq, dwk, None, None, None, None


def qknorm_rope(qkv, wq, wk, cos, sin, heads, eps=1e-6):
    return QkNormRopeFn.apply(qkv, wq, wk, cos, sin, heads, eps)


class AttnSmallFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(hd)) v per frame; qk [N,T,1,2D] (rotated q | k), qkv [N,T,1,3D] (v = last third)."""

    @staticmethod
    def forward(ctx, qk, qkv, heads):
        n, t, rows = _tok(qk)
        d = qk.shape[-1] // 2
        hd = d // heads
        o = torch.empty((n, t, 1, d), dtype=torch.bfloat16, device=qk.device)
        lse = torch.empty((n * heads * t,), dtype=torch.float32, device=qk.device)
        q, k, v = qk[..., :d], qk[..., d:], qkv[..., 2 * d:]
        L.call("jg_attn_small_fwd", L.ptr(q), _ld(qk), L.ptr(k), _ld(qk), L.ptr(v), _ld(qkv), L.ptr(o), _ld(o), L.ptr(lse),
               n, t, heads, hd, L.stream())
        ctx.save_for_backward(qk, qkv, o, lse)
        ctx.heads = heads
        return o

    @staticmethod
    def backward(ctx, d_o):
        qk, qkv, o, lse = ctx.saved_tensors
        n, t, rows = _tok(qk)
        d = qk.shape[-1] // 2
        hd = d // ctx.heads
        d_o = d_o.contiguous()
        dqk = torch.empty_like(qk)
        dqkv = torch.zeros_like(qkv)
        q, k, v = qk[..., :d], qk[..., d:], qkv[..., 2 * d:]
        L.call("jg_attn_small_bwd", L.ptr(q), _ld(qk), L.ptr(k), _ld(qk), L.ptr(v), _ld(qkv), L.ptr(o), _ld(o), L.ptr(d_o),
               _ld(d_o), L.ptr(lse), L.ptr(dqk[..., :d]), _ld(dqk), L.ptr(dqk[..., d:]), _ld(dqk),
               L.ptr(dqkv[..., 2 * d:]), _ld(dqkv), n, t, ctx.heads, hd, L.stream())
        return dqk, dqkv, None


def attn_small(qk, qkv, heads):
    return AttnSmallFn.apply(qk, qkv, heads)


class SwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, t, rows = _tok(x)
        h = x.shape[-1] // 2
        y = torch.empty((n, t, 1, h), dtype=torch.bfloat16, device=x.device)
        L.call("jg_swiglu_fwd", L.ptr(x), _ld(x), L.ptr(y), _ld(y), rows, h, L.stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, t, rows = _tok(x)
        h = x.shape[-1] // 2
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        L.call("jg_swiglu_bwd", L.ptr(x), _ld(x), L.ptr(dy), _ld(dy), L.ptr(dx), _ld(dx), rows, h, L.stream())
        return dx


def swiglu(x):
    return SwigluFn.apply(x)


class GatedResidualFn(torch.autograd.Function):
    """out = x + gate * y  (vit_vid.py:270-279); gate fp32 [N, C] (a view of the adaLN output)."""

    @staticmethod
    def forward(ctx, x, y, gate):
        n, t, rows = _tok(x)
        c = x.shape[-1]
        out = torch.empty_like(x)
        L.call("jg_gated_residual_fwd", L.ptr(x), _ld(x), L.ptr(y), _ld(y), L.ptr(gate), _ldm(gate), L.ptr(out), _ld(out), rows,
               c, t, L.stream())
        ctx.save_for_backward(y, gate)
        return out

    @staticmethod
    def backward(ctx, d):
        y, gate = ctx.saved_tensors
        n, t, rows = _tok(y)
        c = y.shape[-1]
        d = d.contiguous()
        dy = torch.empty_like(y)
        dgate = torch.empty((n, c), dtype=torch.float32, device=y.device)
        L.call("jg_gated_residual_bwd", L.ptr(d), _ld(d), L.ptr(y), _ld(y), L.ptr(gate), _ldm(gate), L.ptr(dy), _ld(dy),
               L.ptr(dgate), c, rows, c, t, L.stream())
        return d, dy, dgate


def gated_residual(x, y, gate):
    return GatedResidualFn.apply(x, y, gate)
