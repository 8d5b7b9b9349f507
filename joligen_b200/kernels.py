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


def pack_conv_weight(w_oihw, want_dgrad=True, out=None):
    """fp32 OIHW -> (bf16 [Cout8][RS][Cin8], bf16 [Cin8][RS][Cout8] or None); out = (wf, wd) buffers to refill."""
    cout, cin, r, s = w_oihw.shape
    w = w_oihw.detach().contiguous().float()
    cin8, cout8 = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    if out is not None:
        wf, wd = out
    else:
        # rows cout..cout8-1 of wf stay zero (the kernels see round_up(Cout, 8) output channels)
        alloc = torch.zeros if cout8 != cout else torch.empty
        wf = alloc((cout8, r * s, cin8), dtype=torch.bfloat16, device=w.device)
        alloc_d = torch.zeros if cin8 != cin else torch.empty
        wd = alloc_d((cin8, r * s, cout8), dtype=torch.bfloat16, device=w.device) if want_dgrad else None
    L.call("jg_pack_conv_weight", L.ptr(w), L.ptr(wf), L.ptr(wd), cout, cin, r, s, L.stream())
    return wf, wd


def conv2d_fwd(x, w_packed, bias, cout, r, s, stride=1, pad=None, act=L.ACT_NONE, residual=None, res_scale=1.0,
               out=None, stats=None, gn=None):
    """x: NHWC bf16 (C multiple of 8). Returns NHWC bf16 [N,Ho,Wo,cout]; `out` may be a channel slice.
    GroupNorm work fused into the epilogue (jg_conv_epilogue):
      stats: fp32 [N, cout, 2] ZEROED buffer that receives the per-(image, channel) sum / sum of squares of the output;
      gn = (x_gn, ab, act, sums): this call is the dgrad of the conv that follows act(GN(x_gn)): sums (fp32 [N, cout, 2],
           zeroed) receives (sum du, sum du * x_gn) with du = out * act'(a * x_gn + b)."""
    n, h, w, _ = x.shape
    if pad is None:
        pad = (r - 1) // 2
    ho, wo = conv_out_size(h, w, r, s, stride, pad)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    d = make_conv_desc(x, cout, r, s, stride, pad, ldy=_ld(out), act=act,
                       ldres=(_ld(residual) if residual is not None else 0), res_scale=res_scale)
    if stats is None and gn is None:
        L.call("jg_conv2d_fwd", ctypes.byref(d), L.ptr(x), L.ptr(w_packed), L.ptr(bias), L.ptr(residual), L.ptr(out),
               L.stream())
        return out
    e = L.ConvEpilogue()
    if stats is not None:
        assert stats.is_contiguous() and stats.dtype == torch.float32 and stats.numel() == n * cout * 2
        e.stats = stats.data_ptr()
    if gn is not None:
        x_gn, ab, gn_act, sums = gn
        assert tuple(x_gn.shape) == (n, ho, wo, cout) and ab.numel() == n * cout * 2 and sums.numel() == n * cout * 2
        e.gn_sums, e.gn_x, e.ldgx, e.gn_ab, e.gn_act = sums.data_ptr(), x_gn.data_ptr(), _ld(x_gn), ab.data_ptr(), gn_act
    L.call("jg_conv2d_fwd_ex", ctypes.byref(d), ctypes.byref(e), L.ptr(x), L.ptr(w_packed), L.ptr(bias),
           L.ptr(residual), L.ptr(out), L.stream())
    return out


def chan_stats(x):
    """fp32 [N, C, 2]: per-(image, channel) sum and sum of squares of an NHWC bf16 tensor (stand-alone pass)."""
    n, h, w, c = x.shape
    stats = torch.zeros((n, c, 2), dtype=torch.float32, device=x.device)
    L.call("jg_chan_stats", L.ptr(x), _ld(x), n, h * w, c, L.ptr(stats), L.stream())
    return stats


def conv2d_cropped(x, w_packed, bias, cout, r, s, pad, out_hw, act=L.ACT_NONE):
    """stride-1 conv whose output is cropped to out_hw (<= the full correlation size): the transposed
    convolution / stride-2 dgrad after zero insertion."""
    n, h, w, cin = x.shape
    ho, wo = out_hw
    out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    d = L.ConvDesc()
    d.N, d.H, d.W, d.Cin, d.ldx = n, h, w, cin, _ld(x)
    d.Ho, d.Wo, d.Cout, d.ldy = ho, wo, cout, cout
    d.R, d.S, d.stride, d.pad, d.up2x, d.act = r, s, 1, pad, 0, act
    d.ldres, d.res_scale = 0, 1.0
    L.call("jg_conv2d_fwd", ctypes.byref(d), L.ptr(x), L.ptr(w_packed), L.ptr(bias), 0, L.ptr(out), L.stream())
    return out


def conv2d_wgrad(x, dy, cout, r, s, stride=1, pad=None, out=None, beta=0.0):
    """fp32 OIHW weight gradient [cout, Cin, r, s] (Cin = x channels): out = beta*out + dW."""
    cin = x.shape[-1]
    d = make_conv_desc(x, cout, r, s, stride, pad)
    ws = torch.empty((cout * r * s * cin,), dtype=torch.float32, device=x.device)
    if out is None:
        out = torch.empty((cout, cin, r, s), dtype=torch.float32, device=x.device)
        beta = 0.0
    L.call("jg_conv2d_wgrad", ctypes.byref(d), L.ptr(x), L.ptr(dy), _ld(dy), L.ptr(ws), L.ptr(out), float(beta),
           L.stream())
    return out


def conv2d_wgrad_acc(x, dy, cout, r, s, acc, stride=1, pad=None):
    """Raw split-K accumulation into the persistent fp32 accumulator `acc` (flat, R*S*Cin*Cout).  Returns the layout
    code of the accumulator (0: [RS][Cin][Cout], 1: [Cout][RS][Cin])."""
    if pad is None:
        pad = (r - 1) // 2
    d = make_conv_desc(x, cout, r, s, stride, pad)
    layout = ctypes.c_int(-1)
    L.call("jg_conv2d_wgrad_acc", ctypes.byref(d), L.ptr(x), L.ptr(dy), _ld(dy), L.ptr(acc), ctypes.byref(layout),
           L.stream())
    return layout.value


def _device_table(structs, device):
    """ctypes struct array -> uint8 device tensor (kept alive by the caller)."""
    arr = (type(structs[0]) * len(structs))(*structs)
    raw = bytes(arr)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


class WeightTable:
    """Device-side item table + tile prefix sums for the batched pack / unpack launches."""

    def __init__(self, items, dims, device):
        lib = L.load()
        starts, tot = [], 0
        for (cout, cin, rs) in dims:
            starts.append(tot)
            tot += lib.jg_weight_tiles(cout, cin, rs)
        self.n = len(items)
        self.total_tiles = tot
        self.items = _device_table(items, device)
        self.tile_start = torch.tensor(starts, dtype=torch.int32, device=device)


def pack_conv_weights_batched(table):
    L.call("jg_pack_conv_weights_batched", L.ptr(table.items), L.ptr(table.tile_start), table.n, table.total_tiles,
           L.stream())


def wgrad_unpack_batched(table):
    L.call("jg_wgrad_unpack_batched", L.ptr(table.items), L.ptr(table.tile_start), table.n, table.total_tiles,
           L.stream())


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


# ---------------------------------------------------------------------------------------------
# GroupNorm (+FiLM)(+SiLU)
# ---------------------------------------------------------------------------------------------
def groupnorm_fwd(x, gamma, beta, groups, film=None, act=L.ACT_NONE, eps=1e-5, out=None, chan_stats=None):
    """x NHWC bf16.  Returns (y, stats[N,G,2], ab[N,C,2]).  chan_stats: fp32 [N,C,2] per-channel sums of x that its
    producer already accumulated (conv2d_fwd(stats=...)): the statistics pass is skipped."""
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((n, groups, 2), dtype=torch.float32, device=x.device)
    ab = torch.empty((n, c, 2), dtype=torch.float32, device=x.device)
    ws = torch.empty((L.load().jg_groupnorm_fwd_ws_floats(n, c, groups),), dtype=torch.float32, device=x.device)
    if chan_stats is not None:
        assert chan_stats.is_contiguous() and chan_stats.numel() == n * c * 2 and chan_stats.dtype == torch.float32
    L.call("jg_groupnorm_fwd", L.ptr(x), _ld(x), L.ptr(out), _ld(out), n, h * w, c, groups, eps, L.ptr(gamma),
           L.ptr(beta), L.ptr(film), act, L.ptr(stats), L.ptr(ab), L.ptr(ws), L.ptr(chan_stats), L.stream())
    return out, stats, ab


def groupnorm_bwd(x, dy, gamma, beta, groups, film, act, stats, ab, need_param_grads=True, need_film_grad=False,
                  dx=None, addend=None, colsum=None, addend2=None, sums_pre=None, dfilm_out=None):
    """addend: optional NHWC bf16 tensor summed into dx in the same pass (gradient of another consumer of x).
    colsum: optional fp32 [C] output = per-channel sum of dx (the bias gradient of the conv that produced x).
    sums_pre: fp32 [N,C,2] = (sum du, sum du*x) already accumulated by the dgrad that produced dy
    (conv2d_fwd(gn=...)): the reduction pass over (x, dy) is skipped."""
    n, h, w, c = x.shape
    if dx is None:
        dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    dgamma = torch.empty((c,), dtype=torch.float32, device=x.device) if need_param_grads else None
    dbeta = torch.empty((c,), dtype=torch.float32, device=x.device) if need_param_grads else None
    dfilm = None
    if need_film_grad:
        ok = dfilm_out is not None and dfilm_out.is_contiguous() and tuple(dfilm_out.shape) == (n, 2 * c)
        dfilm = dfilm_out if ok else torch.empty((n, 2 * c), dtype=torch.float32, device=x.device)
    ws = torch.empty((L.load().jg_groupnorm_bwd_ws_floats(n, c, groups),), dtype=torch.float32, device=x.device)
    L.call("jg_groupnorm_bwd", L.ptr(x), _ld(x), L.ptr(dy), _ld(dy), L.ptr(dx), _ld(dx), L.ptr(addend),
           _ld(addend) if addend is not None else 0, L.ptr(addend2), _ld(addend2) if addend2 is not None else 0,
           n, h * w, c,
           groups, L.ptr(gamma), L.ptr(beta), L.ptr(film), act, L.ptr(stats), L.ptr(ab), L.ptr(dgamma), L.ptr(dbeta),
           L.ptr(dfilm), L.ptr(colsum), L.ptr(ws), L.ptr(sums_pre), L.stream())
    return dx, dgamma, dbeta, dfilm


# ---------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------
def attn_fwd(qkv, heads, ch, layout=0, out=None):
    """qkv NHWC [N,H,W,3*heads*ch]; layout 0 = legacy per-head q|k|v interleave, 1 = (q | k | v) chunks.
    Returns (out [N,H,W,heads*ch], lse); `out` may be a channel slice of a wider buffer."""
    n, h, w, _ = qkv.shape
    t = h * w
    if out is None:
        out = torch.empty((n, h, w, heads * ch), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((n * heads, t), dtype=torch.float32, device=qkv.device)
    L.call("jg_attn_fwd", L.ptr(qkv), _ld(qkv), L.ptr(out), _ld(out), L.ptr(lse), n, t, heads, ch, layout, L.stream())
    return out, lse


def attn_bwd(qkv, out, d_out, lse, heads, ch, layout=0):
    n, h, w, c3 = qkv.shape
    t = h * w
    dqkv = torch.empty((n, h, w, c3), dtype=torch.bfloat16, device=qkv.device)
    ws = torch.empty((n * heads * t,), dtype=torch.float32, device=qkv.device)
    L.call("jg_attn_bwd", L.ptr(qkv), _ld(qkv), L.ptr(out), _ld(out), L.ptr(d_out), _ld(d_out), L.ptr(lse),
           L.ptr(dqkv), _ld(dqkv), L.ptr(ws), n, t, heads, ch, layout, L.stream())
    return dqkv


# ---------------------------------------------------------------------------------------------
# MotionModule kernels (video UNet): LayerNorm(+PE), temporal attention, GEGLU
# ---------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps=1e-5, pe=None, frames=1):
    """x NHWC bf16 [N,H,W,C] (N = B*frames).  pe: fp32 [frames, C] added after the affine.  Returns (y, stats)."""
    n, h, w, c = x.shape
    rows = n * h * w
    y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
    L.call("jg_layernorm_fwd", L.ptr(x), _ld(x), L.ptr(y), _ld(y), rows, c, eps, L.ptr(gamma), L.ptr(beta), L.ptr(pe),
           h * w, frames, L.ptr(stats), L.stream())
    return y, stats


def layernorm_bwd(x, dy, gamma, stats, addend=None, colsum=None):
    """-> (dx, dgamma, dbeta); dx += addend (bf16 rows); colsum ([C] fp32 buffer, optional) = column sums of dx."""
    n, h, w, c = x.shape
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    dgamma = torch.empty((c,), dtype=torch.float32, device=x.device)
    dbeta = torch.empty((c,), dtype=torch.float32, device=x.device)
    L.call("jg_layernorm_bwd", L.ptr(x), _ld(x), L.ptr(dy), _ld(dy), L.ptr(dx), _ld(dx), L.ptr(addend),
           0 if addend is None else _ld(addend), n * h * w, c, L.ptr(gamma), L.ptr(stats), L.ptr(dgamma),
           L.ptr(dbeta), L.ptr(colsum), L.stream())
    return dx, dgamma, dbeta


def temporal_attn_fwd(qkv, frames, heads):
    """qkv NHWC [B*frames,H,W,3*C] = (q | k | v).  Returns out [B*frames,H,W,C]."""
    n, h, w, c3 = qkv.shape
    c = c3 // 3
    out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=qkv.device)
    L.call("jg_temporal_attn_fwd", L.ptr(qkv), _ld(qkv), L.ptr(out), _ld(out), n // frames, frames, h * w, heads,
           c // heads, L.stream())
    return out


def temporal_attn_bwd(qkv, d_out, frames, heads):
    n, h, w, c3 = qkv.shape
    c = c3 // 3
    dqkv = torch.empty((n, h, w, c3), dtype=torch.bfloat16, device=qkv.device)
    L.call("jg_temporal_attn_bwd", L.ptr(qkv), _ld(qkv), L.ptr(d_out), _ld(d_out), L.ptr(dqkv), _ld(dqkv),
           n // frames, frames, h * w, heads, c // heads, L.stream())
    return dqkv


def geglu_fwd(x):
    """x NHWC [N,H,W,2*Cout] = (a | gate) -> a * gelu(gate)."""
    n, h, w, c2 = x.shape
    y = torch.empty((n, h, w, c2 // 2), dtype=torch.bfloat16, device=x.device)
    L.call("jg_geglu_fwd", L.ptr(x), _ld(x), L.ptr(y), _ld(y), n * h * w, c2 // 2, L.stream())
    return y


def geglu_bwd(x, dy):
    n, h, w, c2 = x.shape
    dx = torch.empty((n, h, w, c2), dtype=torch.bfloat16, device=x.device)
    L.call("jg_geglu_bwd", L.ptr(x), _ld(x), L.ptr(dy), _ld(dy), L.ptr(dx), _ld(dx), n * h * w, c2 // 2, L.stream())
    return dx


# ---------------------------------------------------------------------------------------------
# small fp32 linear, prologue, loss, optimiser
# ---------------------------------------------------------------------------------------------
def linear_fwd(x, w, b, act_in=L.ACT_NONE, act_out=L.ACT_NONE):
    bsz, i = x.shape
    o = w.shape[0]
    y = torch.empty((bsz, o), dtype=torch.float32, device=x.device)
    L.call("jg_linear_fwd", L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), bsz, i, o, act_in, act_out, L.stream())
    return y


def linear_bwd(x, w, dy, act_in=L.ACT_NONE, need_dx=True):
    bsz, i = x.shape
    o = w.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    db = torch.empty((o,), dtype=torch.float32, device=x.device)
    L.call("jg_linear_bwd", L.ptr(x), L.ptr(w), L.ptr(dy), L.ptr(dx), 0, L.ptr(dw), L.ptr(db), bsz, i, o, act_in,
           L.stream())
    return dx, dw, db


class LinearBank:
    """Device table for the batched Linear launches: n Linears with the same input width applied to the same input."""

    def __init__(self, linears, device):
        lib = L.load()
        self.widths = [lin.weight.shape[0] for lin in linears]
        self.in_features = linears[0].weight.shape[1]
        items, starts, off, tiles = [], [], 0, 0
        for lin in linears:
            w = lin.weight
            assert w.dtype == torch.float32 and w.is_contiguous() and w.shape[1] == self.in_features
            items.append(L.LinearItem(w.data_ptr(), 0 if lin.bias is None else lin.bias.data_ptr(), w.shape[0], off))
            starts.append(tiles)
            tiles += lib.jg_linear_batched_tiles(w.shape[0])
            off += w.shape[0]
        self.offsets = [it.off for it in items]
        self.total_out, self.total_tiles, self.n = off, tiles, len(items)
        self.key = tuple((it.w, it.b) for it in items)
        self.items = _device_table(items, device)
        self.tile_start = torch.tensor(starts, dtype=torch.int32, device=device)


def linear_batched_fwd(x, bank, act_in=L.ACT_NONE):
    """-> Y flat fp32 [B * sum O]; item i is Y[B*off_i : B*(off_i+O_i)].view(B, O_i)."""
    bsz, i = x.shape
    y = torch.empty((bsz * bank.total_out,), dtype=torch.float32, device=x.device)
    L.call("jg_linear_batched_fwd", L.ptr(x), L.ptr(bank.items), L.ptr(bank.tile_start), bank.n, bank.total_tiles,
           L.ptr(y), bsz, i, act_in, L.stream())
    return y


def linear_batched_bwd(x, bank, dy, act_in=L.ACT_NONE, need_dx=True):
    """dy: flat like the forward's Y.  -> (dx [B,I] or None, dW [sum O, I], dB [sum O])"""
    bsz, i = x.shape
    dw = torch.empty((bank.total_out, i), dtype=torch.float32, device=x.device)
    db = torch.empty((bank.total_out,), dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x) if need_dx else None
    L.call("jg_linear_batched_bwd", L.ptr(x), L.ptr(bank.items), L.ptr(bank.tile_start), bank.n, bank.total_tiles,
           L.ptr(dy), L.ptr(dw), L.ptr(db), L.ptr(dx), bsz, i, act_in, L.stream())
    return dx, dw, db


def _mask_ptrs(mask):
    if mask is None:
        return 0, 0
    assert mask.is_contiguous()
    if mask.dtype == torch.int64:
        return 0, mask.data_ptr()
    assert mask.dtype == torch.float32, "mask must be int64 or float32"
    return mask.data_ptr(), 0


def noise_pack(y0, ycond, noise, mask, gammas, ld=8):
    b, c, h, w = y0.shape
    out = torch.empty((b, h, w, ld), dtype=torch.bfloat16, device=y0.device)
    mf, mi = _mask_ptrs(mask)
    L.call("jg_noise_pack_fwd", L.ptr(y0), L.ptr(ycond), L.ptr(noise), mf, mi, L.ptr(gammas), L.ptr(out), b, c, h, w,
           ld, L.stream())
    return out


def ddpm_step(eps, y_t, y_cond, y_0, mask, noise, coef, ld=8, want_next_input=True, ddim=False):
    """One reverse-diffusion step + mask blend + next-input pack (jg_ddpm_step).  Returns (y_next, x_next or None)."""
    b, c, h, w = y_t.shape
    y_next = torch.empty_like(y_t)
    x_next = torch.empty((b, h, w, ld), dtype=torch.bfloat16, device=y_t.device) if want_next_input else None
    mf, mi = _mask_ptrs(mask)
    L.call("jg_ddpm_step", L.ptr(eps), _ld(eps), L.ptr(y_t), L.ptr(y_cond), L.ptr(y_0), mf, mi, L.ptr(noise),
           L.ptr(coef), L.ptr(y_next), L.ptr(x_next), b, c, h, w, ld, int(ddim), L.stream())
    return y_next, x_next


def palette_loss_fwd(noise, noise_hat, mask, w_b, lambda_g=1.0, l1=False):
    b, c, h, w = noise.shape
    loss = torch.empty((), dtype=torch.float32, device=noise.device)
    mf, mi = _mask_ptrs(mask)
    L.call("jg_palette_loss_fwd", L.ptr(noise), L.ptr(noise_hat), _ld(noise_hat), mf, mi, L.ptr(w_b), b, c, h * w,
           float(lambda_g), int(l1), L.ptr(loss), L.stream())
    return loss


def palette_loss_bwd(noise, noise_hat, mask, w_b, grad_out, lambda_g=1.0, l1=False):
    b, c, h, w = noise.shape
    d = torch.empty(noise_hat.shape, dtype=torch.bfloat16, device=noise.device)
    mf, mi = _mask_ptrs(mask)
    L.call("jg_palette_loss_bwd", L.ptr(noise), L.ptr(noise_hat), _ld(noise_hat), mf, mi, L.ptr(w_b), b, c, h * w,
           float(lambda_g), int(l1), L.ptr(grad_out), L.ptr(d), _ld(d), L.stream())
    return d


def adamw_ema_step(p, g, m, v, ema, lr, beta1, beta2, eps, weight_decay, adamw, step, grad_scale=1.0, ema_beta=0.999,
                   ema_init=False, step_dev=None):
    """step_dev: optional int32 device scalar holding the step count (incremented by the call)."""
    L.call("jg_adamw_ema_step", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(ema), p.numel(), float(lr),
           float(beta1), float(beta2), float(eps), float(weight_decay), int(adamw), int(step), L.ptr(step_dev),
           float(grad_scale), float(ema_beta), int(ema_init), L.stream())


# ---------------------------------------------------------------------------------------------
# GAN generator / discriminator helpers
# ---------------------------------------------------------------------------------------------
def pad2d(x, pad, mode=0):
    n, h, w, c = x.shape
    out = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=torch.bfloat16, device=x.device)
    L.call("jg_pad2d_fwd", L.ptr(x), _ld(x), L.ptr(out), c, n, h, w, c, pad, mode, L.stream())
    return out


def pad2d_bwd(dpad, pad, mode=0):
    n, hp, wp, c = dpad.shape
    h, w = hp - 2 * pad, wp - 2 * pad
    out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dpad.device)
    L.call("jg_pad2d_bwd", L.ptr(dpad), _ld(dpad), L.ptr(out), c, n, h, w, c, pad, mode, L.stream())
    return out


def dilate2x(x):
    """zero insertion: [N,H,W,C] -> [N,2H,2W,C] with x at even positions"""
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
    L.call("jg_dilate2x", L.ptr(x), _ld(x), L.ptr(out), c, n, 2 * h, 2 * w, c, 0, L.stream())
    return out


def undilate2x(x):
    """adjoint of dilate2x: [N,2H,2W,C] -> [N,H,W,C] taking even positions"""
    n, h2, w2, c = x.shape
    out = torch.empty((n, h2 // 2, w2 // 2, c), dtype=torch.bfloat16, device=x.device)
    L.call("jg_dilate2x", L.ptr(x), _ld(x), L.ptr(out), c, n, h2 // 2, w2 // 2, c, 1, L.stream())
    return out


def act_bwd(y, dy, act):
    n, h, w, c = y.shape
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=y.device)
    L.call("jg_act_bwd", L.ptr(y), _ld(y), L.ptr(dy), _ld(dy), L.ptr(dx), c, n * h * w, c, act, L.stream())
    return dx


GAN_LSGAN, GAN_HINGE, GAN_LINEAR = 0, 1, 2


def gan_loss_fwd(pred, c_real, mode, target, sign):
    n, h, w, _ = pred.shape
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    L.call("jg_gan_loss_fwd", L.ptr(pred), _ld(pred), n * h * w, c_real, mode, float(target), float(sign), L.ptr(loss),
           L.stream())
    return loss


def gan_loss_bwd(pred, c_real, mode, target, sign, grad_out):
    n, h, w, c = pred.shape
    d = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=pred.device)
    L.call("jg_gan_loss_bwd", L.ptr(pred), _ld(pred), n * h * w, c_real, mode, float(target), float(sign),
           L.ptr(grad_out), L.ptr(d), c, L.stream())
    return d


# ---- CUT contrastive path (csrc/nce.cu; tests/test_gpu_widen_cut.py) -------------------------------------------------
def gather_rows(feat, ids):
    """feat NHWC bf16 [B,H,W,C]; ids int64 [P] on the device (distinct positions in [0, H*W)).
    -> [B*P, C] bf16: the same positions of every image (PatchSampleF.forward, cut_networks.py:44-56)."""
    b, h, w, c = feat.shape
    p = ids.numel()
    out = torch.empty((b * p, c), dtype=torch.bfloat16, device=feat.device)
    L.call("jg_gather_rows", L.ptr(feat), _ld(feat), L.ptr(ids), L.ptr(out), c, b, h * w, p, c, L.stream())
    return out


def gather_rows_bwd(d_out, ids, shape):
    b, h, w, c = shape
    p = ids.numel()
    d_feat = torch.empty((b, h, w, c), dtype=torch.bfloat16, device=d_out.device)
    L.call("jg_gather_rows_bwd", L.ptr(d_out), d_out.stride(0), L.ptr(ids), L.ptr(d_feat), c, b, h * w, p, c,
           L.stream())
    return d_feat


def l2norm_fwd(x, eps=1e-7):
    """x bf16 [rows, D] -> (y fp32 [rows, D], norms fp32 [rows]) = F.normalize(x, dim=1, eps)."""
    rows, d = x.shape
    y = torch.empty((rows, d), dtype=torch.float32, device=x.device)
    norms = torch.empty((rows,), dtype=torch.float32, device=x.device)
    L.call("jg_l2norm_fwd", L.ptr(x), x.stride(0), L.ptr(y), L.ptr(norms), rows, d, float(eps), L.stream())
    return y, norms


def l2norm_bwd(y, dy, norms, eps=1e-7):
    rows, d = y.shape
    dx = torch.empty((rows, d), dtype=torch.bfloat16, device=y.device)
    L.call("jg_l2norm_bwd", L.ptr(y), L.ptr(dy), L.ptr(norms), L.ptr(dx), d, rows, d, float(eps), L.stream())
    return dx


def patch_nce_fwd(q, k, groups, temperature):
    """q, k fp32 [groups*P, D] -> (loss [groups*P], lse [groups*P])."""
    rows, d = q.shape
    loss = torch.empty((rows,), dtype=torch.float32, device=q.device)
    lse = torch.empty((rows,), dtype=torch.float32, device=q.device)
    L.call("jg_patch_nce_fwd", L.ptr(q), L.ptr(k), groups, rows // groups, d, float(temperature), L.ptr(loss),
           L.ptr(lse), L.stream())
    return loss, lse


def patch_nce_bwd(q, k, lse, grad_loss, groups, temperature, need_dq=True, need_dk=True):
    rows, d = q.shape
    dq = torch.empty_like(q) if need_dq else None
    dk = torch.empty_like(k) if need_dk else None
    L.call("jg_patch_nce_bwd", L.ptr(q), L.ptr(k), L.ptr(lse), L.ptr(grad_loss), groups, rows // groups, d,
           float(temperature), L.ptr(dq), L.ptr(dk), L.stream())
    return dq, dk


def monce_fwd(q, k, groups, temperature, num_patches_opt, iters=50):
    """MoNCE loss per patch; returns (loss, lse, ws) — ws keeps C, K and the Sinkhorn scaling history for the backward."""
    rows, d = q.shape
    p = rows // groups
    lib = L.load()
    ws = torch.empty((lib.jg_monce_ws_floats(groups, p, iters, 1),), dtype=torch.float32, device=q.device)
    loss = torch.empty((rows,), dtype=torch.float32, device=q.device)
    lse = torch.empty((rows,), dtype=torch.float32, device=q.device)
    L.call("jg_monce_fwd", L.ptr(q), L.ptr(k), groups, p, d, float(temperature), int(num_patches_opt), int(iters),
           L.ptr(ws), L.ptr(loss), L.ptr(lse), L.stream())
    return loss, lse, ws


def monce_bwd(q, k, lse, grad_loss, ws, groups, temperature, num_patches_opt, iters=50, need_dk=True):
    rows, d = q.shape
    dq = torch.empty_like(q)
    dk = torch.empty_like(k) if need_dk else None
    L.call("jg_monce_bwd", L.ptr(q), L.ptr(k), L.ptr(lse), L.ptr(grad_loss), groups, rows // groups, d,
           float(temperature), int(num_patches_opt), int(iters), L.ptr(ws), L.ptr(dq), L.ptr(dk), L.stream())
    return dq, dk


# ---- on-GPU input preparation and Haar wavelets (csrc/prep.cu) -------------------------
def fill_mask_random(img, mask, noise, cls=-1):
    """data/online_creation.fill_mask_with_random on the device: img, noise fp32 [N,C,H,W]; mask [N,1,H,W] int64/fp32."""
    n, c, h, w = img.shape
    img, noise = img.contiguous().float(), noise.contiguous().float()
    out = torch.empty_like(img)
    mf, mi = _mask_ptrs(mask.contiguous())
    L.call("jg_fill_mask_random", L.ptr(img), mf, mi, L.ptr(noise), L.ptr(out), n, c, h * w, int(cls), L.stream())
    return out


def u8_to_f32_normalized(x_nhwc_u8, mean=0.5, std=0.5):
    """ToTensor + Normalize: uint8 [N,H,W,C] -> fp32 [N,C,H,W]."""
    n, h, w, c = x_nhwc_u8.shape
    assert x_nhwc_u8.dtype == torch.uint8 and x_nhwc_u8.is_contiguous()
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x_nhwc_u8.device)
    L.call("jg_u8_to_f32_normalized", L.ptr(x_nhwc_u8), L.ptr(out), n, c, h, w, float(mean), float(std), L.stream())
    return out


def mask_class_dropout(mask, drop_u, prob, fill):
    """palette_model.py:565-584: samples with drop_u[n] < prob get the unconditioned class `fill` everywhere."""
    mask = mask.contiguous()
    n = mask.shape[0]
    out = torch.empty_like(mask)
    mf, mi = _mask_ptrs(mask)
    of, oi = _mask_ptrs(out)
    L.call("jg_mask_class_dropout", mf, mi, L.ptr(drop_u.contiguous().float()), float(prob), int(fill), of, oi, n,
           mask.numel() // n, L.stream())
    return out


def haar(x, mode):
    """mode 0: DWT fwd, 1: DWT bwd, 2: IWT fwd, 3: IWT bwd (freq_utils.HaarTransform / InverseHaarTransform)."""
    x = x.contiguous().float()
    n, c, hh, ww = x.shape
    if mode in (0, 3):   # full resolution in, bands out
        c_img, h, w = c, hh // 2, ww // 2
        out = torch.empty((n, 4 * c, h, w), dtype=torch.float32, device=x.device)
    else:                # bands in, full resolution out
        c_img, h, w = c // 4, hh, ww
        out = torch.empty((n, c_img, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
    L.call("jg_haar", L.ptr(x), L.ptr(out), n, c_img, h, w, mode, L.stream())
    return out


def embed_rows(table, idx, out, col0):
    """out[..., col0:col0+E] = bf16(table[idx]) for an NHWC bf16 tensor `out` ([N,H,W,ld]); idx [N,1,H,W] int64/fp32."""
    k, e = table.shape
    mf, mi = _mask_ptrs(idx.contiguous())
    L.call("jg_embed_rows", L.ptr(table), mf, mi, L.ptr(out), _ld(out), col0, idx.numel(), e, k, L.stream())
    return out


def embed_rows_bwd(d, idx, col0, k, e):
    """-> (dtable [K,E] = scatter-sum of d[..., col0:col0+E] by label, counts [K])"""
    dtable = torch.empty((k, e), dtype=torch.float32, device=d.device)
    counts = torch.empty((k,), dtype=torch.float32, device=d.device)
    mf, mi = _mask_ptrs(idx.contiguous())
    L.call("jg_embed_rows_bwd", L.ptr(d), _ld(d), col0, mf, mi, idx.numel(), e, k, L.ptr(dtable), L.ptr(counts),
           L.stream())
    return dtable, counts
