"""torch.autograd.Function wrappers: each op's forward AND backward run in libjg_b200.so.

All feature-map tensors are NHWC bf16.  Parameter gradients are returned in the reference layout
(fp32 OIHW / [C]) so that `.grad`, DDP, `state_dict` and optimisers see what joliGEN expects.
"""
import torch

from . import kernels as K
from . import lib as L


def _colsum_buffer(x):
    """[C] fp32 buffer for the per-channel sums of dx that GroupNorm-backward produces on the side."""
    return torch.empty((x.shape[-1],), dtype=torch.float32, device=x.device)


def _needs(ctx, i):
    return ctx.needs_input_grad[i]


# ---------------------------------------------------------------------------------------------------------------------
# GroupNorm work that travels with tensors (python attributes on the tensor OBJECT, validated by its version counter):
#   t._jg_stats  = (fp32 [N,C,2] per-(image, channel) sum / sum of squares of t, version)   set by the conv that wrote t
#                  -> the GroupNorm reading t skips its statistics pass (jg_conv_epilogue.stats)
#   y._jg_gn     = (x, ab, act, version)   set by the GroupNorm that produced y = act(a*x + b)
#                  -> the conv reading y hands (x, ab, act) to its dgrad, which accumulates the GroupNorm-backward sums
#   dy._jg_gnsums = (fp32 [N,C,2] (sum du, sum du*x), version)   set by that dgrad on its output
#                  -> the GroupNorm backward receiving dy skips its reduction pass (jg_conv_epilogue.gn_sums)
# A missing or stale stamp only means the stand-alone pass runs.  JG_FUSE_GN=0 turns the fusion off.
# ---------------------------------------------------------------------------------------------------------------------
import os as _os

# JG_FUSE_GN: "stats" (default) / "sums" one of the two reductions, "1" both, "0" none
# Default "stats": measured on B200 (profiles/r02_gn_fusion_ab.md) the statistics fused into the forward convs win
# ~1 ms per step, while the backward sums in the dgrad epilogue cost the convs more than the pass they remove.
_FUSE_MODE = _os.environ.get("JG_FUSE_GN", "stats")
FUSE_GN = [_FUSE_MODE != "0"]
FUSE_STATS = [_FUSE_MODE in ("1", "stats")]
FUSE_SUMS = [_FUSE_MODE in ("1", "sums")]
# the dgrad epilogue takes the GroupNorm-backward sums only for GroupNorms with at least this many channels
FUSE_SUMS_MIN_C = [int(_os.environ.get("JG_FUSE_SUMS_MIN_C", "0"))]
# ... and at most this many: the 64-channel kernels get the GroupNorm input tile by TMA next to the output staging; the
# wider ones read x with per-lane loads.  (Only with JG_FUSE_GN=1/sums: measured on B200 the fused sums lose in both
# forms — the derivative recompute makes the epilogue the bottleneck)
FUSE_SUMS_MAX_C = [int(_os.environ.get("JG_FUSE_SUMS_MAX_C", "64"))]


def _stamp(t, name):
    st = getattr(t, name, None)
    if st is not None and st[-1] == t._version:
        return st
    return None


def _carry_stats(src, *dst):
    st = _stamp(src, "_jg_stats")
    if st is not None:
        for d in dst:
            if d is not None:
                d._jg_stats = (st[0], d._version)


def _rows(t):
    """t as the kernels can address it: an NHWC tensor or a channel slice of one (unit channel stride, uniform row
    stride that is a multiple of 8 elements, 16-byte aligned start).  Anything else is copied."""
    if t.is_contiguous():
        return t
    if t.dim() == 4 and t.stride(3) == 1:
        n, h, w, _ = t.shape
        ld = t.stride(2)
        if (ld % 8 == 0 and t.stride(1) == w * ld and (n == 1 or t.stride(0) == h * w * ld)
                and t.storage_offset() % 8 == 0):
            return t
    return t.contiguous()


_MAX_FWD_COUT = 4096  # kMaxCout of csrc/conv_common.cuh (bias staging buffer of the forward kernels)


class Conv2dFn(torch.autograd.Function):
    """y = conv(x, W) + b (+ res_scale*residual).  x NHWC bf16 with channels padded to a multiple of 8;
    W fp32 OIHW (its bf16 packed copies wf / wd and the 8-padded bias are passed in).  The output has
    round_up(Cout, 8) channels (padding channels are exactly zero)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, wf, wd, bias_p, stride, pad, res_scale, sink, out, aux):
        cout, cin, r, s = weight.shape
        cout8 = (cout + 7) // 8 * 8
        # aux (dict or None): {"want_stats": bool, "gn": (x_gn, ab, act) of the GroupNorm that produced x}; receives
        # "stats" = the per-(image, channel) sums of y accumulated by the conv epilogue
        stats = None
        if aux is not None and aux.get("want_stats"):
            stats = torch.zeros((x.shape[0], cout8, 2), dtype=torch.float32, device=x.device)
            aux["stats"] = stats
        # out: optional destination (a channel slice of a wider NHWC buffer, e.g. the next block's concat input)
        if cout8 > _MAX_FWD_COUT and residual is None and stats is None:
            # wider than the kernels' bias staging (the b2b backbone's GEGLU projection: 768 -> 6144): output-channel
            # chunks of the same packed weights (rows of wf), each written into its slice of y
            y = out if out is not None else torch.empty(tuple(x.shape[:-1]) + (cout8,), dtype=torch.bfloat16,
                                                        device=x.device)
            for c0 in range(0, cout8, _MAX_FWD_COUT):
                c1 = min(cout8, c0 + _MAX_FWD_COUT)
                K.conv2d_fwd(x, wf[c0:c1], None if bias_p is None else bias_p[c0:c1], c1 - c0, r, s, stride=stride,
                             pad=pad, out=y[..., c0:c1])
        else:
            y = K.conv2d_fwd(x, wf, bias_p, cout8, r, s, stride=stride, pad=pad, residual=residual,
                             res_scale=res_scale, out=out, stats=stats)
        if out is not None:
            y = y.view(y.shape)  # a fresh tensor object: autograd must not see an input returned as an output
        gn = aux.get("gn") if aux is not None else None
        if gn is not None and stride == 1:
            ctx.save_for_backward(x, wd, gn[0], gn[1])
            ctx.gn_act = gn[2]
        else:
            ctx.save_for_backward(x, wd)
            ctx.gn_act = None
        ctx.geom = (cout, cin, r, s, stride, pad, res_scale, bias is not None, residual is not None)
        ctx.sink = sink  # (weight Parameter,) or None: accumulate dW straight into its .grad when possible
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wd = ctx.saved_tensors[:2]
        cout, cin, r, s, stride, pad, res_scale, has_bias, has_res = ctx.geom
        cout8 = (cout + 7) // 8 * 8
        # a GroupNorm backward that produced this very tensor has already summed it over (n, pixel)
        # (autograd may accumulate a second gradient INTO that tensor in place: the stamp carries its version)
        stamp = getattr(dy, "_jg_colsum", None)
        colsum = stamp[0] if stamp is not None and stamp[1] == dy._version else None
        dy = _rows(dy)
        dx = dw = db = dres = None
        if _needs(ctx, 0):
            if stride != 1:
                raise RuntimeError("Conv2dFn: dgrad for stride %d is not implemented" % stride)
            if ctx.gn_act is not None:
                # x = act(a * x_gn + b): dx is that GroupNorm's incoming gradient; its reduction pass (sum du,
                # sum du * x_gn per (image, channel)) runs in this dgrad's epilogue
                x_gn, ab = ctx.saved_tensors[2:4]
                sums = torch.zeros((x.shape[0], x.shape[-1], 2), dtype=torch.float32, device=x.device)
                dx = K.conv2d_fwd(dy, wd, None, x.shape[-1], r, s, stride=1, pad=r - 1 - pad,
                                  gn=(x_gn, ab, ctx.gn_act, sums))
                dx._jg_gnsums = (sums, dx._version)
            else:
                dx = K.conv2d_fwd(dy, wd, None, x.shape[-1], r, s, stride=1, pad=r - 1 - pad)
        if _needs(ctx, 1):
            # Writing dW behind autograd's back (no AccumulateGrad: DDP reducer hooks, register_hook and
            # post-accumulate-grad hooks never fire for that parameter) is strictly opt-in: only a trainer that owns
            # the parameter's flat .grad and its own all-reduce installs a `_jg_wstage` slot (trainer.WgradStage) and
            # enables it for the duration of ITS backward pass.  Plain / accelerate() / DDP use always gets dw returned.
            g = None
            stage = getattr(ctx.sink[0], "_jg_wstage", None) if ctx.sink is not None else None
            if stage is not None and stage.enabled and cout8 == cout and x.shape[-1] == cin:
                g = ctx.sink[0].grad
                if g is not None and not (g.is_contiguous() and g.dtype == torch.float32 and g.numel() == cout * cin * r * s):
                    g = None
            if g is not None:
                # trainer-owned persistent split-K accumulator: raw accumulation now, ONE batched unpack into .grad
                # for all convolutions at the end of the backward pass (WgradStage.flush)
                stage.layout = K.conv2d_wgrad_acc(x, dy, cout8, r, s, stage.acc, stride=stride, pad=pad)
                stage.notify()  # gradient final: the trainer may ship its bucket (overlapped all-reduce)
            else:
                dw = K.conv2d_wgrad(x, dy, cout8, r, s, stride=stride, pad=pad)
                if dw.shape[0] != cout or dw.shape[1] != cin:  # zero-padded channels (e.g. 6 -> 8, 3 -> 8)
                    dw = dw[:cout, :cin].contiguous()
        if has_bias and _needs(ctx, 2):
            db = colsum if colsum is not None and colsum.numel() == cout8 else K.bias_grad(dy)
            if cout8 != cout:
                db = db[:cout].contiguous()
        if has_res and _needs(ctx, 3):
            dres = dy if res_scale == 1.0 else (dy.float() * res_scale).to(torch.bfloat16)
        return dx, dw, db, dres, None, None, None, None, None, None, None, None, None


class GroupNormFn(torch.autograd.Function):
    """y = act(GN(x; gamma, beta) * (1 + scale) + shift), film = [N, 2C] fp32 (scale | shift) or None."""

    @staticmethod
    def forward(ctx, x, gamma, beta, film, groups, act, eps=1e-5, chan_stats=None, aux=None):
        y, stats, ab = K.groupnorm_fwd(x, gamma, beta, groups, film=film, act=act, eps=eps, chan_stats=chan_stats)
        ctx.save_for_backward(x, gamma, beta, film, stats, ab)
        ctx.cfg = (groups, act)
        ctx.film_slot = getattr(film, "_jg_grad_slot", None) if film is not None else None
        if aux is not None:
            aux["ab"] = ab
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, film, stats, ab = ctx.saved_tensors
        groups, act = ctx.cfg
        pre = _stamp(dy, "_jg_gnsums")
        dy = _rows(dy)
        need_p = gamma is not None and (_needs(ctx, 1) or _needs(ctx, 2))
        need_f = film is not None and _needs(ctx, 3)
        colsum = _colsum_buffer(x)
        dx, dgamma, dbeta, dfilm = K.groupnorm_bwd(x, dy, gamma, beta, groups, film, act, stats, ab,
                                                   need_param_grads=need_p, need_film_grad=need_f, colsum=colsum,
                                                   sums_pre=None if pre is None else pre[0],
                                                   dfilm_out=ctx.film_slot)
        if colsum is not None:
            # picked up by Conv2dFn.backward when x is a conv output (its bias gradient)
            dx._jg_colsum = (colsum, dx._version)
        return dx, dgamma, dbeta, dfilm, None, None, None, None, None


class GroupNormTapFn(torch.autograd.Function):
    """GroupNorm(+FiLM)(+act) that also hands its input through: returns (y, x_tap, x_tap2).  x has several consumers
    in the UNet (the norm, the ResBlock / AttentionBlock skip path, the decoder's concat of the same tensor);
    routing the other consumers through the taps lets the backward sum all gradients inside the GN-backward apply
    pass (dx = GN'(dy) + d_tap + d_tap2) instead of separate add kernels."""

    @staticmethod
    def forward(ctx, x, gamma, beta, film, groups, act, eps=1e-5, chan_stats=None, aux=None):
        y, stats, ab = K.groupnorm_fwd(x, gamma, beta, groups, film=film, act=act, eps=eps, chan_stats=chan_stats)
        ctx.save_for_backward(x, gamma, beta, film, stats, ab)
        ctx.cfg = (groups, act)
        ctx.film_slot = getattr(film, "_jg_grad_slot", None) if film is not None else None
        ctx.set_materialize_grads(False)  # an unused tap must arrive as None, not as a zero-filled tensor
        if aux is not None:
            aux["ab"] = ab
        return y, x.detach(), x.detach()

    @staticmethod
    def backward(ctx, dy, dtap, dtap2):
        x, gamma, beta, film, stats, ab = ctx.saved_tensors
        groups, act = ctx.cfg
        if dtap is None:
            dtap, dtap2 = dtap2, None
        if dy is None:
            if dtap2 is not None:
                dtap = dtap + dtap2
            return dtap, None, None, None, None, None, None, None, None
        pre = _stamp(dy, "_jg_gnsums")
        dy = _rows(dy)
        need_p = gamma is not None and (_needs(ctx, 1) or _needs(ctx, 2))
        need_f = film is not None and _needs(ctx, 3)
        colsum = _colsum_buffer(x)
        dx, dgamma, dbeta, dfilm = K.groupnorm_bwd(x, dy, gamma, beta, groups, film, act, stats, ab,
                                                   need_param_grads=need_p, need_film_grad=need_f,
                                                   addend=None if dtap is None else _rows(dtap),
                                                   addend2=None if dtap2 is None else _rows(dtap2),
                                                   colsum=colsum, sums_pre=None if pre is None else pre[0],
                                                   dfilm_out=ctx.film_slot)
        if colsum is not None:
            dx._jg_colsum = (colsum, dx._version)
        return dx, dgamma, dbeta, dfilm, None, None, None, None, None


class AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, ch, layout, out):
        o, lse = K.attn_fwd(qkv, heads, ch, layout, out=out)
        if out is not None:
            o = o.view(o.shape)  # a fresh tensor object (see Conv2dFn.forward)
        ctx.save_for_backward(qkv, o, lse)
        ctx.cfg = (heads, ch, layout)
        return o

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        heads, ch, layout = ctx.cfg
        return K.attn_bwd(qkv, out, _rows(d_out), lse, heads, ch, layout), None, None, None, None


class MixQkvFn(torch.autograd.Function):
    """AttentionBlockRef._forward (unet_generator_attn.py:1111-1118): chunk(3) of the block's own qkv and of the
    reference UNet's qkv along channels, then cat([q, k_ref, v_ref]): the first third of the channels comes from
    `qkv`, the rest from `qkv_ref` (whatever the attention order does with those channels afterwards)."""

    @staticmethod
    def forward(ctx, qkv, qkv_ref):
        c = qkv.shape[-1] // 3
        out = torch.empty_like(qkv_ref, memory_format=torch.contiguous_format)
        K.copy_channels(qkv[..., :c], out[..., :c])
        K.copy_channels(qkv_ref[..., c:], out[..., c:])
        ctx.c = c
        return out

    @staticmethod
    def backward(ctx, d):
        c = ctx.c
        d = _rows(d)
        dq = torch.zeros(d.shape, dtype=d.dtype, device=d.device)
        dr = torch.zeros(d.shape, dtype=d.dtype, device=d.device)
        K.copy_channels(d[..., :c], dq[..., :c])
        K.copy_channels(d[..., c:], dr[..., c:])
        return dq, dr


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over channels of NHWC tokens (+ the frame positional encoding of VersatileAttention)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, pe, frames, eps):
        y, stats = K.layernorm_fwd(x, gamma, beta, eps=eps, pe=pe, frames=frames)
        ctx.save_for_backward(x, gamma, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, stats = ctx.saved_tensors
        dx, dgamma, dbeta = K.layernorm_bwd(x, _rows(dy), gamma, stats)
        return dx, dgamma, dbeta, None, None, None


class LayerNormTapFn(torch.autograd.Function):
    """LayerNorm that also hands its input through: returns (y, x_tap).  In the temporal transformer x feeds the norm
    AND the residual add after the attention / feed-forward branch; the residual reads x_tap, so both gradients meet
    inside the LayerNorm backward (dx = LN'(dy) + d_tap) and its column sums come out on the side — the bias gradient
    of the Linear that produced x."""

    @staticmethod
    def forward(ctx, x, gamma, beta, pe, frames, eps):
        y, stats = K.layernorm_fwd(x, gamma, beta, eps=eps, pe=pe, frames=frames)
        ctx.save_for_backward(x, gamma, stats)
        ctx.set_materialize_grads(False)
        return y, x.detach()

    @staticmethod
    def backward(ctx, dy, dtap):
        x, gamma, stats = ctx.saved_tensors
        if dy is None:
            return dtap, None, None, None, None, None
        colsum = _colsum_buffer(x)
        dx, dgamma, dbeta = K.layernorm_bwd(x, _rows(dy), gamma, stats, addend=None if dtap is None else _rows(dtap),
                                            colsum=colsum)
        dx._jg_colsum = (colsum, dx._version)
        return dx, dgamma, dbeta, None, None, None


class TemporalAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, frames, heads):
        ctx.save_for_backward(qkv)
        ctx.cfg = (frames, heads)
        return K.temporal_attn_fwd(qkv, frames, heads)

    @staticmethod
    def backward(ctx, d_out):
        (qkv,) = ctx.saved_tensors
        frames, heads = ctx.cfg
        return K.temporal_attn_bwd(qkv, _rows(d_out), frames, heads), None, None


class GegluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return K.geglu_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return K.geglu_bwd(x, _rows(dy))


class LinearFn(torch.autograd.Function):
    """fp32 y = act_out(act_in(x) @ W^T + b) on [B, I] embeddings (act_out only without grad)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act_in):
        x = x.contiguous().float()
        y = K.linear_fwd(x, weight, bias, act_in=act_in)
        ctx.save_for_backward(x, weight)
        ctx.act_in = act_in
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx, dw, db = K.linear_bwd(x, weight, dy.contiguous().float(), act_in=ctx.act_in, need_dx=_needs(ctx, 0))
        return dx, dw, db, None


class LinearBankFn(torch.autograd.Function):
    """All emb_layers Linears of a UNet on the same SiLU(emb) in ONE launch each way (kernels.LinearBank).  Returns one
    [B, O_i] tensor per Linear (views of one buffer).  Every output carries `_jg_grad_slot`: the matching view of the
    backward's dY buffer — a GroupNorm whose FiLM input is that output writes its d(film) straight into the slot, so
    the backward needs no gather."""

    @staticmethod
    def forward(ctx, x, bank, act_in, holder, *params):
        x = x.contiguous().float()
        y = K.linear_batched_fwd(x, bank, act_in)
        bsz = x.shape[0]
        ctx.save_for_backward(x)
        ctx.bank, ctx.act_in = bank, act_in
        ctx.dy = torch.empty_like(y)
        ctx.set_materialize_grads(False)
        holder["slots"] = [ctx.dy[bsz * off:bsz * (off + o)].view(bsz, o) for off, o in zip(bank.offsets, bank.widths)]
        return tuple(y[bsz * off:bsz * (off + o)].view(bsz, o) for off, o in zip(bank.offsets, bank.widths))

    @staticmethod
    def backward(ctx, *grads):
        (x,) = ctx.saved_tensors
        bank, bsz = ctx.bank, x.shape[0]
        for g, off, o in zip(grads, bank.offsets, bank.widths):
            slot = ctx.dy[bsz * off:bsz * (off + o)].view(bsz, o)
            if g is None:
                slot.zero_()
            elif g.data_ptr() != slot.data_ptr():
                slot.copy_(g)
        dx, dw, db = K.linear_batched_bwd(x, bank, ctx.dy, ctx.act_in, need_dx=_needs(ctx, 0))
        out = [dx, None, None, None]
        for off, o in zip(bank.offsets, bank.widths):
            out.append(dw[off:off + o])
            out.append(db[off:off + o])
        return tuple(out)


def linear_bank(x, bank, linears, act_in=L.ACT_NONE):
    """[lin(act_in(x)) for lin in linears] in one launch; see LinearBankFn."""
    holder = {}
    params = []
    for lin in linears:
        params += [lin.weight, lin.bias]
    outs = LinearBankFn.apply(x, bank, act_in, holder, *params)
    for y, slot in zip(outs, holder["slots"]):
        y._jg_grad_slot = slot
    return outs


class Resample2xFn(torch.autograd.Function):
    """mode 0: nearest 2x upsample; mode 1: 2x2 average pool."""

    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return K.resample2x(x, mode)

    @staticmethod
    def backward(ctx, dy):
        return K.resample2x(_rows(dy), 2 if ctx.mode == 0 else 3), None


class CatChannelsFn(torch.autograd.Function):
    """torch.cat([a, b], dim=channel) for NHWC tensors."""

    @staticmethod
    def forward(ctx, a, b):
        n, h, w, ca = a.shape
        cb = b.shape[-1]
        out = torch.empty((n, h, w, ca + cb), dtype=torch.bfloat16, device=a.device)
        K.copy_channels(a, out[..., :ca])
        K.copy_channels(b, out[..., ca:])
        ctx.split = (ca, cb)
        return out

    @staticmethod
    def backward(ctx, d):
        # channel-slice views: every backward kernel addresses its incoming gradient with a row stride
        ca, cb = ctx.split
        return _slice_with_stamp(d, 0, ca), _slice_with_stamp(d, ca, ca + cb)


def _cat_stats(out, a, b):
    """per-channel statistics of a channel concatenation = the concatenation of the parts' statistics"""
    sa, sb = _stamp(a, "_jg_stats"), _stamp(b, "_jg_stats")
    if sa is not None and sb is not None:
        out._jg_stats = (torch.cat([sa[0], sb[0]], dim=1), out._version)


def _slice_with_stamp(d, c0, c1):
    """d[..., c0:c1] as a view; the per-channel sums that a GroupNorm backward attached to d (see Conv2dFn.backward)
    are sliced along with it, so the conv that produced this part of a concat input gets its bias gradient for free."""
    v = d[..., c0:c1]
    stamp = getattr(d, "_jg_colsum", None)
    if stamp is not None and stamp[1] == d._version:
        v._jg_colsum = (stamp[0][c0:c1], v._version)
    return v


class CatIntoFn(torch.autograd.Function):
    """torch.cat([a, b], dim=channel) where `a` already IS buf[..., :Ca] (its producer wrote it there through
    conv2d(out=...)): only b is copied.  Returns buf."""

    @staticmethod
    def forward(ctx, a, b, buf):
        ca, cb = a.shape[-1], b.shape[-1]
        assert buf.shape[-1] == ca + cb and a.data_ptr() == buf.data_ptr() and a.stride(2) == buf.stride(2)
        K.copy_channels(b, buf[..., ca:])
        ctx.split = (ca, cb)
        return buf.view(buf.shape)

    @staticmethod
    def backward(ctx, d):
        ca, cb = ctx.split
        return _slice_with_stamp(d, 0, ca), _slice_with_stamp(d, ca, ca + cb), None


class ToNHWCFn(torch.autograd.Function):
    """fp32 NCHW -> bf16 NHWC (channels zero-padded to a multiple of 8)."""

    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return K.nchw_to_nhwc(x)

    @staticmethod
    def backward(ctx, d):
        return K.nhwc_to_nchw(d.contiguous(), ctx.c)


class ToNCHWFn(torch.autograd.Function):
    """bf16 NHWC -> fp32 NCHW, keeping the first c channels."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.ld = x.shape[-1]
        return K.nhwc_to_nchw(x, c)

    @staticmethod
    def backward(ctx, d):
        return K.nchw_to_nhwc(d, ctx.ld), None


class PaletteLossFn(torch.autograd.Function):
    """lambda * mean((w*m*(noise - noise_hat))^2) with noise_hat NHWC bf16 (only it receives a gradient)."""

    @staticmethod
    def forward(ctx, noise_hat, noise, mask, w_b, lambda_g, l1):
        loss = K.palette_loss_fwd(noise, noise_hat, mask, w_b, lambda_g, l1)
        ctx.save_for_backward(noise_hat, noise, mask, w_b)
        ctx.cfg = (lambda_g, l1)
        return loss

    @staticmethod
    def backward(ctx, g):
        noise_hat, noise, mask, w_b = ctx.saved_tensors
        lambda_g, l1 = ctx.cfg
        g = g.contiguous().float()
        return K.palette_loss_bwd(noise, noise_hat, mask, w_b, g, lambda_g, l1), None, None, None, None, None


def conv2d(x, weight, bias, packed, stride=1, pad=None, residual=None, res_scale=1.0, grad_sink=None, out=None,
           want_stats=False):
    """packed = (wf, wd, bias_padded) from nets.ConvPack.get(); grad_sink = the weight nn.Parameter whose
    (pre-existing, fp32, contiguous) .grad the weight gradient is accumulated into directly; out = optional
    destination view (a channel slice of a wider NHWC buffer); want_stats: the output feeds a GroupNorm — accumulate
    its per-(image, channel) statistics in the epilogue and attach them to the returned tensor."""
    r = weight.shape[2]
    if pad is None:
        pad = (r - 1) // 2
    wf, wd, bias_p = packed
    aux = None
    if FUSE_GN[0]:
        # the GroupNorm that produced x rides along for the dgrad; a 3x3 conv feeding a GroupNorm emits its statistics
        gn = _stamp(x, "_jg_gn") if (FUSE_SUMS[0] and stride == 1 and x.requires_grad
                                     and FUSE_SUMS_MIN_C[0] <= x.shape[-1] <= FUSE_SUMS_MAX_C[0]) else None
        want_stats = bool(want_stats) and FUSE_STATS[0]
        if gn is not None or want_stats:
            aux = {"want_stats": want_stats, "gn": None if gn is None else gn[:3]}
    y = Conv2dFn.apply(x, weight, bias, residual, wf, wd, bias_p, stride, pad, res_scale,
                       None if grad_sink is None else (grad_sink,), out, aux)
    if aux is not None and "stats" in aux:
        y._jg_stats = (aux["stats"], y._version)
    return y


def _gn_apply(fn, x, gamma, beta, film, groups, act, eps):
    if gamma is not None and gamma.numel() != x.shape[-1]:
        # a width that is not a multiple of 8 is stored zero-padded: the kernel would normalise over the padded
        # channels (wrong groups) and read gamma / beta past their end
        raise NotImplementedError("GroupNorm over %d channels stored as %d (channel counts must be multiples of 8)"
                                  % (gamma.numel(), x.shape[-1]))
    st = _stamp(x, "_jg_stats") if FUSE_GN[0] else None
    aux = {} if FUSE_GN[0] else None
    out = fn.apply(x, gamma, beta, film, groups, act, eps, None if st is None else st[0], aux)
    y = out[0] if isinstance(out, tuple) else out
    if aux is not None and "ab" in aux:
        y._jg_gn = (x.detach(), aux["ab"], act, y._version)
    return out


def group_norm(x, gamma, beta, groups, film=None, act=L.ACT_NONE, eps=1e-5):
    return _gn_apply(GroupNormFn, x, gamma, beta, film, groups, act, eps)


def group_norm_tap(x, gamma, beta, groups, film=None, act=L.ACT_NONE, eps=1e-5):
    """-> (y, x_tap): use x_tap for every other consumer of x (see GroupNormTapFn)."""
    y, tap, _ = _gn_apply(GroupNormTapFn, x, gamma, beta, film, groups, act, eps)
    _carry_stats(x, tap)
    return y, tap


def group_norm_tap2(x, gamma, beta, groups, film=None, act=L.ACT_NONE, eps=1e-5):
    """-> (y, x_tap, x_tap2): two independent hand-throughs of x (block skip path + decoder concat)."""
    y, tap, tap2 = _gn_apply(GroupNormTapFn, x, gamma, beta, film, groups, act, eps)
    _carry_stats(x, tap, tap2)
    return y, tap, tap2


def attention(qkv, heads, ch, layout=0, out=None):
    """layout 0: QKVAttentionLegacy channel order, 1: QKVAttention (q | k | v); out = optional destination slice."""
    return AttentionFn.apply(qkv, heads, ch, layout, out)


def mix_qkv(qkv, qkv_ref):
    return MixQkvFn.apply(qkv, qkv_ref)


def layer_norm(x, gamma, beta, pe=None, frames=1, eps=1e-5):
    return LayerNormFn.apply(x, gamma, beta, pe, frames, eps)


def layer_norm_tap(x, gamma, beta, pe=None, frames=1, eps=1e-5):
    """-> (y, x_tap): the residual add that follows the normed branch must read x_tap (see LayerNormTapFn)."""
    return LayerNormTapFn.apply(x, gamma, beta, pe, frames, eps)


def temporal_attention(qkv, frames, heads):
    return TemporalAttentionFn.apply(qkv, frames, heads)


def geglu(x):
    return GegluFn.apply(x)


def linear(x, weight, bias, act_in=L.ACT_NONE):
    return LinearFn.apply(x, weight, bias, act_in)


def upsample2x(x):
    return Resample2xFn.apply(x, 0)


def avgpool2x(x):
    return Resample2xFn.apply(x, 1)


def cat_channels(a, b):
    out = CatChannelsFn.apply(a, b)
    _cat_stats(out, a, b)
    return out


class JoinSlicesFn(torch.autograd.Function):
    """parts[i] already ARE consecutive channel slices of buf (written there through conv2d(out=...)): returns buf as
    a function of the parts; the backward hands out the matching channel-slice views."""

    @staticmethod
    def forward(ctx, buf, *parts):
        ctx.widths = [p.shape[-1] for p in parts]
        assert sum(ctx.widths) == buf.shape[-1] and parts[0].data_ptr() == buf.data_ptr()
        return buf.view(buf.shape)

    @staticmethod
    def backward(ctx, d):
        outs, c0 = [None], 0
        for w in ctx.widths:
            outs.append(d[..., c0:c0 + w])
            c0 += w
        return tuple(outs)


def join_slices(buf, *parts):
    return JoinSlicesFn.apply(buf, *parts)


def cat_into(buf, a, b):
    """a = buf[..., :Ca] already in place; copy b behind it and return the full buffer."""
    out = CatIntoFn.apply(a, b, buf)
    _cat_stats(out, a, b)
    return out


def to_nhwc(x):
    return ToNHWCFn.apply(x)


def to_nchw(x, c):
    return ToNCHWFn.apply(x, c)


def palette_loss(noise_hat, noise, mask, w_b=None, lambda_g=1.0, l1=False):
    return PaletteLossFn.apply(noise_hat, noise, mask, w_b, lambda_g, l1)


# ---------------------------------------------------------------------------------------------
# GAN generator / discriminator ops
# ---------------------------------------------------------------------------------------------
class ConvActFn(torch.autograd.Function):
    """Generalised conv for the GAN nets: y = act(conv(x, W, stride) + b), act in {none, lrelu, tanh}.

    stride 2: forward = the strided implicit GEMM; dgrad = zero-insertion of dy then a stride-1 implicit GEMM
    with the flipped weights (cropped to the input size); wgrad = strided-X implicit GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, wf, wd, bias_p, stride, pad, act):
        cout, cin, r, s = weight.shape
        cout8 = (cout + 7) // 8 * 8
        y = K.conv2d_fwd(x, wf, bias_p, cout8, r, s, stride=stride, pad=pad, act=act)
        ctx.save_for_backward(x, wd, y if act != L.ACT_NONE else None)
        ctx.geom = (cout, cin, r, s, stride, pad, act, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wd, y = ctx.saved_tensors
        cout, cin, r, s, stride, pad, act, has_bias = ctx.geom
        cout8 = (cout + 7) // 8 * 8
        dy = dy.contiguous()
        if act != L.ACT_NONE:
            dy = K.act_bwd(y, dy, act)
        dx = dw = db = None
        n, h, w, cx = x.shape
        if _needs(ctx, 0):
            if stride == 1:
                dx = K.conv2d_fwd(dy, wd, None, cx, r, s, stride=1, pad=r - 1 - pad)
            else:
                dyd = K.dilate2x(dy)
                dx = K.conv2d_cropped(dyd, wd, None, cx, r, s, pad=r - 1 - pad, out_hw=(h, w))
        if _needs(ctx, 1):
            dw = K.conv2d_wgrad(x, dy, cout8, r, s, stride=stride, pad=pad)
            if dw.shape[0] != cout or dw.shape[1] != cin:
                dw = dw[:cout, :cin].contiguous()
        if has_bias and _needs(ctx, 2):
            db = K.bias_grad(dy)
            if cout8 != cout:
                db = db[:cout].contiguous()
        return dx, dw, db, None, None, None, None, None, None


class ConvTranspose2dFn(torch.autograd.Function):
    """nn.ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1) (resnet_generator.py:306-318):
    y[2H,2W] = the dgrad of a stride-2 3x3 convolution whose OIHW weight is the transposed-conv weight
    [Cin, Cout, 3, 3] read as O = Cin, I = Cout."""

    @staticmethod
    def forward(ctx, x, weight, bias, wf, wd, bias_p, pad):
        cin, cout, r, s = weight.shape
        n, h, w, _ = x.shape
        cout8 = (cout + 7) // 8 * 8
        xd = K.dilate2x(x)
        y = K.conv2d_cropped(xd, wd, bias_p, cout8, r, s, pad=r - 1 - pad, out_hw=(2 * h, 2 * w))
        ctx.save_for_backward(x, wf)
        ctx.geom = (cin, cout, r, s, pad, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wf = ctx.saved_tensors
        cin, cout, r, s, pad, has_bias = ctx.geom
        dy = dy.contiguous()
        dx = dw = db = None
        if _needs(ctx, 0):
            # adjoint of the transposed conv = the stride-2 conv with the same weight (O = Cin_ct, I = Cout_ct)
            dx = K.conv2d_fwd(dy, wf, None, x.shape[-1], r, s, stride=2, pad=pad)
        if _needs(ctx, 1):
            # wgrad of that conv: its input is dy (2H x 2W, Cout_ct channels), its output gradient is x
            dw = K.conv2d_wgrad(dy, x, x.shape[-1], r, s, stride=2, pad=pad)  # [Cin_ct8, Cout_ct8, r, s]
            if dw.shape[0] != cin or dw.shape[1] != cout:
                dw = dw[:cin, :cout].contiguous()
        if has_bias and _needs(ctx, 2):
            db = K.bias_grad(dy)
            if db.shape[0] != cout:
                db = db[:cout].contiguous()
        return dx, dw, db, None, None, None, None


class Pad2dFn(torch.autograd.Function):
    """nn.ReflectionPad2d (mode 0) / nn.ReplicationPad2d (mode 1)"""

    @staticmethod
    def forward(ctx, x, pad, mode=0):
        ctx.pad, ctx.mode = pad, mode
        return K.pad2d(x, pad, mode)

    @staticmethod
    def backward(ctx, d):
        return K.pad2d_bwd(d.contiguous(), ctx.pad, ctx.mode), None, None


class AddFn(torch.autograd.Function):
    """a + b for NHWC bf16 tensors (ResnetBlock skip, resnet_generator.py:92-95)."""

    @staticmethod
    def forward(ctx, a, b):
        out = a.clone()
        K.copy_channels(b, out, accumulate=True)
        return out

    @staticmethod
    def backward(ctx, d):
        return d, d


class GanLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, c_real, mode, target, sign):
        ctx.save_for_backward(pred)
        ctx.cfg = (c_real, mode, target, sign)
        return K.gan_loss_fwd(pred, c_real, mode, target, sign)

    @staticmethod
    def backward(ctx, g):
        (pred,) = ctx.saved_tensors
        c_real, mode, target, sign = ctx.cfg
        return K.gan_loss_bwd(pred, c_real, mode, target, sign, g.contiguous().float()), None, None, None, None


def conv_act(x, weight, bias, packed, stride=1, pad=0, act=L.ACT_NONE):
    wf, wd, bias_p = packed
    return ConvActFn.apply(x, weight, bias, wf, wd, bias_p, stride, pad, act)


def conv_transpose2d(x, weight, bias, packed, pad=1):
    wf, wd, bias_p = packed
    return ConvTranspose2dFn.apply(x, weight, bias, wf, wd, bias_p, pad)


def reflection_pad(x, pad):
    return Pad2dFn.apply(x, pad, 0)


def replication_pad(x, pad):
    return Pad2dFn.apply(x, pad, 1)


def add(a, b):
    return AddFn.apply(a, b)


def gan_loss(pred, c_real, mode, target=1.0, sign=1.0):
    return GanLossFn.apply(pred, c_real, mode, target, sign)


# ---- CUT contrastive path (tests/test_gpu_widen_cut.py) --------------------------------------------------------------
class GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, ids):
        ctx.save_for_backward(ids)
        ctx.shape = tuple(feat.shape)
        return K.gather_rows(feat.contiguous(), ids)

    @staticmethod
    def backward(ctx, d_out):
        (ids,) = ctx.saved_tensors
        return K.gather_rows_bwd(d_out.contiguous(), ids, ctx.shape), None


class L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        y, norms = K.l2norm_fwd(x.contiguous(), eps)
        ctx.save_for_backward(y, norms)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        y, norms = ctx.saved_tensors
        return K.l2norm_bwd(y, dy.contiguous().float(), norms, ctx.eps), None


class PatchNceFn(torch.autograd.Function):
    """loss per patch of PatchNCELoss; q and k both receive gradient (k through the negatives only)."""

    @staticmethod
    def forward(ctx, q, k, groups, temperature):
        q, k = q.contiguous().float(), k.contiguous().float()
        loss, lse = K.patch_nce_fwd(q, k, groups, temperature)
        ctx.save_for_backward(q, k, lse)
        ctx.cfg = (groups, temperature)
        return loss

    @staticmethod
    def backward(ctx, g):
        q, k, lse = ctx.saved_tensors
        groups, temperature = ctx.cfg
        dq, dk = K.patch_nce_bwd(q, k, lse, g.contiguous().float(), groups, temperature, _needs(ctx, 0), _needs(ctx, 1))
        return dq, dk, None, None


class MonceFn(torch.autograd.Function):
    """loss per patch of MoNCELoss (Sinkhorn-weighted negatives); q gets the gradient through the OT iterations too."""

    @staticmethod
    def forward(ctx, q, k, groups, temperature, num_patches_opt):
        q, k = q.contiguous().float(), k.contiguous().float()
        loss, lse, ws = K.monce_fwd(q, k, groups, temperature, num_patches_opt)
        ctx.save_for_backward(q, k, lse, ws)
        ctx.cfg = (groups, temperature, num_patches_opt)
        return loss

    @staticmethod
    def backward(ctx, g):
        q, k, lse, ws = ctx.saved_tensors
        groups, temperature, num_patches_opt = ctx.cfg
        dq, dk = K.monce_bwd(q, k, lse, g.contiguous().float(), ws, groups, temperature, num_patches_opt,
                             need_dk=_needs(ctx, 1))
        return dq, dk, None, None, None


def gather_rows(feat, ids):
    return GatherRowsFn.apply(feat, ids)


def l2_normalize(x, eps=1e-7):
    return L2NormFn.apply(x, eps)


def patch_nce(q, k, groups, temperature):
    return PatchNceFn.apply(q, k, groups, temperature)


def monce(q, k, groups, temperature, num_patches_opt):
    return MonceFn.apply(q, k, groups, temperature, num_patches_opt)


# ---- Haar wavelets (freq_utils.HaarTransform / InverseHaarTransform) -------------------------------------------------
class HaarFn(torch.autograd.Function):
    """inverse=False: x [N,C,H,W] -> [N,4C,H/2,W/2] = cat(ll, lh, hl, hh); inverse=True: the synthesis transform."""

    @staticmethod
    def forward(ctx, x, inverse):
        ctx.inverse = inverse
        return K.haar(x, 2 if inverse else 0)

    @staticmethod
    def backward(ctx, d):
        return K.haar(d, 3 if ctx.inverse else 1), None


def haar_dwt(x):
    return HaarFn.apply(x, False)


def haar_iwt(x):
    return HaarFn.apply(x, True)


# ---- label embeddings of PaletteDenoiseFn (palette_denoise_fn.py:14-31, 118-136) ---------------------------------------
@torch.no_grad()
def _embedding_renorm_(table, idx, max_norm=1.0):
    """nn.Embedding(max_norm=...): the rows that are looked up are rescaled in place to norm <= max_norm
    (torch.embedding_renorm_: scale = max_norm / (norm + 1e-7)).  Static shapes only (CUDA-graph capturable)."""
    k = table.shape[0]
    present = torch.zeros(k, dtype=torch.bool, device=table.device)
    present.index_fill_(0, idx.reshape(-1).long().clamp_(0, k - 1), True)
    norms = table.norm(dim=1)
    scale = torch.where(present & (norms > max_norm), max_norm / (norms + 1e-7), torch.ones_like(norms))
    table.mul_(scale[:, None])


class LabelEmbedFn(torch.autograd.Function):
    """Class-label lookup [B] -> [B, E] fp32 with nn.Embedding(max_norm=1, scale_grad_by_freq=True) semantics.  O(B):
    plain torch gathers with static shapes (like the t / gamma gathers of the noising prologue)."""

    @staticmethod
    def forward(ctx, table, labels):
        _embedding_renorm_(table, labels)
        ctx.save_for_backward(labels)
        ctx.k = table.shape[0]
        return table.detach().index_select(0, labels)

    @staticmethod
    def backward(ctx, d):
        (labels,) = ctx.saved_tensors
        g = torch.zeros((ctx.k, d.shape[1]), dtype=d.dtype, device=d.device).index_add_(0, labels, d)
        counts = torch.zeros(ctx.k, dtype=d.dtype, device=d.device).index_add_(0, labels, torch.ones_like(d[:, 0]))
        return g / counts.clamp(min=1.0)[:, None], None


class EmbedRowsFn(torch.autograd.Function):
    """Per-pixel label embedding written INTO channels [col0, col0+E) of the NHWC bf16 UNet input (the reference
    concatenates mask_embed to the input, palette_denoise_fn.py:104-108).  Same nn.Embedding semantics as above."""

    @staticmethod
    def forward(ctx, table, mask, x, col0):
        _embedding_renorm_(table, mask)
        K.embed_rows(table.detach(), mask, x, col0)
        ctx.save_for_backward(mask)
        ctx.cfg = (col0, table.shape[0], table.shape[1])
        ctx.mark_dirty(x)
        return x

    @staticmethod
    def backward(ctx, d):
        (mask,) = ctx.saved_tensors
        col0, k, e = ctx.cfg
        dtable, counts = K.embed_rows_bwd(_rows(d), mask, col0, k, e)
        return dtable / counts.clamp(min=1.0)[:, None], None, None, None


def label_embed(table, labels):
    return LabelEmbedFn.apply(table, labels)


def embed_rows_into(table, mask, x, col0):
    return EmbedRowsFn.apply(table, mask, x, col0)
