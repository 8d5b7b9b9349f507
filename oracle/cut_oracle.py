"""CPU oracle for CUT's contrastive path (TEST INFRASTRUCTURE — see oracle/__init__.py): SURVEY.md section 8(f)
rank 3, the next row after the generator / discriminator operators of oracle/gan_oracle.py.

Functional fp32 restatement of

  PatchSampleF.forward      /root/reference/models/modules/cut_networks.py:38-73
      (per NCE layer: NHWC flatten, gather of `num_patches` spatial positions — the same positions for every image of
       the batch —, Linear -> ReLU -> Linear, L2 normalisation with eps 1e-7)
  BaseNCELoss / PatchNCELoss  /root/reference/models/modules/NCE/base_NCE.py:17-77, patchnce.py
  MoNCELoss + Sinkhorn OT     /root/reference/models/modules/NCE/monce.py:12-33, sinkhorn.py  (--alg_cut_nce_loss monce,
                              the default of example_gan_horse2zebra.json)
      (positive logit = <q_i, k_i>, negatives = q_i . k_j over the patches of the same image (or of the whole
       minibatch), diagonal filled with -10, cross entropy of [pos | neg] / T against class 0, no reduction)
  CUTModel.calculate_NCE_loss /root/reference/models/cut_model.py:889-909  (lambda_NCE * mean per layer, / n_layers)

The random patch positions (torch.randperm in the reference) are an input here: the golden generator records the
reference's own draws.  Pinned against the real reference by oracle/gen_golden_cut.py + tests/test_cut_oracle.py.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .palette_oracle import _r, _rw


def mlp_param_shapes(feat_channels: Sequence[int], nc: int = 256) -> Dict[str, Tuple[int, ...]]:
    """state_dict of PatchSampleF after create_mlp (cut_networks.py:23-36): mlp_<i>.{0,2}.{weight,bias}."""
    s: Dict[str, Tuple[int, ...]] = {}
    for i, c in enumerate(feat_channels):
        s["mlp_%d.0.weight" % i] = (nc, c)
        s["mlp_%d.0.bias" % i] = (nc,)
        s["mlp_%d.2.weight" % i] = (nc, nc)
        s["mlp_%d.2.bias" % i] = (nc,)
    return s


def patch_sample(sd: Optional[Dict[str, torch.Tensor]], feats: List[torch.Tensor], num_patches: int,
                 patch_ids: List[torch.Tensor], use_mlp: bool = True):
    """PatchSampleF.forward with explicit patch ids (one 1-D LongTensor per layer).  num_patches > 0 only (the
    num_patches == 0 branch reshapes whole maps and is not used by cut_model)."""
    out = []
    for i, feat in enumerate(feats):
        flat = _r(feat).permute(0, 2, 3, 1).flatten(1, 2)                # [B, H*W, C]  (bf16 NHWC map on the CUDA path)
        pid = patch_ids[i].reshape(-1)[: int(min(num_patches, flat.shape[1]))]
        x = flat[:, pid, :].flatten(0, 1)                                # [B*P, C]
        if use_mlp:
            # with palette_oracle.EMULATE_BF16 set, tensors the CUDA path stores as bf16 are rounded at the same places
            x = _r(F.relu(F.linear(x, _rw(sd["mlp_%d.0.weight" % i]), sd["mlp_%d.0.bias" % i])))
            x = _r(F.linear(x, _rw(sd["mlp_%d.2.weight" % i]), sd["mlp_%d.2.bias" % i]))
        out.append(F.normalize(x, eps=1e-7))
    return out


def sinkhorn_ot(q: torch.Tensor, k: torch.Tensor, eps: float = 1.0, max_iter: int = 50) -> torch.Tensor:
    """NCE/sinkhorn.py OT(q, k, cost_type="hard"): K = exp(q k^T / eps) with the diagonal at exp(-10 / eps), `max_iter`
    Sinkhorn scalings (row sums -> out/in, column sums -> 1); returned in the [group, query, key] orientation that
    MoNCELoss uses after its two transposes.  q keeps its graph (the reference differentiates through the iterations)."""
    n, p, _ = q.shape
    c = torch.einsum("bid,bod->bio", q, k)
    c = c.masked_fill(torch.eye(p, dtype=torch.bool)[None], -10.0)
    kk = torch.exp(c / eps)
    u = kk.new_ones((n, p))
    v = kk.new_ones((n, p))
    for _ in range(max_iter):
        u = 1.0 / torch.bmm(kk, v.view(n, p, 1)).view(n, p)
        v = 1.0 / torch.bmm(u.view(n, 1, p), kk).view(n, p)
    return u.view(n, p, 1) * (kk * v.view(n, 1, p))


def patch_nce_loss(feat_q: torch.Tensor, feat_k: torch.Tensor, batch: int, T: float = 0.07,
                   all_negatives_from_minibatch: bool = False, kind: str = "patchnce",
                   num_patches_opt: int = 256) -> torch.Tensor:
    """PatchNCELoss.forward: feat_q / feat_k [B*P, dim] -> loss per patch [B*P].
    The reference detaches feat_k for the POSITIVE logit only (base_NCE.py:52); the negatives keep the graph, so the
    keys' MLP (and, in cut_model, the generator features of the source image) receive gradient through them."""
    n, dim = feat_q.shape
    l_pos = (feat_q * feat_k.detach()).sum(dim=1, keepdim=True)
    bdim = 1 if all_negatives_from_minibatch else batch
    q = feat_q.view(bdim, -1, dim)
    k = feat_k.view(bdim, -1, dim)
    npatches = q.shape[1]
    l_neg = torch.bmm(q, k.transpose(2, 1))
    if kind == "monce":
        # MoNCELoss.compute_l_neg_curbatch (NCE/monce.py:16-33): optimal-transport weights of the negatives, scaled by
        # (--alg_cut_num_patches - 1), enter the logits as T * log f (added BEFORE the diagonal is masked)
        f = sinkhorn_ot(q, k.detach()) * (num_patches_opt - 1) + 1e-8
        l_neg = l_neg + torch.log(f) * T
    elif kind != "patchnce":
        raise NotImplementedError(kind)
    eye = torch.eye(npatches, dtype=torch.bool)[None]
    l_neg = l_neg.masked_fill(eye, -10.0).view(-1, npatches)
    out = torch.cat((l_pos, l_neg), dim=1) / T
    return F.cross_entropy(out, torch.zeros(n, dtype=torch.long), reduction="none")


def nce_loss_total(q_pool: List[torch.Tensor], k_pool: List[torch.Tensor], batch: int, T: float = 0.07,
                   lambda_nce: float = 1.0, all_negatives_from_minibatch: bool = False, kind: str = "patchnce",
                   num_patches_opt: int = 256, n_layers: Optional[int] = None) -> torch.Tensor:
    """CUTModel.calculate_NCE_loss without weights.  n_layers: the reference divides by len(--alg_cut_nce_layers)
    (cut_model.py:892, 909) even when the encoder is too short to provide all of them; default = the pools given."""
    total = 0.0
    for fq, fk in zip(q_pool, k_pool):
        total = total + (patch_nce_loss(fq, fk, batch, T, all_negatives_from_minibatch, kind, num_patches_opt)
                         * lambda_nce).mean()
    return total / (len(q_pool) if n_layers is None else n_layers)


def cut_train_step(state_G, state_F, state_D, oc_G, oc_F, oc_D, real_A, real_B, ids_A, ids_B, nce_layers, n_blocks=9,
                   n_layers=3, lambda_gan=1.0, lambda_nce=1.0, T=0.07, num_patches=256, mode="lsgan", nce_idt=True,
                   nce_kind="patchnce"):
    """One CUTModel.optimize_parameters(): the (G_A, F) group then the (D_B_basic) group
    (cut_model.py:406-437: forward_cut :608-640, compute_G_loss_GAN base_gan_model.py:467-503, compute_G_loss_cut
    :715-845, compute_D_loss base_gan_model.py:382-419).  ids_A / ids_B: the randperm draws of the two
    calculate_feats calls (one LongTensor per NCE layer), in the reference's order.
    Returns the losses {G_tot, G_GAN, G_NCE, G_NCE_Y, D_tot}."""
    from . import gan_oracle as G
    from .palette_oracle import adam_update
    gl = {k: v.detach().clone().requires_grad_(True) for k, v in state_G.params.items()}
    fl = {k: v.detach().clone().requires_grad_(True) for k, v in state_F.params.items()}
    dl = {k: v.detach().clone() for k, v in state_D.params.items()}
    b = real_A.shape[0]
    real = torch.cat([real_A, real_B], dim=0) if nce_idt else real_A
    fake = G.resnet_generator(gl, real, n_blocks)
    fake_B, idt_B = fake[:b], fake[b:]
    loss_gan = lambda_gan * G.gan_loss(G.nlayer_discriminator(dl, fake_B, n_layers), True, mode, relu=False)

    def feats(x):
        rec = {}
        G.resnet_encoder(gl, x, n_blocks, feats=rec)
        return [rec[i] for i in nce_layers if i in rec]   # layers beyond the encoder are skipped, like compute_feats

    def nce(src, tgt, ids):
        # calculate_feats (:848-887): queries from the translated image, keys from the source, same positions
        fq, fk = feats(tgt), feats(src)
        k_pool = patch_sample(fl, fk, num_patches, ids)
        q_pool = patch_sample(fl, fq, num_patches, ids)
        return nce_loss_total(q_pool, k_pool, b, T, lambda_nce, kind=nce_kind, num_patches_opt=num_patches,
                              n_layers=len(nce_layers))

    loss_nce = nce(real_A, fake_B, ids_A)
    if nce_idt:
        loss_nce_y = nce(real_B, idt_B, ids_B)
        both = (loss_nce + loss_nce_y) * 0.5
    else:
        loss_nce_y = torch.zeros(())
        both = loss_nce
    loss_G = loss_gan + both
    loss_G.backward()
    with torch.no_grad():
        adam_update(state_G, {k: v.grad for k, v in gl.items()}, oc_G)
        adam_update(state_F, {k: v.grad for k, v in fl.items()}, oc_F)
    dl = {k: v.detach().clone().requires_grad_(True) for k, v in state_D.params.items()}
    loss_D = 0.5 * (G.gan_loss(G.nlayer_discriminator(dl, real_B, n_layers), True, mode)
                    + G.gan_loss(G.nlayer_discriminator(dl, fake_B.detach(), n_layers), False, mode))
    loss_D.backward()
    with torch.no_grad():
        adam_update(state_D, {k: v.grad for k, v in dl.items()}, oc_D)
    return {"G_tot": float(loss_G), "G_GAN": float(loss_gan), "G_NCE": float(loss_nce), "G_NCE_Y": float(loss_nce_y),
            "D_tot": float(loss_D)}


def monce_explicit(q, k, gout, groups, T=0.07, num_patches_opt=256, iters=50):
    """The MoNCE forward and its HAND-DERIVED backward exactly as csrc/nce.cu computes them (monce_fwd_kernel /
    monce_bwd_kernel): scalings stored per iteration, softmax weights, then the reverse sweep over the iterations —
    in fp64 on the CPU, to be held against autograd through patch_nce_loss(kind="monce").  Returns (loss, dq, dk)."""
    rows, d = q.shape
    p = rows // groups
    q = q.detach().double().view(groups, p, d)
    k = k.detach().double().view(groups, p, d)
    g = gout.double().view(groups, p)
    eye = torch.eye(p, dtype=torch.bool)
    cm = torch.einsum("gid,gjd->gij", q, k)
    kk = torch.exp(cm.masked_fill(eye, -10.0))
    us, vs = [], [torch.ones(groups, p, dtype=torch.double)]
    u, v = None, vs[0]
    for _ in range(iters):
        u = 1.0 / torch.einsum("gij,gj->gi", kk, v)
        us.append(u)
        v = 1.0 / torch.einsum("gij,gi->gj", kk, u)
        vs.append(v)
    f = u[:, :, None] * kk * v[:, None, :] * (num_patches_opt - 1) + 1e-8
    pos = (q * k).sum(-1) / T
    neg = (cm / T + torch.log(f)).masked_fill(eye, -10.0 / T)
    m = torch.maximum(pos, neg.max(-1).values)
    lse = m + torch.log(torch.exp(pos - m) + torch.exp(neg - m[..., None]).sum(-1))
    p0 = torch.exp(pos - lse)
    pij = torch.exp(neg - lse[..., None]).masked_fill(eye, 0.0)
    w = g[..., None] * pij / T                                           # direct softmax term
    dq = (g * (p0 - 1))[..., None] * k / T + torch.einsum("gij,gjd->gid", w, k)
    dk = torch.einsum("gij,gid->gjd", w, q)
    fbar = g[..., None] * pij * (num_patches_opt - 1) / f                # through T log f
    ubar = (fbar * kk * v[:, None, :]).sum(-1)
    vbar = (fbar * u[:, :, None] * kk).sum(-2)
    kbar = fbar * u[:, :, None] * v[:, None, :]
    for t in reversed(range(iters)):
        u_t, v_t, v_prev = us[t], vs[t + 1], vs[t]
        sbar = -vbar * v_t * v_t                                          # v_t = 1 / (K^T u_t)
        ubar = ubar + torch.einsum("gij,gj->gi", kk, sbar)
        kbar = kbar + u_t[:, :, None] * sbar[:, None, :]
        rbar = -ubar * u_t * u_t                                          # u_t = 1 / (K v_prev)
        vbar = torch.einsum("gij,gi->gj", kk, rbar)
        kbar = kbar + rbar[:, :, None] * v_prev[:, None, :]
        ubar = torch.zeros_like(ubar)
    dq = dq + torch.einsum("gij,gjd->gid", (kbar * kk).masked_fill(eye, 0.0), k)   # the OT branch reaches q only
    return (lse - pos).reshape(-1), dq.reshape(rows, d), dk.reshape(rows, d)
