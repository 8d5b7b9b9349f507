"""CPU oracle for CUT's contrastive path (TEST INFRASTRUCTURE — see oracle/__init__.py): SURVEY.md section 8(f)
rank 3, the next row after the generator / discriminator operators of oracle/gan_oracle.py.

Functional fp32 restatement of

  PatchSampleF.forward      /root/reference/models/modules/cut_networks.py:38-73
      (per NCE layer: NHWC flatten, gather of `num_patches` spatial positions — the same positions for every image of
       the batch —, Linear -> ReLU -> Linear, L2 normalisation with eps 1e-7)
  BaseNCELoss / PatchNCELoss  /root/reference/models/modules/NCE/base_NCE.py:17-77, patchnce.py
      (positive logit = <q_i, k_i>, negatives = q_i . k_j over the patches of the same image (or of the whole
       minibatch), diagonal filled with -10, cross entropy of [pos | neg] / T against class 0, no reduction)
  CUTModel.calculate_NCE_loss /root/reference/models/cut_model.py:889-909  (lambda_NCE * mean per layer, / n_layers)

The random patch positions (torch.randperm in the reference) are an input here: the golden generator records the
reference's own draws.  Pinned against the real reference by oracle/gen_golden_cut.py + tests/test_cut_oracle.py.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


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
        flat = feat.permute(0, 2, 3, 1).flatten(1, 2)                    # [B, H*W, C]
        pid = patch_ids[i].reshape(-1)[: int(min(num_patches, flat.shape[1]))]
        x = flat[:, pid, :].flatten(0, 1)                                # [B*P, C]
        if use_mlp:
            x = F.linear(x, sd["mlp_%d.0.weight" % i], sd["mlp_%d.0.bias" % i])
            x = F.linear(F.relu(x), sd["mlp_%d.2.weight" % i], sd["mlp_%d.2.bias" % i])
        out.append(F.normalize(x, eps=1e-7))
    return out


def patch_nce_loss(feat_q: torch.Tensor, feat_k: torch.Tensor, batch: int, T: float = 0.07,
                   all_negatives_from_minibatch: bool = False) -> torch.Tensor:
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
    eye = torch.eye(npatches, dtype=torch.bool)[None]
    l_neg = l_neg.masked_fill(eye, -10.0).view(-1, npatches)
    out = torch.cat((l_pos, l_neg), dim=1) / T
    return F.cross_entropy(out, torch.zeros(n, dtype=torch.long), reduction="none")


def nce_loss_total(q_pool: List[torch.Tensor], k_pool: List[torch.Tensor], batch: int, T: float = 0.07,
                   lambda_nce: float = 1.0, all_negatives_from_minibatch: bool = False) -> torch.Tensor:
    """CUTModel.calculate_NCE_loss without weights."""
    total = 0.0
    for fq, fk in zip(q_pool, k_pool):
        total = total + (patch_nce_loss(fq, fk, batch, T, all_negatives_from_minibatch) * lambda_nce).mean()
    return total / len(q_pool)
