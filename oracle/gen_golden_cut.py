"""Generate the CUT contrastive-path golden vectors from the UNMODIFIED reference (/root/reference) imported on CPU
in the build container (TEST INFRASTRUCTURE — see oracle/__init__.py).

    python -m oracle.gen_golden_cut        # writes tests/golden/cut_nce.pt

Fixture: PatchSampleF(use_mlp=True, nc=256) + PatchNCELoss (T 0.07, negatives from the same image) on five seeded
feature maps shaped like ResnetGenerator.get_feats(nce_layers 0,4,8,12,16) at reduced resolution, batch 2, 16
patches per layer; seeded MLP weights; the reference's own randperm draws are stored.  Values: pooled features,
per-layer per-patch losses, total loss, gradients w.r.t. the query features (full) and the MLP parameters (sum, L2,
first 16 values).
"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cut_oracle as C  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FEATS = [(3, 22, 22), (128, 8, 8), (256, 4, 4), (256, 4, 4), (256, 4, 4)]   # (C, H, W) per NCE layer
BATCH, NUM_PATCHES, NC, T, LAMBDA = 2, 16, 256, 0.07, 1.0


def feature_maps(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(BATCH, c, h, w, generator=g) for (c, h, w) in FEATS]


def main():
    ref_stubs.install()
    from models.modules.cut_networks import PatchSampleF
    from models.modules.NCE.patchnce import PatchNCELoss
    netF = PatchSampleF(use_mlp=True, init_type="normal", init_gain=0.02, nc=NC)
    netF.set_device(torch.device("cpu"))
    feat_k = feature_maps(21)
    feat_q = [f.clone().requires_grad_(True) for f in feature_maps(22)]
    netF.data_dependent_initialize(feat_k)
    shapes = [(k, tuple(v.shape)) for k, v in netF.named_parameters()]
    assert dict(shapes) == C.mlp_param_shapes([c for c, _, _ in FEATS], NC)
    params = init_params_from_shapes(shapes, seed=4)
    netF.load_state_dict(params)
    opt = SimpleNamespace(alg_cut_nce_T=T, alg_cut_nce_includes_all_negatives_from_minibatch=False,
                          alg_cut_num_patches=NUM_PATCHES)
    crit = [PatchNCELoss(opt) for _ in FEATS]
    torch.manual_seed(99)
    k_pool, ids = netF(feat_k, NUM_PATCHES, None)
    q_pool, _ = netF(feat_q, NUM_PATCHES, ids)
    per_layer = [c(feat_q=fq, feat_k=fk, current_batch=BATCH, weight=None) * LAMBDA
                 for c, fq, fk in zip(crit, q_pool, k_pool)]
    total = sum(p.mean() for p in per_layer) / len(per_layer)
    total.backward()
    out = {"feats": FEATS, "batch": BATCH, "num_patches": NUM_PATCHES, "nc": NC, "T": T, "lambda_NCE": LAMBDA,
           "kseed": 21, "qseed": 22, "wseed": 4, "shapes": shapes, "torch_version": str(torch.__version__),
           "ids": [i.reshape(-1).clone() for i in ids],
           "k_pool": [k.detach().clone() for k in k_pool], "q_pool": [q.detach().clone() for q in q_pool],
           "per_layer": [p.detach().clone() for p in per_layer], "loss": float(total.detach()),
           "dfeat_q": [f.grad.clone() for f in feat_q],
           "grads": {k: {"sum": float(p.grad.double().sum()), "l2": float(p.grad.double().norm()),
                         "head": p.grad.flatten()[:16].clone()} for k, p in netF.named_parameters()}}
    torch.save(out, os.path.join(GOLDEN, "cut_nce.pt"))
    # the restatement against the reference, right here
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    fq2 = [f.detach().clone().requires_grad_(True) for f in feat_q]
    k2 = C.patch_sample(leaves, feat_k, NUM_PATCHES, out["ids"])  # with grad: the negatives are not detached
    q2 = C.patch_sample(leaves, fq2, NUM_PATCHES, out["ids"])
    tot2 = C.nce_loss_total(q2, k2, BATCH, T, LAMBDA)
    tot2.backward()
    gerr = max(float((leaves[k].grad - p.grad).abs().max() / (p.grad.abs().max() + 1e-12))
               for k, p in netF.named_parameters())
    ferr = max(float((a.grad - b.grad).abs().max() / (b.grad.abs().max() + 1e-12)) for a, b in zip(fq2, feat_q))
    print("cut_nce.pt: loss %.6f (oracle %.6f), pooled max err %.2e, MLP grad rel err %.2e, dfeat rel err %.2e" % (
        float(total), float(tot2), max(float((a - b).abs().max()) for a, b in zip(q2, q_pool)), gerr, ferr))


if __name__ == "__main__":
    main()
