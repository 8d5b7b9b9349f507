"""CUT contrastive path, CPU side: the oracle restatement of PatchSampleF +
PatchNCELoss + calculate_NCE_loss against the golden vectors of the unmodified reference (oracle/gen_golden_cut.py)."""
import os

import torch

from oracle import cut_oracle as C
from oracle.vid_oracle import init_params_from_shapes


def test_cut_nce_oracle_matches_reference(golden_dir):
    from oracle.gen_golden_cut import feature_maps
    gold = torch.load(os.path.join(golden_dir, "cut_nce.pt"))
    params = init_params_from_shapes(gold["shapes"], gold["wseed"])
    assert {k: tuple(v.shape) for k, v in params.items()} == C.mlp_param_shapes([c for c, _, _ in gold["feats"]],
                                                                                 gold["nc"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    feat_k = feature_maps(gold["kseed"])
    feat_q = [f.requires_grad_(True) for f in feature_maps(gold["qseed"])]
    k_pool = C.patch_sample(leaves, feat_k, gold["num_patches"], gold["ids"])
    q_pool = C.patch_sample(leaves, feat_q, gold["num_patches"], gold["ids"])
    for mine, ref in zip(k_pool + q_pool, gold["k_pool"] + gold["q_pool"]):
        assert mine.shape == ref.shape and float((mine - ref).abs().max()) < 1e-6
        assert float((mine.norm(dim=1) - 1).abs().max()) < 1e-5   # rows are L2-normalised
    for i, ref in enumerate(gold["per_layer"]):
        mine = C.patch_nce_loss(q_pool[i], k_pool[i], gold["batch"], gold["T"]) * gold["lambda_NCE"]
        assert float((mine - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    total = C.nce_loss_total(q_pool, k_pool, gold["batch"], gold["T"], gold["lambda_NCE"])
    assert abs(float(total) - gold["loss"]) < 1e-6 * abs(gold["loss"])
    total.backward()
    for mine, ref in zip(feat_q, gold["dfeat_q"]):
        assert float((mine.grad - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    for k, ref in gold["grads"].items():
        # the keys' MLP pass receives gradient through the (undetached) negatives, like the reference
        assert abs(float(leaves[k].grad.double().norm()) - ref["l2"]) < 1e-5 * ref["l2"], k
        assert float((leaves[k].grad.flatten()[:16] - ref["head"]).abs().max()) < 1e-5 * ref["l2"], k


def test_cut_nce_all_negatives_and_small_maps():
    """Edge cases the reference handles: fewer positions than num_patches (the id list is clamped), negatives from the
    whole minibatch (one bmm group); with identical q and k the positive logit is 1 and the loss is small, and more
    negatives can only raise it."""
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(2, 8, 2, 2, generator=g)]
    ids = [torch.randperm(4, generator=g)]
    pooled = C.patch_sample(None, feats, 16, ids, use_mlp=False)
    assert pooled[0].shape == (2 * 4, 8)
    same = C.patch_nce_loss(pooled[0], pooled[0], 2, T=0.07)
    assert same.shape == (8,) and float(same.max()) < 5e-2
    allneg = C.patch_nce_loss(pooled[0], pooled[0], 2, T=0.07, all_negatives_from_minibatch=True)
    assert allneg.shape == (8,) and bool((allneg >= same - 1e-7).all()) and float(allneg.max()) < 5e-2


import pytest  # noqa: E402


@pytest.mark.parametrize("name", ["cut_plumbing.pt", "cut_plumbing_patchnce.pt"])
def test_cut_train_steps_match_reference_plumbing(golden_dir, name):
    """BASELINE.json config 3 through the reference's own control path (options -> create_model -> two
    optimize_parameters(): G_A + F group, then D group) with MoNCE (Sinkhorn weights, the example's default) and with
    plain PatchNCE (what the CUDA path implements): the oracle's cut_train_step reproduces every loss of both steps and
    the parameters afterwards."""
    from oracle import palette_oracle as O
    from oracle.gen_golden_cut_plumbing import batch, patch_ids
    gold = torch.load(os.path.join(golden_dir, name))
    opt, cut = gold["optim"], gold["cut"]
    assert cut["nce_loss"] == ("patchnce" if "patchnce" in name else "monce") and cut["iter_size"] == 1
    assert not cut["G_ema"]
    sG, sF, sD = (O.TrainState(params=init_params_from_shapes(gold[k], seed))
                  for k, seed in zip(("shapes_G", "shapes_F", "shapes_D"), gold["seeds"]))
    mk = lambda lr: O.OptimCfg(lr=lr, beta1=opt["beta1"], beta2=opt["beta2"], eps=opt["eps"],  # noqa: E731
                               weight_decay=opt["weight_decay"], kind=opt["kind"], ema_beta=0.0)
    for step in range(2):
        a, b = batch(gold["data_seeds"][step])
        ids_a, ids_b = patch_ids(gold["rng_seeds"][step], cut["hw"])
        lo = C.cut_train_step(sG, sF, sD, mk(opt["G_lr"]), mk(opt["G_lr"]), mk(opt["D_lr"]), a, b, ids_a, ids_b,
                              cut["nce_layers"], n_blocks=gold["n_blocks"], n_layers=3, lambda_gan=cut["lambda_GAN"],
                              lambda_nce=cut["lambda_NCE"], T=cut["T"], num_patches=cut["num_patches"],
                              mode=cut["gan_mode"], nce_idt=cut["nce_idt"], nce_kind=cut["nce_loss"])
        ref = gold["losses"][step]
        for mine, key in ((lo["G_tot"], "G_tot"), (lo["G_GAN"], "G_GAN_D_B_basic"), (lo["G_NCE"], "G_NCE"),
                          (lo["G_NCE_Y"], "G_NCE_Y"), (lo["D_tot"], "D_tot")):
            assert abs(mine - ref[key]) < 1e-5 * abs(ref[key]), (step, key, mine, ref[key])
    for state, stats in ((sG, gold["stats_G"]), (sF, gold["stats_F"]), (sD, gold["stats_D"])):
        for k, (s, n) in stats.items():
            # a bias in front of an InstanceNorm has a zero gradient in exact arithmetic: Adam turns its rounding noise
            # into +-lr steps that no two implementations share
            tol = 5e-2 if k.endswith(".bias") else 1e-4
            assert abs(float(state.params[k].double().norm()) - n) <= tol * n + 1e-9, k


def test_monce_hand_derived_backward_equals_autograd():
    """The reverse sweep over the Sinkhorn scalings that csrc/nce.cu implements (restated in fp64 by
    cut_oracle.monce_explicit) against autograd through the reference-pinned patch_nce_loss(kind="monce")."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    groups, p, d, popt = 2, 24, 32, 16
    q = F.normalize(torch.randn(groups * p, d, generator=g)).requires_grad_(True)
    k = F.normalize(torch.randn(groups * p, d, generator=g)).requires_grad_(True)
    gout = torch.rand(groups * p, generator=g)
    ref = C.patch_nce_loss(q, k, groups, T=0.07, kind="monce", num_patches_opt=popt)
    ref.backward(gout)
    loss, dq, dk = C.monce_explicit(q, k, gout, groups, T=0.07, num_patches_opt=popt)
    assert float((loss.float() - ref.detach()).abs().max()) < 1e-4
    assert float((dq.float() - q.grad).abs().max()) < 1e-5 * float(q.grad.abs().max())
    assert float((dk.float() - k.grad).abs().max()) < 1e-5 * float(k.grad.abs().max())
