"""Projected discriminator, trainable part (MultiScaleD's spectral-norm
mini-discriminators), CPU side: the oracle restatement against the golden vectors of the unmodified reference
(oracle/gen_golden_projd.py)."""
import os

import torch
import torch.nn.functional as F

from oracle import projd_oracle as P


def test_multi_scale_d_oracle_matches_reference(golden_dir):
    from oracle.gen_golden_projd import features, seeded_state
    gold = torch.load(os.path.join(golden_dir, "projd_small.pt"))
    sd = seeded_state(gold["shapes"], gold["wseed"])
    is_uv = lambda k: k.endswith(("weight_u", "weight_v"))  # noqa: E731
    leaves = {k: (v.clone() if is_uv(k) else v.clone().requires_grad_(True)) for k, v in sd.items()}
    feats = {k: v.requires_grad_(True) for k, v in features(gold["fseed"]).items()}
    new_state = {}
    logits = P.multi_scale_d(leaves, feats, gold["channels"], gold["resolutions"], training=True, new_state=new_state)
    assert logits.shape == gold["logits"].shape
    assert float((logits - gold["logits"]).abs().max()) < 1e-5 * float(gold["logits"].abs().max())
    loss = F.relu(torch.ones_like(logits) - logits).mean()
    assert abs(float(loss) - gold["loss"]) < 1e-6
    loss.backward()
    for k, g in gold["grads"].items():
        assert abs(float(leaves[k].grad.double().norm()) - g["l2"]) < 1e-4 * g["l2"] + 1e-9, k
        assert float((leaves[k].grad.flatten()[:16] - g["head"]).abs().max()) < 1e-4 * g["l2"] + 1e-9, k
    for k, ref in gold["dfeats"].items():
        assert float((feats[k].grad - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    for k, ref in gold["uv_after"].items():        # one power iteration per training forward
        assert float((new_state[k] - ref).abs().max()) < 1e-6, k
    # eval mode: no power iteration, the stored vectors are used as they are
    with torch.no_grad():
        ev = P.multi_scale_d({**sd, **new_state}, features(gold["fseed"]), gold["channels"], gold["resolutions"],
                             training=False)
    assert float((ev - gold["logits"]).abs().max()) < 1e-5 * float(gold["logits"].abs().max())


def test_disc_plan_follows_the_channel_table():
    assert P.disc_plan(64, 64) == [(64, 128), (128, 256), (256, 512)]
    assert P.disc_plan(512, 8) == []
    assert P.disc_plan(16, 32) == [(16, 256), (256, 512)]


def test_b200_multi_scale_d_keys_and_spectral_norm_match(golden_dir):
    """The part of joligen_b200.nets_projd that needs no GPU: the module tree has the reference's state_dict, and the
    spectrally normalised weight (value, power-iteration update, gradient w.r.t. weight_orig) equals the oracle's."""
    from joligen_b200 import nets_projd
    from oracle.gen_golden_projd import seeded_state
    gold = torch.load(os.path.join(golden_dir, "projd_small.pt"))
    net = nets_projd.MultiScaleD(channels=gold["channels"], resolutions=gold["resolutions"], conv=True, feats=None,
                                 num_discs=len(gold["channels"]))
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s in gold["shapes"]]
    sd = seeded_state(gold["shapes"], gold["wseed"])
    net.load_state_dict(sd)
    net.train()
    name = "mini_discs.0.main.1.main.0"
    conv = net.mini_discs["0"].main[1].main[0]
    w = nets_projd._sn_weight(conv)
    leaf = sd[name + ".weight_orig"].clone().requires_grad_(True)
    w_ref, u_ref, v_ref = P.spectral_normalize({**sd, name + ".weight_orig": leaf}, name, training=True)
    assert torch.equal(w.detach(), w_ref.detach())
    assert torch.equal(conv.weight_u, u_ref) and torch.equal(conv.weight_v, v_ref)
    g = torch.randn(w.shape, generator=torch.Generator().manual_seed(0))
    (w * g).sum().backward()
    (w_ref * g).sum().backward()
    assert torch.equal(conv.weight_orig.grad, leaf.grad)
    net.eval()
    u0 = conv.weight_u.clone()
    nets_projd._sn_weight(conv)
    assert torch.equal(conv.weight_u, u0)      # no power iteration outside training
