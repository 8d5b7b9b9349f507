"""Generate the projected-discriminator golden vectors from the UNMODIFIED reference (/root/reference) imported on CPU
in the build container (TEST INFRASTRUCTURE — see oracle/__init__.py).

    python -m oracle.gen_golden_projd        # writes tests/golden/projd_small.pt

Fixture: MultiScaleD(channels [16, 32], resolutions [32, 16]) — two spectral-norm mini-discriminators — on seeded
feature maps, batch 2, training mode (one power iteration), hinge loss of the D step on "real" features
(loss.py:70-75: relu(1 - logits).mean()); seeded weights and u / v vectors.  Values: logits, loss, per-parameter
gradient (sum, L2, first 16 values), d loss / d features, the updated u / v.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import projd_oracle as P  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CHANNELS, RESOLUTIONS, BATCH = [16, 32], [32, 16], 2


def features(seed):
    g = torch.Generator().manual_seed(seed)
    return {str(i): torch.randn(BATCH, c, r, r, generator=g) for i, (c, r) in enumerate(zip(CHANNELS, RESOLUTIONS))}


def seeded_state(shapes, seed):
    """Weights like the other fixtures; the power-iteration vectors u / v are unit random vectors."""
    sd = init_params_from_shapes(shapes, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k, shape in shapes:
        if k.endswith("weight_u") or k.endswith("weight_v"):
            sd[k] = F.normalize(torch.randn(shape, generator=g), dim=0)
    return sd


def main():
    ref_stubs.install()
    from models.modules.projected_d.discriminator import MultiScaleD
    net = MultiScaleD(channels=CHANNELS, resolutions=RESOLUTIONS, conv=True, feats=None, num_discs=2, proj_type=2,
                      cond=0)
    net.train()
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = seeded_state(shapes, seed=17)
    net.load_state_dict(sd)
    feats = {k: v.requires_grad_(True) for k, v in features(23).items()}
    logits = net(feats)
    loss = F.relu(torch.ones_like(logits) - logits).mean()
    loss.backward()
    after = {k: v.detach().clone() for k, v in net.state_dict().items() if k.endswith(("weight_u", "weight_v"))}
    grads = {k: {"sum": float(p.grad.double().sum()), "l2": float(p.grad.double().norm()),
                 "head": p.grad.flatten()[:16].clone()} for k, p in net.named_parameters()}
    torch.save({"channels": CHANNELS, "resolutions": RESOLUTIONS, "batch": BATCH, "wseed": 17, "fseed": 23,
                "shapes": shapes, "torch_version": str(torch.__version__), "logits": logits.detach().clone(),
                "loss": float(loss.detach()), "grads": grads, "dfeats": {k: v.grad.clone() for k, v in feats.items()},
                "uv_after": after}, os.path.join(GOLDEN, "projd_small.pt"))
    # the restatement against the reference, right here
    leaves = {k: (v.clone().requires_grad_(True) if not k.endswith(("weight_u", "weight_v")) else v.clone())
              for k, v in sd.items()}
    feats2 = {k: v.requires_grad_(True) for k, v in features(23).items()}
    new_state = {}
    lg2 = P.multi_scale_d(leaves, feats2, CHANNELS, RESOLUTIONS, training=True, new_state=new_state)
    lo = F.relu(torch.ones_like(lg2) - lg2).mean()
    lo.backward()
    named = dict(net.named_parameters())
    gerr = max(float((leaves[k].grad - p.grad).norm() / (p.grad.norm() + 1e-12)) for k, p in named.items())
    uerr = max(float((new_state[k] - v).abs().max()) for k, v in after.items())
    print("projd_small.pt: %d tensors, loss %.6f (oracle %.6f), logits max err %.2e, grad rel L2 err %.2e, u/v err %.2e"
          % (len(shapes), float(loss), float(lo), float((lg2 - logits).abs().max()), gerr, uerr))


if __name__ == "__main__":
    main()
