"""oracle/prep_oracle.py (input preparation + Haar wavelets) against the unmodified reference's vectors
(tests/golden/prep_small.pt, oracle/gen_golden_prep.py).  CPU only."""
import os

import torch


def test_prep_oracle_matches_reference_golden(golden_dir):
    from oracle import prep_oracle as P
    g = torch.load(os.path.join(golden_dir, "prep_small.pt"))
    xr = g["x"].clone().requires_grad_(True)
    y = P.haar_dwt(xr)
    assert torch.allclose(y, g["dwt"], atol=1e-6)
    y.backward(g["d_dwt"])
    assert torch.allclose(xr.grad, g["dx_dwt"], atol=1e-6)
    br = g["bands"].clone().requires_grad_(True)
    z = P.haar_iwt(br)
    assert torch.allclose(z, g["iwt"], atol=1e-6)
    z.backward(g["d_iwt"])
    assert torch.allclose(br.grad, g["dbands_iwt"], atol=1e-6)
    assert torch.allclose(P.haar_iwt(P.haar_dwt(g["x"])), g["x"], atol=2e-6)  # perfect reconstruction
    for cls, rec in g["fills"].items():
        assert torch.equal(P.fill_mask_with_random(g["img"], g["mask"], cls, rec["noise"]), rec["out"])
    # conditioning dropout: dropped samples carry the unconditioned class everywhere, the others are untouched
    drop_u = torch.tensor([0.05, 0.9])
    out = P.mask_class_dropout(g["mask"], drop_u, 0.1, num_classes=5)
    assert bool((out[0] == 4).all()) and torch.equal(out[1], g["mask"][1])
