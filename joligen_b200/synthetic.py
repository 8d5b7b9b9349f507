"""Synthetic inputs and seeded (de-zeroed) weights for benchmarks and smoke runs.

No datasets or checkpoints are reachable offline, so the benchmark uses inputs of the BASELINE
config-2 shape: gt ~ N(0, 0.5^2) clamped to [-1, 1], one random box mask per image
covering 10–40 % of the area (int64 0/1), cond = gt*(1-m) + N(0,1)*m (what joliGEN's
`fill_mask_with_random` produces, data/online_creation.py:1366-1376).  The reference zero-initialises
the second conv of every ResBlock, attention proj_out and the final conv; a benchmark (and a parity
test) on such a net would skip work, so weights are drawn for EVERY tensor.
(tests/test_synthetic.py checks these against the oracle's independent restatement.)
"""
import math

import numpy as np
import torch


def synthetic_batch(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    gt = (0.5 * torch.randn(batch, 3, size, size, generator=g)).clamp(-1, 1)
    mask = torch.zeros(batch, 1, size, size, dtype=torch.int64)
    for i in range(batch):
        frac = 0.1 + 0.3 * float(torch.rand((), generator=g))
        side = max(1, int(round(size * math.sqrt(frac))))
        y0 = int(torch.randint(0, size - side + 1, (), generator=g))
        x0 = int(torch.randint(0, size - side + 1, (), generator=g))
        mask[i, 0, y0:y0 + side, x0:x0 + side] = 1
    rnd = torch.randn(batch, 3, size, size, generator=g)
    cond = gt * (1 - mask) + rnd * mask
    return {"A": cond, "B": gt, "B_label_mask": mask}


def synthetic_rect_batch(batch, h, w, seed):
    """The same inputs for non-square crops (config 4's 512 x 384 VITON frames): boxes cover 10-40 % of the area."""
    g = torch.Generator().manual_seed(seed)
    gt = (0.5 * torch.randn(batch, 3, h, w, generator=g)).clamp(-1, 1)
    mask = torch.zeros(batch, 1, h, w, dtype=torch.int64)
    for i in range(batch):
        frac = math.sqrt(0.1 + 0.3 * float(torch.rand((), generator=g)))
        bh, bw = max(1, int(round(h * frac))), max(1, int(round(w * frac)))
        y0 = int(torch.randint(0, h - bh + 1, (), generator=g))
        x0 = int(torch.randint(0, w - bw + 1, (), generator=g))
        mask[i, 0, y0:y0 + bh, x0:x0 + bw] = 1
    rnd = torch.randn(batch, 3, h, w, generator=g)
    return {"A": gt * (1 - mask) + rnd * mask, "B": gt, "B_label_mask": mask}


@torch.no_grad()
def dezero_init_(module, seed, scale=1.0):
    """In-place seeded init of every parameter, in named_parameters() order: fan-in scaled normal
    weights, norm weights 1 + 0.1 N(0,1), biases 0.05 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        shape = tuple(p.shape)
        if name.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            t = scale * torch.randn(shape, generator=g) / math.sqrt(fan_in)
        p.copy_(t.to(p.device))
    return module
