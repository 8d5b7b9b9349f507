"""Run GroupNorm(+FiLM)+SiLU fwd/bwd on one shape (for ncu captures / bandwidth timing).
    python tools/gpu_gn_one.py N HW_side C [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import kernels as K  # noqa: E402
from joligen_b200 import lib as L  # noqa: E402

n, side, c = [int(v) for v in sys.argv[1:4]]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, side, side, c, device="cuda", generator=g).to(torch.bfloat16)
dy = torch.randn(n, side, side, c, device="cuda", generator=g).to(torch.bfloat16)
gamma = torch.ones(c, device="cuda")
beta = torch.zeros(c, device="cuda")
film = 0.1 * torch.randn(n, 2 * c, device="cuda", generator=g)
elems = x.numel()
addend = torch.randn(n, side, side, c, device="cuda", generator=g).to(torch.bfloat16)
for what in ("fwd", "bwd", "bwd+addend"):
    y, stats, ab = K.groupnorm_fwd(x, gamma, beta, 32, film=film, act=L.ACT_SILU)
    fn = (lambda: K.groupnorm_fwd(x, gamma, beta, 32, film=film, act=L.ACT_SILU)) if what == "fwd" else (
        lambda: K.groupnorm_bwd(x, dy, gamma, beta, 32, film, L.ACT_SILU, stats, ab, need_film_grad=True,
                                addend=addend if what == "bwd+addend" else None))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bpe = {"fwd": 6, "bwd": 10, "bwd+addend": 12}[what]
    print("GN %s %s: %.3f ms, %.2f TB/s of issued traffic (%d B/elem; algorithmic %d B/elem -> %.2f TB/s)" % (
        what, sys.argv[1:4], ms, elems * bpe / ms / 1e9, bpe, {"fwd": 4, "bwd": 6, "bwd+addend": 8}[what],
        elems * {"fwd": 4, "bwd": 6, "bwd+addend": 8}[what] / ms / 1e9))
