"""Attention forward on one shape: parity vs an fp32 torch reference + timing.
    [JG_ATTN_TC=0] python tools/gpu_attn_one.py N T heads ch [layout] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import kernels as K  # noqa: E402

n, t, heads, ch = [int(v) for v in sys.argv[1:5]]
layout = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 10
g = torch.Generator(device="cuda").manual_seed(0)
c = heads * ch
qkv = (1.5 * torch.randn(n, t, 3 * c, device="cuda", generator=g)).to(torch.bfloat16)
qkv4 = qkv.view(n, t, 1, 3 * c)
out, lse = K.attn_fwd(qkv4, heads, ch, layout)
out = out.view(n, t, c)
torch.cuda.synchronize()
# reference (fp32 on the bf16-rounded inputs), a few (image, head) pairs
f = qkv.float()
worst = 0.0
for (i, h) in [(0, 0), (n - 1, heads - 1), (n // 2, heads // 2)]:
    if layout == 0:
        q, k, v = [f[i, :, h * 3 * ch + j * ch: h * 3 * ch + (j + 1) * ch] for j in range(3)]
    else:
        q, k, v = [f[i, :, j * c + h * ch: j * c + (h + 1) * ch] for j in range(3)]
    w = torch.softmax((q @ k.t()) * ch ** -0.5, dim=-1)
    ref = w @ v
    got = out[i, :, h * ch:(h + 1) * ch].float()
    err = float((got - ref).norm() / ref.norm())
    l2 = torch.logsumexp((q @ k.t()) * ch ** -0.5, dim=-1) * 1.4426950408889634
    lerr = float((lse.view(n, heads, t)[i, h] - l2).abs().max())
    worst = max(worst, err)
    print("image %d head %d: out rel-l2 %.2e, lse max-abs %.2e" % (i, h, err, lerr))
for _ in range(3):
    K.attn_fwd(qkv4, heads, ch, layout)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    K.attn_fwd(qkv4, heads, ch, layout)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("attn_fwd tc=%s %s: %.3f ms  %.1f TFLOP/s  worst rel-l2 %.2e %s" % (
    os.environ.get("JG_ATTN_TC", "1"), sys.argv[1:5], ms, 4.0 * n * heads * t * t * ch / ms / 1e9, worst,
    "OK" if worst < 1e-2 else "MISMATCH"))
