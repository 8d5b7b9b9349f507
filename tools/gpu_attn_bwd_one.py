"""Attention backward on one shape: parity vs torch autograd (fp32 on the bf16-rounded inputs) + timing.
    [JG_ATTN_TC=0] python tools/gpu_attn_bwd_one.py N T heads ch [layout] [reps]"""
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
qkv = (1.5 * torch.randn(n, t, 1, 3 * c, device="cuda", generator=g)).to(torch.bfloat16)
d_out = torch.randn(n, t, 1, c, device="cuda", generator=g).to(torch.bfloat16)
out, lse = K.attn_fwd(qkv, heads, ch, layout)
dqkv = K.attn_bwd(qkv, out, d_out, lse, heads, ch, layout)
torch.cuda.synchronize()
worst = 0.0
for (i, h) in [(0, 0), (n - 1, heads - 1), (n // 2, heads // 2)]:
    f = qkv[i, :, 0].float()
    if layout == 0:
        sl = [slice(h * 3 * ch + j * ch, h * 3 * ch + (j + 1) * ch) for j in range(3)]
    else:
        sl = [slice(j * c + h * ch, j * c + (h + 1) * ch) for j in range(3)]
    q, k, v = [f[:, s].clone().requires_grad_(True) for s in sl]
    w = torch.softmax((q @ k.t()) * ch ** -0.5, dim=-1)
    (w @ v).backward(d_out[i, :, 0, h * ch:(h + 1) * ch].float())
    errs = []
    for name, ref, s in zip("qkv", (q.grad, k.grad, v.grad), sl):
        got = dqkv[i, :, 0, s].float()
        errs.append(float((got - ref).norm() / ref.norm()))
    worst = max(worst, max(errs))
    print("image %d head %d: dq %.2e dk %.2e dv %.2e" % ((i, h) + tuple(errs)))
for _ in range(3):
    K.attn_bwd(qkv, out, d_out, lse, heads, ch, layout)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    K.attn_bwd(qkv, out, d_out, lse, heads, ch, layout)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("attn_bwd tc=%s %s: %.3f ms  %.1f TFLOP/s (5 GEMMs)  worst rel-l2 %.2e %s" % (
    os.environ.get("JG_ATTN_TC", "1"), sys.argv[1:5], ms, 10.0 * n * heads * t * t * ch / ms / 1e9, worst,
    "OK" if worst < 1.5e-2 else "MISMATCH"))
