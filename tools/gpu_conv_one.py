"""Run one conv shape a few times (for ncu captures / quick timing).
    python tools/gpu_conv_one.py N H W Cin Cout K [what=fwd|fwdres|fwdstats|dgradgn|wgrad] [reps]
fwdres = forward with the residual-add epilogue; fwdstats = forward + the fused GroupNorm statistics; dgradgn = forward
kernel with the fused GroupNorm-backward sums; inputs rotate over 3 buffers so that nothing is L2-resident."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import kernels as K  # noqa: E402

n, h, w, cin, cout, k = [int(v) for v in sys.argv[1:7]]
what = sys.argv[7] if len(sys.argv) > 7 else "fwd"
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, h, w, cin, device="cuda", generator=g).to(torch.bfloat16)
wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
bias = torch.zeros(cout, device="cuda")
wf, wd = K.pack_conv_weight(wt)
dy = torch.randn(n, h, w, cout, device="cuda", generator=g).to(torch.bfloat16)
xs = [x, x.clone(), x.clone()]
dys = [dy, dy.clone(), dy.clone()]
stats = torch.zeros(n, cout, 2, device="cuda")
ab = torch.stack([1 + 0.1 * torch.randn(n, cout, device="cuda"), 0.1 * torch.randn(n, cout, device="cuda")], -1).contiguous()
it = [0]


def fn():
    it[0] += 1
    xi, di = xs[it[0] % 3], dys[it[0] % 3]
    if what == "fwd":
        return K.conv2d_fwd(xi, wf, bias, cout, k, k)
    if what == "fwdres":
        return K.conv2d_fwd(xi, wf, bias, cout, k, k, residual=di)
    if what == "fwdstats":
        return K.conv2d_fwd(xi, wf, bias, cout, k, k, stats=stats)
    if what == "dgradgn":
        return K.conv2d_fwd(xi, wf, None, cout, k, k, gn=(di, ab, 4, stats))
    return K.conv2d_wgrad(xi, di, cout, k, k)


for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("%s %s: %.3f ms  %.1f TFLOP/s" % (what, sys.argv[1:7], ms, 2.0 * n * h * w * cin * cout * k * k / ms / 1e9))
