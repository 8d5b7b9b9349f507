"""Golden vectors for the input-prep / wavelet ops from the UNMODIFIED reference (run in the build container):

    python -m oracle.gen_golden_prep        -> tests/golden/prep_small.pt

The reference modules cannot simply be imported: models/modules/op/upfirdn2d.py JIT-compiles its CUDA extension at
import time and data/online_creation.py drags in the whole data pipeline.  The two pure-torch functions are therefore
executed from their own, unmodified source text (ast-extracted), and HaarTransform / InverseHaarTransform are the
reference classes themselves, bound to that upfirdn2d_native.
"""
import ast
import os
import sys

import torch
import torch.nn.functional as F  # noqa: F401  (used by the extracted source)

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _extract(path, names, env):
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
    return env


def main():
    env = {"torch": torch, "F": F, "nn": torch.nn}
    _extract("models/modules/op/upfirdn2d.py", {"upfirdn2d_native"}, env)
    native = env["upfirdn2d_native"]

    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
        # the dispatcher of upfirdn2d.py on the CPU: upfirdn2d_native with symmetric up/down and (x0,x1,y0,y1) pads
        if len(pad) == 2:
            pad = (pad[0], pad[1], pad[0], pad[1])
        return native(input, kernel, up, up, down, down, *pad)

    env["upfirdn2d"] = upfirdn2d
    _extract("models/modules/freq_utils.py", {"get_haar_wavelet", "HaarTransform", "InverseHaarTransform"}, env)
    _extract("data/online_creation.py", {"fill_mask_with_random"}, env)

    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 16, 24, generator=g)
    dwt, iwt = env["HaarTransform"](3), env["InverseHaarTransform"](3)
    xr = x.clone().requires_grad_(True)
    y = dwt(xr)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    yr = torch.randn(2, 12, 8, 12, generator=g).requires_grad_(True)
    z = iwt(yr)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    # fill_mask_with_random draws its noise with torch.randn_like: replay the draw
    img = torch.randn(2, 3, 16, 24, generator=g)
    mask = torch.randint(0, 4, (2, 1, 16, 24), generator=g)
    fills = {}
    for cls in (-1, 2):
        torch.manual_seed(77 + cls)
        noise = torch.randn_like(img)
        torch.manual_seed(77 + cls)
        fills[cls] = {"noise": noise, "out": env["fill_mask_with_random"](img, mask, cls)}
    out = {"x": x, "dwt": y.detach(), "d_dwt": dy, "dx_dwt": xr.grad.clone(), "bands": yr.detach().clone(),
           "iwt": z.detach(), "d_iwt": dz, "dbands_iwt": yr.grad.clone(), "roundtrip": iwt(dwt(x)).detach(),
           "img": img, "mask": mask, "fills": fills, "torch_version": str(torch.__version__)}
    torch.save(out, os.path.join(GOLDEN, "prep_small.pt"))
    print("prep_small.pt: dwt absmax %.4f, roundtrip err %.2e" % (float(y.abs().max()),
                                                                 float((out["roundtrip"] - x).abs().max())))


if __name__ == "__main__":
    sys.exit(main())
