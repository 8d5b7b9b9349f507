"""GPU bring-up check of the implicit-GEMM conv kernels against torch's fp32 conv on the same
bf16-rounded operands.  Each case runs in its own subprocess (a device trap poisons the context).

    python tools/gpu_check_conv.py            # all cases -> gpurun_out/conv_check.jsonl
    python tools/gpu_check_conv.py --case 3   # one case in-process
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, N, H, W, Cin, Cout, R, stride, pad, flags)
CASES = [
    ("gemm_1x1_64_64", 1, 16, 16, 64, 64, 1, 1, 0, ""),
    ("c3_64_64_16", 2, 16, 16, 64, 64, 3, 1, 1, ""),
    ("c3_128_256_32", 2, 32, 32, 128, 256, 3, 1, 1, ""),
    ("c3_512_512_32", 2, 32, 32, 512, 512, 3, 1, 1, ""),
    ("c3_8_64_64", 2, 64, 64, 8, 64, 3, 1, 1, ""),
    ("c3_64_8_64", 2, 64, 64, 64, 8, 3, 1, 1, ""),
    ("c4_ragged_30", 2, 30, 30, 64, 128, 4, 1, 1, ""),
    ("c3_small_8x8", 4, 8, 8, 128, 128, 3, 1, 1, ""),
    ("c3_s2_64_128", 2, 32, 32, 64, 128, 3, 2, 1, "nowgrad"),
    ("c3_epi_192_64", 2, 32, 32, 192, 64, 3, 1, 1, "bias,res,silu"),
    ("c1_1024_512", 2, 32, 32, 1024, 512, 1, 1, 0, "bias"),
    ("c3_64_64_256_b4", 4, 256, 256, 64, 64, 3, 1, 1, "bias,time"),
    ("c3_256_256_64_b8", 8, 64, 64, 256, 256, 3, 1, 1, "bias,time"),
    ("c3_512_512_32_b32", 32, 32, 32, 512, 512, 3, 1, 1, "bias,time"),
    ("c3_128_128_128_b16", 16, 128, 128, 128, 128, 3, 1, 1, "bias,time"),
    ("c3_64_64_256_b32", 32, 256, 256, 64, 64, 3, 1, 1, "bias,res,time"),
    ("c3_8_64_256_b8", 8, 256, 256, 8, 64, 3, 1, 1, "bias,time"),
    ("c3_192_64_256_b8", 8, 256, 256, 192, 64, 3, 1, 1, "bias,time"),
    ("c3_64_128_64_b4_5x5", 4, 64, 64, 64, 128, 5, 1, 2, ""),
]


def run_case(i):
    import torch
    import torch.nn.functional as F
    from joligen_b200 import kernels as K
    from joligen_b200 import lib as L

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name, n, h, w, cin, cout, r, stride, pad, flags = CASES[i]
    flags = set(f for f in flags.split(",") if f)
    g = torch.Generator(device="cuda").manual_seed(100 + i)
    dev = "cuda"
    x = torch.randn(n, cin, h, w, device=dev, generator=g)
    wt = torch.randn(cout, cin, r, r, device=dev, generator=g) / (cin * r * r) ** 0.5
    bias = torch.randn(cout, device=dev, generator=g) if "bias" in flags else None
    x_nhwc = K.nchw_to_nhwc(x)
    xb = K.nhwc_to_nchw(x_nhwc)  # bf16-rounded copy of x in NCHW fp32
    res = {"case": name}
    res["layout_roundtrip_maxerr"] = float((xb - x.to(torch.bfloat16).float()).abs().max())
    wf, wd = K.pack_conv_weight(wt)
    wb = wt.to(torch.bfloat16).float()
    ho, wo = K.conv_out_size(h, w, r, r, stride, pad)
    resid = None
    ref = F.conv2d(xb, wb, bias, stride=stride, padding=pad)
    if "res" in flags:
        rr = torch.randn(n, cout, ho, wo, device=dev, generator=g)
        resid = K.nchw_to_nhwc(rr)
        ref = ref + 0.5 * K.nhwc_to_nchw(resid)
    act = L.ACT_NONE
    if "silu" in flags:
        act = L.ACT_SILU
        ref = F.silu(ref)
    y = K.conv2d_fwd(x_nhwc, wf, bias, cout, r, r, stride=stride, pad=pad, act=act, residual=resid, res_scale=0.5)
    torch.cuda.synchronize()
    yn = K.nhwc_to_nchw(y)
    scale = float(ref.abs().max()) + 1e-6
    err = (yn - ref).abs()
    res["fwd_max_err_rel"] = float(err.max()) / scale
    res["fwd_mean_err_rel"] = float(err.mean()) / scale
    res["fwd_ok"] = bool(res["fwd_max_err_rel"] < 1.5e-2)
    if not res["fwd_ok"]:
        # locate the error: which output channels / rows are wrong
        e = err.amax(dim=(0, 2, 3))
        res["bad_channels"] = int((e > 2e-2 * scale).sum())
        e2 = err.amax(dim=(0, 1))
        res["bad_pixels"] = int((e2 > 2e-2 * scale).sum())
        res["sample_y"] = yn.flatten()[:8].tolist()
        res["sample_ref"] = ref.flatten()[:8].tolist()

    if stride == 1 and "nowgrad" not in flags:
        # dgrad (as a forward conv of dy with the dgrad-packed weights) and wgrad
        dy = torch.randn(n, cout, ho, wo, device=dev, generator=g)
        dy_nhwc = K.nchw_to_nhwc(dy)
        dyb = K.nhwc_to_nchw(dy_nhwc)
        dx_ref = torch.nn.grad.conv2d_input(xb.shape, wb, dyb, stride=1, padding=pad)
        dx = K.conv2d_fwd(dy_nhwc, wd, None, cin, r, r, stride=1, pad=r - 1 - pad)
        dxn = K.nhwc_to_nchw(dx)
        s2 = float(dx_ref.abs().max()) + 1e-6
        res["dgrad_max_err_rel"] = float((dxn - dx_ref).abs().max()) / s2
        res["dgrad_ok"] = bool(res["dgrad_max_err_rel"] < 1.5e-2)
        dw_ref = torch.nn.grad.conv2d_weight(xb, wb.shape, dyb, stride=1, padding=pad)
        dw = K.conv2d_wgrad(x_nhwc, dy_nhwc, cout, r, r, stride=1, pad=pad)
        torch.cuda.synchronize()
        s3 = float(dw_ref.abs().max()) + 1e-6
        res["wgrad_max_err_rel"] = float((dw - dw_ref).abs().max()) / s3
        res["wgrad_ok"] = bool(res["wgrad_max_err_rel"] < 5e-3)
        db = K.bias_grad(dy_nhwc)
        db_ref = dyb.sum(dim=(0, 2, 3))
        res["bias_grad_err_rel"] = float((db - db_ref).abs().max()) / (float(db_ref.abs().max()) + 1e-6)

    if "time" in flags:
        flops = 2.0 * n * ho * wo * cout * cin * r * r
        for what in ("fwd", "wgrad"):
            if what == "wgrad" and stride != 1:
                continue
            fn = (lambda: K.conv2d_fwd(x_nhwc, wf, bias, cout, r, r, stride=stride, pad=pad)) if what == "fwd" else (
                lambda: K.conv2d_wgrad(x_nhwc, dy_nhwc, cout, r, r, stride=1, pad=pad))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[what + "_ms"] = ms
            res[what + "_tflops"] = flops / ms / 1e9
        # cuDNN bf16 channels_last for orientation
        xc = xb.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wc = wb.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bc = bias.to(torch.bfloat16) if bias is not None else None
        for _ in range(3):
            F.conv2d(xc, wc, bc, stride=stride, padding=pad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            F.conv2d(xc, wc, bc, stride=stride, padding=pad)
        e1.record()
        torch.cuda.synchronize()
        res["cudnn_bf16_fwd_tflops"] = flops / (e0.elapsed_time(e1) / 10) / 1e9
    return res


def main():
    if "--case" in sys.argv:
        i = int(sys.argv[sys.argv.index("--case") + 1])
        try:
            r = run_case(i)
        except Exception as e:  # noqa
            r = {"case": CASES[i][0], "error": repr(e)[:500]}
        print("RESULT " + json.dumps(r))
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    sel = range(len(CASES))
    if "--only" in sys.argv:
        sel = [int(v) for v in sys.argv[sys.argv.index("--only") + 1].split(",")]
    tag = os.environ.get("JG_CHECK_TAG", "")
    with open(os.path.join(out_dir, "conv_check%s.jsonl" % tag), "w") as f:
        for i in sel:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True,
                                   text=True, timeout=300)
                lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                if lines:
                    r = json.loads(lines[-1][7:])
                else:
                    r = {"case": CASES[i][0], "crash": p.returncode, "stdout": p.stdout[-600:], "stderr": p.stderr[-1200:]}
            except subprocess.TimeoutExpired:
                r = {"case": CASES[i][0], "timeout": True}
            r["secs"] = round(time.time() - t0, 1)
            f.write(json.dumps(r) + "\n")
            f.flush()
            print(json.dumps(r))


if __name__ == "__main__":
    main()
