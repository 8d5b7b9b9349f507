"""Debug aid (test infrastructure: it lives under tests/ because it calls the oracle).  Per-block error localisation:
every UNet block is fed the SAME (bf16-rounded) input on the CPU oracle (bf16-storage emulation and fp32) and on the
CUDA path; prints rel-L2 of the block outputs.        python tests/debug_blocks.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import nets, ops  # noqa: E402
from oracle import palette_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    cfg = O.UNetCfg(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,),
                    num_head_channels=16)
    params = O.init_params(cfg, 7)
    net = nets.build_palette_generator(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1),
                                       attn_res=(2,), num_head_channels=16)
    net.load_state_dict(params, strict=False)
    net = net.cuda()
    unet = net.denoise_fn.model
    g = torch.Generator().manual_seed(0)
    inp, mid, outb = O.unet_structure(cfg)
    emb = torch.randn(2, 32, generator=g)

    def blocks():
        for i, layers in enumerate(inp):
            for j, b in enumerate(layers):
                yield "denoise_fn.model.input_blocks.%d.%d" % (i, j), b, unet.input_blocks[i][j]
        for j, b in enumerate(mid):
            yield "denoise_fn.model.middle_block.%d" % j, b, unet.middle_block[j]
        for i, layers in enumerate(outb):
            for j, b in enumerate(layers):
                yield "denoise_fn.model.output_blocks.%d.%d" % (i, j), b, unet.output_blocks[i][j]

    for name, b, mod in blocks():
        hw = 32 if (b.cin in (6,) or "input_blocks.0" in name or "input_blocks.1" in name or "output_blocks.2" in name
                    or "output_blocks.3" in name) else 16
        cin = b.cin
        x = torch.randn(2, cin, hw, hw, generator=g).to(torch.bfloat16).float()
        res = {}
        for emu in (False, True):
            O.EMULATE_BF16[0] = emu
            if b.kind == "conv":
                ref = O._r(O._conv2d(x, params[name + ".weight"], params[name + ".bias"], padding=1))
            elif b.kind == "res":
                ref = O.res_block(params, name, x, emb, b, cfg)
            else:
                ref = O.attention_block(params, name, x, b)
            res[emu] = ref
        O.EMULATE_BF16[0] = False
        xd = ops.to_nhwc(x.cuda())
        with torch.no_grad():
            if b.kind == "conv":
                y = mod.forward_nhwc(xd)
            elif b.kind == "res":
                y = mod.forward_nhwc(xd, emb.cuda())
            else:
                y = mod.forward_nhwc(xd)
        yn = ops.to_nchw(y, b.cout)
        print("%-45s %-4s cin %4d cout %4d up %d down %d | cuda-vs-emu %.2e  cuda-vs-fp32 %.2e  emu-vs-fp32 %.2e" % (
            name, b.kind, b.cin, b.cout, b.up, b.down, rel(yn, res[True]), rel(yn, res[False]),
            rel(res[True], res[False])))


if __name__ == "__main__":
    main()
