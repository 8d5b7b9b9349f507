"""Informative timings of the other SURVEY.md section 8(d) configurations through PaletteTrainer (not bench lines):
cfg 4 = UNetGeneratorRefAttn (example_ddpm_unetref_viton.json: res_blocks [2,4,4,2], attention at ds 4 and 8, 128^2,
batch 16), cfg 5 = UNetVid (example_ddpm_vid_mario.json: 8 frames; 64^2 native and 128^2, one clip per GPU).
    python tools/gpu_time_configs.py [steps] [all|ref|vid] [graph]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from joligen_b200 import lib as L, nets, nets_ref, nets_vid, synthetic  # noqa: E402
from joligen_b200.trainer import PaletteTrainer  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
COMMON = dict(in_channel=6, inner_channel=64, out_channel=3, tanh=False, n_timestep_train=2000, n_timestep_test=1000,
              norm="groupnorm", group_norm_size=32, cond_embed_dim=32, channel_mults=(1, 2, 4, 8), num_heads=1,
              num_head_channels=32)


def run(name, unet, data, images, graph):
    g = nets.DiffusionGenerator(nets.PaletteDenoiseFn(unet, 32), image_size=unet.image_size, G_ngf=64)
    synthetic.dezero_init_(g, 3)
    nparam = sum(p.numel() for p in g.parameters())
    tr = PaletteTrainer(g.cuda(), lr=1e-4, optim="adamw", ema=True, device="cuda", cuda_graph=graph, graph_warmup=2)
    tr.set_input(data)
    for _ in range(4):
        loss = tr.optimize_parameters()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        loss = tr.optimize_parameters()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    print("%-34s params %.1f M  graph %d  %8.2f ms/step  %8.1f img/s  launches/step %d  loss %.4f  peak mem %.1f GB" % (
        name, nparam / 1e6, graph, ms, images / ms * 1e3, tr.launches_per_step, float(loss),
        torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
    del tr, g
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


def clip_batch(frames, size, seed):
    d = synthetic.synthetic_batch(frames, size, seed)
    return {k: v.unsqueeze(0) for k, v in d.items()}


if __name__ == "__main__":
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    modes = (True,) if (len(sys.argv) > 3 and sys.argv[3] == "graph") else (False, True)
    for graph in modes:
        if which in ("all", "ref"):
            unet = nets_ref.UNetGeneratorRefAttn(image_size=128, res_blocks=[2, 4, 4, 2], attn_res=[4, 8], **COMMON)
            d = synthetic.synthetic_batch(16, 128, 1)
            d["ref_A"] = 0.5 * torch.randn(16, 3, 128, 128, generator=torch.Generator().manual_seed(2))
            try:
                run("cfg4 UNetRefAttn 128^2 b16", unet, d, 16, graph)
            except Exception as e:  # noqa: BLE001
                print("cfg4 graph=%d failed: %r" % (graph, e), flush=True)
        if which in ("all", "vid"):
            for size in (64, 128):
                unet = nets_vid.UNetVid(image_size=size, res_blocks=[2, 2, 2, 2], attn_res=[16], **COMMON)
                try:
                    run("cfg5 UNetVid 8x%d^2 1 clip" % size, unet, clip_batch(8, size, 1), 8, graph)
                except Exception as e:  # noqa: BLE001
                    print("cfg5 %d graph=%d failed: %r" % (size, graph, e), flush=True)
