"""Generate the b2b / JiTViD golden vectors from the UNMODIFIED reference (/root/reference) imported on CPU in the
build container (TEST INFRASTRUCTURE — see oracle/__init__.py).

    python -m oracle.gen_golden_jit        # writes tests/golden/jit_small.pt

Fixture: B2BGenerator(JiTViD(32x32, patch 8, hidden 96, depth 6, 6 heads, 4 in-context tokens from block 2, one
MotionModule after the last block)) on a 2 x 3-frame clip, 6 input channels (condition | noisy image), seeded
de-zeroed weights; the random draws of b2b_forward (t per clip, e) are replayed from the seed.  Values: x_pred, the
masked pseudo-Huber loss, per-parameter gradient (sum, L2, first 16 values).
"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import jit_oracle as J  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CFG = dict(input_size=32, patch_size=8, in_channels=6, out_channels=3, hidden_size=96, depth=6, num_heads=6,
           num_classes=1, in_context_len=4, in_context_start=2, max_frames=8, motion_num_heads=8, motion_num_layers=2)


def build_reference(cfg: J.JitCfg):
    from models.modules.vit.vit_vid import JiTViD
    return JiTViD(input_size=cfg.input_size, patch_size=cfg.patch_size, in_channels=cfg.in_channels,
                  out_channels=cfg.out_channels, hidden_size=cfg.hidden_size, depth=cfg.depth,
                  num_heads=cfg.num_heads, num_classes=cfg.num_classes, bottleneck_dim=cfg.hidden_size,
                  in_context_len=cfg.in_context_len, in_context_start=cfg.in_context_start,
                  max_frames=cfg.max_frames, motion_num_heads=cfg.motion_num_heads,
                  motion_num_layers=cfg.motion_num_layers, motion_every=0)


def inputs(cfg, batch, frames, seed):
    g = torch.Generator().manual_seed(seed)
    shp = (batch, frames, 3, cfg.input_size, cfg.input_size)
    gt = (0.5 * torch.randn(shp, generator=g)).clamp(-1, 1)
    mask = (torch.rand((batch, frames, 1, cfg.input_size, cfg.input_size), generator=g) > 0.6).float()
    cond = gt * (1 - mask) + torch.randn(shp, generator=g) * mask
    label = torch.zeros(batch, dtype=torch.long)
    return gt, cond, mask, label


# a second configuration whose head sizes the B200 kernels take (JiT head dim 32, temporal head dim 16): the CUDA side is
# tested against it (tests/test_gpu_jit.py)
# (hidden 192: SwiGLU width int(4 * 192 * 2 / 3) = 512, JiT head dim 32, temporal head dim 24 — all multiples of 8)
CFG_B200 = dict(input_size=32, patch_size=8, in_channels=6, out_channels=3, hidden_size=192, depth=4, num_heads=6,
                num_classes=1, in_context_len=4, in_context_start=2, max_frames=8, motion_num_heads=8,
                motion_num_layers=1)


def main(name="jit_small.pt", CFG=CFG):
    ref_stubs.install()
    from models.modules.b2b_generator import B2BGenerator
    cfg = J.JitCfg(**CFG)
    opt = SimpleNamespace(alg_b2b_P_mean=-0.8, alg_b2b_P_std=0.8, alg_b2b_timestep_uniform_mix_prob=0.0,
                          alg_b2b_noise_scale=-1.0, alg_b2b_t_eps=0.05, alg_b2b_cfg_scale=1.0,
                          alg_b2b_clip_denoised=False, alg_b2b_disable_inference_clipping=True,
                          alg_b2b_denoise_timesteps=[2], G_vit_num_classes=1, alg_diffusion_dropout_prob=0.0)
    net = B2BGenerator(build_reference(cfg), sampling_method="", image_size=cfg.input_size, G_ngf=64, opt=opt)
    net.train()
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters() if v.requires_grad]
    params = init_params_from_shapes(shapes, seed=12)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    frozen = {k: v.detach().clone() for k, v in net.named_parameters() if not v.requires_grad}   # pos_embed
    batch, frames, rseed = 2, 3, 31
    gt, cond, mask, label = inputs(cfg, batch, frames, seed=14)
    torch.manual_seed(rseed)
    v_pred, v, x_pred = net(gt, mask, cond, label=label, return_x_pred=True)
    # the same draws, in the reference's order (sample_t: randn(B); then randn_like(x))
    torch.manual_seed(rseed)
    t_base = torch.sigmoid(torch.randn(batch) * opt.alg_b2b_P_std + opt.alg_b2b_P_mean)
    e = torch.randn_like(gt)
    import math
    c = 0.00054 * math.sqrt(math.prod(v_pred.shape[1:]))
    mb = torch.clamp(mask, 0, 1)     # one channel: broadcast in the numerator only, like B2BModel._masked_region_loss
    le = torch.sqrt((v_pred - v) ** 2 + c ** 2) - c
    dims = tuple(range(1, le.ndim))
    loss = ((le * mb).sum(dim=dims) / mb.sum(dim=dims).clamp_min(1e-8)).mean()     # _masked_region_loss
    loss.backward()
    grads = {}
    for k, p in net.named_parameters():
        if not p.requires_grad:
            continue
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        grads[k] = {"sum": float(g.double().sum()), "l2": float(g.double().norm()), "head": g.flatten()[:16].clone(),
                    "none": p.grad is None}
    # sampling: 3 denoising steps (2 Heun + the final Euler step), guidance neutral
    net.eval()
    torch.manual_seed(rseed + 1)
    init_noise = torch.randn_like(gt)
    restored = net.restoration(gt, cond, denoise_timesteps=3, mask=mask, labels=label, init_noise=init_noise)
    net.train()
    torch.save({"cfg": CFG, "restored": restored.clone(), "denoise_timesteps": 3, "batch": batch, "frames": frames, "wseed": 12, "dseed": 14, "rseed": rseed,
                "t_base": t_base, "torch_version": str(torch.__version__), "shapes": shapes, "frozen": frozen,
                "x_pred": x_pred.detach().clone(), "loss": float(loss.detach()), "grads": grads},
               os.path.join(GOLDEN, name))
    # the restatement against the reference, right here
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    sd = J.add_buffers({**leaves, **frozen}, cfg)
    vp2, v2, xp2 = J.b2b_forward(sd, gt, mask, cond, label, t_base, e, cfg)
    lo = J.masked_region_loss(vp2, v2, mb)
    lo.backward()
    rest2 = J.restoration(J.add_buffers({**params, **frozen}, cfg), gt, cond, mask, label, init_noise, cfg, steps=3)
    print("restoration: oracle vs reference rel max err %.2e" % float((rest2 - restored).abs().max() /
                                                                        restored.abs().max()))
    zero = lambda g, p: g if g is not None else torch.zeros_like(p)  # noqa: E731
    named = dict(net.named_parameters())
    gerr = max(float((zero(leaves[k].grad, leaves[k]) - zero(named[k].grad, named[k])).norm() /
                     (zero(named[k].grad, named[k]).norm() + 1e-9)) for k in leaves)
    print(name + ": %d parameter tensors, loss %.6f (oracle %.6f), x_pred rel max err %.2e, v err %.2e, "
          "worst grad rel L2 err %.2e, params without grad %d" % (
              len(shapes), float(loss), float(lo), float((xp2 - x_pred).abs().max() / x_pred.abs().max()),
              float((v2 - v).abs().max()), gerr, sum(g["none"] for g in grads.values())))


if __name__ == "__main__":
    main()
    main("jit_b200.pt", CFG_B200)
