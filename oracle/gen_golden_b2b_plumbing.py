"""BASELINE.json config 5 as written (`b2b_model`, `G_netG vit_vid`, JiTVid-B/16) through the reference's OWN control
path, on CPU in the build container (TEST INFRASTRUCTURE — see oracle/__init__.py):

    python -m oracle.gen_golden_b2b_plumbing        # writes tests/golden/b2b_plumbing.pt

options (example_b2b_vid_mario.json at 32x32, 3 frames, batch 2; perceptual losses off — LPIPS / DISTS are third-party
frozen networks —, the autoregressive reference-frame trick off) -> create_model -> setup -> two
optimize_parameters() (AdamW betas (0.9, 0.95), EMA).  Stored: both losses, per-tensor (sum, L2) of parameters and EMA.
The draws of B2BGenerator.b2b_forward are replayed from the seed: sample_t = randn(B), rand(B), rand(B) (logit-normal
with a 10 % uniform mix, b2b_generator.py:190-206), then e = randn_like(x).
"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import jit_oracle as J  # noqa: E402
from oracle import palette_oracle as O  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SIZE, BATCH, FRAMES = 32, 2, 3
CFG = dict(input_size=SIZE, patch_size=16, in_channels=3, out_channels=3, hidden_size=768, depth=12, num_heads=12,
           num_classes=1, in_context_len=32, in_context_start=4, max_frames=8, motion_num_heads=8, motion_num_layers=2)


def batch(seed):
    g = torch.Generator().manual_seed(seed)
    gt = (0.5 * torch.randn(BATCH, FRAMES, 3, SIZE, SIZE, generator=g)).clamp(-1, 1)
    mask = (torch.rand(BATCH, FRAMES, 1, SIZE, SIZE, generator=g) > 0.6).long()
    cond = gt * (1 - mask) + torch.randn(gt.shape, generator=g) * mask
    return {"A": cond, "B": gt, "B_label_mask": mask}


def draws(seed, p_mean, p_std, mix):
    """(t_base [B], e) exactly as b2b_forward draws them after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    t = torch.sigmoid(torch.randn(BATCH) * p_std + p_mean)
    if mix > 0.0:
        t_uniform = torch.rand_like(t)
        t = torch.where(torch.rand_like(t) < mix, t_uniform, t)
    e = torch.randn(BATCH, FRAMES, 3, SIZE, SIZE)
    return t, e


def create_reference_model():
    """options -> create_model -> setup of the reference's b2b_model on CPU, seeded weights loaded
    -> (model, opt, shapes, params, frozen)"""
    ref_stubs.install()
    import train as ref_train
    from models import create_model
    from options.train_options import TrainOptions
    with open(os.path.join(ref_stubs.REFERENCE_ROOT, "examples", "example_b2b_vid_mario.json")) as f:
        nested = json.load(f)

    def flatten(d, prefix=""):
        flat = {}
        for k, v in d.items():
            if isinstance(v, dict):
                flat.update(flatten(v, prefix + k + "_"))
            else:
                flat[prefix + k] = v
        return flat

    flat = flatten(nested)
    tmp = tempfile.mkdtemp()
    flat.update({"gpu_ids": "-1", "data_crop_size": SIZE, "data_load_size": SIZE, "train_batch_size": BATCH,
                 "dataroot": tmp, "checkpoints_dir": tmp, "name": "golden", "output_no_html": True,
                 "train_iter_size": 1, "data_temporal_number_frames": FRAMES, "alg_b2b_perceptual_loss": [""],
                 "alg_b2b_autoregressive": False, "train_G_lr": 1e-3, "train_G_ema": True, "train_G_ema_beta": 0.9})
    opt = TrainOptions().parse_json(flat, save_config=False)
    opt.use_cuda = False
    opt.optim = ref_train.optim
    opt.jg_dir = ref_stubs.REFERENCE_ROOT
    opt.total_iters = 0
    opt.num_test_images = 0
    torch.manual_seed(5)
    model = create_model(opt, 0)
    model.setup(opt)
    net = model.netG_A
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters() if v.requires_grad]
    params = init_params_from_shapes(shapes, 71)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected
    frozen = {k: v.detach().clone() for k, v in net.named_parameters() if not v.requires_grad}
    return model, opt, shapes, params, frozen


def main():
    torch.set_num_threads(8)
    model, opt, shapes, params, frozen = create_reference_model()
    net = model.netG_A
    gen = dict(P_mean=net.P_mean, P_std=net.P_std, mix=net.timestep_uniform_mix_prob, noise_scale=net.noise_scale,
               t_eps=net.t_eps)
    losses = []
    for step in range(2):
        data = batch(300 + step)
        model.set_input(dict(data, A_img_paths=["a"] * BATCH, B_label_cls=torch.zeros(BATCH, dtype=torch.long)))
        assert model.cond_image is None and getattr(model, "use_gt", None) is None
        torch.manual_seed(3000 + step)
        model.optimize_parameters()
        losses.append(float(model.loss_G_tot))
    # BaseModel.optimize_parameters calls set_requires_grad(net, True) on the whole network (base_model.py:1317-1322),
    # so the "fixed" sin-cos pos_embed (requires_grad=False at construction) is trained from the first step on
    assert all(p.requires_grad for p in net.parameters())
    stat = lambda n: {k: (float(p.double().sum()), float(p.double().norm())) for k, p in n.named_parameters()}  # noqa: E731
    optim = dict(lr=1e-3, beta1=opt.train_beta1, beta2=opt.train_beta2, eps=opt.train_optim_eps,
                 weight_decay=opt.train_optim_weight_decay, kind=opt.train_optim, ema_beta=0.9, iter_size=1)
    out = {"cfg": CFG, "batch": BATCH, "frames": FRAMES, "size": SIZE, "wseed": 71, "data_seeds": [300, 301],
           "rng_seeds": [3000, 3001], "shapes": shapes, "frozen": frozen, "optim": optim, "gen": gen,
           "loss_kind": opt.alg_b2b_loss, "masked_region_only": bool(opt.alg_b2b_loss_masked_region_only),
           "lambda_G": opt.alg_diffusion_lambda_G, "losses": losses, "param_stats": stat(net),
           "ema_stats": stat(model.netG_A_ema), "torch_version": str(torch.__version__)}
    torch.save(out, os.path.join(GOLDEN, "b2b_plumbing.pt"))
    print("b2b_plumbing.pt", gen, optim, "reference losses", losses)
    # the restatement against the reference, right here
    cfg = J.JitCfg(**CFG, t_eps=gen["t_eps"], noise_scale=gen["noise_scale"])
    state = O.TrainState(params={**{k: v.clone() for k, v in params.items()}, **{k: v.clone() for k, v in frozen.items()}})
    mine = []
    for step in range(2):
        data = batch(300 + step)
        t_base, e = draws(3000 + step, gen["P_mean"], gen["P_std"], gen["mix"])
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in state.params.items()}
        sd = J.add_buffers(leaves, cfg)
        loss = J.b2b_loss(sd, data["B"], data["B_label_mask"].float(), None, torch.zeros(BATCH, dtype=torch.long), t_base,
                          e, cfg, kind=out["loss_kind"], lambda_G=out["lambda_G"],
                          masked_region_only=out["masked_region_only"])
        loss.backward()
        with torch.no_grad():
            O.adam_update(state, {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()},
                          O.OptimCfg(**optim))
        mine.append(float(loss))
    rows = sorted(((abs(float(state.params[k].double().norm()) - n) / (n + 1e-12), k)
                   for k, (s, n) in out["param_stats"].items()), reverse=True)
    print("oracle losses", mine, "worst parameter-norm differences", rows[:3])


if __name__ == "__main__":
    main()
