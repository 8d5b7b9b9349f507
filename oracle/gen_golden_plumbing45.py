"""BASELINE.json configs 4 and 5 through the reference's OWN control path, on CPU in the build container (TEST
INFRASTRUCTURE — see oracle/__init__.py):

    python -m oracle.gen_golden_plumbing45        # writes tests/golden/refattn_plumbing.pt, vid_plumbing.pt

options (example_ddpm_unetref_viton.json / example_ddpm_vid_mario.json, reduced: ngf 64, mults (1, 2), one res block
per level, attention at ds 2, 16x16, batch 2, 3 frames) -> create_model -> setup -> two optimize_parameters() with
AdamW + weight decay + EMA (iter_size 1).  Stored: both losses, per-tensor (sum, L2) of the parameters and of the EMA
afterwards.  The step's random draws are replayed from the seed: t ~ randint, u ~ rand, noise ~ randn_like(y_0) — for
clips in the reference's folded "b c (f h) w" layout (diffusion_generator.py:460-466).
"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import palette_oracle as O  # noqa: E402
from oracle import ref_oracle as R  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle import vid_oracle as V  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SIZE, BATCH, FRAMES = 16, 2, 3
NET = dict(image_size=SIZE, in_channel=6, inner_channel=64, out_channel=3, res_blocks=(1, 1), attn_res=(2,),
           channel_mults=(1, 2), num_head_channels=32)
EXAMPLE = {"ref": "example_ddpm_unetref_viton.json", "vid": "example_ddpm_vid_mario.json"}


def batch(which, seed):
    g = torch.Generator().manual_seed(seed)
    lead = (BATCH, FRAMES) if which == "vid" else (BATCH,)
    gt = (0.5 * torch.randn(*lead, 3, SIZE, SIZE, generator=g)).clamp(-1, 1)
    mask = (torch.rand(*lead, 1, SIZE, SIZE, generator=g) > 0.6).long()
    cond = gt * (1 - mask) + torch.randn(gt.shape, generator=g) * mask
    data = {"A": cond, "B": gt, "B_label_mask": mask}
    if which == "ref":
        data["ref_A"] = 0.5 * torch.randn(BATCH, 3, SIZE, SIZE, generator=g)
    return data


def draws(which, cfg, seed):
    """(t, u, noise) as the step draws them after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    t, u = O.sample_t_gamma(cfg, BATCH)
    if which == "vid":
        n = torch.randn(BATCH, 3, FRAMES * SIZE, SIZE)            # randn_like of the folded clip
        noise = n.reshape(BATCH, 3, FRAMES, SIZE, SIZE).permute(0, 2, 1, 3, 4).contiguous()
    else:
        noise = torch.randn(BATCH, 3, SIZE, SIZE)
    return t, u, noise


def oracle_cfg(which):
    return V.VidCfg(**NET) if which == "vid" else O.UNetCfg(**NET)


def oracle_forward(which, cfg, data, noise, t, u):
    if which == "vid":
        def fwd(leaves):
            n, nh = V.diffusion_forward_vid(V.add_buffers(leaves, cfg), data["B"], data["A"], data["B_label_mask"],
                                            noise, t, u, cfg)
            return n, nh, None
        return fwd
    return lambda leaves: O.diffusion_forward(leaves, data["B"], data["A"], data["B_label_mask"], noise, t, u, cfg,
                                              unet=R.denoiser(data["ref_A"]))


def create_reference_model(which):
    """options -> create_model -> setup of the reference's palette_model for cfg 4 ("ref") / cfg 5 ("vid") on CPU, seeded
    weights loaded -> (model, opt, shapes, params, wseed)"""
    import train as ref_train
    from models import create_model
    from options.train_options import TrainOptions
    with open(os.path.join(ref_stubs.REFERENCE_ROOT, "examples", EXAMPLE[which])) as f:
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
                 "dataroot": tmp, "checkpoints_dir": tmp, "name": "golden", "alg_diffusion_cond_embed": "",
                 "G_ngf": 64, "G_unet_mha_channel_mults": [1, 2], "G_unet_mha_res_blocks": [1, 1],
                 "G_unet_mha_attn_res": [2], "G_unet_mha_num_head_channels": 32, "train_optim": "adamw",
                 "train_G_lr": 1e-3, "train_G_ema": True, "train_G_ema_beta": 0.9, "train_optim_weight_decay": 0.01,
                 "output_no_html": True, "train_iter_size": 1})
    if which == "vid":
        flat["data_temporal_number_frames"] = FRAMES
    opt = TrainOptions().parse_json(flat, save_config=False)
    opt.use_cuda = False
    opt.optim = ref_train.optim
    opt.jg_dir = ref_stubs.REFERENCE_ROOT
    opt.total_iters = 0
    opt.num_test_images = 0
    torch.manual_seed(5)
    model = create_model(opt, 0)
    model.setup(opt)
    shapes = [(k, tuple(v.shape)) for k, v in model.netG_A.named_parameters()]
    wseed = 61 if which == "vid" else 62
    params = init_params_from_shapes(shapes, wseed)
    missing, unexpected = model.netG_A.load_state_dict(params, strict=False)
    assert not unexpected
    return model, opt, shapes, params, wseed


def run(which):
    model, opt, shapes, params, wseed = create_reference_model(which)
    cfg = oracle_cfg(which)
    cfg.n_timestep_train, cfg.n_timestep_test = opt.G_diff_n_timestep_train, opt.G_diff_n_timestep_test
    losses = []
    for step in range(2):
        data = batch(which, 200 + step)
        model.set_input(dict(data, A_img_paths=["a"] * BATCH, B_label_cls=torch.zeros(BATCH, dtype=torch.long)))
        torch.manual_seed(2000 + step)
        model.optimize_parameters()
        losses.append(float(model.loss_G_tot))
    stat = lambda net: {k: (float(p.double().sum()), float(p.double().norm())) for k, p in net.named_parameters()}  # noqa: E731
    optim = dict(lr=1e-3, beta1=opt.train_beta1, beta2=opt.train_beta2, eps=opt.train_optim_eps, weight_decay=0.01,
                 kind="adamw", ema_beta=0.9, iter_size=1)
    out = {"which": which, "net": NET, "n_timestep_train": cfg.n_timestep_train, "n_timestep_test": cfg.n_timestep_test,
           "batch": BATCH, "frames": FRAMES, "size": SIZE, "wseed": wseed, "data_seeds": [200, 201],
           "rng_seeds": [2000, 2001], "shapes": shapes, "optim": optim, "lambda_G": opt.alg_diffusion_lambda_G,
           "losses": losses, "param_stats": stat(model.netG_A), "ema_stats": stat(model.netG_A_ema),
           "torch_version": str(torch.__version__)}
    name = "vid_plumbing.pt" if which == "vid" else "refattn_plumbing.pt"
    torch.save(out, os.path.join(GOLDEN, name))
    # the restatement against the reference, right here
    state = O.TrainState(params={k: v.clone() for k, v in params.items()})
    mine = []
    for step in range(2):
        data = batch(which, 200 + step)
        t, u, noise = draws(which, cfg, 2000 + step)
        lo, _, _ = O.train_step(state, cfg, O.OptimCfg(**optim), data["B"], data["A"], data["B_label_mask"], noise, t, u,
                                lambda_G=out["lambda_G"], forward=oracle_forward(which, cfg, data, noise, t, u))
        mine.append(float(lo))
    rows = sorted(((abs(float(state.params[k].double().norm()) - n) / (n + 1e-12), k)
                   for k, (s, n) in out["param_stats"].items()), reverse=True)
    erows = sorted(((abs(float(state.ema[k].double().norm()) - n) / (n + 1e-12), k)
                    for k, (s, n) in out["ema_stats"].items()), reverse=True)
    print(name, "reference losses", losses, "oracle", mine)
    print("   worst parameter-norm differences", rows[:3], "EMA", erows[:2])


if __name__ == "__main__":
    ref_stubs.install()
    torch.set_num_threads(8)
    for w in (sys.argv[1:] or ["ref", "vid"]):
        run(w)
