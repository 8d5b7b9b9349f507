"""CUT (BASELINE.json config 3) through the reference's OWN control path, on CPU in the build container (TEST
INFRASTRUCTURE — see oracle/__init__.py):

    python -m oracle.gen_golden_cut_plumbing            # writes tests/golden/cut_plumbing.pt (MoNCE, the example's loss)
    python -m oracle.gen_golden_cut_plumbing patchnce   # writes tests/golden/cut_plumbing_patchnce.pt

options (example_gan_horse2zebra.json, reduced: resnet 2 blocks ngf 16, D_netDs ["basic"] ndf 16, 32x32, batch 2, 16
patches) -> create_model -> data_dependent_initialize -> setup -> two optimize_parameters().  Stored: the losses of
both steps and per-tensor (sum, L2) of every G / F / D parameter afterwards, plus the option values the oracle needs.
The random patch positions are replayed from the seed set before each step (nothing else in the step draws).
"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cut_oracle as C  # noqa: E402
from oracle import gan_oracle as G  # noqa: E402
from oracle import palette_oracle as O  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SIZE, BATCH, NGF, NB, NDF, P = 32, 2, 16, 2, 16, 16


def batch(seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(BATCH, 3, SIZE, SIZE, generator=g) * 2 - 1, torch.rand(BATCH, 3, SIZE, SIZE, generator=g) * 2 - 1)


def patch_ids(seed, hw_per_layer):
    """The draws of the two calculate_feats calls of one step, in order: randperm(H*W)[:P] per NCE layer."""
    torch.manual_seed(seed)
    ids_a = [torch.randperm(hw)[: min(P, hw)] for hw in hw_per_layer]
    ids_b = [torch.randperm(hw)[: min(P, hw)] for hw in hw_per_layer]
    return ids_a, ids_b


def seeded(net, seed):
    shapes = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    return shapes, init_params_from_shapes(shapes, seed)


def create_reference_model(nce_loss=None):
    """options -> create_model -> data_dependent_initialize -> setup of the reference's cut_model on CPU, seeded weights
    loaded -> (model, opt, seeded parameter dicts)"""
    ref_stubs.install()
    import train as ref_train
    from models import create_model
    from options.train_options import TrainOptions

    with open(os.path.join(ref_stubs.REFERENCE_ROOT, "examples", "example_gan_horse2zebra.json")) as f:
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
                 "dataroot": tmp, "checkpoints_dir": tmp, "name": "golden", "G_netG": "resnet", "G_nblocks": NB,
                 "G_ngf": NGF, "D_netDs": ["basic"], "D_ndf": NDF, "output_no_html": True, "alg_cut_num_patches": P,
                 "train_G_lr": 1e-3, "train_D_lr": 5e-4, "train_iter_size": 1})  # the example accumulates 8 iterations
    if nce_loss is not None:
        flat["alg_cut_nce_loss"] = nce_loss
    opt = TrainOptions().parse_json(flat, save_config=False)
    opt.use_cuda = False
    opt.optim = ref_train.optim
    opt.jg_dir = ref_stubs.REFERENCE_ROOT
    opt.total_iters = 0
    opt.num_test_images = 0
    torch.manual_seed(5)
    model = create_model(opt, 0)
    a0, b0 = batch(40)
    model.data_dependent_initialize({"A": a0, "B": b0, "A_img_paths": ["a"] * BATCH, "B_img_paths": ["b"] * BATCH})
    model.setup(opt)
    shapes_G, pG = seeded(model.netG_A, 51)
    shapes_F, pF = seeded(model.netF, 52)
    shapes_D, pD = seeded(model.netD_B_basic, 53)
    assert dict(shapes_G) == G.resnet_param_shapes(3, 3, NGF, NB) and dict(shapes_D) == G.nlayer_d_param_shapes(3, NDF, 3)
    model.netG_A.load_state_dict(pG)
    model.netF.load_state_dict(pF)
    model.netD_B_basic.load_state_dict(pD)
    return model, opt, a0, (shapes_G, pG), (shapes_F, pF), (shapes_D, pD)


def main(nce_loss=None):
    """nce_loss None: the example's own --alg_cut_nce_loss (monce) -> cut_plumbing.pt;
    "patchnce" -> cut_plumbing_patchnce.pt (the variant the CUDA path implements)."""
    model, opt, a0, (shapes_G, pG), (shapes_F, pF), (shapes_D, pD) = create_reference_model(nce_loss)
    nce_layers = list(model.nce_layers)
    with torch.no_grad():
        hw = [f.shape[2] * f.shape[3] for f in model.netG_A.get_feats(a0, nce_layers)]
    losses, names = [], ["G_tot", "G_GAN_D_B_basic", "G_NCE", "G_NCE_Y", "D_tot"]
    for step in range(2):
        a, b = batch(100 + step)
        model.set_input({"A": a, "B": b, "A_img_paths": ["a"] * BATCH, "B_img_paths": ["b"] * BATCH})
        torch.manual_seed(1000 + step)
        model.optimize_parameters()
        losses.append({n: float(getattr(model, "loss_" + n)) for n in names})
    stat = lambda net: {k: (float(p.double().sum()), float(p.double().norm())) for k, p in net.named_parameters()}  # noqa: E731
    optim = dict(beta1=opt.train_beta1, beta2=opt.train_beta2, eps=opt.train_optim_eps,
                 weight_decay=opt.train_optim_weight_decay, kind=opt.train_optim, G_lr=opt.train_G_lr,
                 D_lr=opt.train_D_lr)
    cutopt = dict(nce_layers=nce_layers, hw=hw, T=opt.alg_cut_nce_T, lambda_NCE=opt.alg_cut_lambda_NCE,
                  lambda_GAN=opt.alg_gan_lambda if hasattr(opt, "alg_gan_lambda") else 1.0,
                  nce_idt=bool(opt.alg_cut_nce_idt), gan_mode=opt.train_gan_mode, num_patches=P,
                  all_negatives=bool(opt.alg_cut_nce_includes_all_negatives_from_minibatch), netF=opt.alg_cut_netF,
                  nce_loss=opt.alg_cut_nce_loss, G_ema=bool(opt.train_G_ema), iter_size=opt.train_iter_size)
    out = {"size": SIZE, "batch": BATCH, "ngf": NGF, "n_blocks": NB, "ndf": NDF, "seeds": (51, 52, 53),
           "data_seeds": [100, 101], "rng_seeds": [1000, 1001], "shapes_G": shapes_G, "shapes_F": shapes_F,
           "shapes_D": shapes_D, "optim": optim, "cut": cutopt, "losses": losses, "torch_version": str(torch.__version__),
           "stats_G": stat(model.netG_A), "stats_F": stat(model.netF), "stats_D": stat(model.netD_B_basic)}
    name = "cut_plumbing.pt" if nce_loss is None else "cut_plumbing_%s.pt" % nce_loss
    torch.save(out, os.path.join(GOLDEN, name))
    print(name, json.dumps(cutopt), json.dumps(optim))
    print("reference losses", losses)
    # the restatement against the reference, right here
    mk = lambda lr: O.OptimCfg(lr=lr, beta1=optim["beta1"], beta2=optim["beta2"], eps=optim["eps"],  # noqa: E731
                               weight_decay=optim["weight_decay"], kind=optim["kind"], ema_beta=0.0)
    sG, sF, sD = O.TrainState(params=pG), O.TrainState(params=pF), O.TrainState(params=pD)
    for step in range(2):
        a, b = batch(100 + step)
        ids_a, ids_b = patch_ids(1000 + step, hw)
        lo = C.cut_train_step(sG, sF, sD, mk(optim["G_lr"]), mk(optim["G_lr"]), mk(optim["D_lr"]), a, b, ids_a, ids_b,
                              nce_layers, n_blocks=NB, n_layers=3, lambda_gan=cutopt["lambda_GAN"],
                              lambda_nce=cutopt["lambda_NCE"], T=cutopt["T"], num_patches=P, mode=cutopt["gan_mode"],
                              nce_idt=cutopt["nce_idt"], nce_kind=cutopt["nce_loss"])
        print("oracle step", step, lo)
    rows = []
    for tag, state, stats in (("G", sG, out["stats_G"]), ("F", sF, out["stats_F"]), ("D", sD, out["stats_D"])):
        for k, (s, n) in stats.items():
            rows.append((abs(float(state.params[k].double().norm()) - n) / (n + 1e-12), tag, k))
    rows.sort(reverse=True)
    # biases in front of an InstanceNorm have a zero gradient in exact arithmetic: Adam turns their rounding noise
    # into +-lr steps, which no two implementations share
    print("oracle vs reference after 2 steps, largest relative parameter-norm differences:", rows[:4])
    print("... weights only:", max(r for r in rows if not r[2].endswith(".bias"))[:3])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
