"""PaletteTrainer on BASELINE.json configs 4 and 5 against the reference's own control path (refattn_plumbing.pt,
vid_plumbing.pt: options -> create_model -> two optimize_parameters(), oracle/gen_golden_plumbing45.py) — the analogue
of test_gpu_palette.py::test_train_steps_match_reference_plumbing for the other denoisers.

Ran green on a B200 (profiles/r02_unverified_tests_first_run.log).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(which, gold, params):
    from joligen_b200 import nets, nets_ref, nets_vid
    kw = dict(tanh=False, n_timestep_train=gold["n_timestep_train"], n_timestep_test=gold["n_timestep_test"],
              norm="groupnorm", group_norm_size=32, cond_embed_dim=32, num_heads=1, **gold["net"])
    kw["res_blocks"], kw["attn_res"] = list(kw["res_blocks"]), list(kw["attn_res"])
    unet = (nets_vid.UNetVid if which == "vid" else nets_ref.UNetGeneratorRefAttn)(**kw)
    g = nets.DiffusionGenerator(nets.PaletteDenoiseFn(unet, 32), image_size=gold["size"], G_ngf=gold["net"]["inner_channel"])
    missing, unexpected = g.load_state_dict(params, strict=False)
    assert not unexpected and not [m for m in missing if not any(t in m for t in ("gammas", "posterior", "pos_encoder.pe"))]
    return g


@pytest.mark.parametrize("which", ["ref", "vid"])
def test_train_steps_match_reference_plumbing(golden_dir, which):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200.trainer import PaletteTrainer
    from oracle import palette_oracle as O
    from oracle.gen_golden_plumbing45 import batch, draws, oracle_cfg, oracle_forward
    from oracle.vid_oracle import init_params_from_shapes
    gold = torch.load(os.path.join(golden_dir, "vid_plumbing.pt" if which == "vid" else "refattn_plumbing.pt"))
    p0 = init_params_from_shapes(gold["shapes"], gold["wseed"])
    net = _build(which, gold, p0)
    assert [(k, tuple(v.shape)) for k, v in net.named_parameters()] == [(k, tuple(s)) for k, s in gold["shapes"]]
    oc = gold["optim"]
    tr = PaletteTrainer(net, lr=oc["lr"], beta1=oc["beta1"], beta2=oc["beta2"], eps=oc["eps"],
                        weight_decay=oc["weight_decay"], optim=oc["kind"], ema=True, ema_beta=oc["ema_beta"],
                        iter_size=oc["iter_size"], lambda_G=gold["lambda_G"])
    cfg = oracle_cfg(which)
    cfg.n_timestep_train, cfg.n_timestep_test = gold["n_timestep_train"], gold["n_timestep_test"]
    emu = O.TrainState(params={k: v.clone() for k, v in p0.items()})
    for step in range(2):
        data = batch(which, gold["data_seeds"][step])
        t, u, noise = draws(which, cfg, gold["rng_seeds"][step])
        tr.set_input(data)
        loss = tr.optimize_parameters(noise=noise.cuda(), t=t.cuda(), u=u.cuda())
        O.EMULATE_BF16[0] = True
        try:
            emu_loss, _, _ = O.train_step(emu, cfg, O.OptimCfg(**oc), data["B"], data["A"], data["B_label_mask"], noise, t,
                                          u, lambda_G=gold["lambda_G"],
                                          forward=oracle_forward(which, cfg, data, noise, t, u))
        finally:
            O.EMULATE_BF16[0] = False
        assert abs(float(loss) - float(emu_loss)) < 2e-2 * abs(float(emu_loss)), step
        assert abs(float(loss) - gold["losses"][step]) < 2e-2 * abs(gold["losses"][step]), step
    sd = net.state_dict()
    # Adam's first steps move every weight by ~lr whatever the gradient scale: compare the UPDATES in aggregate
    num = den = 0.0
    for k in p0:
        upd = sd[k].cpu().double() - p0[k].double()
        upd_ref = emu.params[k].double() - p0[k].double()
        num += float((upd - upd_ref).norm()) ** 2
        den += float(upd_ref.norm()) ** 2
    assert (num / den) ** 0.5 < 0.35
    ema = tr.ema_state_dict()
    for k, (s, n) in gold["param_stats"].items():
        assert abs(float(sd[k].double().norm()) - n) <= (1e-2 if sd[k].dim() > 1 else 2e-2) * n + 1e-6, k
    for k, (s, n) in gold["ema_stats"].items():
        assert abs(float(ema[k].double().norm()) - n) <= (1e-2 if ema[k].dim() > 1 else 2e-2) * n + 1e-6, k
