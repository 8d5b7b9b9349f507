"""CUT (BASELINE.json config 3) on the CPU through the kernel TEST DOUBLE: trainer_cut.CutTrainer (G on
cat(real_A, real_B), GAN + NCE + identity NCE for the (G, F) group, then D; flat Adam per group) and trainer_gan vs the
reference's own two optimize_parameters() (tests/golden/cut_plumbing*.pt) — the comparison of
tests/test_gpu_widen_cut.py — and the gradient exchange of its three parameter groups over two gloo ranks."""
import os
import socket
from unittest import mock

import pytest
import torch

import kernel_double as KD


def _build(golden_dir, golden_name, **kw):
    from joligen_b200 import nets_cut, nets_gan
    from joligen_b200.trainer_cut import CutTrainer
    from oracle.gen_golden_cut_plumbing import batch
    from oracle.vid_oracle import init_params_from_shapes
    gold = torch.load(os.path.join(golden_dir, golden_name))
    opt, cut = gold["optim"], gold["cut"]
    pG, pF, pD = (init_params_from_shapes(gold[k], seed)
                  for k, seed in zip(("shapes_G", "shapes_F", "shapes_D"), gold["seeds"]))
    netG = nets_gan.ResnetGenerator(3, 3, gold["ngf"], n_blocks=gold["n_blocks"])
    netD = nets_gan.NLayerDiscriminator(3, gold["ndf"], n_layers=3)
    netG.load_state_dict(pG)
    netD.load_state_dict(pD)
    netF = nets_cut.PatchSampleF(use_mlp=True, nc=256)
    a0, _ = batch(40)
    netF.data_dependent_initialize(netG.get_feats(a0, cut["nce_layers"]))
    assert [(k, tuple(v.shape)) for k, v in netF.named_parameters()] == [(k, tuple(s)) for k, s in gold["shapes_F"]]
    netF.load_state_dict(pF)
    with mock.patch("torch.cuda.is_available", return_value=True):
        tr = CutTrainer(netG, netF, netD, nce_layers=cut["nce_layers"], num_patches=cut["num_patches"], nce_T=cut["T"],
                        lambda_NCE=cut["lambda_NCE"], nce_idt=cut["nce_idt"], nce_loss=cut["nce_loss"],
                        gan_mode=cut["gan_mode"], lambda_gan=cut["lambda_GAN"], G_lr=opt["G_lr"], D_lr=opt["D_lr"],
                        beta1=opt["beta1"], beta2=opt["beta2"], eps=opt["eps"], weight_decay=opt["weight_decay"],
                        optim=opt["kind"], device="cpu", **kw)
    return gold, tr, (netG, netF, netD)


@pytest.mark.parametrize("golden_name", ["cut_plumbing_patchnce.pt", "cut_plumbing.pt"])
def test_cut_trainer_vs_reference_plumbing_on_the_double(golden_dir, golden_name):
    from oracle import cut_oracle as C
    from oracle import palette_oracle as O
    from oracle.gen_golden_cut_plumbing import batch, patch_ids
    from oracle.vid_oracle import init_params_from_shapes
    with KD.installed():
        gold, tr, (netG, netF, netD) = _build(golden_dir, golden_name)
        opt, cut = gold["optim"], gold["cut"]
        mk = lambda lr: O.OptimCfg(lr=lr, beta1=opt["beta1"], beta2=opt["beta2"], eps=opt["eps"],  # noqa: E731
                                   weight_decay=opt["weight_decay"], kind=opt["kind"], ema_beta=0.0)
        sG, sF, sD = (O.TrainState(params=init_params_from_shapes(gold[k], seed))
                      for k, seed in zip(("shapes_G", "shapes_F", "shapes_D"), gold["seeds"]))
        for step in range(2):
            a, b = batch(gold["data_seeds"][step])
            ids_a, ids_b = patch_ids(gold["rng_seeds"][step], cut["hw"])
            tr.set_input({"A": a, "B": b})
            tr.optimize_parameters(patch_ids_A=ids_a, patch_ids_B=ids_b)
            # the bf16-storage floor of this step: the oracle with the CUDA path's rounding points vs the fp32 golden
            O.EMULATE_BF16[0] = True
            try:
                lo = C.cut_train_step(sG, sF, sD, mk(opt["G_lr"]), mk(opt["G_lr"]), mk(opt["D_lr"]), a, b, ids_a, ids_b,
                                      cut["nce_layers"], n_blocks=gold["n_blocks"], n_layers=3,
                                      lambda_gan=cut["lambda_GAN"], lambda_nce=cut["lambda_NCE"], T=cut["T"],
                                      num_patches=cut["num_patches"], mode=cut["gan_mode"], nce_idt=cut["nce_idt"],
                                      nce_kind=cut["nce_loss"])
            finally:
                O.EMULATE_BF16[0] = False
            ref = gold["losses"][step]
            for mine, key, emu in ((tr.loss_G_tot, "G_tot", lo["G_tot"]), (tr.loss_G_GAN, "G_GAN_D_B_basic", lo["G_GAN"]),
                                   (tr.loss_G_NCE, "G_NCE", lo["G_NCE"]), (tr.loss_G_NCE_Y, "G_NCE_Y", lo["G_NCE_Y"]),
                                   (tr.loss_D_tot, "D_tot", lo["D_tot"])):
                floor = abs(float(emu) - ref[key]) / abs(ref[key])
                # step 0: a pure forward comparison; step 1 sits behind one Adam update of a GAN (sign-like steps from
                # near-zero gradients): the bounds of tests/test_gpu_widen_cut.py
                bound = max(3e-2, 3 * floor) if step == 0 else max(7e-2, 5 * floor)
                assert abs(float(mine.detach()) - ref[key]) < bound * abs(ref[key]), (step, key, float(mine.detach()), floor)
    for net, stats in ((netG, gold["stats_G"]), (netF, gold["stats_F"]), (netD, gold["stats_D"])):
        sd = net.state_dict()
        for k, (_, n) in stats.items():
            assert abs(float(sd[k].double().norm()) - n) <= (5e-2 if k.endswith(".bias") else 1e-2) * n + 1e-6, k


def _dp_worker(rank, world, port, golden_dir, out):
    import torch.distributed as dist
    from oracle.gen_golden_cut_plumbing import batch, patch_ids
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with KD.installed():
        gold, tr, nets_ = _build(golden_dir, "cut_plumbing.pt", process_group=dist.group.WORLD)
        losses = []
        for step in range(2):
            a, b = batch(700 + 10 * step + rank)
            ids_a, ids_b = patch_ids(800 + step, gold["cut"]["hw"])
            tr.set_input({"A": a, "B": b})
            tr.optimize_parameters(patch_ids_A=ids_a, patch_ids_B=ids_b)
            losses.append(float(tr.loss_G_tot))
        out[rank] = {"losses": losses, "flat": [torch.cat([p.detach().reshape(-1) for p in n.parameters()]) for n in nets_]}
    dist.barrier()
    dist.destroy_process_group()


def test_cut_trainer_two_ranks_over_gloo(golden_dir):
    """CutTrainer(process_group=...): the G, F and D gradients are averaged over the ranks before their Adam steps —
    replicas that see different images stay bit-identical."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, golden_dir, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a["losses"] != b["losses"]
    for x, y in zip(a["flat"], b["flat"]):
        assert torch.equal(x, y)
