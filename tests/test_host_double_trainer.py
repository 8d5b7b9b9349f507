"""PaletteTrainer (flat parameter / gradient / Adam / EMA buffers, staged split-K weight gradients unpacked by one
batched launch, batched bf16 weight re-pack after the optimizer, checkpoint state, gradient accumulation, the bucketed
gradient exchange) on the CPU through the kernel TEST DOUBLE: the trainer's raw-pointer tables are followed in host
memory (tests/kernel_double.py).  The trainer itself still refuses to be BUILT without CUDA — the tests lift that one
guard (`torch.cuda.is_available`) during construction; see tests/test_host_double.py for what a pass proves."""
import os
import socket
from unittest import mock

import pytest
import torch

import kernel_double as KD

CFG = dict(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,), num_head_channels=16)


def rel_l2(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _net(seed):
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    net = nets.build_palette_generator(**CFG)
    net.load_state_dict(O.init_params(O.UNetCfg(**CFG), seed), strict=False)
    return net


def _trainer(net, **kw):
    from joligen_b200.trainer import PaletteTrainer
    with mock.patch("torch.cuda.is_available", return_value=True):
        return PaletteTrainer(net, device="cpu", **kw)


def _draw(seed, batch=2):
    from oracle import palette_oracle as O
    cfg = O.UNetCfg(**CFG)
    data = O.synthetic_batch(batch, cfg.image_size, seed)
    torch.manual_seed(seed + 7)
    t, u = O.sample_t_gamma(cfg, batch)
    return {"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"]}, torch.randn_like(data["gt"]), t, u


def _step(tr, seed, batch=2):
    data, noise, t, u = _draw(seed, batch)
    tr.set_input(data)
    return float(tr.optimize_parameters(noise=noise, t=t, u=u))


def test_trainer_matches_reference_plumbing_on_the_double(golden_dir):
    """Two optimize_parameters() (AdamW + weight decay + EMA) vs the reference's own control path (golden)."""
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    gold = torch.load(os.path.join(golden_dir, "palette_plumbing.pt"))
    cfg = O.UNetCfg(**gold["cfg"])
    net = nets.build_palette_generator(image_size=cfg.image_size, inner_channel=cfg.inner_channel,
                                       res_blocks=cfg.res_blocks, attn_res=cfg.attn_res,
                                       channel_mults=cfg.channel_mults, num_head_channels=cfg.num_head_channels)
    net.load_state_dict(O.init_params(cfg, gold["wseed"]), strict=False)
    oc = gold["optim"]
    with KD.installed():
        tr = _trainer(net, lr=oc["lr"], beta1=oc["beta1"], beta2=oc["beta2"], eps=oc["eps"],
                      weight_decay=oc["weight_decay"], optim=oc["kind"], ema=True, ema_beta=oc["ema_beta"],
                      iter_size=oc["iter_size"], lambda_G=gold["lambda_G"], use_minsnr=gold["minsnr"])
        assert len(tr.wstage.slots) > 10 and len(tr.packset.packs) > 10  # the staged / batched paths are the ones running
        for step in range(2):
            data = O.synthetic_batch(gold["batch"], gold["size"], gold["data_seeds"][step])
            torch.manual_seed(gold["rng_seeds"][step])
            t, u = O.sample_t_gamma(cfg, gold["batch"])
            noise = torch.randn_like(data["gt"])
            tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"]})
            loss = tr.optimize_parameters(noise=noise, t=t, u=u)
            assert abs(float(loss) - gold["losses"][step]) < 2e-2 * abs(gold["losses"][step]), step
        sd, ema = net.state_dict(), tr.ema_state_dict()
    for k, (_, n) in gold["param_stats"].items():
        assert abs(float(sd[k].double().norm()) - n) <= (1e-2 if sd[k].dim() > 1 else 2e-2) * n + 1e-6, k
    for k, (_, n) in gold["ema_stats"].items():
        assert abs(float(ema[k].double().norm()) - n) <= (1e-2 if ema[k].dim() > 1 else 2e-2) * n + 1e-6, k


def test_staged_weight_gradients_equal_plain_autograd():
    """The trainer's gradient (persistent split-K accumulators, one batched unpack, loose gradients folded in) equals
    the gradient plain autograd puts into .grad for the same net, data and draws."""
    with KD.installed():
        tr = _trainer(_net(3), lr=1e-3)
        data, noise, t, u = _draw(40)
        tr.set_input(data)
        g = tr.reduced_gradient(noise=noise, t=t, u=u)
        flat = tr.flat.unflatten(g)
        plain = _net(3)
        loss = plain.forward_loss(data["B"], data["A"], data["B_label_mask"], noise=noise, t=t, u=u)
        loss.backward()
    for k, p in plain.named_parameters():
        assert torch.allclose(flat[k], p.grad, rtol=1e-5, atol=1e-7 * float(p.grad.abs().max()) + 1e-12), k


def test_checkpoint_resume_equals_continuous_run_and_late_weight_loads_repack():
    """(a) a resumed run (net + trainer state_dict into a FRESH trainer) continues exactly like the
    uninterrupted one; (b) weights written through torch AFTER the trainer exists (load_state_dict of a checkpoint)
    reach the convolutions — the managed bf16 copies are re-packed."""
    with KD.installed():
        a = _trainer(_net(5), lr=1e-3, ema_beta=0.9)
        for s in (60, 61):
            _step(a, s)
        net_sd = {k: v.clone() for k, v in a.netG_A.state_dict().items()}
        tr_sd = a.state_dict()
        want = _step(a, 62)
        b = _trainer(_net(99), lr=1e-3, ema_beta=0.9)       # different weights: everything must come from the checkpoint
        stale = _step(b, 70)
        b.netG_A.load_state_dict(net_sd)
        b.load_state_dict(tr_sd)
        got = _step(b, 62)
        assert got == want and got != stale
        for k, v in a.netG_A.state_dict().items():
            assert torch.equal(v, b.netG_A.state_dict()[k]), k
        assert torch.equal(a.exp_avg, b.exp_avg) and torch.equal(a.exp_avg_sq, b.exp_avg_sq) and torch.equal(a.ema, b.ema)
        assert a.step == b.step == 3


def test_gradient_accumulation_equals_one_step_on_the_mean_gradient():
    """train_iter_size = 2 (base_model.py:1256-1282): two micro-steps of loss / 2, one optimizer step."""
    with KD.installed():
        acc = _trainer(_net(8), lr=1e-3, iter_size=2)
        for s in (80, 81):
            _step(acc, s)
        assert acc.step == 1
        one = _trainer(_net(8), lr=1e-3)
        g = None
        for s in (80, 81):
            data, noise, t, u = _draw(s)
            one.set_input(data)
            gi = one.reduced_gradient(noise=noise, t=t, u=u)
            g = gi if g is None else g + gi
        one.flat.grad.copy_(g / 2)
        one._optimizer_step()
        assert rel_l2(acc.flat.data, one.flat.data) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# two ranks over gloo: the bucketed gradient exchange of the trainer (dp.GradBuckets started from inside the backward
# pass, dp.Comm on torch.distributed when the communicator is not NCCL)
# ---------------------------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    solo_groups = [dist.new_group([r]) for r in range(world)]   # (process_group=None would mean the default group)
    with KD.installed():
        tr = _trainer(_net(11 + rank), lr=1e-3, ema_beta=0.9, process_group=dist.group.WORLD, comm_min_bucket=1 << 12)
        tr.broadcast_parameters()            # rank 0's weights everywhere
        res = {"buckets": len(tr.buckets.buckets), "overlap": tr.overlap}
        data, noise, t, u = _draw(500 + rank)
        tr.set_input(data)
        g = tr.reduced_gradient(noise=noise, t=t, u=u) / world
        if rank == 0:
            solo = _trainer(_net(11), lr=1e-3, process_group=solo_groups[0])
            parts = [_draw(500 + r) for r in range(world)]
            big = {k: torch.cat([p[0][k] for p in parts]) for k in parts[0][0]}
            solo.set_input(big)
            gs = solo.reduced_gradient(noise=torch.cat([p[1] for p in parts]), t=torch.cat([p[2] for p in parts]),
                                       u=torch.cat([p[3] for p in parts]))
            res["grad_rel"] = rel_l2(g, gs)
        for s in range(3):
            _step(tr, 600 + 10 * s + rank)
        res.update(params=tr.flat.data.clone(), exp_avg=tr.exp_avg.clone(), ema=tr.ema.clone())
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_trainer_gradient_and_replicas_on_the_double():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a["buckets"] >= 2 and a["overlap"]
    # DDP's mean gradient of two ranks on their shards == the one-process gradient on the concatenated batch
    # (per-sample work is independent: GroupNorm / attention statistics are per image)
    assert a["grad_rel"] < 1e-4, a["grad_rel"]
    for k in ("params", "exp_avg", "ema"):
        assert torch.equal(a[k], b[k]), "replicas differ in %s" % k


def test_production_architecture_train_step_vs_oracle_on_the_double():
    """The config-2 ARCHITECTURE (ngf 64, mults 1-2-4-8, two ResBlocks per level, attention at ds 16: 59 M parameters,
    22 batched emb Linears, 75 managed convolutions, four concat levels) at 128^2, one PaletteTrainer step vs the oracle
    in fp32 and in bf16-storage emulation — tests/test_gpu_config2_full.py's comparison, sized for the CPU."""
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    cfg = O.UNetCfg(image_size=128)
    assert (cfg.inner_channel, tuple(cfg.channel_mults), tuple(cfg.res_blocks), tuple(cfg.attn_res)) == \
        (64, (1, 2, 4, 8), (2, 2, 2, 2), (16,))
    params = O.init_params(cfg, 2024)
    net = nets.build_palette_generator(image_size=128)
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all("gammas" in m or "posterior" in m for m in missing)
    data = O.synthetic_batch(1, 128, 31)
    torch.manual_seed(77)
    t, u = O.sample_t_gamma(cfg, 1)
    noise = torch.randn_like(data["gt"])
    with KD.installed():
        tr = _trainer(net, lr=1e-4, optim="adamw", ema=True, ema_beta=0.999)
        assert len(tr.packset.packs) == 75 and len(tr.wstage.slots) >= 70
        with torch.no_grad():
            _, nh, _ = tr.netG_A(data["gt"], data["cond"], data["mask"], noise, t=t, u=u)
        tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": data["mask"]})
        loss = tr.optimize_parameters(noise=noise, t=t, u=u)
        sd = {k: v.clone() for k, v in tr.netG_A.state_dict().items()}
    oc = O.OptimCfg(lr=1e-4, ema_beta=0.999)
    fp = O.TrainState(params={k: v.clone() for k, v in params.items()})
    fp_loss, fp_hat, _ = O.train_step(fp, cfg, oc, data["gt"], data["cond"], data["mask"], noise, t, u)
    emu = O.TrainState(params={k: v.clone() for k, v in params.items()})
    O.EMULATE_BF16[0] = True
    try:
        emu_loss, emu_hat, _ = O.train_step(emu, cfg, oc, data["gt"], data["cond"], data["mask"], noise, t, u)
    finally:
        O.EMULATE_BF16[0] = False
    assert abs(float(loss) - float(emu_loss)) < 2e-2 * abs(float(emu_loss)), (float(loss), float(emu_loss))
    assert abs(float(loss) - float(fp_loss)) < 2e-2 * abs(float(fp_loss)), (float(loss), float(fp_loss))
    floor = rel_l2(emu_hat, fp_hat)
    assert rel_l2(nh, fp_hat) < max(3e-2, 2.5 * floor), (rel_l2(nh, fp_hat), floor)
    num = den = 0.0
    for k in params:
        upd = sd[k].double() - params[k].double()
        upd_ref = emu.params[k].double() - params[k].double()
        num += float((upd - upd_ref).norm()) ** 2
        den += float(upd_ref.norm()) ** 2
    assert (num / den) ** 0.5 < 0.35


def _build_cfg45(which, gold, params):
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
def test_trainer_cfg4_cfg5_match_reference_plumbing_on_the_double(golden_dir, which):
    """PaletteTrainer with the reference-image UNet (`ref_A` through set_input) and with the video UNet (5-D clips folded
    to B*F frames, per-clip draws) vs the reference's own two steps (refattn_plumbing.pt / vid_plumbing.pt)."""
    from oracle.gen_golden_plumbing45 import batch, draws, oracle_cfg
    from oracle.vid_oracle import init_params_from_shapes
    gold = torch.load(os.path.join(golden_dir, "vid_plumbing.pt" if which == "vid" else "refattn_plumbing.pt"))
    p0 = init_params_from_shapes(gold["shapes"], gold["wseed"])
    net = _build_cfg45(which, gold, p0)
    assert [(k, tuple(v.shape)) for k, v in net.named_parameters()] == [(k, tuple(s)) for k, s in gold["shapes"]]
    oc = gold["optim"]
    cfg = oracle_cfg(which)
    cfg.n_timestep_train, cfg.n_timestep_test = gold["n_timestep_train"], gold["n_timestep_test"]
    with KD.installed():
        tr = _trainer(net, lr=oc["lr"], beta1=oc["beta1"], beta2=oc["beta2"], eps=oc["eps"],
                      weight_decay=oc["weight_decay"], optim=oc["kind"], ema=True, ema_beta=oc["ema_beta"],
                      iter_size=oc["iter_size"], lambda_G=gold["lambda_G"])
        for step in range(2):
            data = batch(which, gold["data_seeds"][step])
            t, u, noise = draws(which, cfg, gold["rng_seeds"][step])
            tr.set_input(data)
            loss = tr.optimize_parameters(noise=noise, t=t, u=u)
            assert abs(float(loss) - gold["losses"][step]) < 2e-2 * abs(gold["losses"][step]), step
        sd, ema = net.state_dict(), tr.ema_state_dict()
    for k, (_, n) in gold["param_stats"].items():
        assert abs(float(sd[k].double().norm()) - n) <= (1e-2 if sd[k].dim() > 1 else 2e-2) * n + 1e-6, k
    for k, (_, n) in gold["ema_stats"].items():
        assert abs(float(ema[k].double().norm()) - n) <= (1e-2 if ema[k].dim() > 1 else 2e-2) * n + 1e-6, k


def test_trainer_with_class_conditioning_and_dropout_on_the_double():
    """cond_embed "class": B_label_cls travels through set_input; conditioning dropout (palette_model.py:565-584)
    replaces the dropped samples' class (and mask) by num_classes - 1 — equal to the same step with the labels edited by
    hand (exactly: the double is deterministic)."""
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    from oracle.gen_golden_cond import BASE, cond_batch, cond_cfg, cond_params
    cfg = cond_cfg("class", 4)
    params = cond_params(cfg, 3)
    data = cond_batch(cfg, 4, 9)
    torch.manual_seed(1)
    t, u = O.sample_t_gamma(cfg, 4)
    noise = torch.randn_like(data["gt"])
    drop_u = torch.tensor([0.01, 0.9, 0.05, 0.5])
    losses = []
    with KD.installed():
        for mode in ("dropout", "by_hand"):
            net = nets.build_palette_generator(conditioning="class", nclasses=4, **BASE)
            net.load_state_dict(params, strict=False)
            cls, mask = data["cls"].clone(), data["mask"].clone()
            if mode == "by_hand":   # the reference fills BOTH the class and the mask of a dropped sample (:571-583)
                cls[drop_u < 0.1] = 3
                mask[drop_u < 0.1] = 3
            tr = _trainer(net, lr=1e-3, dropout_prob=0.1 if mode == "dropout" else 0.0, num_classes=4)
            tr.set_input({"A": data["cond"], "B": data["gt"], "B_label_mask": mask, "B_label_cls": cls})
            losses.append(float(tr.compute_palette_loss(noise=noise, t=t, u=u,
                                                        drop_u=drop_u if mode == "dropout" else None).detach()))
        assert losses[0] == losses[1]
        assert torch.isfinite(tr.optimize_parameters(noise=noise, t=t, u=u))
