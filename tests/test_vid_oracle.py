"""Video UNet, CPU side: the oracle restatement against the golden vectors
of the unmodified reference `UNetVid` (oracle/gen_golden_vid.py), and the B200 module tree against the reference's
parameter list (names and shapes: checkpoints interchange)."""
import os

import torch

from oracle import vid_oracle as V


def _load(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "vid_small.pt"))
    cfg = V.VidCfg(**gold["cfg"])
    params = V.init_params_from_shapes(gold["shapes"], gold["wseed"])
    return gold, cfg, params


def test_vid_oracle_matches_reference_forward_backward(golden_dir):
    from oracle.gen_golden_vid import inputs
    gold, cfg, params = _load(golden_dir)
    x, emb, gy = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y = V.unet_vid_forward(V.add_buffers(leaves, cfg), x, emb, cfg)
    assert float((y - gold["y"]).abs().max()) < 1e-4 * float(gold["y"].abs().max())
    loss = (y * gy).sum()
    assert abs(float(loss) - gold["loss"]) < 1e-4 * max(1.0, abs(gold["loss"]))
    loss.backward()
    scale = max(g["l2"] for g in gold["grads"].values())
    for k, g in gold["grads"].items():
        mine = leaves[k].grad
        assert abs(float(mine.double().norm()) - g["l2"]) < 1e-4 * max(g["l2"], 1e-3 * scale), k
        assert float((mine.flatten()[:16] - g["head"]).abs().max()) < 1e-4 * max(g["l2"], 1e-3 * scale), k
        if g["full"] is not None:
            assert float((mine - g["full"]).norm()) < 1e-4 * max(g["l2"], 1e-3 * scale), k


def test_b200_unetvid_has_reference_parameter_list(golden_dir):
    from joligen_b200 import nets_vid
    gold, cfg, _ = _load(golden_dir)
    net = nets_vid.UNetVid(image_size=cfg.image_size, in_channel=cfg.in_channel, inner_channel=cfg.inner_channel,
                           out_channel=cfg.out_channel, res_blocks=list(cfg.res_blocks), attn_res=list(cfg.attn_res),
                           tanh=False, n_timestep_train=cfg.n_timestep_train, n_timestep_test=cfg.n_timestep_test,
                           norm="groupnorm", group_norm_size=cfg.group_norm_size, cond_embed_dim=cfg.cond_embed_dim,
                           channel_mults=cfg.channel_mults, num_heads=cfg.num_heads,
                           num_head_channels=cfg.num_head_channels, max_sequence_length=cfg.max_sequence_length,
                           num_attention_heads=cfg.num_attention_heads,
                           num_transformer_blocks=cfg.num_transformer_blocks)
    mine = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    assert mine == [(k, tuple(s)) for k, s in gold["shapes"]]
    # buffers: the positional-encoding tables, equal to the reference formula
    sd = net.state_dict()
    ref = V.add_buffers({k: torch.zeros(s) for k, s in gold["shapes"]}, cfg)
    pes = [k for k in sd if k.endswith("pos_encoder.pe")]
    assert pes and all(torch.equal(sd[k], ref[k]) for k in pes)
    # zero-initialised proj_out, like zero_module() in the reference MotionModule
    assert all(float(p.abs().sum()) == 0.0 for k, p in net.named_parameters() if ".temporal_transformer.proj_out." in k)


def test_vid_generator_oracle_matches_reference(golden_dir):
    """cfg 5 end to end on CPU: DiffusionGenerator(PaletteDenoiseFn(UNetVid)) + Palette loss + backward."""
    from oracle.gen_golden_vid import generator_inputs
    gold = torch.load(os.path.join(golden_dir, "vid_generator.pt"))
    cfg = V.VidCfg(**gold["cfg"])
    params = V.init_params_from_shapes(gold["shapes"], gold["wseed"])
    gt, cond, mask, noise = generator_inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    _, nh = V.diffusion_forward_vid(V.add_buffers(leaves, cfg), gt, cond, mask, noise, gold["t"], gold["u"], cfg)
    assert float((nh - gold["noise_hat"]).abs().max()) < 1e-4 * float(gold["noise_hat"].abs().max())
    mb = torch.clamp(mask, min=0, max=1)
    loss = torch.nn.MSELoss()(mb * noise, mb * nh)
    assert abs(float(loss) - gold["loss"]) < 1e-5 * gold["loss"]
    loss.backward()
    scale = max(g["l2"] for g in gold["grads"].values())
    for k, g in gold["grads"].items():
        assert abs(float(leaves[k].grad.double().norm()) - g["l2"]) < 1e-4 * max(g["l2"], 1e-3 * scale), k
