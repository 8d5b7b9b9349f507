"""Drop-in swap: replace the hot-path modules of a REFERENCE-built network (joliGEN's own
`diffusion_networks.define_G` / `models.modules.*` objects) by their B200 mirrors, in place.

    netG_A = diffusion_networks.define_G(**vars(opt))      # joliGEN, unchanged
    netG_A = joligen_b200.accelerate(netG_A)               # this module

The swap is structural, not textual: reference classes are recognised by class name, rebuilt from
their own hyper-parameters, and the new module ADOPTS the reference's nn.Parameter / buffer objects
(no copy), so optimisers, DDP, `state_dict()`, `save_networks` / `load_networks`
(base_model.py:824-868, 957-1103) and `ema_step` (base_model.py:1284-1297) keep working on the same
tensors under the same names.  Modules that are not on the hot path are left untouched.
"""
import torch.nn as nn

from . import nets


def _adopt(dst: nn.Module, src: nn.Module):
    """Make dst's parameters/buffers BE src's (same objects), matched by qualified name."""
    src_params = dict(src.named_parameters())
    src_bufs = dict(src.named_buffers())
    dst_names = [n for n, _ in dst.named_parameters()]
    if sorted(dst_names) != sorted(src_params.keys()):
        missing = set(src_params) ^ set(dst_names)
        raise RuntimeError("accelerate: parameter names differ between reference and B200 module: %s"
                           % sorted(missing)[:6])
    for name in dst_names:
        mod, leaf = _resolve(dst, name)
        if mod._parameters[leaf].shape != src_params[name].shape:
            raise RuntimeError("accelerate: shape mismatch for %s" % name)
        mod._parameters[leaf] = src_params[name]
    _check_structure(dst, src)
    for name, buf in src_bufs.items():
        mod, leaf = _resolve(dst, name, create=True)
        mod._buffers[leaf] = buf
    nets.invalidate_packed_weights()


def _adopt_b2b(dst, src):
    """_adopt for the b2b backbone: the mirror derives the MotionModule's sinusoidal `pe` buffers itself (the reference
    registers them non-persistently or not at all depending on the version), every Parameter is the reference's."""
    src_params = dict(src.named_parameters())
    dst_names = [n for n, _ in dst.named_parameters()]
    if sorted(dst_names) != sorted(src_params):
        raise RuntimeError("accelerate: parameter names differ between the reference B2BGenerator and the B200 mirror: "
                           "%s" % sorted(set(src_params) ^ set(dst_names))[:6])
    for name in dst_names:
        mod, leaf = _resolve(dst, name)
        if mod._parameters[leaf].shape != src_params[name].shape:
            raise RuntimeError("accelerate: shape mismatch for %s" % name)
        mod._parameters[leaf] = src_params[name]
    dst_bufs = dict(dst.named_buffers())
    for name, buf in src.named_buffers():
        if name in dst_bufs:
            if dst_bufs[name].shape != buf.shape:
                raise RuntimeError("accelerate: buffer shape mismatch for %s" % name)
            mod, leaf = _resolve(dst, name)
            mod._buffers[leaf] = buf
    nets.invalidate_packed_weights()


def _resolve(root, name, create=False):
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        mod = getattr(mod, p)
    return mod, parts[-1]


def _norm_spec(ref_unet):
    """The `norm` argument that rebuilds the net's normalisation wrappers (unet_attn_utils.normalization :94-113), read
    off ALL of them: "instancenorm" is GroupNorm(C, C) — a group count that follows each block's width — "layernorm"
    GroupNorm(1, C), "groupnormN" a fixed count (which coincides with the width of the narrowest block for N = ngf)."""
    wrapped = [m.norm for m in ref_unet.modules() if type(m).__name__ == "GroupNorm" and hasattr(m, "norm")]
    if not wrapped:
        raise NotImplementedError("accelerate: no GroupNorm wrapper in the reference UNet (batchnorm / switchablenorm?)")
    counts = {g.num_groups for g in wrapped}
    if len(counts) == 1:
        n = counts.pop()
        return "layernorm" if n == 1 and any(g.num_channels > 1 for g in wrapped) else "groupnorm%d" % n
    if all(g.num_groups == g.num_channels for g in wrapped):
        return "instancenorm"
    raise NotImplementedError("accelerate: mixed GroupNorm group counts %s" % sorted(counts))


def _has_tanh(ref):
    """UNet(tanh=True) ends in (norm, conv, Tanh) instead of (norm, SiLU, conv) (unet_generator_attn.py:634-645)."""
    return any(isinstance(m, nn.Tanh) for m in ref.out)


def _check_structure(dst: nn.Module, src: nn.Module):
    """Beyond parameter names: the two trees must hold the same number of blocks of each hot-path kind and the same
    resampling / dropout configuration, otherwise an option this mirror does not implement would be dropped silently."""
    kinds = ("ResBlock", "AttentionBlock", "AttentionBlockRef", "MotionModule")
    is_kind = lambda m, k: any(c.__name__ == k for c in type(m).__mro__)  # noqa: E731  (VidResBlock is a ResBlock)
    count = lambda root: {k: sum(is_kind(m, k) for m in root.modules()) for k in kinds}  # noqa: E731
    if count(dst) != count(src):
        raise RuntimeError("accelerate: block structure differs: reference %s vs B200 %s" % (count(src), count(dst)))
    gn = lambda root: [(m.num_groups, m.num_channels, m.eps, m.affine) for m in root.modules()  # noqa: E731
                       if isinstance(m, nn.GroupNorm)]
    if gn(dst) != gn(src):
        raise RuntimeError("accelerate: the normalisation layers of the reference net (groups, width, eps, affine) are "
                           "not the ones the B200 mirror was built with")
    other = sorted({type(m).__name__ for m in src.modules()
                    if isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.LayerNorm)) or "SwitchNorm" in type(m).__name__}
                   - {type(m).__name__ for m in dst.modules()})
    if other:
        raise NotImplementedError("accelerate: normalisation %s is not on the B200 path" % other)
    drops = sorted({float(m.p) for m in src.modules() if isinstance(m, nn.Dropout) and m.p > 0})
    if drops:
        raise NotImplementedError("accelerate: the reference net uses dropout %s (not supported on the B200 path)" % drops)
    att = lambda root: [m for m in root.modules() if is_kind(m, "AttentionBlock") or is_kind(m, "AttentionBlockRef")]  # noqa: E731
    for a, b in zip(att(src), att(dst)):
        new_order = type(getattr(a, "attention", None)).__name__ == "QKVAttention"
        if a.num_heads != b.num_heads or a.channels != b.channels or int(new_order) != int(b.attention_layout):
            raise RuntimeError("accelerate: attention configuration (heads / width / q-k-v channel order) differs between "
                               "the reference block and its B200 mirror")
    for a, b in zip((m for m in src.modules() if is_kind(m, "ResBlock")),
                    (m for m in dst.modules() if is_kind(m, "ResBlock"))):
        if (bool(a.updown), bool(a.use_scale_shift_norm), bool(getattr(a, "efficient", False))) != \
                (bool(b.updown), bool(b.use_scale_shift_norm), bool(b.efficient)):
            raise RuntimeError("accelerate: ResBlock configuration differs between reference and B200 module")


def _unet_from_reference(ref):
    first_res = ref.input_blocks[1][0]
    norm = _norm_spec(ref)
    return nets.UNet(
        image_size=ref.image_size, in_channel=ref.in_channel, inner_channel=ref.inner_channel,
        out_channel=ref.out_channel, res_blocks=list(ref.res_blocks), attn_res=list(ref.attn_res),
        tanh=_has_tanh(ref), dropout=getattr(ref, "dropout", 0),
        n_timestep_train=ref.beta_schedule["train"]["n_timestep"],
        n_timestep_test=ref.beta_schedule["test"]["n_timestep"], norm=norm,
        group_norm_size=first_res.in_layers[0].norm.num_groups,
        cond_embed_dim=ref.cond_embed_dim, channel_mults=tuple(ref.channel_mults), num_heads=ref.num_heads,
        num_head_channels=ref.num_head_channels, num_heads_upsample=ref.num_heads_upsample,
        use_scale_shift_norm=first_res.use_scale_shift_norm, efficient=first_res.efficient,
        freq_space=getattr(ref, "freq_space", False))


def _common_unet_kwargs(ref):
    first_res = ref.input_blocks[1][0]
    return dict(
        image_size=ref.image_size, in_channel=ref.in_channel, inner_channel=ref.inner_channel,
        out_channel=ref.out_channel, res_blocks=list(ref.res_blocks), attn_res=list(ref.attn_res),
        tanh=_has_tanh(ref), dropout=getattr(ref, "dropout", 0),
        n_timestep_train=ref.beta_schedule["train"]["n_timestep"],
        n_timestep_test=ref.beta_schedule["test"]["n_timestep"], norm=_norm_spec(ref),
        group_norm_size=first_res.in_layers[0].norm.num_groups, cond_embed_dim=ref.cond_embed_dim,
        channel_mults=tuple(ref.channel_mults),
        use_scale_shift_norm=first_res.use_scale_shift_norm, efficient=first_res.efficient,
        freq_space=getattr(ref, "freq_space", False))


def _first_attention(ref):
    for m in ref.modules():
        if type(m).__name__ in ("AttentionBlock", "AttentionBlockRef"):
            return m
    return None


def _unetvid_from_reference(ref):
    from . import nets_vid
    att = _first_attention(ref)
    new_order = att is not None and type(att.attention).__name__ == "QKVAttention"
    return nets_vid.UNetVid(num_heads=ref.num_heads, num_head_channels=ref.num_head_channels,
                            num_heads_upsample=ref.num_heads_upsample, use_new_attention_order=new_order,
                            max_sequence_length=ref.max_sequence_length, cross_attention_dim=ref.cross_attention_dim,
                            num_attention_heads=ref.num_attention_heads,
                            num_transformer_blocks=ref.num_transformer_blocks, **_common_unet_kwargs(ref))


def _unetref_from_reference(ref):
    from . import nets_ref
    att = _first_attention(ref)
    new_order = att is not None and type(att.attention).__name__ == "QKVAttention"
    heads = att.num_heads if att is not None else 1
    head_ch = att.channels // heads if att is not None else -1
    # the reference class does not keep num_heads / num_head_channels: recover them from its attention blocks
    # (all blocks share num_head_channels when it was given; otherwise num_heads)
    per_block = {m.channels // m.num_heads for m in ref.modules() if type(m).__name__ == "AttentionBlockRef"}
    kw = dict(num_head_channels=head_ch) if len(per_block) == 1 else dict(num_heads=heads)
    return nets_ref.UNetGeneratorRefAttn(use_new_attention_order=new_order, **kw, **_common_unet_kwargs(ref))


def _layer_kinds(seq):
    """Class names of a Sequential's layers, ResnetBlocks expanded: the GAN mirrors must hold the SAME layer sequence as
    the reference net (a BatchNorm / Dropout / spectral-norm / separable-conv variant must not be swapped silently)."""
    out = []
    for m in seq:
        if type(m).__name__ == "ResnetBlock":
            out.append(["ResnetBlock"] + _layer_kinds(m.conv_block))
        else:
            out.append(type(m).__name__)
    return out


def _same_layers(dst_seq, src_seq, what):
    a, b = _layer_kinds(dst_seq), _layer_kinds(src_seq)
    if a != b:
        raise NotImplementedError("accelerate: %s layer sequence of the reference is not the one the B200 mirror "
                                  "implements (reference %s)" % (what, b[:6]))
    for d, r in zip(dst_seq.modules(), src_seq.modules()):
        if isinstance(r, (nn.Conv2d, nn.ConvTranspose2d)):
            if (r.kernel_size, r.stride, r.padding, r.dilation, r.groups) != \
                    (d.kernel_size, d.stride, d.padding, d.dilation, d.groups) or \
                    hasattr(r, "weight_orig") != hasattr(d, "weight_orig"):
                raise NotImplementedError("accelerate: %s convolution geometry / parametrisation differs" % what)
        if isinstance(r, nn.InstanceNorm2d) and (r.affine or r.track_running_stats):
            raise NotImplementedError("accelerate: InstanceNorm2d with affine / running statistics")
        if isinstance(r, nn.LeakyReLU) and abs(r.negative_slope - d.negative_slope) > 1e-12:
            raise NotImplementedError("accelerate: LeakyReLU slope differs")


def _resnet_generator_from_reference(ref):
    """models/modules/resnet_architecture/resnet_generator.py:98-164 (gan_networks.define_G, G_netG resnet_*blocks):
    the class keeps no hyper-parameters, they are read off its layers."""
    from . import nets_gan
    enc_convs = [m for m in ref.encoder.model if isinstance(m, nn.Conv2d)]
    dec_convs = [m for m in ref.decoder.model if isinstance(m, nn.Conv2d)]
    if not enc_convs or not dec_convs:
        raise NotImplementedError("accelerate: ResnetGenerator without plain nn.Conv2d stem / head (mobile variant)")
    n_blocks = sum(type(m).__name__ == "ResnetBlock" for m in ref.encoder.model)
    # --G_padding_type: the pad layer in front of the decoder's 7x7 convolution (none = "zeros")
    kinds = {type(m).__name__ for m in ref.decoder.model}
    padding_type = "reflect" if "ReflectionPad2d" in kinds else ("replicate" if "ReplicationPad2d" in kinds else "zeros")
    new = nets_gan.ResnetGenerator(enc_convs[0].in_channels, dec_convs[-1].out_channels, enc_convs[0].out_channels,
                                   n_blocks=n_blocks, padding_type=padding_type,
                                   use_spectral=hasattr(enc_convs[0], "weight_orig"))
    _same_layers(new.encoder.model, ref.encoder.model, "ResnetGenerator encoder")
    _same_layers(new.decoder.model, ref.decoder.model, "ResnetGenerator decoder")
    return new


def _nlayer_discriminator_from_reference(ref):
    """models/modules/discriminators.py:10-117 (gan_networks.define_D, D_netDs basic / n_layers)."""
    from . import nets_gan
    freq = bool(getattr(ref, "freq_space", False))
    convs = [m for m in ref.model if isinstance(m, nn.Conv2d)]
    if len(convs) < 3:
        raise NotImplementedError("accelerate: NLayerDiscriminator with %d convolutions" % len(convs))
    new = nets_gan.NLayerDiscriminator(convs[0].in_channels // (4 if freq else 1), convs[0].out_channels,
                                       n_layers=len(convs) - 2, freq_space=freq,
                                       use_spectral=all(hasattr(c, "weight_orig") for c in convs))
    _same_layers(new.model, ref.model, "NLayerDiscriminator")
    return new


def _b2b_generator_from_reference(ref):
    """models/modules/b2b_generator.py B2BGenerator around models/modules/vit/vit_vid.py JiTViD (`model_type b2b`,
    `G_netG vit_vid`): hyper-parameters from the attributes the reference keeps and from its layers."""
    from . import nets_jit
    m = ref.b2b_model
    if type(m).__name__ != "JiTViD":
        raise NotImplementedError("accelerate: B2BGenerator backbone %s is not on the B200 path" % type(m).__name__)
    unsupported = [k for k in ("mask_size_conditioning", "temporal_frame_step_conditioning", "global_context_conditioning")
                   if getattr(m, k, False)]
    if getattr(m, "num_register_tokens", 0) or getattr(m, "object_ref_num_images", 0) or getattr(m, "motion_every", 0):
        unsupported.append("register tokens / object references / per-layer motion modules")
    if unsupported:
        raise NotImplementedError("accelerate: JiTViD options not on the B200 path: %s" % unsupported)
    hidden = m.hidden_size
    ffn = m.blocks[0].mlp.w12.out_features // 2                    # = int(int(hidden * mlp_ratio) * 2 / 3)
    wide = [h for h in range(int(ffn * 1.5) - 2, int(ffn * 1.5) + 4) if int(h * 2 / 3) == ffn]
    if not wide:
        raise NotImplementedError("accelerate: cannot recover JiTViD's mlp_ratio from its SwiGLU width %d" % ffn)
    mlp_ratio = 4.0 if int(int(hidden * 4.0) * 2 / 3) == ffn else wide[0] / hidden
    tblocks = m.motion_module.temporal_transformer.transformer_blocks
    net = nets_jit.JiTViD(input_size=m.input_size, patch_size=m.patch_size, in_channels=m.in_channels,
                          out_channels=m.out_channels, hidden_size=hidden, depth=len(m.blocks), num_heads=m.num_heads,
                          mlp_ratio=mlp_ratio, num_classes=m.y_embedder.embedding_table.num_embeddings - 1,
                          bottleneck_dim=m.x_embedder.proj1.out_channels, in_context_len=m.in_context_len,
                          in_context_start=m.in_context_start, max_frames=m.max_frames,
                          motion_num_heads=tblocks[0].attention_blocks[0].heads, motion_num_layers=len(tblocks))
    if abs(float(getattr(ref, "cfg_scale", 1.0)) - 1.0) > 1e-12:
        raise NotImplementedError("accelerate: B2BGenerator with classifier-free guidance (alg_b2b_cfg_scale != 1)")
    gen = nets_jit.B2BGenerator(net, t_eps=ref.t_eps, noise_scale=ref.noise_scale, P_mean=ref.P_mean, P_std=ref.P_std,
                                timestep_uniform_mix_prob=getattr(ref, "timestep_uniform_mix_prob", 0.0),
                                label_drop_prob=getattr(ref, "label_drop_prob", 0.0),
                                num_classes=getattr(ref, "num_classes", 1),
                                denoise_timesteps=getattr(ref, "denoise_timesteps", 50))
    gen.clip_denoised_default = bool(getattr(ref, "clip_denoised_default", False))
    gen.disable_inference_clipping = bool(getattr(ref, "disable_inference_clipping", False))
    return gen


def accelerate(module: nn.Module) -> nn.Module:
    """Returns the accelerated module (the same object with children swapped, or a new root when the
    root itself is a hot-path class)."""
    cls = type(module).__name__
    if isinstance(module, (nets.UNet, nets.DiffusionGenerator, nets.ResBlock, nets.AttentionBlock)) or \
            type(module).__module__.startswith("joligen_b200."):
        return module
    if cls == "DiffusionGenerator":
        ref_unet = module.denoise_fn.model
        builders = {"UNet": _unet_from_reference, "UNetVid": _unetvid_from_reference,
                    "UNetGeneratorRefAttn": _unetref_from_reference}
        if type(ref_unet).__name__ not in builders:
            raise NotImplementedError("accelerate: DiffusionGenerator backbone %s is not on the B200 path yet"
                                      % type(ref_unet).__name__)
        cond = getattr(module.denoise_fn, "conditioning", "")
        nclasses = 2
        for name in ("netl_embedder_class", "netl_embedder_mask"):
            if hasattr(module.denoise_fn, name):
                nclasses = getattr(module.denoise_fn, name).num_classes
        unet = builders[type(ref_unet).__name__](ref_unet)
        dn = nets.PaletteDenoiseFn(model=unet, cond_embed_dim=module.denoise_fn.cond_embed_dim, conditioning=cond,
                                   nclasses=nclasses)  # ("ref" conditioning raises: frozen third-party backbone)
        new = nets.DiffusionGenerator(denoise_fn=dn, sampling_method=module.sampling_method,
                                      image_size=module.image_size)
        _adopt(new, module)
        new.train(module.training)
        return new
    if cls == "UNet" and hasattr(module, "input_blocks") and hasattr(module, "middle_block"):
        new = _unet_from_reference(module)
        _adopt(new, module)
        new.train(module.training)
        return new
    if cls == "UNetVid" and hasattr(module, "input_blocks"):
        new = _unetvid_from_reference(module)
        _adopt(new, module)
        new.train(module.training)
        return new
    if cls == "UNetGeneratorRefAttn" and hasattr(module, "input_blocks_ref"):
        new = _unetref_from_reference(module)
        _adopt(new, module)
        new.train(module.training)
        return new
    if cls == "B2BGenerator" and hasattr(module, "b2b_model"):
        new = _b2b_generator_from_reference(module)
        _adopt_b2b(new, module)
        new.train(module.training)
        return new
    if cls == "MultiScaleD" and isinstance(getattr(module, "mini_discs", None), nn.ModuleDict):
        # the trainable heads of the projected discriminator (models/modules/projected_d/discriminator.py:166-230); the
        # frozen timm feature network in front of them is third-party and stays the reference's
        from . import nets_projd
        kinds = {type(d).__name__ for d in module.mini_discs.values()}
        if kinds != {"SingleDisc"}:
            raise NotImplementedError("accelerate: MultiScaleD mini-discriminators %s (SingleDisc is implemented)"
                                      % sorted(kinds))
        new = nets_projd.MultiScaleD(channels=list(module.disc_in_channels), resolutions=list(module.disc_in_res),
                                     conv=True, feats=None, num_discs=len(module.mini_discs))
        _adopt(new, module)
        new.train(module.training)
        return new
    if cls == "ResnetGenerator" and hasattr(module, "encoder") and hasattr(module, "decoder"):
        new = _resnet_generator_from_reference(module)
        _adopt(new, module)
        new.train(module.training)
        return new
    if cls == "NLayerDiscriminator" and isinstance(getattr(module, "model", None), nn.Sequential):
        new = _nlayer_discriminator_from_reference(module)
        _adopt(new, module)
        new.train(module.training)
        return new
    for name, child in list(module.named_children()):
        swapped = accelerate(child)
        if swapped is not child:
            setattr(module, name, swapped)
    return module
