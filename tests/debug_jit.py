"""Debug aid (test infrastructure: it lives under tests/ because it calls the oracle).  Stage-by-stage comparison of
nets_jit.JiTViD with the oracle on the jit_b200 golden inputs.        python tests/debug_jit.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import nets_jit  # noqa: E402
from oracle import jit_oracle as J  # noqa: E402
from oracle.gen_golden_jit import inputs  # noqa: E402
from oracle.vid_oracle import init_params_from_shapes  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


gold = torch.load(os.path.join(ROOT, "tests", "golden", "jit_b200.pt"))
cfg = J.JitCfg(**gold["cfg"])
params = init_params_from_shapes(gold["shapes"], gold["wseed"])
sd = J.add_buffers({**params, **gold["frozen"]}, cfg)
P = "b2b_model."
m = nets_jit.JiTViD(input_size=cfg.input_size, patch_size=cfg.patch_size, hidden_size=cfg.hidden_size, depth=cfg.depth,
                    num_heads=cfg.num_heads, in_context_len=cfg.in_context_len, in_context_start=cfg.in_context_start,
                    motion_num_heads=cfg.motion_num_heads, motion_num_layers=cfg.motion_num_layers)
net = nets_jit.B2BGenerator(m)
net.load_state_dict({**params, **gold["frozen"]}, strict=False)
net = net.cuda()
gt, cond, mask, label = inputs(cfg, gold["batch"], gold["frames"], gold["dseed"])
b, f = gt.shape[:2]
n = b * f
x5 = torch.cat([cond, gt], dim=2)
t = torch.linspace(0.2, 0.8, n)
with torch.no_grad():
    # patch embed
    x = x5.reshape(n, 6, cfg.input_size, cfg.input_size)
    r = F.conv2d(x, sd[P + "x_embedder.proj1.weight"], None, stride=cfg.patch_size)
    r1 = r.flatten(2).transpose(1, 2)
    r = F.conv2d(r, sd[P + "x_embedder.proj2.weight"], sd[P + "x_embedder.proj2.bias"])
    r = r.flatten(2).transpose(1, 2)
    tok = m.x_embedder.forward_tokens(x.cuda())
    print("patch embed", rel(tok[:, :, 0], r), tuple(tok.shape))
    r = r + sd[P + "pos_embed"]
    tok = (tok.float() + m.pos_embed[:, :, None, :]).to(torch.bfloat16)
    y = label.repeat_interleave(f)
    t_emb = J._lin(sd, P + "t_embedder.mlp.2", F.silu(J._lin(sd, P + "t_embedder.mlp.0", J.timestep_embedding(t))))
    y_emb = sd[P + "y_embedder.embedding_table.weight"][y]
    cvec = t_emb + y_emb
    cv = m.t_embedder(t.cuda()) + m.y_embedder(y.cuda())
    print("cvec", rel(cv, cvec))
    cos0, sin0 = J.rope_tables(cfg, 0)
    cos1, sin1 = J.rope_tables(cfg, cfg.in_context_len)
    for i in range(cfg.depth):
        if i == cfg.in_context_start:
            ctx = y_emb.unsqueeze(1).repeat(1, cfg.in_context_len, 1) + sd[P + "in_context_posemb"]
            r = torch.cat([ctx, r], dim=1)
            ctxd = cv.new_tensor(0)  # placeholder
            cd = m.y_embedder(y.cuda()).unsqueeze(1).repeat(1, cfg.in_context_len, 1) + m.in_context_posemb
            tok = torch.cat([cd[:, :, None, :].to(torch.bfloat16), tok], dim=1).contiguous()
        cos, sin = (cos0, sin0) if i < cfg.in_context_start else (cos1, sin1)
        r_in, tok_in = r, tok
        r = J.jit_block(sd, P + "blocks.%d" % i, r, cvec, cos, sin, cfg.num_heads)
        prefix = cfg.in_context_len if i >= cfg.in_context_start else 0
        cd_, sd_ = m._rope_for(prefix, tok.device)
        tok = m.blocks[i].forward_tokens(tok, cv, cd_, sd_)
        print("block %d" % i, rel(tok[:, :, 0], r), "(input err %.2e)" % rel(tok_in[:, :, 0], r_in))
        # the same block on the ORACLE's input (isolates the block)
        iso = m.blocks[i].forward_tokens(r_in.to(torch.bfloat16).cuda()[:, :, None, :].contiguous(), cvec.cuda(), cd_, sd_)
        print("   isolated", rel(iso[:, :, 0], r))
    r = r[:, cfg.in_context_len:]
    tok = tok[:, cfg.in_context_len:].contiguous()
    hp = cfg.input_size // cfg.patch_size
    d = r.shape[-1]
    from types import SimpleNamespace
    from oracle import vid_oracle as V
    grid = r.reshape(n, hp, hp, d).permute(0, 3, 1, 2)
    mcfg = SimpleNamespace(num_transformer_blocks=cfg.motion_num_layers, num_attention_heads=cfg.motion_num_heads)
    g2 = V.motion_module(sd, P + "motion_module", grid, f, mcfg)
    r2 = g2.permute(0, 2, 3, 1).reshape(n, hp * hp, d)
    m._clip["frames"] = f
    iso = m.motion_module.forward_nhwc(r.to(torch.bfloat16).cuda().reshape(n, hp, hp, d).contiguous())
    print("motion isolated", rel(iso.reshape(n, hp * hp, d), r2))
    shift, scale = J._lin(sd, P + "final_layer.adaLN_modulation.1", F.silu(cvec)).chunk(2, dim=1)
    r3 = J._lin(sd, P + "final_layer.linear", J.modulate(J.rms_norm(r2, sd[P + "final_layer.norm_final.weight"]), shift, scale))
    iso = m.final_layer.forward_tokens(r2.to(torch.bfloat16).cuda()[:, :, None, :].contiguous(), cvec.cuda())
    print("final isolated", rel(iso[:, :, 0], r3))
    full = m(x5.cuda(), t.cuda(), label.cuda())
    ref = J.jit_vid_forward(sd, x5, t, label, cfg, prefix=P)
    print("full forward", rel(full, ref))
