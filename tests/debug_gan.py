"""Debug aid (test infrastructure: it lives under tests/ because it calls the oracle): per-parameter gradient errors
of the B200 ResnetGenerator vs the reference golden.        python tests/debug_gan.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import nets_gan  # noqa: E402
from oracle import gan_oracle as G  # noqa: E402

gold = torch.load(os.path.join(ROOT, "tests", "golden", "gan_resnet.pt"))
net = nets_gan.ResnetGenerator(3, 3, gold["ngf"], n_blocks=gold["n_blocks"])
shapes = G.resnet_param_shapes(3, 3, gold["ngf"], gold["n_blocks"])
net.load_state_dict(G.init_from_shapes(shapes, gold["wseed"]))
net = net.cuda()
y = net(gold["x"].cuda())
print("y rel", float((y.cpu() - gold["y"]).norm() / gold["y"].norm()))
y.backward(gold["dy"].cuda())
for k, p in net.named_parameters():
    gref = gold["grads"][k].double()
    e = float((p.grad.cpu().double() - gref).norm())
    print("%-40s shape %-20s ref_norm %10.3f rel_err %.4f" % (k, tuple(p.shape), float(gref.norm()), e / float(gref.norm())))
# emulated oracle for comparison
from oracle import palette_oracle as O
O.EMULATE_BF16[0] = True
leaves = {k: v.requires_grad_(True) for k, v in G.init_from_shapes(shapes, gold["wseed"]).items()}
ye = G.resnet_generator(leaves, gold["x"], gold["n_blocks"])
ye.backward(gold["dy"])
O.EMULATE_BF16[0] = False
print("emulated oracle vs fp32 golden:")
for k in list(leaves)[:4]:
    gref = gold["grads"][k].double()
    print("%-40s rel_err %.4f" % (k, float((leaves[k].grad.double() - gref).norm() / gref.norm())))
