"""CPU oracle for the projected discriminator's trainable part (TEST INFRASTRUCTURE — see oracle/__init__.py):
SURVEY.md section 8(f) rank 3, `D_netDs ["projected_d"]` of BASELINE.json config 3.

Functional fp32 restatement over {state_dict key: tensor} of

  MultiScaleD / SingleDisc   /root/reference/models/modules/projected_d/discriminator.py:13-77, 166-230
  DownBlock, conv2d = spectral_norm(nn.Conv2d), NormLayer = GroupNorm(c // 2, c)
                             /root/reference/models/modules/projected_d/blocks.py:11-31, 182-200
  torch.nn.utils.spectral_norm (training mode: ONE power iteration per forward on the persistent u / v vectors, then
      W = W_orig / sigma with sigma = u^T W_mat v; u and v are treated as constants by autograd)

The frozen feature network in front of it (timm tf_efficientnet_lite0 + the CCM / CSM projections, projector.py) is a
third-party pretrained backbone without weights in this container: out of scope; the mini-discriminators are fed
feature maps directly, as SURVEY.md section 8(c) prescribes.  Pinned by oracle/gen_golden_projd.py.
"""
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .palette_oracle import _r, _rw

CHANNEL_DICT = {4: 512, 8: 512, 16: 256, 32: 128, 64: 64, 128: 64, 256: 32, 512: 16, 1024: 8}


def disc_plan(nc: int, start_sz: int, end_sz: int = 8) -> List[Tuple[int, int]]:
    """(cin, cout) of the DownBlocks of SingleDisc(nc=nc, start_sz=start_sz, head=None)."""
    nfc = dict(CHANNEL_DICT)
    nfc[start_sz] = nc
    plan = []
    while start_sz > end_sz:
        plan.append((nfc[start_sz], nfc[start_sz // 2]))
        start_sz //= 2
    return plan


def spectral_normalize(sd, name, training=True, eps=1e-12):
    """-> (W, u_new, v_new).  torch.nn.utils.spectral_norm.SpectralNorm.compute_weight with n_power_iterations = 1."""
    w = sd[name + ".weight_orig"]
    u, v = sd[name + ".weight_u"], sd[name + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=eps)
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma, u, v


def single_disc(sd, name, x, nc, start_sz, training=True, new_state=None):
    """SingleDisc.forward: [DownBlock = SN-conv4x4 s2 p1 (+bias) -> GroupNorm(c/2) -> LeakyReLU 0.2]* -> SN-conv4x4 s1 p0
    (no bias, one output channel).  new_state (dict) receives the updated power-iteration vectors."""
    plan = disc_plan(nc, start_sz)
    for i, (cin, cout) in enumerate(plan):
        n = "%s.main.%d.main" % (name, i)
        w, u, v = spectral_normalize(sd, n + ".0", training)
        if new_state is not None:
            new_state[n + ".0.weight_u"], new_state[n + ".0.weight_v"] = u, v
        # with palette_oracle.EMULATE_BF16 set, tensors the CUDA path stores as bf16 are rounded at the same places
        x = _r(F.conv2d(x, _rw(w), sd[n + ".0.bias"], stride=2, padding=1))
        x = _r(F.leaky_relu(F.group_norm(x, cout // 2, sd[n + ".1.weight"], sd[n + ".1.bias"], eps=1e-5), 0.2))
    n = "%s.main.%d" % (name, len(plan))
    w, u, v = spectral_normalize(sd, n, training)
    if new_state is not None:
        new_state[n + ".weight_u"], new_state[n + ".weight_v"] = u, v
    return _r(F.conv2d(x, _rw(w), None, stride=1, padding=0))


def multi_scale_d(sd, feats: Dict[str, torch.Tensor], channels, resolutions, training=True, new_state=None,
                  prefix=""):
    """MultiScaleD.forward (conv mini-discriminators): logits of every scale flattened and concatenated."""
    outs = []
    for i, (c, r) in enumerate(zip(channels, resolutions)):
        x = _r(feats[str(i)])
        outs.append(single_disc(sd, "%smini_discs.%d" % (prefix, i), x, c, r, training, new_state)
                    .reshape(x.shape[0], -1))
    return torch.cat(outs, dim=1)
