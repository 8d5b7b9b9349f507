"""CUT's contrastive modules on the B200 kernels: mirrors of

    PatchSampleF        /root/reference/models/modules/cut_networks.py:6-73   (netF "mlp_sample")
    PatchNCELoss        /root/reference/models/modules/NCE/base_NCE.py, patchnce.py

same constructor arguments, sub-module names (mlp_<i>.0 / .2) and call signatures, so that
`cut_model.calculate_feats` / `calculate_NCE_loss` (models/cut_model.py:848-909) run unchanged.

STATUS: written at the end of round 1; one run on a B200 (profiles/r01_cut_tests_first_run.log): the kernels match
the oracle (1e-4) and the pooled features / total NCE loss match the reference's golden vectors; the gradient checks
of the end-to-end test are re-bounded but not re-run yet (tests/test_gpu_widen_cut.py, `unverified` marker).  MoNCELoss
(the example's default --alg_cut_nce_loss) is compiled but has not run on hardware.
"""
import torch
import torch.nn as nn

from . import lib as L
from . import ops
from .nets import ConvPack


def _token_grid(rows):
    """[rows, C] tokens as an NHWC 'image' [1, rows / w, w, C] with the widest power-of-two w <= 16 dividing rows: the
    1x1 convolutions then see the tile shapes the UNet layers use."""
    w = 16
    while w > 1 and rows % w:
        w //= 2
    return rows // w, w


class PatchSampleF(nn.Module):
    def __init__(self, use_mlp=False, init_type="normal", init_gain=0.02, nc=256):
        super().__init__()
        self.use_mlp = use_mlp
        self.nc = nc
        self.mlp_init = False
        self.init_type = init_type
        self.init_gain = init_gain
        self.device = None
        self._packs = {}

    def set_device(self, device):
        self.device = device

    def data_dependent_initialize(self, feats):
        if self.use_mlp and not self.mlp_init:
            self.create_mlp(feats)

    def create_mlp(self, feats):
        """feats: list of NCHW tensors (only the channel counts are used).  init_net(normal, 0.02) of the reference:
        Linear weights ~ N(0, init_gain), biases 0 (models/modules/utils.py init_weights)."""
        if self.init_type != "normal":
            raise NotImplementedError("B200 PatchSampleF: init_type %r" % self.init_type)
        for mlp_id, feat in enumerate(feats):
            input_nc = feat.shape[1]
            mlp = nn.Sequential(nn.Linear(input_nc, self.nc), nn.ReLU(), nn.Linear(self.nc, self.nc))
            for m in mlp:
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, 0.0, self.init_gain)
                    nn.init.constant_(m.bias, 0.0)
            setattr(self, "mlp_%d" % mlp_id, mlp.to(self.device) if self.device is not None else mlp)
        self.mlp_init = True

    def _pack(self, lin):
        p = self._packs.get(id(lin))
        if p is None:
            p = self._packs[id(lin)] = ConvPack(lin)
        return p

    def _mlp(self, mlp, x):
        """x bf16 [rows, C8] -> bf16 [rows, nc]: Linear -> ReLU -> Linear as two 1x1 tcgen05 convolutions."""
        rows = x.shape[0]
        h, w = _token_grid(rows)
        t = x.reshape(1, h, w, x.shape[1])
        l0, l2 = mlp[0], mlp[2]
        w0 = l0.weight.unsqueeze(-1).unsqueeze(-1)  # [O, I] viewed as a 1x1 filter (the view keeps autograd's link)
        w2 = l2.weight.unsqueeze(-1).unsqueeze(-1)
        t = ops.conv_act(t, w0, l0.bias, self._pack(l0).get(), stride=1, pad=0, act=L.ACT_RELU)
        t = ops.conv_act(t, w2, l2.bias, self._pack(l2).get(), stride=1, pad=0, act=L.ACT_NONE)
        return t.reshape(rows, t.shape[-1])

    def forward_nhwc(self, feats, num_patches=64, patch_ids=None):
        """feats: list of NHWC bf16 maps (channels padded to 8).  -> ([fp32 [B*P, nc]], [ids [1, P]])."""
        if num_patches <= 0:
            raise NotImplementedError("B200 PatchSampleF: num_patches == 0 (whole maps) is not used by cut_model")
        return_ids, return_feats = [], []
        for feat_id, feat in enumerate(feats):
            hw = feat.shape[1] * feat.shape[2]
            if patch_ids is not None:
                patch_id = patch_ids[feat_id].reshape(-1)
            else:
                k = int(min(num_patches, hw))
                if torch.cuda.is_current_stream_capturing():
                    # (graph-safe draw of k distinct positions: randperm's CUDA path may synchronise)
                    patch_id = torch.rand(hw, device=feat.device).argsort()[:k]
                else:
                    patch_id = torch.randperm(hw, device=feat.device)[:k]
            x = ops.gather_rows(feat, patch_id.contiguous())
            if self.use_mlp:
                x = self._mlp(getattr(self, "mlp_%d" % feat_id), x)
            return_ids.append(patch_id.unsqueeze(0))
            return_feats.append(ops.l2_normalize(x[:, : self._width(feat_id, feat)], eps=1e-7))
        return return_feats, return_ids

    def _width(self, feat_id, feat):
        return self.nc if self.use_mlp else feat.shape[-1]

    def forward(self, feats, num_patches=64, patch_ids=None):
        """The reference signature: feats = list of NCHW fp32 maps."""
        return self.forward_nhwc([ops.to_nhwc(f) for f in feats], num_patches, patch_ids)


class PatchNCELoss(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt

    def forward(self, feat_q, feat_k, current_batch, **unused_args):
        if unused_args.get("weight") is not None:
            raise NotImplementedError("B200 PatchNCELoss: per-patch weights (SRC loss)")
        groups = 1 if self.opt.alg_cut_nce_includes_all_negatives_from_minibatch else current_batch
        d = feat_q.shape[1]
        if d % 32 or d > 512:
            raise NotImplementedError("B200 PatchNCELoss: feature width %d (multiples of 32 up to 512)" % d)
        return ops.patch_nce(feat_q, feat_k, groups, self.opt.alg_cut_nce_T)


class MoNCELoss(PatchNCELoss):
    """models/modules/NCE/monce.py: the negatives are re-weighted by Sinkhorn optimal-transport weights (csrc/nce.cu:
    monce_fwd / monce_bwd kernels — compiled, NOT yet run on hardware; checked against autograd on the CPU only)."""

    def forward(self, feat_q, feat_k, current_batch, **unused_args):
        if unused_args.get("weight") is not None:
            raise NotImplementedError("B200 MoNCELoss: per-patch weights (SRC loss)")
        if self.opt.alg_cut_nce_includes_all_negatives_from_minibatch:
            raise NotImplementedError("B200 MoNCELoss: negatives from the whole minibatch (one CTA per group, P <= 1024)")
        d = feat_q.shape[1]
        if d % 32 or d > 512:
            raise NotImplementedError("B200 MoNCELoss: feature width %d (multiples of 32 up to 512)" % d)
        return ops.monce(feat_q, feat_k, current_batch, self.opt.alg_cut_nce_T, self.opt.alg_cut_num_patches)
