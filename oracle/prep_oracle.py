"""CPU restatement of the reference's input preparation and Haar wavelet ops (TEST INFRASTRUCTURE — see
oracle/__init__.py; only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this).

Pinned by tests/golden/prep_small.pt, generated from the UNMODIFIED reference functions by oracle/gen_golden_prep.py:
  fill_mask_with_random       /root/reference/data/online_creation.py:1366-1376
  upfirdn2d_native            /root/reference/models/modules/op/upfirdn2d.py:167-208
  HaarTransform / Inverse...  /root/reference/models/modules/freq_utils.py:9-59
  conditioning dropout        /root/reference/models/palette_model.py:565-584
"""
import torch
import torch.nn.functional as F


def fill_mask_with_random(img, mask, cls, noise):
    """online_creation.py:1366-1376 with the noise draw passed in (the reference calls torch.randn_like)."""
    if cls == -1:
        m = torch.where(mask != 0, 1.0, 0.0)
    else:
        m = torch.where(mask == cls, 1.0, 0.0)
    return img * (1 - m) + noise * m


def haar_kernels():
    """freq_utils.get_haar_wavelet (:9-19): ll, lh, hl, hh as 2x2 tensors."""
    s = 1 / (2 ** 0.5)
    low = s * torch.ones(1, 2)
    high = s * torch.ones(1, 2)
    high[0, 0] = -high[0, 0]
    return low.T * low, high.T * low, low.T * high, high.T * high


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0, 0, 0)):
    """upfirdn2d.py:167-208 for per-channel filtering of [N,C,H,W]: zero-insert upsample, pad (x0, x1, y0, y1),
    correlate with the FLIPPED kernel, keep every `down`-th sample."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    t = x.reshape(n * c, 1, h, w)
    if up > 1:
        z = torch.zeros(n * c, 1, h * up, w * up, dtype=x.dtype)
        z[:, :, ::up, ::up] = t
        t = z
    t = F.pad(t, [pad[0], pad[1], pad[2], pad[3]])
    t = F.conv2d(t, torch.flip(kernel, [0, 1]).view(1, 1, kh, kw).to(x.dtype))
    t = t[:, :, ::down, ::down]
    return t.reshape(n, c, t.shape[2], t.shape[3])


def haar_dwt(x):
    """HaarTransform.forward (freq_utils.py:34-40)."""
    return torch.cat([upfirdn2d(x, k, down=2) for k in haar_kernels()], dim=1)


def haar_iwt(y):
    """InverseHaarTransform.forward (:54-59): kernels (ll, -lh, -hl, hh), up=2, pad=(1,0,1,0), summed in that order."""
    ll, lh, hl, hh = haar_kernels()
    parts = y.chunk(4, 1)
    outs = [upfirdn2d(p, k, up=2, pad=(1, 0, 1, 0)) for p, k in zip(parts, (ll, -lh, -hl, hh))]
    return outs[0] + outs[1] + outs[2] + outs[3]


def mask_class_dropout(mask, drop_u, prob, num_classes):
    """palette_model.py:565-579: the highest class is the unconditioned one."""
    drop_ids = drop_u < prob
    return torch.where(drop_ids.reshape(-1, 1, 1, 1).expand(mask.shape), num_classes - 1, mask)


def to_tensor_normalize(x_u8_nhwc, mean=0.5, std=0.5):
    """transforms.ToTensor (uint8 HWC -> float CHW / 255) followed by transforms.Normalize(mean, std)."""
    t = x_u8_nhwc.permute(0, 3, 1, 2).float().div(255)
    return (t - mean) / std
