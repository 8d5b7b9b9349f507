"""csrc/prep.cu (on-GPU input preparation, Haar DWT / IWT) against oracle/prep_oracle.py
and the unmodified reference's vectors.  Index / mask handling and the elementwise image arithmetic are bit exact; the
2x2 Haar sums are compared at 1e-6 (the reference's conv2d does not fix the order of its four additions)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from joligen_b200 import kernels
    return kernels


def test_haar_forward_backward_vs_reference_golden(K, golden_dir):
    from joligen_b200 import ops
    g = torch.load(os.path.join(golden_dir, "prep_small.pt"))
    x = g["x"].cuda().requires_grad_(True)
    y = ops.haar_dwt(x)
    assert torch.allclose(y.cpu(), g["dwt"], atol=1e-6)
    y.backward(g["d_dwt"].cuda())
    assert torch.allclose(x.grad.cpu(), g["dx_dwt"], atol=1e-6)
    b = g["bands"].cuda().requires_grad_(True)
    z = ops.haar_iwt(b)
    assert torch.equal(z.cpu(), g["iwt"])  # one product per band, summed in the reference's order: bit exact
    z.backward(g["d_iwt"].cuda())
    assert torch.allclose(b.grad.cpu(), g["dbands_iwt"], atol=1e-6)
    assert torch.allclose(ops.haar_iwt(ops.haar_dwt(g["x"].cuda())).cpu(), g["x"], atol=2e-6)


def test_haar_at_the_benchmark_size(K):
    """256 x 256, batch 32: round trip + linearity (size-independent properties)."""
    from oracle import prep_oracle as P
    g = torch.Generator().manual_seed(1)
    x = torch.randn(32, 3, 256, 256, generator=g)
    xd = x.cuda()
    y = K.haar(xd, 0)
    assert y.shape == (32, 12, 128, 128)
    assert torch.allclose(K.haar(y, 2).cpu(), x, atol=2e-6)
    assert torch.allclose(y[:2].cpu(), P.haar_dwt(x[:2]), atol=1e-6)
    x2 = torch.randn(32, 3, 256, 256, generator=g).cuda()
    assert torch.allclose(K.haar(xd + 2 * x2, 0), y + 2 * K.haar(x2, 0), atol=1e-5)


def test_fill_mask_dropout_normalize_bit_exact(K, golden_dir):
    from oracle import prep_oracle as P
    g = torch.load(os.path.join(golden_dir, "prep_small.pt"))
    for cls, rec in g["fills"].items():
        for mask in (g["mask"], g["mask"].float()):
            out = K.fill_mask_random(g["img"].cuda(), mask.cuda(), rec["noise"].cuda(), cls)
            assert torch.equal(out.cpu(), rec["out"]), cls
    gen = torch.Generator().manual_seed(5)
    mask = torch.randint(0, 3, (6, 1, 40, 24), generator=gen)
    drop_u = torch.rand(6, generator=gen)
    for m in (mask, mask.float()):
        out = K.mask_class_dropout(m.cuda(), drop_u.cuda(), 0.4, 3)
        ref = P.mask_class_dropout(m, drop_u, 0.4, 4)
        assert torch.equal(out.cpu(), ref.to(out.dtype))
    u8 = torch.randint(0, 256, (3, 20, 28, 3), generator=gen, dtype=torch.uint8)
    assert torch.equal(K.u8_to_f32_normalized(u8.cuda()).cpu(), P.to_tensor_normalize(u8))
