"""CPU checks of the drop-in boundary: the C-ABI shared library builds/loads and exports every symbol
that include/jg_b200.h declares; the Python binding table covers exactly those symbols; the B200
modules expose the reference's state_dict keys.  No kernels are launched here."""
import ctypes
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "jg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from joligen_b200 import build, lib
    path = build.build()
    assert os.path.exists(path)
    so = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(so, s), "libjg_b200.so does not export %s" % s
    assert sorted(lib.exported_symbols()) == syms, (set(syms) ^ set(lib.exported_symbols()))
    assert lib.load().jg_version() >= 100


def test_bad_arguments_fail_loudly_without_gpu():
    """Argument validation happens before any CUDA call: errors come back as codes + messages."""
    from joligen_b200 import lib
    l = lib.load()
    d = lib.ConvDesc()
    d.N, d.H, d.W, d.Cin, d.ldx, d.Ho, d.Wo, d.Cout, d.ldy = 1, 8, 8, 6, 6, 8, 8, 64, 64
    d.R, d.S, d.stride, d.pad = 3, 3, 1, 1
    rc = l.jg_conv2d_fwd(ctypes.byref(d), 16, 16, 0, 0, 16, 0)
    assert rc == -1 and b"multiples of 8" in l.jg_last_error()
    rc = l.jg_attn_fwd(16, 96, 16, 32, 16, 1, 100, 2, 16, 0, 0)
    assert rc == -1 and b"multiple of 64" in l.jg_last_error()


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device (no silent eager fallback)."""
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from joligen_b200 import nets
    from joligen_b200.trainer import PaletteTrainer
    g = nets.build_palette_generator(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1),
                                     attn_res=(2,), num_head_channels=16)
    with pytest.raises(RuntimeError):
        PaletteTrainer(g)
    with pytest.raises(Exception):
        g(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32), None, torch.zeros(1, 3, 32, 32))


def test_no_cpu_fallback_in_the_widening_rows():
    """The b2b backbone and its trainer, the CUT trainer and the new autograd ops: hard errors without a CUDA device."""
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from joligen_b200 import nets_jit, ops_jit
    from joligen_b200.trainer_b2b import B2BTrainer
    net = nets_jit.B2BGenerator(nets_jit.JiTViD(input_size=32, patch_size=8, hidden_size=192, depth=1, num_heads=6,
                                                in_context_len=0, in_context_start=0, motion_num_layers=1))
    with pytest.raises(RuntimeError):
        B2BTrainer(net)
    with pytest.raises(Exception):
        net(torch.zeros(1, 2, 3, 32, 32), torch.ones(1, 2, 1, 32, 32), None, torch.zeros(1, dtype=torch.long))
    with pytest.raises(Exception):
        ops_jit.swiglu(torch.zeros(1, 4, 1, 16, dtype=torch.bfloat16))


def test_b200_modules_have_reference_state_dict_keys():
    from joligen_b200 import nets
    from oracle import palette_oracle as O
    cfg = O.UNetCfg()
    g = nets.build_palette_generator()
    assert [(k, tuple(v.shape)) for k, v in g.named_parameters()] == list(O.generator_param_shapes(cfg).items())
    assert len(g.state_dict()) == 338  # 338 entries for the default net
    bufs = dict(g.named_buffers())
    sched = O.schedule_buffers(cfg, "train")
    for k, v in sched.items():
        assert torch.equal(bufs["denoise_fn.model." + k], v)  # schedule tables: bit exact
