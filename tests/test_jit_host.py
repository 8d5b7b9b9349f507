"""Host-side checks of the b2b backbone mirror (no GPU): nets_jit builds the reference's parameter set (names, shapes,
trainable flags) for both the small test configuration and JiTVid-B/16 as the plumbing golden has it, and its rotary tables
are the oracle's."""
import os

import pytest
import torch


@pytest.mark.parametrize("name", ["jit_b200.pt", "b2b_plumbing.pt"])
def test_parameter_set_matches_reference(golden_dir, name):
    from joligen_b200 import nets_jit
    gold = torch.load(os.path.join(golden_dir, name))
    c = dict(gold["cfg"])
    with torch.device("meta"):
        net = nets_jit.B2BGenerator(nets_jit.JiTViD(**c))
    mine = {k: tuple(v.shape) for k, v in net.named_parameters() if v.requires_grad}
    assert mine == dict(gold["shapes"])
    frozen = {k: tuple(v.shape) for k, v in net.named_parameters() if not v.requires_grad}
    assert frozen == {k: tuple(v.shape) for k, v in gold["frozen"].items()}


def test_rope_tables_match_oracle():
    from joligen_b200.nets_jit import rope_tables
    from oracle import jit_oracle as J
    for hd, grid, prefix in ((16, 4, 4), (32, 4, 0), (64, 8, 32)):
        cfg = J.JitCfg(input_size=grid * 16, patch_size=16, hidden_size=hd * 2, num_heads=2, in_context_len=prefix)
        c0, s0 = J.rope_tables(cfg, prefix)
        c1, s1 = rope_tables(hd, grid, prefix, "cpu")
        assert torch.equal(c0, c1) and torch.equal(s0, s1)


def test_unsupported_shapes_are_hard_errors():
    from joligen_b200 import nets_jit
    with pytest.raises(NotImplementedError):
        nets_jit.JiTViD(input_size=32, patch_size=8, hidden_size=96, depth=1, num_heads=8)      # head dim 12
    with pytest.raises(NotImplementedError):
        nets_jit.JiTViD(input_size=32, patch_size=8, hidden_size=128, depth=1, num_heads=4)     # SwiGLU width 341
    with pytest.raises(NotImplementedError):
        nets_jit.JiTViD(input_size=32, patch_size=8, hidden_size=192, depth=1, num_heads=6, motion_every=2)
