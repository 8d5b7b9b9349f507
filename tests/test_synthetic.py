"""The product's synthetic-data / seeded-init helpers equal the oracle's independent restatement."""
import torch

from joligen_b200 import nets, synthetic
from oracle import palette_oracle as O


def test_synthetic_batch_and_init_match_oracle():
    a = synthetic.synthetic_batch(3, 32, 99)
    b = O.synthetic_batch(3, 32, 99)
    assert torch.equal(a["A"], b["cond"]) and torch.equal(a["B"], b["gt"]) and torch.equal(a["B_label_mask"], b["mask"])
    m = a["B_label_mask"].float().mean(dim=(1, 2, 3))
    assert float(m.min()) >= 0.08 and float(m.max()) <= 0.45  # one box of 10-40 % per image
    cfg = O.UNetCfg(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1), attn_res=(2,),
                    num_head_channels=16)
    g = nets.build_palette_generator(image_size=32, inner_channel=32, channel_mults=(1, 2), res_blocks=(1, 1),
                                     attn_res=(2,), num_head_channels=16)
    synthetic.dezero_init_(g, 5)
    ref = O.init_params(cfg, 5)
    for k, p in g.named_parameters():
        assert torch.equal(p.detach(), ref[k]), k
        assert float(p.abs().sum()) > 0  # nothing left at the reference's zero init
