"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch fp32 functional ops + explicit loops, no nn.Module graph) of the
reference's training inner loop for the Palette diffusion UNet and the GAN G/D operators.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it, and only as the checker / CPU baseline — never as the product path.

Parity status: PINNED BY GENERATED GOLDENS.  The reference's own tests hold no numeric golden
vectors for this path; the restatement is pinned against outputs of the real
reference modules imported in the build container (`oracle/gen_golden.py` → `tests/golden/*.pt`,
checked by `tests/test_oracle_golden.py`).
"""
