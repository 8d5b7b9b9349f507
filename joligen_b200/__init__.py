"""joligen_b200 — B200-native (sm_100a) kernels behind joliGEN's diffusion-UNet / GAN training inner loop.

    from joligen_b200 import accelerate, PaletteTrainer, build_palette_generator

See DESIGN.md (architecture, kernels, rooflines) and INTEGRATION.md (how joliGEN binds to it).
"""
from .accelerate import accelerate  # noqa: F401
from .nets import build_palette_generator  # noqa: F401
from .trainer import PaletteTrainer  # noqa: F401
