"""One eager training step of the north-star config between cudaProfilerStart/Stop (for `ncu --profile-from-start
off`): the launch list and the DRAM traffic of exactly one step.
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/gpu_step_once.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import nets, synthetic  # noqa: E402
from joligen_b200.trainer import PaletteTrainer  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = nets.build_palette_generator(image_size=256)
synthetic.dezero_init_(net, 1234)
tr = PaletteTrainer(net, lr=1e-4, optim="adamw", ema=True, ema_beta=0.999, device="cuda", cuda_graph=False)
dev = {k: v.cuda() for k, v in synthetic.synthetic_batch(batch, 256, 1234).items()}
for _ in range(2):
    tr.set_input(dev)
    tr.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.set_input(dev)
tr.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
