"""One eager training step of cfg 5 (UNetVid, 8 frames) between cudaProfilerStart/Stop, for the ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/vid_step_launches.csv python tools/gpu_vid_step_once.py [size]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joligen_b200 import nets, nets_vid, synthetic  # noqa: E402
from joligen_b200.trainer import PaletteTrainer  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
unet = nets_vid.UNetVid(image_size=size, in_channel=6, inner_channel=64, out_channel=3, res_blocks=[2, 2, 2, 2],
                        attn_res=[16], tanh=False, n_timestep_train=2000, n_timestep_test=1000, norm="groupnorm",
                        group_norm_size=32, cond_embed_dim=32, channel_mults=(1, 2, 4, 8), num_heads=1,
                        num_head_channels=32)
net = nets.DiffusionGenerator(nets.PaletteDenoiseFn(unet, 32), image_size=size, G_ngf=64)
synthetic.dezero_init_(net, 3)
tr = PaletteTrainer(net.cuda(), lr=1e-4, optim="adamw", ema=True, device="cuda", cuda_graph=False)
data = {k: v.unsqueeze(0).cuda() for k, v in synthetic.synthetic_batch(8, size, 1).items()}
for _ in range(2):
    tr.set_input(data)
    tr.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.set_input(data)
tr.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
