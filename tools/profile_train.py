"""Runs training steps of the B = 32, 256 x 256 trainer — the command ncu wraps for the training row's launch list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noisediff_b200 as nd                      # noqa: E402
from noisediff_b200 import tiles, training       # noqa: E402
from types import SimpleNamespace                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
net = nd.NoiseDiffNet(SimpleNamespace(dim=64, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False)).eval().requires_grad_(False).cuda()
gd = nd.GaussianDiffusion(net, image_size=256, timesteps=1000, beta_schedule="sigmoid2", objective="pred_v").cuda()
tr = training.DiffusionTrainer(gd, batch_size=B, lr=1e-4)
cond = {k: v.cuda() for k, v in tiles.synthetic_condition(B, 256).items()}
img = torch.randn(B, 4, 256, 256, device="cuda") * 0.05
for i in range(steps):
    loss = tr.step(img, cond)
torch.cuda.synchronize()
print("ok", loss)
