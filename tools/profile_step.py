"""Runs a few reverse steps of the 256x256 engine (eager launches) — the command ncu wraps for profiles/."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noisediff_b200 as nd                      # noqa: E402
from noisediff_b200 import _lib                  # noqa: E402
from noisediff_b200 import tiles                 # noqa: E402
from tests.util import seeded_net                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps_n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = copy.deepcopy(seeded_net()).cuda()
gd = nd.GaussianDiffusion(net, image_size=256, timesteps=1000, beta_schedule="sigmoid2").cuda()
eng = nd.Engine(dim=64, batch=B, height=256, width=256, flags=_lib.FLAG_NO_GRAPH)
eng.load_state_dict(net.state_dict())
cond = {k: v.cuda() for k, v in tiles.synthetic_condition(B, 256).items()}
eng.set_condition(cond["clean_img"], cond["position"], cond["iso_ratio_idx"])
eng.chain_begin(gd.ddpm_steps(), None, 1)
eng.chain_run(steps_n)
torch.cuda.synchronize()
print("ok", float(eng.chain_read().std()))
