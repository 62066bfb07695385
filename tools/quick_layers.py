"""Quick per-layer timing table at the bench geometry (B patches of 256x256): python tools/quick_layers.py [B] [filter]"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noisediff_b200 as nd                      # noqa: E402
from noisediff_b200 import tiles                 # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flt = sys.argv[2] if len(sys.argv) > 2 else ""
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = nd.NoiseDiffNet(SimpleNamespace(dim=64, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False))
net = net.eval().requires_grad_(False).to(dev)
gd = nd.GaussianDiffusion(net, image_size=256, timesteps=1000, beta_schedule="sigmoid2", objective="pred_v").to(dev)
cond = {k: v.to(dev) for k, v in tiles.synthetic_condition(B, 256, seed=1).items()}
eng = net.engine_for(B, 256, 256, dev)
eng.set_condition(cond["clean_img"], cond["position"], cond["iso_ratio_idx"])
eng.chain_begin(gd.ddpm_steps(), None, 7)
eng.chain_run(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.chain_run(10); e1.record(); torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / 10:.3f} ms  (B={B})")
rows = eng.time_layers(5)
tot = sum(r[1] for r in rows)
print(f"sum of layers {tot:.3f} ms")
for n, t, f, _b in rows:
    if flt in n:
        print(f"{n:48s} {t * 1e3:9.1f} us {f / t / 1e9 if t > 0 else 0:9.1f} TF/s")
