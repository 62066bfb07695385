"""Per-layer rel-L2 of the engine's activations vs the CPU oracle at an arbitrary geometry (development aid; a checker, so it lives under tests/: only tests, smoke() and the bench CPU leg touch oracle/)."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noisediff_b200 as nd                      # noqa: E402
from noisediff_b200 import _lib                  # noqa: E402
from oracle import noisediff_oracle as O         # noqa: E402
from tests.util import seeded_sd, rel_l2         # noqa: E402
from tests.test_gpu_net import TAPS              # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sd = seeded_sd()
cond = O.synthetic_condition(B, S, S, seed=1)
x = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(9))
t = torch.full((B,), 500, dtype=torch.long)
taps = {}
torch.set_num_threads(os.cpu_count())
ref = O.net_forward(sd, x, t, cond, taps=taps)
eng = nd.Engine(dim=64, batch=B, height=S, width=S, flags=flags | _lib.FLAG_KEEP_ACTIVATIONS)
eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
eng.set_condition(cond["clean_img"].cuda(), cond["position"].cuda(), cond["iso_ratio_idx"].cuda())
out = eng.forward(x.cuda(), t.cuda())
torch.cuda.synchronize()
for name in TAPS:
    g = eng.debug_tensor(name)
    print(f"{name:20s} rel {rel_l2(g, taps[name]):.3e} nan {int(torch.isnan(g).sum())}")
print(f"out rel {rel_l2(out, ref):.3e} nan {int(torch.isnan(out).sum())}")
