"""Generates tests/golden/losses.npz: training-loss VALUES of the UNMODIFIED reference (``GaussianDiffusion.p_losses``,
models/denoising_diffusion_pytorch.py:481-531) for the three objectives on seeded inputs with per-sample timesteps.
TEST INFRASTRUCTURE; run in the build container only:  python -m oracle.make_golden_losses
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import noisediff_oracle as O   # noqa: E402
from oracle import ref_shim                # noqa: E402
from oracle.make_golden import sd_hash     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "losses.npz")
CASES = {"pred_v": ("sigmoid2", 1000, [17, 803]), "pred_noise": ("cosine", 100, [3, 96]), "pred_x0": ("linear", 100, [50, 9])}


def main():
    torch.set_num_threads(os.cpu_count())
    net, _ = ref_shim.build(dim=64, seed=0, image_size=64, timesteps=8)
    _, GD = ref_shim.load()
    sd = {k: v.detach() for k, v in net.module.state_dict().items()}
    cond = O.synthetic_condition(2, 64, 64, seed=3)
    g = torch.Generator().manual_seed(41)
    x_start = torch.randn(2, 4, 64, 64, generator=g) * 0.05
    noise = torch.randn(2, 4, 64, 64, generator=g)
    out = dict(x_start=x_start.numpy(), noise=noise.numpy(), clean=cond["clean_img"].numpy(), position=cond["position"].numpy(),
               iso=cond["iso_ratio_idx"].numpy(), weights_sha256=sd_hash(sd))
    for obj, (sched, T, ts) in CASES.items():
        gd = GD(net, image_size=64, timesteps=T, beta_schedule=sched, objective=obj)
        t = torch.tensor(ts)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):      # the pred_x0 branch prints
            loss = gd.p_losses(x_start.clone(), t, cond, noise=noise.clone())
        out[obj + "/loss"] = np.float32(loss.item())
        out[obj + "/t"] = t.numpy()
        out[obj + "/schedule"], out[obj + "/T"] = sched, T
        print(obj, sched, T, ts, float(loss))
    np.savez(OUT, **out)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
