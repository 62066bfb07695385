"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shim.py) on seeded inputs.  TEST INFRASTRUCTURE; run in the build container only:

    python -m oracle.make_golden

The reference ships no tests or golden vectors, so these files are the parity pin: the oracle
restatement (oracle/noisediff_oracle.py) must reproduce them, and the CUDA path is then checked against the oracle.
Weights are NOT stored (150 MB): both the reference and noisediff_b200.NoiseDiffNet create them from
``torch.manual_seed(0)`` with identical RNG consumption; the fixture stores a SHA-256 of the state_dict instead.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import noisediff_oracle as O   # noqa: E402
from oracle import ref_shim                # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sd_hash(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


class record_draws:
    """Records the LOGICAL values of every torch.randn_like the reference sampler makes (the draw of each noisy
    step, denoising_diffusion_pytorch.py:371,:433).  Recording beats replaying the seed: randn_like fills in memory
    order, so when the running image is channels-last (pred_x0 on CPU: x inherits the conv output's strides) the
    same RNG stream lands on different logical elements."""

    def __enter__(self):
        self.draws, self._orig = [], torch.randn_like
        def rec(x, *a, **k):
            r = self._orig(x, *a, **k)
            self.draws.append(r.detach().clone().contiguous())
            return r
        torch.randn_like = rec
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    net, _ = ref_shim.build(dim=64, seed=0, image_size=64, timesteps=8)
    sd = {k: v.detach() for k, v in net.module.state_dict().items()}
    meta = dict(weights_sha256=sd_hash(sd), n_keys=len(sd), n_params=int(sum(v.numel() for v in sd.values())))
    print(meta)

    # ---- A: single forward, 64x64, B=2, per-sample t and iso index -----------------------------------------------
    cond = O.synthetic_condition(2, 64, 64, seed=1)
    cond["iso_ratio_idx"] = torch.tensor([24, 3])
    x = torch.randn(2, 4, 64, 64, generator=torch.Generator().manual_seed(5))
    t = torch.tensor([3, 977])
    with torch.no_grad():
        v = net(x, t, cond)
    np.savez(os.path.join(OUT, "fwd_64.npz"), x=x.numpy(), t=t.numpy(), clean=cond["clean_img"].numpy(),
             position=cond["position"].numpy(), iso=cond["iso_ratio_idx"].numpy(), out=v.numpy(), **meta)

    # ---- B: DDPM chain T=8 (sigmoid2, pred_v), all timesteps, RNG replay stored --------------------------------------
    Net, GD = ref_shim.load()
    cond2 = O.synthetic_condition(2, 64, 64, seed=2)
    gd = GD(net, image_size=64, timesteps=8, beta_schedule="sigmoid2", objective="pred_v")
    torch.manual_seed(123)
    with torch.no_grad(), record_draws() as r:
        allx = gd.sample(batch_size=2, condition=cond2, return_all_timesteps=True)      # (B, T+1, 4, 64, 64)
    x_T, zs = allx[:, 0], r.draws
    assert len(zs) == 7
    np.savez(os.path.join(OUT, "chain_ddpm_T8.npz"), xs=allx.numpy(), x_T=x_T.numpy(), noises=torch.stack(zs).numpy(),
             clean=cond2["clean_img"].numpy(), position=cond2["position"].numpy(), iso=cond2["iso_ratio_idx"].numpy(),
             **meta)

    # ---- C: DDIM, T=50, S=5, eta=0.5 -----------------------------------------------------------------------------
    gdi = GD(net, image_size=64, timesteps=50, sampling_timesteps=5, ddim_sampling_eta=0.5, beta_schedule="sigmoid2",
             objective="pred_v")
    torch.manual_seed(321)
    with torch.no_grad(), record_draws() as r:
        alli = gdi.sample(batch_size=2, condition=cond2, return_all_timesteps=True)
    x_Ti, zsi = alli[:, 0], r.draws
    assert len(zsi) == 4
    np.savez(os.path.join(OUT, "chain_ddim_T50_S5.npz"), xs=alli.numpy(), x_T=x_Ti.numpy(),
             noises=torch.stack(zsi).numpy(), eta=0.5, **meta)

    # ---- D: other objectives, DDPM T=4 -------------------------------------------------------------------------------
    extra = {}
    for obj in ("pred_noise", "pred_x0"):
        g = GD(net, image_size=64, timesteps=4, beta_schedule="cosine", objective=obj)
        torch.manual_seed(77)
        with torch.no_grad(), record_draws() as r:
            extra[obj] = g.sample(batch_size=2, condition=cond2, return_all_timesteps=True).numpy()
        extra[obj + "_noises"] = torch.stack(r.draws).numpy()
    np.savez(os.path.join(OUT, "chain_objectives_T4.npz"), x_T=extra["pred_x0"][:, 0], **extra, **meta)

    # ---- E: one full-size forward (the BASELINE shape), B=1, 256x256 -----------------------------------------------
    cond3 = O.synthetic_condition(1, 256, 256, seed=1)
    x3 = torch.randn(1, 4, 256, 256, generator=torch.Generator().manual_seed(9))
    t3 = torch.tensor([500])
    with torch.no_grad():
        v3 = net(x3, t3, cond3)
    np.savez(os.path.join(OUT, "fwd_256.npz"), x=x3.numpy().astype(np.float32), t=t3.numpy(), out=v3.numpy(), **meta)

    # ---- F: schedule tables of every schedule the reference defines, T=1000 and T=50 -----------------------------------
    tabs = {}
    for name in ("linear", "cosine", "sigmoid1", "sigmoid2", "sigmoid3"):
        for T in (1000, 50):
            g = GD(net, image_size=64, timesteps=T, beta_schedule=name, objective="pred_v")
            for k, b in g.named_buffers(recurse=False):
                tabs[f"{name}/{T}/{k}"] = b.numpy()
    np.savez(os.path.join(OUT, "schedules.npz"), **tabs)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
