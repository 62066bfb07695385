"""Generates tests/golden/stats_1k.npz: distribution statistics of 1024 noise samples drawn by the UNMODIFIED reference
(imported from /root/reference through oracle/ref_shim.py).  TEST INFRASTRUCTURE; run in the build container only:

    python -m oracle.make_golden_stats

BASELINE.json asks that the per-channel mean/variance and the 2-D noise power spectrum of 1k samples match the reference.
The reference's CPU path needs ~17 min per 256x256 chain of 1000 steps, so the 1k-sample comparison runs on reduced geometry
through the identical code (``GaussianDiffusion.sample`` -> ``p_sample_loop`` -> ``NoiseDiffNet.forward``,
models/denoising_diffusion_pytorch.py:446,375; models/archs/Diffusion_arch.py:577): 32x32 patches, T = 24 DDPM steps (sigmoid2,
pred_v), seed-0 weights, ONE condition shared by all patches so every sample is an i.i.d. draw of one distribution.

Two independent sets of 1024 samples are drawn (different torch seeds).  Set A is the fixture the CUDA path and the oracle are
compared with; set B measures the sampling noise of the statistics themselves (``self_*`` entries), which the tests'
tolerances are stated against.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import noisediff_oracle as O   # noqa: E402
from oracle import ref_shim                # noqa: E402
from oracle.make_golden import sd_hash     # noqa: E402
from tests.util import noise_stats         # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "stats_1k.npz")
N, S, T, CHUNK, COND_SEED = 1024, 32, 24, 256, 21


def shared_condition(n: int):
    one = O.synthetic_condition(1, S, S, seed=COND_SEED)
    return {k: v.expand(n, *v.shape[1:]).contiguous() for k, v in one.items()}


def draw(gd, n: int, seed0: int) -> torch.Tensor:
    outs = []
    for i in range(0, n, CHUNK):
        torch.manual_seed(seed0 + i // CHUNK)
        with torch.no_grad():
            outs.append(gd.sample(batch_size=CHUNK, condition=shared_condition(CHUNK)))
    return torch.cat(outs)


def main():
    torch.set_num_threads(os.cpu_count())
    net, gd = ref_shim.build(dim=64, seed=0, image_size=S, timesteps=T)
    sd = {k: v.detach() for k, v in net.module.state_dict().items()}
    t0 = time.time()
    a = draw(gd, N, 1000)
    print(f"set A: {time.time() - t0:.0f} s, std {float(a.std()):.4f}", flush=True)
    b = draw(gd, N, 2000)
    print(f"set B: {time.time() - t0:.0f} s", flush=True)
    sa, sb = noise_stats(a), noise_stats(b)
    self_mean = np.abs(sa["mean"] - sb["mean"]) / np.sqrt(sa["var"])
    self_var = np.abs(sa["var"] - sb["var"]) / sa["var"]
    self_radial = np.abs(sa["radial"] - sb["radial"]) / sa["radial"]
    nz = sa["psd2d"] > 0                                   # the DC bin is zero by construction (patch mean removed)
    self_psd = np.abs(sa["psd2d"] - sb["psd2d"])[nz] / sa["psd2d"][nz]
    print("self |dmean|/std", self_mean, "\nself |dvar|/var", self_var, "\nself radial max", self_radial.max(),
          "\nself psd2d max", self_psd.max(), "rms", np.sqrt((self_psd ** 2).mean()), flush=True)
    np.savez(OUT, n=N, size=S, timesteps=T, cond_seed=COND_SEED, weights_sha256=sd_hash(sd),
             mean=sa["mean"], var=sa["var"], psd2d=sa["psd2d"].astype(np.float32), radial=sa["radial"],
             mean_b=sb["mean"], var_b=sb["var"], psd2d_b=sb["psd2d"].astype(np.float32), radial_b=sb["radial"],
             self_mean=self_mean, self_var=self_var, self_radial_max=self_radial.max(), self_psd2d_max=self_psd.max(),
             self_psd2d_rms=np.sqrt((self_psd ** 2).mean()), first=a[:2].numpy())
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
