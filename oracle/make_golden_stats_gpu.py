"""Generates tests/golden/stats_1k_64_T1000.npz: distribution statistics of 1024 noise samples at 64 x 64 with the FULL T = 1000
chain, drawn by the oracle (oracle/noisediff_oracle.py, the restatement pinned bit-exactly to the unmodified reference by
tests/test_oracle.py) running in fp32 — TF32 off — on a B200.  TEST INFRASTRUCTURE; run on a GPU box:

    gpurun -- 'python -m oracle.make_golden_stats_gpu gpurun_out/stats_1k_64_T1000.npz'      (then copy it to tests/golden/)

BASELINE.json asks that the per-channel mean / variance and the 2-D noise power spectrum of 1k samples match the reference.
The reference's CPU path needs ~1 min per 64 x 64 chain of 1000 steps (17 min at 256 x 256), so the 1k-sample reference set is
drawn on the GPU through the identical code path; the reduced-geometry fixture minted from the reference ITSELF
(stats_1k.npz: 32 x 32, T = 24) stays alongside.  /root/reference does not exist on the GPU box, hence the oracle.
Seed-0 weights, ONE condition shared by all patches (every sample is an i.i.d. draw of one distribution), DDPM, sigmoid2,
pred_v.  The two halves of the set (512 + 512) are also summarised: their difference is the sampling-noise yardstick
(a 1024-vs-1024 comparison has 1/sqrt(2) of it).  Budget: ~7 GPU-minutes.
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
from tests.util import noise_stats, sd_hash, seeded_sd         # noqa: E402

N, S, T, CHUNK, COND_SEED = 1024, 64, 1000, 512, 21


def main(out_path):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd_cpu = seeded_sd()
    sd = {k: v.cuda() for k, v in sd_cpu.items()}
    one = O.synthetic_condition(1, S, S, seed=COND_SEED)
    cond = {k: v.expand(CHUNK, *v.shape[1:]).contiguous().cuda() for k, v in one.items()}
    tab = O.schedule_tables("sigmoid2", T)
    outs = []
    t0 = time.time()
    with torch.no_grad():
        for c in range(N // CHUNK):
            g = torch.Generator(device="cuda").manual_seed(5000 + c)
            x = torch.randn(CHUNK, 4, S, S, generator=g, device="cuda")
            for t in reversed(range(T)):
                out = O.net_forward(sd, x, torch.full((CHUNK,), t, dtype=torch.long, device="cuda"), cond)
                z = torch.randn(CHUNK, 4, S, S, generator=g, device="cuda") if t > 0 else None
                x, _ = O.ddpm_step(tab, "pred_v", x, t, out, z)
            outs.append(x.cpu())
            print(f"chunk {c}: {time.time() - t0:.0f} s, std {float(x.std()):.4f}", flush=True)
    a = torch.cat(outs)
    sa, s1, s2 = noise_stats(a), noise_stats(a[: N // 2]), noise_stats(a[N // 2:])
    self_mean = np.abs(s1["mean"] - s2["mean"]) / np.sqrt(sa["var"])
    self_var = np.abs(s1["var"] - s2["var"]) / sa["var"]
    self_radial = np.abs(s1["radial"] - s2["radial"]) / sa["radial"]
    nz = sa["psd2d"] > 1e-6 * np.median(sa["psd2d"])          # the DC bins are zero up to rounding (patch mean removed)
    self_psd = np.abs(s1["psd2d"] - s2["psd2d"])[nz] / sa["psd2d"][nz]
    print("halves |dmean|/std", self_mean, "\nhalves |dvar|/var", self_var, "\nhalves radial max", self_radial.max(),
          "\nhalves psd2d max", self_psd.max(), "rms", np.sqrt((self_psd ** 2).mean()), flush=True)
    np.savez(out_path, n=N, size=S, timesteps=T, cond_seed=COND_SEED, weights_sha256=sd_hash(sd_cpu),
             mean=sa["mean"], var=sa["var"], psd2d=sa["psd2d"].astype(np.float32), radial=sa["radial"],
             halves_mean=self_mean, halves_var=self_var, halves_radial_max=self_radial.max(), halves_psd2d_max=self_psd.max(),
             halves_psd2d_rms=np.sqrt((self_psd ** 2).mean()), first=a[:2].numpy(), seconds=time.time() - t0)
    print(out_path, os.path.getsize(out_path) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "stats_1k_64_T1000.npz"))
