"""Shared helpers for the test-suite (the oracle is the checker, never the thing under test on the GPU side)."""
from __future__ import annotations

import hashlib
import os
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def net_args(dim=64):
    return SimpleNamespace(dim=dim, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False)


def sd_hash(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


_net_cache = {}


def seeded_net(dim=64, seed=0):
    """noisediff_b200.NoiseDiffNet with the same weights the reference gets from torch.manual_seed(seed)."""
    import noisediff_b200 as nd
    key = (dim, seed)
    if key not in _net_cache:
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        net = nd.NoiseDiffNet(net_args(dim)).eval()
        torch.random.set_rng_state(state)
        for p in net.parameters():
            p.requires_grad_(False)
        _net_cache[key] = net
    return _net_cache[key]


def seeded_sd(dim=64, seed=0):
    """CPU fp32 state_dict (the oracle runs on the host even when the module was moved to the GPU)."""
    return {k: v.detach().cpu() for k, v in seeded_net(dim, seed).state_dict().items()}


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def noise_stats(x: torch.Tensor, nbins: int = 8):
    """x: [n, C, S, S] -> dict of float64 arrays: per-channel mean / variance over (n, S, S); psd2d [C, S, S] = mean
    periodogram |FFT2(x - patch mean)|^2 / S^2 (a noise-spectrum estimate removes each patch's DC); radial [C, nbins] = psd2d
    averaged over annuli of |f|, normalised to sum 1 per channel."""
    x = x.double().cpu()
    mean, var = x.mean(dim=(0, 2, 3)), x.var(dim=(0, 2, 3))
    xc = x - x.mean(dim=(-1, -2), keepdim=True)
    psd = torch.fft.fft2(xc).abs().pow(2).mean(dim=0) / (x.shape[-1] * x.shape[-2])
    f = torch.fft.fftfreq(x.shape[-1]).abs()
    rad = torch.sqrt(f[:, None] ** 2 + f[None, :] ** 2)
    edges = torch.linspace(0, float(rad.max()) + 1e-9, nbins + 1)
    radial = torch.stack([psd[:, (rad >= edges[i]) & (rad < edges[i + 1])].mean(dim=1) for i in range(nbins)], dim=1)
    radial = radial / radial.sum(dim=1, keepdim=True)
    return {"mean": mean.numpy(), "var": var.numpy(), "psd2d": psd.numpy(), "radial": radial.numpy()}
