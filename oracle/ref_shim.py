"""Import shim for the UNMODIFIED reference at /root/reference (build container only; TEST INFRASTRUCTURE).

The reference imports four packages that are absent here and unused on the sampling path (``ema_pytorch``,
``matplotlib``, ``tensorboardX``, ``rawpy`` — SURVEY.md §8c).  We register empty stand-ins in ``sys.modules`` and
import the reference modules in place.  Nothing is copied.  ``/root/reference`` does not exist on the GPU box, so
only ``oracle/make_golden.py`` and container-side tests (skipped when the tree is absent) use this file.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

REF_ROOT = os.environ.get("NOISEDIFF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "denoising_diffusion_pytorch.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns (NoiseDiffNet, GaussianDiffusion) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _stub("ema_pytorch", EMA=type("EMA", (), {}))
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("tensorboardX", SummaryWriter=type("SummaryWriter", (), {}))
    _stub("rawpy")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models.archs.Diffusion_arch import NoiseDiffNet            # noqa: E402
    from models.denoising_diffusion_pytorch import GaussianDiffusion  # noqa: E402
    return NoiseDiffNet, GaussianDiffusion


def net_args(dim: int = 64):
    return SimpleNamespace(dim=dim, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False)


def build(dim: int = 64, seed: int = 0, image_size: int = 64, timesteps: int = 1000, **gd_kwargs):
    """Seeded reference network (wrapped in DataParallel — mandatory, denoising_diffusion_pytorch.py:189) and
    its GaussianDiffusion."""
    import torch
    from torch import nn
    Net, GD = load()
    torch.manual_seed(seed)
    net = nn.DataParallel(Net(net_args(dim))).eval()
    gd_kwargs.setdefault("beta_schedule", "sigmoid2")
    gd_kwargs.setdefault("objective", "pred_v")
    gd = GD(net, image_size=image_size, timesteps=timesteps, **gd_kwargs)
    return net, gd
