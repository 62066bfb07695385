"""noisediff_b200 — B200-native (sm_100a) implementation of NoiseDiff's reverse-diffusion sampling hot path.

Public surface mirrors the reference's two classes on that path:
  * :class:`NoiseDiffNet`        (reference ``models/archs/Diffusion_arch.py:447``)
  * :class:`GaussianDiffusion`   (reference ``models/denoising_diffusion_pytorch.py:167``)
Both call the C-ABI CUDA library in ``noisediff_b200/csrc`` (see ``include/noisediff_b200.h``).
"""
from .arch import NoiseDiffNet
from .diffusion import GaussianDiffusion, ModelPrediction, make_betas
from .engine import Engine
from . import frames  # noqa: F401
from . import _lib, tiles

__all__ = ["NoiseDiffNet", "GaussianDiffusion", "ModelPrediction", "make_betas", "Engine"]
__version__ = "0.1.0"
