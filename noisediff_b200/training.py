"""Diffusion training step on the B200 library (BASELINE configs[4]).

Host-side mirror of what ``Trainer.train`` does per batch (``models/trainer_diffusion.py:176-191``)::

    loss = diffusion(noise_gt, condition=...)   # GaussianDiffusion.forward -> p_losses (denoising_diffusion_pytorch.py:481-542)
    loss.backward(); optimizer_G.step(); ema.update()

with the arithmetic behind the C ABI (``ndiff_trainer_*``, ``include/noisediff_b200.h``): forward with saved activations,
backward, Adam and the EMA lerp are CUDA kernels; this module is plumbing — the three elementwise lines of ``q_sample`` / the
regression target / the loss weight on the caller's tensors, the schedules (``CosineAnnealingLR``, the EMA warm-up of
``ema_pytorch``), and the data-parallel gradient all-reduce (ONE ``torch.distributed.all_reduce`` over the library's flat fp32
gradient buffer — NCCL over NVLink on a B200 box, gloo in the CPU tests).

EMA: the reference uses ``ema_pytorch.EMA(net, beta=0.995, update_after_step=500, update_every=20)`` — a pip dependency that is
NOT vendored in the reference tree and is unpinned in ``install.sh:12`` (parity unpinned for this part: restated from the
package's published algorithm, ema-pytorch 0.2-0.7 ``EMA.update / get_current_decay`` with its defaults inv_gamma = 1,
power = 2/3, min_value = 0): see :class:`EmaSchedule`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib
from .arch import NoiseDiffNet

def train_lib() -> C.CDLL:
    """The shared library (the training entry points live in the same .so as the sampling path; signatures in ``_lib._SIGS``)."""
    return _lib.lib()


def cosine_lr(base_lr: float, epoch: int, t_max: int, eta_min: float = 0.0) -> float:
    """``torch.optim.lr_scheduler.CosineAnnealingLR(T_max=max_iter)`` in closed form (models/trainer_diffusion.py:93; the reference
    steps it once per epoch, :152-154): eta_min + (base - eta_min) * (1 + cos(pi * epoch / T_max)) / 2."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / t_max)) / 2.0


class EmaSchedule:
    """When and how strongly ``ema_pytorch.EMA.update()`` moves the average (models/trainer_diffusion.py:63-69,191).

    update() -> one of ("skip", 0), ("copy", 1) or ("lerp", 1 - decay):
      step = self.step; self.step += 1
      step % update_every != 0          -> nothing
      step <= update_after_step         -> copy the parameters
      first update after that           -> copy, then lerp
      decay = clamp(1 - (1 + epoch / inv_gamma) ** -power, min_value, beta), epoch = max(self.step - update_after_step - 1, 0),
              0 when epoch <= 0;  ema.lerp_(param, 1 - decay)
    """

    def __init__(self, beta: float = 0.995, update_after_step: int = 500, update_every: int = 20, inv_gamma: float = 1.0,
                 power: float = 2.0 / 3.0, min_value: float = 0.0):
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.step, self.initted = 0, False

    def current_decay(self) -> float:
        epoch = max(self.step - self.update_after_step - 1, 0)
        if epoch <= 0:
            return 0.0
        value = 1.0 - (1.0 + epoch / self.inv_gamma) ** -self.power
        return min(max(value, self.min_value), self.beta)

    def update(self):
        """Returns the list of actions for this call, in order: [] | [("copy", 1.0)] | [("copy", 1.0), ("lerp", w)] | [("lerp", w)]."""
        step = self.step
        self.step += 1
        if step % self.update_every != 0:
            return []
        if step <= self.update_after_step:
            return [("copy", 1.0)]
        acts = []
        if not self.initted:
            acts.append(("copy", 1.0))
            self.initted = True
        acts.append(("lerp", 1.0 - self.current_decay()))
        return acts


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> float:
    """DDP's gradient averaging over the ranks as ONE collective on the flat gradient buffer; returns the scale the optimizer
    applies (1 / world size; the SUM stays in the buffer).  No-op (scale 1) without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DiffusionTrainer:
    """One training engine (fixed batch / crop / device) around a ``GaussianDiffusion``: ``step(img, condition)`` is
    ``loss = diffusion(img, condition); loss.backward(); optimizer.step(); ema.update()`` of the reference's loop, returning the
    loss value.  Weights are read from ``diffusion.model`` once (``state_dict``) and then live in the library's flat fp32 buffer;
    ``state_dict()`` / ``ema_state_dict()`` read them back in the reference's layout (what ``Trainer.save_networks`` writes)."""

    def __init__(self, diffusion, *, batch_size: int, lr: float = 1e-4, weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 ema: Optional[EmaSchedule] = None, device: Optional[torch.device] = None):
        self.lib = train_lib()
        if not torch.cuda.is_available():
            raise RuntimeError("noisediff_b200 training needs a CUDA device (sm_100a); there is no CPU path")
        self.diffusion = diffusion
        net = diffusion._net()
        if not isinstance(net, NoiseDiffNet):
            raise TypeError("DiffusionTrainer drives noisediff_b200.NoiseDiffNet")
        if diffusion.objective not in ("pred_v", "pred_noise"):
            raise NotImplementedError("the training kernels implement the pred_v / pred_noise objectives (the reference trains pred_v)")
        self.device = device if device is not None else diffusion.device
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.batch, self.size = int(batch_size), int(diffusion.image_size)
        self.base_lr, self.lr, self.weight_decay, self.betas, self.eps = lr, lr, weight_decay, betas, eps
        self.ema = ema if ema is not None else EmaSchedule()
        cfg = _lib.Config(net.dim, self.batch, self.size, self.size, self.device_index, 0)
        h = C.c_void_p()
        with torch.cuda.device(self.device_index):
            _lib.check(self.lib.ndiff_trainer_create(C.byref(cfg), C.byref(h)))
            self._h = h
            self._eng = C.c_void_p(self.lib.ndiff_trainer_engine(h))
            st = self._stream()
            keep = []
            for k, v in net.state_dict().items():
                t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(self.lib.ndiff_load_param_async(self._eng, k.encode(), _ptr(t), t.dim(), shape, st))
            _lib.check(self.lib.ndiff_trainer_finalize(h, st))
        self._names = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        self.steps_done = 0
        self._keep = []

    # ---- lifetime ---------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            with torch.cuda.device(self.device_index):
                self.lib.ndiff_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device_index).cuda_stream)

    # ---- views of the library's flat buffers -------------------------------------------------------------------
    def _flat(self, which: int) -> torch.Tensor:
        """Zero-copy fp32 view of flat buffer `which` (0 parameters, 1 gradients, 2 Adam m, 3 Adam v, 4 EMA) — library memory."""
        cache = self.__dict__.setdefault("_flat_cache", {})
        if which in cache:
            return cache[which][1]
        p, n_live, n_tot = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.ndiff_trainer_flat(self._h, which, C.byref(p), C.byref(n_live), C.byref(n_tot)))
        n = n_tot.value

        class _Holder:      # __cuda_array_interface__ over the raw pointer
            pass
        hold = _Holder()
        hold.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (p.value, False), "version": 3, "strides": None}
        t = torch.as_tensor(hold, device=self.device)
        cache[which] = (hold, t)
        return t

    def _slot(self, name: str):
        off, n = C.c_int64(), C.c_int64()
        _lib.check(self.lib.ndiff_trainer_slot(self._h, name.encode(), C.byref(off), C.byref(n)))
        return off.value, n.value

    def _named(self, which: int) -> Dict[str, torch.Tensor]:
        flat = self._flat(which)
        out = {}
        for k, shape in self._names.items():
            off, n = self._slot(k)
            out[k] = flat[off:off + n].reshape(shape).clone()
        return out

    def state_dict(self):
        return self._named(0)

    def gradients(self):
        return self._named(1)

    def ema_state_dict(self):
        return self._named(4)

    # ---- one training step --------------------------------------------------------------------------------------
    def forward_backward(self, img: torch.Tensor, condition: Dict[str, torch.Tensor], t: Optional[torch.Tensor] = None,
                         noise: Optional[torch.Tensor] = None) -> float:
        """``GaussianDiffusion.forward`` + ``loss.backward()`` (ref :481-542): draws t and the noise like the reference unless they
        are given (tests inject both), leaves the gradients in the flat gradient buffer, returns the loss value."""
        gd = self.diffusion
        B = img.shape[0]
        if tuple(img.shape) != (self.batch, 4, self.size, self.size):
            raise ValueError(f"this trainer was built for batches of shape {(self.batch, 4, self.size, self.size)}, got {tuple(img.shape)}")
        dev = self.device
        img = gd.normalize(img.to(dev, torch.float32))
        if t is None:
            t = torch.randint(0, gd.num_timesteps, (B,), device=dev).long()                # ref :540
        t = t.to(dev).long()
        if noise is None:
            noise = torch.randn_like(img)                                                   # ref :486
        if gd.offset_noise_strength > 0.:                                                   # ref :490-492
            noise = noise + gd.offset_noise_strength * torch.randn(img.shape[:2], device=dev)[:, :, None, None]
        x = gd.q_sample(x_start=img, t=t, noise=noise).contiguous()                          # ref :496
        target = (noise if gd.objective == "pred_noise" else gd.predict_v(img, t, noise)).contiguous()     # ref :508-514
        w = gd.loss_weight.gather(-1, t).to(torch.float32).contiguous()                      # ref :520
        with torch.cuda.device(self.device_index):
            st = self._stream()
            c = condition["clean_img"].to(dev, torch.float32).contiguous()
            p = condition["position"].to(dev, torch.float32).contiguous()
            i = condition["iso_ratio_idx"].to(dev, torch.int64).contiguous()
            _lib.check(self.lib.ndiff_set_condition(self._eng, _ptr(c), _ptr(p), _ptr(i), st))
            loss = C.c_double()
            _lib.check(self.lib.ndiff_trainer_forward_backward(self._h, _ptr(x), _ptr(t), _ptr(target), _ptr(w), C.byref(loss), st))
        self._keep = [c, p, i, x, t, target, w]
        return float(loss.value)

    def optimizer_step(self, group=None):
        """All-reduce (data parallel), ``Adam.step()``, ``ema.update()``."""
        scale = allreduce_gradients(self._flat(1), group)
        with torch.cuda.device(self.device_index):
            st = self._stream()
            _lib.check(self.lib.ndiff_trainer_adam_step(self._h, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, scale, st))
            for kind, wgt in self.ema.update():
                _lib.check(self.lib.ndiff_trainer_ema_update(self._h, 1.0 if kind == "copy" else wgt, st))
        self.steps_done += 1

    def step(self, img, condition, t=None, noise=None, group=None) -> float:
        loss = self.forward_backward(img, condition, t, noise)
        self.optimizer_step(group)
        return loss

    def set_epoch(self, epoch: int, max_iter: int):
        """``scheduler.step()`` at the top of every epoch (ref :152-154)."""
        self.lr = cosine_lr(self.base_lr, epoch, max_iter)

    def time_families(self):
        """Per-kernel-family durations of the last step's launch lists: [(pass, family, launches, ms)]."""
        buf = C.create_string_buffer(16 * 1024)
        with torch.cuda.device(self.device_index):
            _lib.check(self.lib.ndiff_trainer_time(self._h, buf, len(buf), self._stream()))
        rows = []
        for line in buf.value.decode().strip().split("\n"):
            ps, fam, n, ms = line.split(";")
            rows.append((ps, fam, int(n), float(ms)))
        return rows

    @property
    def activation_bytes(self) -> int:
        return int(self.lib.ndiff_trainer_activation_bytes(self._h))

    @property
    def launches(self):
        return int(self.lib.ndiff_trainer_launches(self._h, 0)), int(self.lib.ndiff_trainer_launches(self._h, 1))
