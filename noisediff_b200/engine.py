"""Python handle on one C-ABI engine (fixed micro-batch / crop / device).  PyTorch here is plumbing only: it owns the
device buffers whose raw pointers cross the C boundary, and names the current CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device_index: int):
    return C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Engine:
    def __init__(self, *, dim: int, batch: int, height: int, width: int, device: int = 0, flags: int = 0):
        self._lib = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("noisediff_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.dim, self.batch, self.height, self.width, self.device_index = dim, batch, height, width, int(device)
        self.device = torch.device("cuda", self.device_index)
        cfg = _lib.Config(dim, batch, height, width, self.device_index, flags)
        h = C.c_void_p()
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.weights_version = None
        self._cond_key = None
        self._keep = []          # tensors whose pointers the library may still read asynchronously

    # ---- lifetime ---------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            with torch.cuda.device(self.device_index):
                self._lib.ndiff_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ----------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Mirror of ``Trainer.load_networks`` (ref models/trainer_diffusion.py:333-349): strips 'module.' and is
        strict about the live keys."""
        with torch.cuda.device(self.device_index):
            st = _stream(self.device_index)
            keep = []
            for k, v in sd.items():
                if k.startswith("module."):
                    k = k[7:]
                t = v.detach().to(dtype=torch.float32).contiguous()
                keep.append(t)                      # the copies are asynchronous on `st`: sources stay alive until finalize returns
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(self._lib.ndiff_load_param_async(self._h, k.encode(), _ptr(t), t.dim(), shape, st))
            _lib.check(self._lib.ndiff_finalize_params(self._h, st))    # synchronises `st` before returning
        self._cond_key = None

    # ---- condition --------------------------------------------------------------------------------------------
    def set_condition(self, clean_img: torch.Tensor, position: torch.Tensor, iso_ratio_idx: torch.Tensor):
        B, H, W = self.batch, self.height, self.width
        if tuple(clean_img.shape) != (B, 4, H, W) or tuple(position.shape) != (B, 2, H, W) or iso_ratio_idx.numel() != B:
            raise ValueError(f"condition shapes do not match the engine geometry (B={B}, H={H}, W={W}): "
                             f"{tuple(clean_img.shape)}, {tuple(position.shape)}, {tuple(iso_ratio_idx.shape)}")
        try:
            key = (clean_img.data_ptr(), clean_img._version, position.data_ptr(), position._version,
                   iso_ratio_idx.data_ptr(), iso_ratio_idx._version, self.weights_version)
        except RuntimeError:          # inference tensors carry no version counter: no safe identity, never cached
            key = None
        if key is not None and key == self._cond_key:
            return
        c = _f32c(clean_img, self.device)
        p = _f32c(position, self.device)
        i = iso_ratio_idx.detach().to(device=self.device, dtype=torch.int64).contiguous()
        if int(i.min()) < 0 or int(i.max()) >= 100:
            raise IndexError("iso_ratio_idx out of range for Embedding(100, 16)")
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_set_condition(self._h, _ptr(c), _ptr(p), _ptr(i), _stream(self.device_index)))
        # holding the caller's tensors keeps their storage from being recycled, so (data_ptr, _version) stays a
        # sound identity for the cache above
        self._keep = [c, p, i, clean_img, position, iso_ratio_idx]
        self._cond_key = key

    # ---- one network evaluation ---------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        xx = _f32c(x, self.device)
        tt = time.detach().to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(xx)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_forward(self._h, _ptr(xx), _ptr(tt), _ptr(out), _stream(self.device_index)))
        self._keep_io = [xx, tt]
        return out

    # ---- chain ------------------------------------------------------------------------------------------------
    def chain_begin(self, steps: Sequence[_lib.Step], x_init: Optional[torch.Tensor], seed: int):
        arr = (_lib.Step * len(steps))(*steps)
        xi = _f32c(x_init, self.device) if x_init is not None else None
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_chain_begin(self._h, arr, len(steps), _ptr(xi), C.c_uint64(seed & (2**64 - 1)),
                                                   _stream(self.device_index)))
        self._keep_chain = [xi]

    def chain_run(self, n: int, noise: Optional[torch.Tensor] = None, teacher: Optional[torch.Tensor] = None,
                  snapshots: Optional[torch.Tensor] = None):
        for t in (noise, teacher, snapshots):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                assert tuple(t.shape) == (n, self.batch, 4, self.height, self.width), tuple(t.shape)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_chain_run(self._h, n, _ptr(noise), _ptr(teacher), _ptr(snapshots),
                                                 _stream(self.device_index)))
        self._keep_run = [noise, teacher, snapshots]

    def chain_seek(self, step: int, x: torch.Tensor, seed: int = 0):
        xx = _f32c(x, self.device)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_chain_seek(self._h, int(step), _ptr(xx), C.c_uint64(seed & (2**64 - 1)),
                                                  _stream(self.device_index)))
        self._keep_seek = [xx]

    def chain_read(self) -> torch.Tensor:
        out = torch.empty((self.batch, 4, self.height, self.width), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_chain_read(self._h, _ptr(out), _stream(self.device_index)))
        return out

    def sample_host(self, clean, position, iso_idx, steps: Sequence[_lib.Step], seed: int) -> torch.Tensor:
        """End-to-end with HOST tensors (copies inside), returns a CPU tensor."""
        arr = (_lib.Step * len(steps))(*steps)
        c = clean.detach().to("cpu", torch.float32).contiguous()
        p = position.detach().to("cpu", torch.float32).contiguous()
        i = iso_idx.detach().to("cpu", torch.int64).contiguous()
        out = torch.empty((self.batch, 4, self.height, self.width), dtype=torch.float32)
        _lib.check(self._lib.ndiff_sample_host(self._h, _ptr(c), _ptr(p), _ptr(i), arr, len(steps),
                                               C.c_uint64(seed & (2**64 - 1)), _ptr(out)))
        self._cond_key = None
        return out

    # ---- introspection ------------------------------------------------------------------------------------------
    def debug_tensor(self, name: str) -> torch.Tensor:
        shape = (C.c_int64 * 4)()
        _lib.check(self._lib.ndiff_debug_tensor(self._h, name.encode(), C.c_void_p(0), shape, C.c_void_p(0)))
        out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_debug_tensor(self._h, name.encode(), _ptr(out), shape, _stream(self.device_index)))
        return out

    @property
    def launches_per_step(self) -> int:
        return int(self._lib.ndiff_launches_per_step(self._h))

    @property
    def conv_flops_per_step(self) -> float:
        return float(self._lib.ndiff_conv_flops_per_step(self._h))

    def time_layers(self, iters: int = 5):
        n = C.c_int32()
        _lib.check(self._lib.ndiff_time_layers(self._h, iters, None, None, 0, C.byref(n), C.c_void_p(0)))
        ms = (C.c_float * n.value)()
        names = C.create_string_buffer(64 * 1024)
        with torch.cuda.device(self.device_index):
            _lib.check(self._lib.ndiff_time_layers(self._h, iters, ms, names, len(names), C.byref(n),
                                                   _stream(self.device_index)))
        rows = []
        for line, t in zip(names.value.decode().strip().split("\n"), list(ms)):
            nm, fl, by = line.rsplit(";", 2)
            rows.append((nm, float(t), float(fl), float(by)))      # (name, median ms inside the step, executed FLOPs, algorithmic bytes)
        return rows
