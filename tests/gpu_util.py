"""GPU-side helpers for the parity tests: call the C ABI's single-operator entry points on torch buffers."""
from __future__ import annotations

import ctypes as C

import torch

from noisediff_b200 import _lib

MODE_DIRECT, MODE_S2D, MODE_HALO1, MODE_HALO2, MODE_HALO_UP = 0, 2, 3, 4, 5


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_nhwc_bf16(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def from_nhwc(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def pack_weight(w: torch.Tensor, s2d: bool = False) -> torch.Tensor:
    """(Cout, Cin, kh, kw) fp32 -> bf16 [Cout][Cin/64][taps][64]  (the library's K order).  For the space-to-depth conv
    the stored weight is (Cout, 4*C, 1, 1) with input channel = c*4 + p1*2 + p2 (ref Diffusion_arch.py:78-82)."""
    co = w.shape[0]
    if s2d:
        c = w.shape[1] // 4
        wt = w.reshape(co, c // 64, 64, 4).permute(0, 1, 3, 2)
    else:
        ci, kh, kw = w.shape[1:]
        wt = w.reshape(co, ci // 64, 64, kh * kw).permute(0, 1, 3, 2)
    return wt.contiguous().to(torch.bfloat16)


def pack_upconv_weight(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, 3, 3) fp32 -> bf16 [Cout][Cin/64][phase(4)][tap(4)][64]: nearest-x2 upsample + conv3x3 (ref
    Diffusion_arch.py:72-76) as four 2x2 phase convolutions on the low-resolution input with pre-summed taps."""
    co, ci = w.shape[:2]
    sets = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = torch.zeros(co, ci // 64, 4, 4, 64, device=w.device)
    for py in range(2):
        for px in range(2):
            for dy in range(2):
                for dx in range(2):
                    ws = w[:, :, sets[py][dy]][:, :, :, sets[px][dx]].sum((2, 3))
                    out[:, :, py * 2 + px, dy * 2 + dx, :] = ws.reshape(co, ci // 64, 64)
    return out.contiguous().to(torch.bfloat16)


def conv(mode, src0, w_packed, cout, *, src1=None, taps=(1, 1), pad=(0, 0), bias=None, vec=None, res=None, act=0,
         stats=None, groups=0, force_nt=0, tile_w=0, out_hw=None, alloc_hw=None):
    """src*: bf16 NHWC.  Returns bf16 NHWC output."""
    B, H, W, C0 = src0.shape
    Ho, Wo = out_hw if out_hw else (H, W)
    Ha, Wa = alloc_hw if alloc_hw else (Ho, Wo)
    out = torch.empty((B, Ha, Wa, cout), dtype=torch.bfloat16, device=src0.device)
    _lib.check(_lib.lib().ndiff_op_conv(
        mode, B, Ho, Wo, P(src0), C0, P(src1), src1.shape[3] if src1 is not None else 0, taps[0], taps[1], pad[0], pad[1],
        P(w_packed), cout, P(bias), P(vec), vec.shape[1] if vec is not None else 0, P(res), act, P(stats), groups,
        P(out), force_nt, tile_w, stream()))
    torch.cuda.synchronize()
    return out
