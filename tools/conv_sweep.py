"""Times every 3x3-conv shape of the dim=64 network at micro-batch B under each kernel variant (halo1 / halo2 tiles, N tile
64 / 128) through ndiff_op_conv_time; results to gpurun_out/conv_sweep.json.  Development aid for the variant table in
engine.cu (not part of the product or the test-suite).

    python tools/conv_sweep.py [B] [case_index mode nt]     # the 3-argument form runs ONE variant (for ncu)
"""
import ctypes as C
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_util as G          # noqa: E402
from noisediff_b200 import _lib          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
# (H, C0, C1, Cout, groups, kind)
CASES = [(256, 64, 0, 64, 8, "c3"), (256, 64, 64, 64, 8, "c3"), (128, 64, 0, 64, 8, "c3"), (128, 128, 64, 128, 8, "c3"),
         (128, 128, 0, 128, 8, "c3"), (64, 128, 0, 128, 8, "c3"), (64, 256, 128, 256, 8, "c3"), (64, 256, 0, 256, 8, "c3"),
         (32, 256, 0, 256, 8, "c3"), (32, 256, 0, 512, 0, "c3"), (32, 512, 0, 512, 8, "c3"), (32, 512, 256, 512, 8, "c3"),
         (32, 512, 0, 256, 0, "up"), (64, 256, 0, 128, 0, "up"), (128, 128, 0, 64, 0, "up")]


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def run(case, mode, nt, iters=20):
    H, c0, c1, co, groups, kind = case
    x0 = rnd((B, H, H, c0), 1)
    x1 = rnd((B, H, H, c1), 2) if c1 else None
    if kind == "up":
        w = rnd((co, (c0 + c1) // 64, 16, 64), 3, 0.05)
        out = torch.empty((B, 2 * H, 2 * H, co), dtype=torch.bfloat16, device="cuda")
        flops = 2.0 * B * 4 * H * H * co * 4 * c0
        flops_direct = 2.0 * B * 4 * H * H * co * 9 * c0
    else:
        w = rnd((co, (c0 + c1) // 64, 9, 64), 3, 0.05)
        out = torch.empty((B, H, H, co), dtype=torch.bfloat16, device="cuda")
        flops = flops_direct = 2.0 * B * H * H * co * 9 * (c0 + c1)
    bias = torch.zeros(co, device="cuda")
    stats = torch.zeros((B, max(groups, 1), 2), dtype=torch.int64, device="cuda") if groups else None
    ms = C.c_float(0)
    _lib.check(_lib.lib().ndiff_op_conv_time(mode, B, H, H, G.P(x0), c0, G.P(x1), c1, 3, 3, 1, 1, G.P(w), co, G.P(bias),
                                             G.P(stats), groups, G.P(out), nt, 0, iters, C.byref(ms), G.stream()))
    torch.cuda.synchronize()
    return {"ms": ms.value, "tflops": flops / ms.value / 1e9, "tflops_direct": flops_direct / ms.value / 1e9}


def main():
    if len(sys.argv) > 4:
        case = CASES[int(sys.argv[2])]
        print(case, run(case, int(sys.argv[3]), int(sys.argv[4]), iters=2))
        return
    res = {}
    for case in CASES:
        H, c0, c1, co, groups, kind = case
        modes = [G.MODE_HALO_UP] if kind == "up" else [G.MODE_HALO1, G.MODE_HALO2]
        for mode in modes:
            for nt in ([64, 128] if co % 128 == 0 else [64]):
                key = f"{kind}_{H}_{c0}+{c1}->{co}_mode{mode}_nt{nt}"
                try:
                    res[key] = run(case, mode, nt)
                except Exception as ex:       # noqa: BLE001
                    res[key] = {"error": str(ex)[:200]}
                print(key, res[key], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "conv_sweep.json"), "w") as f:
        json.dump({"B": B, "results": res}, f, indent=1)


if __name__ == "__main__":
    main()
