"""Development probe run on the B200 box: kernel experiments + micro-benchmarks, results to gpurun_out/probe.json.
Not part of the product or the test-suite."""
import json
import math
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_util as G          # noqa: E402
import noisediff_b200 as nd              # noqa: E402
from noisediff_b200 import _lib          # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
res = {}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16).float()


def conv_case(mode, B, H, W, c0, co, tile_w=0, force_nt=0, check=True, iters=20, taps=(1, 1), pad=(0, 0)):
    x = rnd((B, c0, H, W), 1)
    k = 3 if mode == 3 or taps == (3, 3) else 1
    w = rnd((co, c0, k, k), 2, 1.0 / math.sqrt(k * k * c0))
    xb, wp = G.to_nhwc_bf16(x), G.pack_weight(w)
    out = G.conv(mode, xb, wp, co, tile_w=tile_w, force_nt=force_nt, taps=taps, pad=pad)
    r = {}
    if check:
        ref = F.conv2d(x, w, None, padding=k // 2)
        got = G.from_nhwc(out)
        r["rel"] = float((got - ref).norm() / ref.norm())
        r["maxabs"] = float((got - ref).abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib = _lib.lib()
    args = (mode, B, H, W, G.P(xb), c0, None, 0, taps[0], taps[1], pad[0], pad[1], G.P(wp), co, None, None, 0, None, 0,
            None, 0, G.P(out), force_nt, tile_w, G.stream())
    for _ in range(3):
        _lib.check(lib.ndiff_op_conv(*args))
    e0.record()
    for _ in range(iters):
        _lib.check(lib.ndiff_op_conv(*args))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * W * co * c0 * k * k
    r["ms"] = ms
    r["tflops"] = fl / ms / 1e9
    return r


def guarded(name, fn):
    try:
        res[name] = fn()
    except Exception as e:   # noqa: BLE001
        res[name] = {"error": repr(e)[:400], "tb": traceback.format_exc()[-600:]}
    print(name, json.dumps(res[name])[:300], flush=True)
    with open(os.path.join(OUT, "probe.json"), "w") as f:
        json.dump(res, f, indent=1)


def main():
    which = sys.argv[1:] or ["exp", "conv", "net"]
    res["device"] = torch.cuda.get_device_name(0)
    if "exp" in which:
        # single-copy halo experiments: does a 128-B row-shifted UMMA start address swizzle consistently?
        guarded("halo1_abs_64", lambda: conv_case(3, 1, 32, 32, 64, 64, iters=3))
        guarded("halo1_baseoff_64", lambda: conv_case(4, 1, 32, 32, 64, 64, iters=3))
    if "conv" in which:
        for nm, mode, kw in [("halo3", 1, {}), ("direct3", 0, dict(taps=(3, 3), pad=(1, 1))), ("halo1", 3, {})]:
            for (B, H, c0, co) in [(4, 256, 64, 64), (8, 256, 64, 64), (4, 256, 128, 64), (8, 32, 512, 512),
                                   (8, 64, 256, 256), (8, 128, 128, 128)]:
                guarded(f"conv3_{nm}_B{B}_{H}_{c0}_{co}", lambda: conv_case(mode, B, H, H, c0, co, check=(B * H <= 1024), **kw))
        for (B, H, c0, co) in [(4, 256, 64, 64), (4, 256, 64, 128), (4, 256, 128, 64), (8, 32, 512, 1024), (8, 32, 1024, 512)]:
            guarded(f"gemm_B{B}_{H}_{c0}_{co}", lambda: conv_case(0, B, H, H, c0, co, check=False))
        for tw in (8, 16, 32):
            guarded(f"conv3_halo3_tw{tw}", lambda: conv_case(1, 4, 256, 256, 64, 64, tile_w=tw, check=False))
    if "net" in which:
        from tests.util import seeded_net
        from noisediff_b200 import tiles
        import copy
        net = copy.deepcopy(seeded_net()).cuda()
        for B in (2, 4, 8):
            def run(B=B):
                gd = nd.GaussianDiffusion(net, image_size=256, timesteps=1000, beta_schedule="sigmoid2").cuda()
                eng = net.engine_for(B, 256, 256, torch.device("cuda", 0))
                cond = {k: v.cuda() for k, v in tiles.synthetic_condition(B, 256).items()}
                eng.set_condition(cond["clean_img"], cond["position"], cond["iso_ratio_idx"])
                steps = gd.ddpm_steps()
                eng.chain_begin(steps, None, 1)
                eng.chain_run(5)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.chain_run(20)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 20
                x = eng.chain_read()
                r = {"ms_per_step": ms, "us_per_patch_step": ms * 1e3 / B, "patches_per_s_1000steps": B / ms,
                     "finite": bool(torch.isfinite(x).all()), "launches": eng.launches_per_step,
                     "conv_tflops": eng.conv_flops_per_step / ms / 1e9}
                if B in (4, 8):
                    rows = eng.time_layers(5)
                    r["layers"] = [(n, round(t * 1e3, 1), round(f / max(t, 1e-9) / 1e9, 1)) for n, t, f, _b in rows]
                    r["sum_layers_ms"] = sum(r[1] for r in rows)
                return r
            guarded(f"net_B{B}", run)


if __name__ == "__main__":
    main()
