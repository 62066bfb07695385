#!/usr/bin/env python
"""bench.py — headline metric of the NoiseDiff hot path on B200: noise patches/s (4x256x256, full 1000-step reverse chain).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A *step* is one reverse-diffusion timestep (network evaluation + posterior update — one pass of the hot path) over this
rank's batch of 64 synthetic packed-Bayer patches (BASELINE.json configs[1]); the chain has T = 1000 cost-identical steps,
so patches/s = patches / (1000 x seconds per step).  K steps are timed with CUDA events between barriers, max over ranks.
Patches are independent, so ranks shard them with no collective (weak scaling: 64 patches per GPU, per-rank seeds).
`e2e` runs the WHOLE 1000-step chain through the public API with host buffers (pinned host -> device copies of the
condition and device -> host copy of the result inside the timed region).
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_CHAIN = 1000
PATCH = 256
BATCH = 64
METRIC = "noise patches/s (4x256x256, full 1000-step reverse chain)"
UNIT = "patches/s"
# 3x3-conv FLOPs per patch per step.  The reference executes 234.34 GFLOP (direct form); the three
# Upsample convs run here as four 2x2 phase convolutions (12.89 instead of 28.99 GFLOP), and §8(d)'s rule is to count the
# FLOPs actually needed, never more than the reference executes: 234.34 - 28.99 + 12.89.
CONV3_FLOPS_DIRECT = 234.34e9
CONV3_FLOPS_PER_PATCH_STEP = 218.24e9


LIVE_FLOPS_PER_PATCH_STEP = 265.52e9     # 3x3 234.34 + 7x7 1.64 + live 1x1 15.03 + live token-linear 14.50


def is_conv3(name: str, flops: float) -> bool:
    return flops > 0 and (".proj" in name and "block" in name or name.endswith(".3.1") and name.startswith("ups")
                          or name in ("downs.3.3", "ups.3.3"))


def family_of(name: str, flops: float) -> str:
    """Kernel family of one plan row (engine.cu names its ops after the reference's module paths)."""
    if "fused chain" in name:
        return "pixel_chain"
    if name.endswith(".norm") and "block" in name:
        return "gn_apply"
    if name.endswith(".norm2"):
        return "layernorm"
    if name.startswith("init_conv"):
        return "init_conv"
    if name.startswith("final("):
        return "heads_update"
    if name == "step prologue":
        return "prologue"
    return "conv_gemm" if flops > 0 else "other"


def families(rows):
    out = {}
    for n, ms, fl, by in rows:
        d = out.setdefault(family_of(n, fl), {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["n"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += by
    out.setdefault("conv_gemm", {"n": 0, "ms": 1e-9, "flops": 0.0, "bytes": 0.0})
    return out


def plan_signature(rows) -> str:
    """Identity of the layer plan a profile belongs to: op names, FLOPs and algorithmic bytes in launch order."""
    import hashlib
    return hashlib.sha256(json.dumps([[n, round(f), round(b)] for n, _, f, b in rows]).encode()).hexdigest()[:16]


def committed_traffic(rows):
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_gemm launch comes from an `ncu --set full` pass, which cannot run
    inside a timed bench.  The committed summary (profiles/ncu_step_summary.json, written by tools/ncu_step_summary.py) carries the
    signature of the plan it was captured on; a different plan (any kernel / fusion change since) reports null instead of a stale
    number."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_step_summary.json")) as f:
            z = json.load(f)
        if z.get("plan_signature") != plan_signature(rows):
            return None
        return z["conv_gemm"]["dram_bytes_per_launch"]
    except Exception:
        return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained bf16 figure: kernel timed inside a long step)")
    return dict(tflops=1590.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_seconds(n_steps: int, warmup: int, threads: int):
    """The reference's CPU path (restated by oracle/noisediff_oracle.py — the Python reference itself does not travel to the
    GPU box): p_sample steps, B=1, 4x256x256, dim=64, fp32, all host threads."""
    import torch
    from oracle import noisediff_oracle as O
    from tests.util import seeded_sd
    torch.set_num_threads(threads)
    sd = seeded_sd()
    cond = O.synthetic_condition(1, PATCH, PATCH, seed=1)
    tab = O.schedule_tables("sigmoid2", T_CHAIN)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(1, 4, PATCH, PATCH, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_steps):
            t = T_CHAIN - 1 - i
            z = torch.randn(1, 4, PATCH, PATCH, generator=g)
            t0 = time.perf_counter()
            out = O.net_forward(sd, x, torch.full((1,), t, dtype=torch.long), cond)
            x, _ = O.ddpm_step(tab, "pred_v", x, t, out, z)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    times.sort()
    return times[len(times) // 2], sum(times) / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    med, mean = cpu_reference_step_seconds(max(args.steps, 1), max(args.warmup, 1), threads)
    value = 1.0 / (T_CHAIN * mean)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference CPU path: DDPM p_sample steps of ONE 4x256x256 patch (dim=64, sigmoid2, pred_v), "
                               "patches/s extrapolated over the 1000 cost-identical steps", "batch": 1, "timesteps": T_CHAIN},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} timed p_sample steps (B=1) after {max(args.warmup, 1)} warm-up, median {med:.3f} s/step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_split(args, rank, local_rank, world):
    """BASELINE configs[2]: a fixed list of --total-patches patches, split statically (contiguous slices, sizes differ by at most
    one) over the ranks; every rank runs the FULL 1000-step chain of its slice in batches of --batch through the public
    GaussianDiffusion.sample() with per-rank Philox seeds and no collective on the data path.  value = total patches / the slowest
    rank's wall time (strong scaling: total work fixed as N grows)."""
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    import noisediff_b200 as nd
    from noisediff_b200 import tiles

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    torch.manual_seed(0)
    net = nd.NoiseDiffNet(SimpleNamespace(dim=args.dim, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False))
    net = net.eval().requires_grad_(False).to(dev)
    gd = nd.GaussianDiffusion(net, image_size=args.patch, timesteps=T_CHAIN, beta_schedule="sigmoid2", objective="pred_v").to(dev)
    gd.micro_batch, gd.noise_source = args.batch, "philox"
    mine = tiles.shard(args.total_patches, world, rank)
    batches = [list(mine)[lo:lo + args.batch] for lo in range(0, len(mine), args.batch)]
    # warm-up: engine creation, weight packing, graph capture (a short chain on the first batch's geometry)
    warm = nd.GaussianDiffusion(net, image_size=args.patch, timesteps=8, beta_schedule="sigmoid2", objective="pred_v").to(dev)
    warm.micro_batch, warm.noise_source = args.batch, "philox"
    n0 = len(batches[0]) if batches else 0
    if n0:
        c = {k: v.to(dev) for k, v in tiles.synthetic_condition(n0, args.patch, seed=1, first_tile=mine[0]).items()}
        warm.sample(batch_size=n0, condition=c)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.manual_seed(4242 + rank)            # Philox base seeds come from the CPU generator: distinct per rank
    out_std, finite = [], True
    # synthetic inputs of the whole slice, generated ahead of the timed region into pinned host memory (the H2D copies are timed)
    pinned = [{k: v.pin_memory() for k, v in tiles.synthetic_condition(len(part), args.patch, seed=1000 + part[0], first_tile=part[0]).items()}
              for part in batches]
    if world > 1:
        dist.barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for part, cond_host in zip(batches, pinned):
            cond = {k: v.to(dev, non_blocking=True) for k, v in cond_host.items()}
            out = gd.sample(batch_size=len(part), condition=cond)
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            torch.cuda.synchronize(dev)
            finite = finite and bool(torch.isfinite(host).all())
            out_std.append(float(host.std()))
        wall = time.perf_counter() - t0
    tt = torch.tensor([wall], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall_max = float(tt.item())
    fin = torch.tensor([1.0 if finite else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    if rank == 0:
        eng = net.engine_for(min(args.batch, max(n0, 1)), args.patch, args.patch, dev)
        n_steps_total = sum(1 for _ in batches) * T_CHAIN
        line = {
            "metric": METRIC, "value": args.total_patches / wall_max, "unit": UNIT, "n_gpus": world, "steps": n_steps_total,
            "warmup": 8, "ms_per_step": wall_max * 1e3 / max(n_steps_total, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: {args.total_patches} patches of 4x{args.patch}x{args.patch} (dim={args.dim}), full "
                                   f"1000-step DDPM chains, static split over {world} rank(s) in batches of {args.batch}, per-rank Philox "
                                   f"seeds, no collective; wall clock incl. pinned H2D of every condition and D2H of every result",
                       "total_patches": args.total_patches, "patches_rank0": len(mine), "batch": args.batch, "timesteps": T_CHAIN,
                       "parallelism": f"independent shards x{world}, no collective"},
            "e2e": {"value": args.total_patches / wall_max, "unit": UNIT,
                    "h2d_bytes_per_step": args.batch * args.patch * args.patch * (4 + 2) * 4 / T_CHAIN,
                    "d2h_bytes_per_step": args.batch * args.patch * args.patch * 4 * 4 / T_CHAIN, "wall_s": wall_max},
            "gpu_launches": int(eng.launches_per_step * n_steps_total), "clocks": clk.summary(), "finite": bool(fin.item() > 0),
            "out_std_rank0": out_std[:3],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="patches per GPU")
    ap.add_argument("--dim", type=int, default=64, help="model width (64 = BASELINE configs; 48 = the reference's shipped checkpoint)")
    ap.add_argument("--patch", type=int, default=PATCH, help="crop size (256 = BASELINE configs; 512 = the reference's script.sh)")
    ap.add_argument("--total-patches", type=int, default=0,
                    help="BASELINE configs[2]: synthesise this many patches in total (e.g. 4096), split statically over the ranks, "
                         "full 1000-step chains in batches of --batch; reports total patches / max rank wall time (strong scaling)")
    ap.add_argument("--micro-batch", type=int, default=int(os.environ.get("NDIFF_MICRO_BATCH", "64")))
    ap.add_argument("--dump-layers", default=None, help="write the per-layer CUDA-event table (name, ms, flops) to this JSON file")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=40, help="timed CPU reverse steps of the cpu_baseline leg (~0.28 s each on 16 cores)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.total_patches > 0:
        return run_split(args, rank, local_rank, world)
    # FLOP model for other widths / crops: every 3x3 / 1x1 / token-linear layer is C x C (scales with (dim/64)^2), the 7x7 input
    # conv is 4 x C (scales with dim/64); everything scales with the pixel count
    fd, fp = args.dim / 64.0, (args.patch / 256.0) ** 2
    global CONV3_FLOPS_DIRECT, CONV3_FLOPS_PER_PATCH_STEP, LIVE_FLOPS_PER_PATCH_STEP
    CONV3_FLOPS_DIRECT *= fd * fd * fp
    CONV3_FLOPS_PER_PATCH_STEP *= fd * fd * fp
    LIVE_FLOPS_PER_PATCH_STEP = ((234.34e9 + 15.03e9 + 14.50e9) * fd * fd + 1.64e9 * fd) * fp
    PATCH_ = args.patch

    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    import noisediff_b200 as nd
    from noisediff_b200 import tiles

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to the C-level stdout when the communicator comes up; stdout must carry exactly one
        # JSON line, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B, mb = args.batch, min(args.micro_batch, args.batch)
    n_mb = (B + mb - 1) // mb

    torch.manual_seed(0)                                      # reference-identical random init (same RNG consumption)
    net = nd.NoiseDiffNet(SimpleNamespace(dim=args.dim, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False))
    net = net.eval().requires_grad_(False).to(dev)
    gd = nd.GaussianDiffusion(net, image_size=PATCH_, timesteps=T_CHAIN, beta_schedule="sigmoid2", objective="pred_v").to(dev)
    gd.micro_batch, gd.noise_source = mb, "philox"
    cond_cpu = tiles.synthetic_condition(B, PATCH_, seed=1 + rank, first_tile=rank * B)
    cond_pinned = {k: v.pin_memory() for k, v in cond_cpu.items()}
    steps = gd.ddpm_steps()
    eng = net.engine_for(mb, PATCH_, PATCH_, dev)

    # ---- device-resident timing: K reverse steps over the whole batch (n_mb micro-batches per step) -------------------
    conds = [{k: v[i * mb:(i + 1) * mb].to(dev) for k, v in cond_cpu.items()} for i in range(n_mb)]
    states = [None] * n_mb

    def run_steps(first, n):
        """n consecutive reverse steps for every micro-batch, scheduled as GaussianDiffusion.sample() does it: one
        micro-batch at a time, its condition set once per chunk of steps."""
        for j in range(n_mb):
            c = conds[j]
            eng.set_condition(c["clean_img"], c["position"], c["iso_ratio_idx"])
            if states[j] is None:
                eng.chain_begin(steps, None, tiles.rank_seed(2024, rank, j))
            else:
                eng.chain_seek(first, states[j], tiles.rank_seed(2024, rank, j))
            eng.chain_run(n)
            states[j] = eng.chain_read()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run_steps(0, W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        run_steps(W, K)
        e1.record()
        barrier()
    ms_step = e0.elapsed_time(e1) / K
    t = torch.tensor([ms_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    value = world * B / (T_CHAIN * ms_step * 1e-3)
    finite = bool(torch.isfinite(states[0]).all())

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv): every launch timed INSIDE the running step ----------
    # eng.time_layers enqueues whole chain steps with a CUDA event between consecutive ops and returns the per-op median, so each
    # kernel runs behind its real predecessor at the clocks of the long-running chain: the denominator is the SUSTAINED bf16 peak
    # of MEASURED_PEAKS.json (the burst figure is reported next to it).
    peaks = _peaks()
    rows = eng.time_layers(7)
    if args.dump_layers and rank == 0:
        with open(args.dump_layers, "w") as f:
            json.dump({"micro_batch": mb, "ms_per_step": ms_step, "plan_signature": plan_signature(rows),
                       "layers": [list(r) for r in rows]}, f)
    fam = families(rows)
    conv_ms, conv_fl, n_conv = fam["conv_gemm"]["ms"], fam["conv_gemm"]["flops"], fam["conv_gemm"]["n"]
    executed_conv_fl = conv_fl
    conv_fl *= fd * fd          # narrower models run zero-padded in the 64-channel kernels: count the reference's FLOPs, not the padding
    c3_ms = sum(t_ for n, t_, f, _ in rows if is_conv3(n, f))
    layers_ms = sum(r[1] for r in rows)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12
    hbm = {k: {"launches": v["n"], "ms": round(v["ms"], 4), "algorithmic_gb": round(v["bytes"] / 1e9, 4),
               "achieved_gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
               "frac_of_hbm_peak": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm"], 4) if v["ms"] > 0 else None}
           for k, v in fam.items() if k != "conv_gemm"}
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": committed_traffic(rows),
                "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, all 3x3/1x1/2x2s2 launches of one step)",
                "timing": "median of 7 in-step iterations, CUDA events between consecutive ops of the running chain step",
                "launches_per_step": n_conv, "avg_launch_us": conv_ms * 1e3 / max(n_conv, 1),
                "flops_per_launch_avg": conv_fl / max(n_conv, 1), "executed_flops_per_launch_avg": executed_conv_fl / max(n_conv, 1),
                "peak_source": peaks["source"],
                "frac_of_burst_peak": achieved / peaks["burst"], "burst_peak": peaks["burst"],
                "conv3x3_frac_of_sustained_peak": CONV3_FLOPS_PER_PATCH_STEP * mb / (c3_ms * 1e-3) / 1e12 / peaks["tflops"],
                "conv3x3_frac_of_burst_peak": CONV3_FLOPS_PER_PATCH_STEP * mb / (c3_ms * 1e-3) / 1e12 / peaks["burst"],
                "conv3x3_frac_of_burst_peak_direct_form_flops": CONV3_FLOPS_DIRECT * mb / (c3_ms * 1e-3) / 1e12 / peaks["burst"],
                "conv_share_of_step": conv_ms * n_mb / ms_step if world == 1 else None,
                "layers_ms_sum": layers_ms, "conv_ms": conv_ms, "conv3x3_ms": c3_ms,
                # whole step against the tensor roofline: SURVEY §8(d)'s live dense FLOPs (265.52 GFLOP per patch-step) over the
                # device-timed step (graph replay), sustained and burst
                "whole_step": {"flops_per_patch_step": LIVE_FLOPS_PER_PATCH_STEP,
                               "achieved_tflops": LIVE_FLOPS_PER_PATCH_STEP * B / (ms_step * 1e-3) / 1e12,
                               "frac_of_sustained_peak": LIVE_FLOPS_PER_PATCH_STEP * B / (ms_step * 1e-3) / 1e12 / peaks["tflops"],
                               "frac_of_burst_peak": LIVE_FLOPS_PER_PATCH_STEP * B / (ms_step * 1e-3) / 1e12 / peaks["burst"]},
                "hbm_families": hbm, "hbm_peak_gbs": peaks["hbm"]}

    # ---- end to end through the public API with host buffers: the WHOLE chain --------------------------------------------
    e2e = None
    if not args.no_e2e:
        barrier()
        torch.manual_seed(1234 + rank)
        t0 = time.perf_counter()
        cond_dev = {k: v.to(dev, non_blocking=True) for k, v in cond_pinned.items()}       # H2D from pinned host memory
        out = gd.sample(batch_size=B, condition=cond_dev)                                    # 1000 graph replays / micro-batch
        out_host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        out_host.copy_(out, non_blocking=True)                                               # D2H of the result
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall = float(tt.item())
        h2d = sum(v.numel() * v.element_size() for v in cond_pinned.values())
        d2h = out_host.numel() * out_host.element_size()
        e2e = {"value": world * B / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / T_CHAIN, "d2h_bytes_per_step": d2h / T_CHAIN,
               "wall_s": wall, "note": "full 1000-step chain of the batch via GaussianDiffusion.sample(); bytes are per chain / 1000",
               "finite": bool(torch.isfinite(out_host).all()), "out_std": float(out_host.std())}

    # ---- the reference's CPU path on this box's host cores (reported baseline) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        med, mean = cpu_reference_step_seconds(args.cpu_steps, 2, threads)
        cpu = {"value": 1.0 / (T_CHAIN * mean), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_steps} timed reverse steps of one 4x256x256 patch (after 2 warm-up), {mean:.3f} s/step, extrapolated x1000"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"NoiseDiffNet dim={args.dim} random-init, DDPM T=1000 sigmoid2 pred_v, {B} patches of 4x{PATCH_}x{PATCH_} per GPU "
                                   f"({'BASELINE configs[1]' if (args.dim, PATCH_, B) == (64, 256, 64) else 'NOT the BASELINE config: width / crop / batch overridden'}); one step = one reverse timestep over the batch in micro-batches of {mb}",
                       "batch_per_gpu": B, "micro_batch": mb, "timesteps": T_CHAIN, "parallelism": f"independent shards x{world}, no collective",
                       "l2": "per-step activation working set (GBs) far exceeds the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(eng.launches_per_step * n_mb * K),
            "clocks": clk.summary(), "finite": finite,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
