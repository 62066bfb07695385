#!/usr/bin/env python
"""BASELINE.json configs[4]: diffusion training step (forward + backward, batch 32 x 4 x 256 x 256 per GPU) with the NCCL gradient
all-reduce over NVLink at 1/2/4/8 B200.

    python tools/bench_train.py [--batch 32] [--steps 10] [--warmup 3] [--out gpurun_out/train.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py

One step = what the reference's loop does per batch (models/trainer_diffusion.py:176-191): q_sample + forward (per-sample t) +
loss + backward on the library, ONE all-reduce of the flat fp32 gradient buffer (DDP averaging), Adam, weight re-pack, EMA.
Inputs are synthetic and resident on the device; every rank draws its own batch (weak scaling: 32 samples per GPU).  Timed with
CUDA events between barriers, max over ranks; the all-reduce is also timed on its own.  One JSON line (rank 0): evidence for the
N1 row, not the bench.py headline."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--patch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))

    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    import noisediff_b200 as nd
    from noisediff_b200 import tiles, training

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                       # NCCL's banner must not land on stdout
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.manual_seed(0)                    # identical initial weights on every rank (DDP broadcasts rank 0's; same seed here)
    net = nd.NoiseDiffNet(SimpleNamespace(dim=64, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False))
    net = net.eval().requires_grad_(False).to(dev)
    gd = nd.GaussianDiffusion(net, image_size=args.patch, timesteps=1000, beta_schedule="sigmoid2", objective="pred_v").to(dev)
    tr = training.DiffusionTrainer(gd, batch_size=args.batch, lr=1e-4)
    B, S = args.batch, args.patch
    g = torch.Generator(device=dev).manual_seed(1000 + rank)          # every rank trains on its own samples
    cond = {k: v.to(dev) for k, v in tiles.synthetic_condition(B, S, seed=1 + rank, first_tile=rank * B).items()}
    img = torch.randn(B, 4, S, S, generator=g, device=dev) * 0.05

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    losses = []
    for _ in range(max(args.warmup, 1)):
        losses.append(tr.step(img, cond))
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        losses.append(tr.step(img, cond))
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    # the collective alone, on the same flat gradient buffer
    flat = tr._flat(1)
    ar = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ar[0].record()
    for _ in range(10):
        training.allreduce_gradients(flat)
    ar[1].record()
    barrier()
    ar_ms = ar[0].elapsed_time(ar[1]) / 10
    t = torch.tensor([ms, ar_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0]), float(t[1])
    # data-parallel invariant: after all-reduced steps from identical initial weights every rank holds the same parameters
    p = tr._flat(0)
    chk = torch.stack([p.double().sum(), p.double().abs().sum()])
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    fam = tr.time_families() if rank == 0 else None       # (re-runs the last step's launch lists with an event around every launch)
    if rank == 0:
        n_params = int(flat.numel())
        line = {"metric": "diffusion training samples/s (forward + backward + all-reduce + Adam + EMA, 4x%dx%d)" % (S, S),
                "value": world * B / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "bf16 activations / fp32 gradients and optimizer",
                "data": "synthetic", "batch_per_gpu": B, "allreduce_ms": ar_ms, "allreduce_bytes": n_params * 4,
                "allreduce_busbw_gbs": (2.0 * (world - 1) / world) * n_params * 4 / (ar_ms * 1e-3) / 1e9 if world > 1 else None,
                "ranks_hold_identical_parameters": same, "loss_first": losses[0], "loss_last": losses[-1],
                "activation_gib": tr.activation_bytes / 2 ** 30, "launches_fwd_bwd": tr.launches,
                "ms_per_kernel_family": {f"{ps}:{k}": {"launches": n, "ms": round(ms, 3)} for ps, k, n, ms in fam},
                "config": {"workload": "BASELINE configs[4]: NoiseDiffNet dim=64, T=1000 sigmoid2 pred_v, per-sample random t"}}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(line, f)
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
