#!/usr/bin/env python
"""BASELINE.json configs[3]: full-frame Sony SID 4x1424x2128 packed-raw noise synthesis, tiled with overlap.

    python tools/bench_frame.py [--ps 256] [--batch 0] [--timesteps 1000] [--out gpurun_out/frame.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_frame.py

One synthetic frame in pinned host memory -> this rank's share of the reference's crop grid (88 crops at ps = 256,
dataloader/dataset.py:203-219) -> full DDPM chain per crop -> ``<clean>+<noisy>+<x>_<y>.npy`` files on disk
(models/trainer_diffusion.py:296-317), through ``noisediff_b200.frames.synthesize_frame``.  Wall-clocked from the host frame
to the last file closed, max over ranks; no collective on the data path.  Prints one JSON line (rank 0); this is evidence
for the N2 row, not the bench.py headline.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ps", type=int, default=256)
    ap.add_argument("--batch", type=int, default=0, help="crops per sample() call; 0 = equal batches of at most 64")
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--frames", type=int, default=1, help="> 1: pack the crops of several frames into full batches (synthesize_frames)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))

    import numpy as np
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    import noisediff_b200 as nd
    from noisediff_b200 import frames, tiles

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = nd.NoiseDiffNet(SimpleNamespace(dim=64, cond_dim=4, inp_dim=4, self_condition=False, normalize_condition=False))
    net = net.eval().requires_grad_(False).to(dev)
    gd = nd.GaussianDiffusion(net, image_size=args.ps, timesteps=args.timesteps, beta_schedule="sigmoid2", objective="pred_v").to(dev)
    gd.noise_source = "philox"
    origins = tiles.tile_origins(args.ps)
    mine = len(tiles.shard(len(origins) * args.frames, world, rank))
    # default: equal batches no larger than the bench geometry (64 crops per engine): 88 crops on one GPU -> 2 x 44
    batch = args.batch or -(-mine // -(-mine // 64))
    gd.micro_batch = batch
    frame = (torch.rand((4, tiles.FULL_H, tiles.FULL_W), generator=torch.Generator().manual_seed(7)) * 0.3).pin_memory()
    folder = tempfile.mkdtemp(prefix=f"ndiff_frame_r{rank}_")

    # warm-up: engine creation, weight packing, graph capture (a short chain on the same geometry)
    gd_w = nd.GaussianDiffusion(net, image_size=args.ps, timesteps=4, beta_schedule="sigmoid2", objective="pred_v").to(dev)
    gd_w.noise_source, gd_w.micro_batch = "philox", batch
    gd_w.sample(batch_size=min(batch, mine), condition=frames.crop_batch(frame.to(dev), origins[:min(batch, mine)], args.ps, 24))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.manual_seed(100 + rank)
    t0 = time.perf_counter()
    if args.frames > 1:
        jobs = [frames.FrameJob(frame, 24, f"synthetic_{i:02d}_10s.ARW", f"synthetic_{i:02d}_0.04s.ARW") for i in range(args.frames)]
        paths = frames.synthesize_frames(gd, jobs, save_folder=folder, batch_size=batch, rank=rank, world_size=world)
    else:
        paths = frames.synthesize_frame(gd, frame, iso_ratio_idx=24, clean_name="synthetic_00_10s.ARW", noisy_name="synthetic_00_0.04s.ARW",
                                        save_folder=folder, batch_size=batch, rank=rank, world_size=world)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    sample = np.load(paths[0])
    ok = len(paths) == mine and sample.shape == (4, args.ps, args.ps) and sample.dtype == np.float32 and bool(np.isfinite(sample).all())
    nbytes = sum(os.path.getsize(p) for p in paths)
    shutil.rmtree(folder, ignore_errors=True)
    if rank == 0:
        line = {"metric": "full-frame synthesis (4x1424x2128 packed raw, overlapping crops, .npy files written)", "n_gpus": world,
                "crops_per_frame": len(origins), "crop": args.ps, "timesteps": args.timesteps, "crops_per_rank": mine, "batch": batch,
                "frames": args.frames, "wall_s": wall, "crops_per_s": len(origins) * args.frames / wall, "frames_per_hour": 3600.0 * args.frames / wall,
                "bytes_written_rank0": nbytes, "files_ok": ok, "first_crop_std": float(sample.std()),
                "note": "host frame (pinned) -> crops + position maps on device -> full chain -> async .npy writer; wall clock, max over ranks"}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(line, f)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
