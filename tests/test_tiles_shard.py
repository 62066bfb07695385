"""CPU: tile grid / position maps against the oracle's restatement, and the N>1 sharding logic over a real
world_size-2 gloo process group (no data-path collective exists; the group is used only to cross-check the split)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from noisediff_b200 import tiles
from oracle import noisediff_oracle as O


def test_tile_grid_and_positions_match_oracle():
    for ps in (256, 512):
        assert tiles.tile_origins(ps) == O.tile_origins(ps)
    assert len(tiles.tile_origins(256)) == 88 and len(tiles.tile_origins(512)) == 24
    for (x, y) in [(0, 0), (1872, 1168), (192, 384)]:
        assert torch.equal(tiles.position_map(256, 256, x, y), O.make_position(256, 256, x, y))
    a, b = tiles.synthetic_condition(3, 64, seed=1), O.synthetic_condition(3, 64, 64, seed=1)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_shard_covers_everything_once():
    for n, w in [(4096, 8), (88, 8), (7, 4), (3, 8), (64, 1)]:
        parts = [tiles.shard(n, w, r) for r in range(w)]
        flat = [i for p in parts for i in p]
        assert flat == list(range(n)) and max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        tiles.shard(4, 2, 2)
    seeds = {tiles.rank_seed(7, r, j) for r in range(8) for j in range(64)}
    assert len(seeds) == 8 * 64


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = tiles.shard(n_items, world, rank)
    # each rank "synthesises" its own patches: here the patch payload is its global index and its seed
    payload = torch.full((n_items,), -1, dtype=torch.long)
    for i in mine:
        payload[i] = i
    seeds = torch.zeros(world, dtype=torch.long)
    seeds[rank] = tiles.rank_seed(99, rank)
    dist.all_reduce(payload, op=dist.ReduceOp.MAX)      # test-only cross-check, not part of the data path
    dist.all_reduce(seeds, op=dist.ReduceOp.SUM)
    counts = torch.tensor([len(mine)])
    dist.all_reduce(counts)
    if rank == 0:
        q.put((payload.tolist(), seeds.tolist(), int(counts)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    n_items = 88                                          # one full frame's tile grid (BASELINE configs[3])
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    payload, seeds, count = q.get()
    assert payload == list(range(n_items)) and count == n_items and len(set(seeds)) == 2


class _PatternDiffusion:
    """CPU stand-in for GaussianDiffusion in the multi-process scheduling test (the sampler itself needs the GPU)."""
    image_size = 32
    device = torch.device("cpu")

    def sample(self, batch_size, condition):
        return condition["clean_img"] + condition["position"].sum(1, keepdim=True)


def _frames_worker(rank, world, port, folder, q):
    from noisediff_b200 import frames
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)                     # every rank builds the same job list, as bench_frame.py does
    jobs = [frames.FrameJob(torch.rand((4, 96, 160), generator=g), 3, "a_10s.ARW", "a_0.1s.ARW"),
            frames.FrameJob(torch.rand((4, 64, 96), generator=g), 24, "b_10s.ARW", iso=800, ratio=250)]
    paths = frames.synthesize_frames(_PatternDiffusion(), jobs, save_folder=folder, batch_size=5, rank=dist.get_rank(),
                                     world_size=dist.get_world_size())
    n = torch.tensor([len(paths)])
    dist.all_reduce(n)                                       # test-only: the data path itself has no collective
    dist.barrier()
    if rank == 0:
        q.put((int(n), len(frames.plan_crops(jobs, 32))))
    dist.destroy_process_group()


def test_two_rank_gloo_packed_frame_synthesis(tmp_path):
    """Two real processes synthesise a two-frame job list into one folder: together exactly one file per crop."""
    from noisediff_b200 import frames
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_frames_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    written, planned = q.get()
    files = [os.path.join(d, f) for d, _, fs in os.walk(str(tmp_path)) for f in fs if f.endswith(".npy")]
    assert written == planned == len(files)
    names = sorted(frames.parse_npy_name(f) for f in files)
    assert len(set(names)) == planned and {n[0] for n in names} == {"a_10s", "b_10s"}
    assert sum(1 for f in files if os.path.basename(os.path.dirname(f)) == "ISO800_Ratio250") == len(tiles.tile_origins(32, 64, 96))
