"""Tile-grid producer and .npy writer of full-frame synthesis: host logic on CPU, the end-to-end path on a
GPU.  Reference behaviour: dataloader/dataset.py:203-281 (grid, crops, position maps), models/trainer_diffusion.py:296-317
(file names, float32 (4, ps, ps) arrays)."""
import os

import numpy as np
import pytest
import torch

from noisediff_b200 import frames, tiles


def _ref_grid(h, w, ps):
    """dataset.py:203-219 restated with numpy, verbatim logic."""
    step = ps - ps // 4
    hs = np.arange(0, h - ps + 1, step)
    if h - (hs[-1] + ps) < ps:
        hs = np.append(hs, h - ps)
    ws = np.arange(0, w - ps + 1, step)
    if w - (ws[-1] + ps) < ps:
        ws = np.append(ws, w - ps)
    return [(int(x), int(y)) for y in hs for x in ws]


@pytest.mark.parametrize("ps,count", [(256, 88), (512, 24)])
def test_tile_grid_matches_reference(ps, count):
    got = tiles.tile_origins(ps)
    assert got == _ref_grid(tiles.FULL_H, tiles.FULL_W, ps) and len(got) == count


def test_crop_batch_and_names():
    g = torch.Generator().manual_seed(0)
    frame = torch.rand((4, 96, 160), generator=g)
    origins = tiles.tile_origins(32, 96, 160)
    assert origins == _ref_grid(96, 160, 32)
    cond = frames.crop_batch(frame, origins[:5], 32, 24)
    assert cond["clean_img"].shape == (5, 4, 32, 32) and cond["position"].shape == (5, 2, 32, 32)
    x, y = origins[3]
    assert torch.equal(cond["clean_img"][3], frame[:, y:y + 32, x:x + 32])
    # utils/util.py:138-147 make_coord(H, W, rescale=True): row / (H - 1), col / (W - 1)
    assert torch.allclose(cond["position"][3, 0, :, 0], torch.arange(y, y + 32).float() / 95)
    assert torch.allclose(cond["position"][3, 1, 0, :], torch.arange(x, x + 32).float() / 159)
    assert cond["iso_ratio_idx"].tolist() == [24] * 5
    assert frames.npy_name("00001_00_10s.ARW", 192, 1168) == "00001_00_10s+00001_00_10s+192_1168.npy"
    assert frames.npy_name("a.ARW", 0, 0, noisy_name="b.ARW") == "a+b+0_0.npy"


def test_consumer_contract():
    """dataset_denoising.py:47-59,136-153: folder / file-name parsing and the noise + clean composition."""
    assert frames.consumer_subfolder(800, 250) == "ISO800_Ratio250"
    n = frames.npy_name("00001_00_10s.ARW", 192, 1168, noisy_name="00001_00_0.04s.ARW")
    assert frames.parse_npy_name("/x/" + n) == ("00001_00_10s", "00001_00_0.04s", 192, 1168)
    g = torch.Generator().manual_seed(3)
    clean = torch.rand((4, 8, 8), generator=g) * 1.2 - 0.1
    noise = torch.randn((4, 8, 8), generator=g) * 0.8
    c, nz = frames.compose_noisy(clean, noise)
    want = np.clip(np.clip(noise.numpy(), -1.0, 1.0).astype(np.float32) + clean.numpy(), 0.0, 1.0)
    assert np.array_equal(nz.numpy(), want) and np.array_equal(c.numpy(), np.clip(clean.numpy(), 0.0, 1.0))


class _FakeDiffusion:
    """Stands in for GaussianDiffusion on CPU: 'noise' = clean + position-dependent pattern, so files are checkable."""
    image_size = 32
    device = torch.device("cpu")

    def sample(self, batch_size, condition):
        assert batch_size == condition["clean_img"].shape[0]
        return condition["clean_img"] + condition["position"].sum(1, keepdim=True)


def test_two_ranks_cover_the_frame_once(tmp_path):
    frame = torch.rand((4, 96, 160), generator=torch.Generator().manual_seed(1))
    origins = tiles.tile_origins(32, 96, 160)
    all_paths = []
    for rank in range(2):
        all_paths += frames.synthesize_frame(_FakeDiffusion(), frame, iso_ratio_idx=3, clean_name="f.ARW", save_folder=str(tmp_path),
                                             batch_size=4, rank=rank, world_size=2)
    names = sorted(os.path.basename(p) for p in all_paths)
    assert names == sorted(frames.npy_name("f.ARW", x, y) for x, y in origins) and len(set(names)) == len(origins)
    x, y = origins[-1]
    arr = np.load(os.path.join(str(tmp_path), "npy", "generated", frames.npy_name("f.ARW", x, y)))
    assert arr.dtype == np.float32 and arr.shape == (4, 32, 32)
    want = frames.crop_batch(frame, [(x, y)], 32, 3)
    assert np.allclose(arr, (want["clean_img"] + want["position"].sum(1, keepdim=True))[0].numpy())
    # consumer layout: <save_folder>/ISO{iso}_Ratio{ratio}/
    p2 = frames.synthesize_frame(_FakeDiffusion(), frame, iso_ratio_idx=3, clean_name="f.ARW", save_folder=str(tmp_path), batch_size=64,
                                 iso=800, ratio=250)
    assert len(p2) == len(origins) and all(os.path.dirname(p).endswith("ISO800_Ratio250") for p in p2)


def test_writer_surfaces_io_errors_and_rejects_mismatched_names(tmp_path):
    """A failed np.save in the background thread must not be lost: close() re-raises it."""
    blocker = tmp_path / "not_a_dir"
    blocker.write_text("x")
    w = frames.NpyWriter(str(tmp_path))
    with pytest.raises(ValueError):
        w.submit(torch.zeros(2, 4, 8, 8), ["only_one.npy"])
    w.submit(torch.zeros(1, 4, 8, 8), [os.path.join("not_a_dir", "a.npy")])          # parent "directory" is a file
    with pytest.raises(OSError):
        w.close()
    ok = frames.NpyWriter(str(tmp_path / "fine"))
    ok.submit(torch.ones(2, 4, 8, 8), ["a.npy", os.path.join("sub", "b.npy")])
    paths = ok.close()
    assert [os.path.relpath(p, str(tmp_path / "fine")) for p in paths] == ["a.npy", os.path.join("sub", "b.npy")]
    assert np.load(paths[1]).dtype == np.float32 and float(np.load(paths[1]).mean()) == 1.0


def _jobs():
    g = torch.Generator().manual_seed(5)
    return [frames.FrameJob(torch.rand((4, 96, 160), generator=g), 3, "a_10s.ARW", "a_0.1s.ARW"),
            frames.FrameJob(torch.rand((4, 64, 64), generator=g), 24, "b_10s.ARW", iso=800, ratio=250),
            frames.FrameJob(torch.rand((4, 96, 96), generator=g), 7, "c_10s.ARW", dark_frame=True, iso=1600, ratio=100)]


def test_packed_multi_frame_synthesis_covers_every_crop_once(tmp_path):
    """Several frames as ONE crop list: batches span frames, ranks take contiguous slices, every crop is written once with the
    name / folder / content a per-frame run gives it."""
    jobs = _jobs()
    plan = frames.plan_crops(jobs, 32)
    per_frame = [len(tiles.tile_origins(32, *j.clean_frame.shape[1:])) for j in jobs]
    assert len(plan) == sum(per_frame) and [p[0] for p in plan] == sum(([i] * n for i, n in enumerate(per_frame)), [])
    for world in (1, 2, 3):
        root = tmp_path / f"w{world}"
        paths = []
        for rank in range(world):
            paths += frames.synthesize_frames(_FakeDiffusion(), jobs, save_folder=str(root), batch_size=7, rank=rank, world_size=world)
        assert len(paths) == len(plan) == len(set(paths))
        for idx, (j, x, y) in enumerate(plan):
            job = jobs[j]
            sub = frames.consumer_subfolder(job.iso, job.ratio) if job.iso else os.path.join("npy", "generated")
            # dark frames: Trainer.test's running item counter (here: the crop's index in the whole plan) + ISO + ratio
            name = frames.dark_npy_name(idx, job.iso, job.ratio, x, y) if job.dark_frame else frames.npy_name(job.clean_name, x, y, job.noisy_name)
            arr = np.load(os.path.join(str(root), sub, name))
            want = frames.crop_batch(job.clean_frame, [(x, y)], 32, job.iso_ratio_idx)
            clean = torch.zeros_like(want["clean_img"]) if job.dark_frame else want["clean_img"]
            assert arr.dtype == np.float32 and np.allclose(arr, (clean + want["position"].sum(1, keepdim=True))[0].numpy())
    # a single job through the packed path writes what synthesize_frame writes
    one = frames.synthesize_frames(_FakeDiffusion(), jobs[:1], save_folder=str(tmp_path / "p"), batch_size=5)
    ref = frames.synthesize_frame(_FakeDiffusion(), jobs[0].clean_frame, iso_ratio_idx=3, clean_name="a_10s.ARW", noisy_name="a_0.1s.ARW",
                                  save_folder=str(tmp_path / "q"), batch_size=64)
    assert [os.path.basename(p) for p in one] == [os.path.basename(p) for p in ref]
    assert all(np.array_equal(np.load(a), np.load(b)) for a, b in zip(one, ref))


@pytest.mark.gpu
def test_packed_multi_frame_synthesis_on_gpu(tmp_path):
    """Two frames in one batch on the CUDA path: identical to sampling the concatenated condition directly."""
    import copy
    import noisediff_b200 as nd
    from tests.util import seeded_net
    net = copy.deepcopy(seeded_net()).cuda()
    gd = nd.GaussianDiffusion(net, image_size=32, timesteps=5, beta_schedule="sigmoid2", objective="pred_v").cuda()
    gd.noise_source = "philox"
    g = torch.Generator().manual_seed(6)
    jobs = [frames.FrameJob(torch.rand((4, 64, 64), generator=g) * 0.3, 24, "u.ARW"),
            frames.FrameJob(torch.rand((4, 64, 96), generator=g) * 0.3, 3, "v.ARW", iso=800, ratio=250)]
    plan = frames.plan_crops(jobs, 32)
    torch.manual_seed(21)
    paths = frames.synthesize_frames(gd, jobs, save_folder=str(tmp_path), batch_size=len(plan))
    assert len(paths) == len(plan)
    conds = [frames.crop_batch(jobs[j].clean_frame.cuda(), [(x, y)], 32, jobs[j].iso_ratio_idx) for j, x, y in plan]
    cond = {k: torch.cat([c[k] for c in conds]) for k in conds[0]}
    torch.manual_seed(21)
    direct = gd.sample(batch_size=len(plan), condition=cond).cpu().numpy()
    for i, p in enumerate(paths):
        arr = np.load(p)
        assert arr.shape == (4, 32, 32) and np.isfinite(arr).all() and np.array_equal(arr, direct[i])
    assert os.path.dirname(paths[-1]).endswith("ISO800_Ratio250") and os.path.dirname(paths[0]).endswith(os.path.join("npy", "generated"))


@pytest.mark.gpu
def test_frame_synthesis_equals_direct_sampling(tmp_path):
    import copy
    import noisediff_b200 as nd
    from tests.util import seeded_net
    net = copy.deepcopy(seeded_net()).cuda()
    gd = nd.GaussianDiffusion(net, image_size=32, timesteps=6, beta_schedule="sigmoid2", objective="pred_v").cuda()
    gd.noise_source = "philox"
    frame = torch.rand((4, 64, 96), generator=torch.Generator().manual_seed(2)) * 0.3
    origins = tiles.tile_origins(32, 64, 96)
    torch.manual_seed(11)
    paths = frames.synthesize_frame(gd, frame, iso_ratio_idx=24, clean_name="g.ARW", save_folder=str(tmp_path), batch_size=len(origins))
    assert len(paths) == len(origins)
    torch.manual_seed(11)
    direct = gd.sample(batch_size=len(origins), condition=frames.crop_batch(frame.cuda(), origins, 32, 24)).cpu().numpy()
    for i, (x, y) in enumerate(origins):
        arr = np.load(os.path.join(str(tmp_path), "npy", "generated", frames.npy_name("g.ARW", x, y)))
        assert arr.shape == (4, 32, 32) and np.isfinite(arr).all()
        assert np.array_equal(arr, direct[i])


def test_dark_frame_names_follow_trainer_test(tmp_path):
    """models/trainer_diffusion.py:318-322: '%05d' % npy_num + '_' + iso + '_' + ratio + '+' + 'x_y' + '.npy', zero clean image."""
    assert frames.dark_npy_name(7, 800.0, 250.0, 192, 1168) == "00007_800_250+192_1168.npy"
    frame = torch.rand((4, 64, 96), generator=torch.Generator().manual_seed(4))
    origins = tiles.tile_origins(32, 64, 96)
    with pytest.raises(ValueError):
        frames.synthesize_frame(_FakeDiffusion(), frame, iso_ratio_idx=3, clean_name="f.ARW", save_folder=str(tmp_path), dark_frame=True)
    paths = []
    for rank in range(2):
        paths += frames.synthesize_frame(_FakeDiffusion(), frame, iso_ratio_idx=3, clean_name="f.ARW", save_folder=str(tmp_path), batch_size=5,
                                         rank=rank, world_size=2, dark_frame=True, iso=800, ratio=250, index_base=100)
    assert [os.path.basename(p) for p in paths] == [frames.dark_npy_name(100 + i, 800, 250, x, y) for i, (x, y) in enumerate(origins)]
    x, y = origins[4]
    want = frames.crop_batch(frame, [(x, y)], 32, 3)["position"].sum(1, keepdim=True)[0].expand(4, 32, 32)      # clean image is zero
    assert np.allclose(np.load(paths[4]), want.numpy())


class _NoisyFakeDiffusion(_FakeDiffusion):
    """'noise' really drawn from the generators a GaussianDiffusion would use (CPU here): torch.randn for the torch mode and a
    torch.randint base seed for the Philox mode."""
    def sample(self, batch_size, condition):
        base = int(torch.randint(0, 2 ** 62, (1,)).item())
        return torch.randn((batch_size, 4, 32, 32)) + (base % 1000) * 1e-6


def test_identically_seeded_ranks_draw_independent_noise(tmp_path):
    """ranks seeded identically (usual DDP practice) must not synthesise the same noise for their crops.  The rank is
    folded into the stream inside synthesize_frame(s) (frames.rank_noise_stream); a single-rank run is untouched; the caller's
    generator advances identically on every rank."""
    frame = torch.zeros((4, 32, 64))                     # overlapping grid (x = 0, 24, 32; the reference repeats the border row): each rank gets a few crops
    outs, states = [], []
    for rank in range(2):
        torch.manual_seed(1234)                          # every rank seeds the same way
        p = frames.synthesize_frame(_NoisyFakeDiffusion(), frame, iso_ratio_idx=3, clean_name=f"r{rank}.ARW", save_folder=str(tmp_path),
                                    batch_size=4, rank=rank, world_size=2)
        states.append(torch.rand(1))
        assert len(p) >= 1
        outs.append(np.load(p[0]))
    assert not np.allclose(outs[0], outs[1])             # different noise on the two shards
    assert np.corrcoef(outs[0].ravel(), outs[1].ravel())[0, 1] < 0.1
    assert torch.equal(states[0], states[1])             # the callers' generators stay in lock-step
    # packed multi-frame path: same property
    outs = []
    for rank in range(2):
        torch.manual_seed(1234)
        p = frames.synthesize_frames(_NoisyFakeDiffusion(), [frames.FrameJob(frame, 3, "m.ARW")], save_folder=str(tmp_path / "m"),
                                     batch_size=4, rank=rank, world_size=2)
        outs.append(np.load(p[0]))
    assert np.corrcoef(outs[0].ravel(), outs[1].ravel())[0, 1] < 0.1
    # world_size == 1: exactly what a direct sample() call draws
    torch.manual_seed(77)
    p = frames.synthesize_frame(_NoisyFakeDiffusion(), frame[:, :, :32], iso_ratio_idx=3, clean_name="s.ARW", save_folder=str(tmp_path / "s"))
    torch.manual_seed(77)
    want = _NoisyFakeDiffusion().sample(4, None)        # a 32 x 32 frame is the same crop four times in the reference's grid
    assert len(p) == 4 and np.array_equal(np.load(p[0]), want[3].numpy())      # (the repeats overwrite one file: the last one stays)


@pytest.mark.gpu
def test_compose_noisy_on_the_gpu_equals_the_consumers_numpy_lines():
    """N4: the consumer's composition (dataset_denoising.py:140-144) as one library kernel on the generated batch."""
    g = torch.Generator().manual_seed(3)
    clean = torch.rand((5, 4, 32, 32), generator=g) * 1.2 - 0.1            # some values outside [0, 1]
    noise = torch.randn((5, 4, 32, 32), generator=g) * 0.8
    c_gpu, n_gpu = frames.compose_noisy(clean.cuda(), noise.cuda())
    want_noisy = np.clip(np.clip(noise.numpy(), -1.0, 1.0).astype(np.float32) + clean.numpy(), 0.0, 1.0)
    assert n_gpu.is_cuda and np.array_equal(n_gpu.cpu().numpy(), want_noisy)
    assert np.array_equal(c_gpu.cpu().numpy(), np.clip(clean.numpy(), 0.0, 1.0))
