"""CPU: the host-side scheduling of ``GaussianDiffusion._run_chain`` (micro-batches, chunks of steps, padding of the last
micro-batch, torch-RNG draw order, ``return_all_timesteps``, ``preset_mean``, the DDIM draw rule) — exercised with a stand-in
engine that implements the C ABI's chain calls with a toy per-sample "network" on CPU tensors.

This tests PLUMBING only (what the Python host decides), never the product arithmetic: the real engine is the CUDA library,
and without it / off-GPU ``sample()`` raises (``tests/test_host.py::test_engine_create_fails_loudly_without_gpu``).
Reference behaviour being mirrored: ``p_sample_loop`` / ``ddim_sample`` (models/denoising_diffusion_pytorch.py:375-444).
"""
import pytest
import torch

import noisediff_b200 as nd
from tests.util import net_args


class FakeEngine:
    """ndiff_chain_* semantics (include/noisediff_b200.h) with out = toy(x, condition) instead of the U-Net.  Every sample is
    processed independently, like the real engine, so micro-batching must not change any result."""

    def __init__(self, batch, height, width):
        self.batch, self.height, self.width = batch, height, width
        self.log = []

    def _toy(self, x, t):
        bias = self.iso.float().view(-1, 1, 1, 1) * 0.01 + self.clean.mean(dim=(1, 2, 3), keepdim=True) + self.pos.mean(dim=(1, 2, 3), keepdim=True)
        return torch.tanh(x) * 0.3 + bias + 1e-4 * t

    def set_condition(self, clean, pos, iso):
        assert clean.shape == (self.batch, 4, self.height, self.width) and pos.shape == (self.batch, 2, self.height, self.width)
        self.clean, self.pos, self.iso = clean.clone(), pos.clone(), iso.clone()
        self.log.append("cond")

    def chain_begin(self, steps, x_init, seed):
        self.steps, self.k, self.seed = list(steps), 0, seed
        if x_init is None:                                  # library-side x_T: any deterministic function of the seed
            x_init = torch.randn((self.batch, 4, self.height, self.width), generator=torch.Generator().manual_seed(seed % (2 ** 31)))
        self.x = x_init.clone()
        self.log.append("begin")

    def chain_seek(self, step, x, seed):
        self.k, self.x, self.seed = step, x.clone(), seed
        self.log.append(f"seek{step}")

    def chain_run(self, n, noise=None, teacher=None, snapshots=None):
        for i in range(n):
            s = self.steps[self.k]
            x = teacher[i] if teacher is not None else self.x
            out = self._toy(x, s.t)
            x0 = s.p * x + s.q * out
            if s.clip:
                x0 = x0.clamp(-1, 1)
            eps = (s.r1 * x - x0) / s.r2
            z = noise[i] if noise is not None else torch.zeros_like(x)
            self.x = ((s.a * x0 + s.b * x) + s.c * eps) + s.sigma * z
            if snapshots is not None:
                snapshots[i] = self.x
            self.k += 1
        self.log.append(f"run{n}")

    def chain_read(self):
        return self.x.clone()


@pytest.fixture()
def rig(monkeypatch):
    net = nd.NoiseDiffNet(net_args()).requires_grad_(False)
    engines = {}

    def engine_for(batch, height, width, device):
        return engines.setdefault((batch, height, width), FakeEngine(batch, height, width))

    monkeypatch.setattr(net, "engine_for", engine_for)

    def make(T=12, **kw):
        kw.setdefault("beta_schedule", "sigmoid2")
        return nd.GaussianDiffusion(net, image_size=8, timesteps=T, **kw)

    return make, engines


def _cond(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"clean_img": torch.rand((B, 4, 8, 8), generator=g), "position": torch.rand((B, 2, 8, 8), generator=g),
            "iso_ratio_idx": torch.arange(B) % 75}


def _manual_ddpm(gd, cond, B, seed):
    """The reference's draw order (ref :381,:371): x_T = randn(shape), then one randn per step that adds noise."""
    eng = FakeEngine(B, 8, 8)
    eng.set_condition(cond["clean_img"], cond["position"], cond["iso_ratio_idx"])
    torch.manual_seed(seed)
    x = torch.randn(B, 4, 8, 8)
    steps = gd.ddpm_steps()
    eng.chain_begin(steps, x, 0)
    xs = [x]
    for s in steps:
        z = torch.randn(B, 4, 8, 8) if s.sigma != 0.0 else torch.zeros(B, 4, 8, 8)
        eng.chain_run(1, z[None])
        xs.append(eng.chain_read())
    return torch.stack(xs, dim=1)


def test_torch_rng_stream_is_the_references_order(rig):
    make, _ = rig
    gd = make(T=12)
    cond = _cond(3)
    want = _manual_ddpm(gd, cond, 3, seed=7)
    torch.manual_seed(7)
    got = gd.sample(batch_size=3, condition=cond)
    assert torch.equal(got, want[:, -1])
    after = torch.rand(1)
    torch.manual_seed(7)
    torch.randn(3, 4, 8, 8)
    for _ in range(11):                                   # T - 1 noisy steps; t = 0 draws nothing (ref :371)
        torch.randn(3, 4, 8, 8)
    assert torch.equal(after, torch.rand(1))              # and nothing else was consumed from the global generator


@pytest.mark.parametrize("micro_batch,chunk", [(64, 25), (2, 25), (2, 5), (1, 1), (3, 7), (5, 12)])
def test_micro_batches_and_chunks_do_not_change_results(rig, micro_batch, chunk):
    make, engines = rig
    gd = make(T=12)
    cond = _cond(5)
    want = _manual_ddpm(gd, cond, 5, seed=3)
    gd.micro_batch, gd.chunk_steps = micro_batch, chunk
    torch.manual_seed(3)
    got = gd.sample(batch_size=5, condition=cond, return_all_timesteps=True)
    assert got.shape == (5, 13, 4, 8, 8) and torch.equal(got, want)
    mb = min(5, micro_batch)
    assert list(engines) == [(mb, 8, 8)]                  # one engine geometry; a short last micro-batch is padded, not re-planned
    log = engines[(mb, 8, 8)].log
    if mb >= 5:
        assert log.count("cond") == 1 and not any(e.startswith("seek") for e in log)      # single micro-batch: set once, never swapped
    else:
        assert log.count("begin") == -(-5 // mb) and any(e.startswith("seek") for e in log) == (chunk < 12)


def test_preset_mean_replaces_x_T_but_still_consumes_the_draw(rig):
    make, _ = rig
    gd = make(T=6)
    cond = _cond(2)
    preset = torch.full((2, 4, 8, 8), 0.25)
    torch.manual_seed(5)
    got = gd.sample(batch_size=2, condition=cond, preset_mean=preset, return_all_timesteps=True)
    assert torch.equal(got[:, 0], preset)
    # reference :381-387: img = randn(shape) is drawn first and then overwritten
    eng = FakeEngine(2, 8, 8)
    eng.set_condition(cond["clean_img"], cond["position"], cond["iso_ratio_idx"])
    torch.manual_seed(5)
    torch.randn(2, 4, 8, 8)
    steps = gd.ddpm_steps()
    eng.chain_begin(steps, preset, 0)
    for s in steps:
        eng.chain_run(1, (torch.randn(2, 4, 8, 8) if s.sigma != 0.0 else torch.zeros(2, 4, 8, 8))[None])
    assert torch.equal(got[:, -1], eng.chain_read())


@pytest.mark.parametrize("eta", [0.0, 0.7])
def test_ddim_draws_noise_for_every_non_final_pair(rig, eta):
    """ref :418-439: ``noise = torch.randn_like(img)`` sits outside any eta / sigma test, so every pair except the final
    (t, -1) consumes one draw — also at the default eta = 0, where sigma * noise adds nothing.  What must match the reference
    is the sample AND the state of the global generator afterwards (the x_T of the next sample() in a seeded loop)."""
    make, _ = rig
    gd = make(T=20, sampling_timesteps=5, ddim_sampling_eta=eta)
    assert gd.is_ddim_sampling and len(gd.ddim_steps()) == 5 and gd.ddim_time_pairs()[-1][1] == -1
    cond = _cond(2)
    torch.manual_seed(9)
    got = gd.sample(batch_size=2, condition=cond, preset_mean=torch.ones(2, 4, 8, 8))     # ddim_sample ignores preset_mean (ref :404-444)
    after = torch.rand(1)
    torch.manual_seed(9)
    x = torch.randn(2, 4, 8, 8)
    eng = FakeEngine(2, 8, 8)
    eng.set_condition(cond["clean_img"], cond["position"], cond["iso_ratio_idx"])
    steps = gd.ddim_steps()
    eng.chain_begin(steps, x, 0)
    for (t, tn), s in zip(gd.ddim_time_pairs(), steps):
        eng.chain_run(1, (torch.randn(2, 4, 8, 8) if tn >= 0 else torch.zeros(2, 4, 8, 8))[None])
    assert torch.equal(got, eng.chain_read()) and torch.equal(after, torch.rand(1))
    assert sum(1 for s in steps if s.sigma != 0.0) == (0 if eta == 0.0 else 4)
    assert [s.reserved[0] for s in steps] == [1, 1, 1, 1, 0]


def test_philox_mode_hands_seeds_to_the_library_and_reads_x_T_back(rig):
    make, engines = rig
    gd = make(T=6)
    gd.noise_source, gd.micro_batch = "philox", 2
    cond = _cond(4)
    torch.manual_seed(1)
    a = gd.sample(batch_size=4, condition=cond, return_all_timesteps=True)
    torch.manual_seed(1)
    b = gd.sample(batch_size=4, condition=cond, return_all_timesteps=True)
    torch.manual_seed(2)
    c = gd.sample(batch_size=4, condition=cond, return_all_timesteps=True)
    assert a.shape == (4, 7, 4, 8, 8) and torch.equal(a, b) and not torch.equal(a[:, 0], c[:, 0])
    assert not torch.equal(a[:2, 0], a[2:, 0])            # the two micro-batches got different seeds
    gd.noise_source = "nonsense"
    with pytest.raises(ValueError):
        gd.sample(batch_size=4, condition=cond)


def test_empty_batch_returns_empty_like_the_reference(rig):
    make, engines = rig
    gd = make(T=4)
    cond = {k: v[:0] for k, v in _cond(1).items()}
    assert gd.sample(batch_size=0, condition=cond).shape == (0, 4, 8, 8)
    assert gd.sample(batch_size=0, condition=cond, return_all_timesteps=True).shape == (0, 5, 4, 8, 8)
    assert not engines                                    # no engine is planned for nothing
