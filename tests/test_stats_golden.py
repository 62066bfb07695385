"""CPU: the 1k-sample statistics fixture (tests/golden/stats_1k.npz, minted from the UNMODIFIED reference by
oracle/make_golden_stats.py) against the oracle restatement.

The GPU side of this comparison (1024 Philox-driven chains of the CUDA path against the same fixture) is
tests/test_gpu_stats.py::test_1k_noise_statistics_match_reference_fixture."""
import numpy as np
import torch

from oracle import noisediff_oracle as O
from tests.util import load, noise_stats, rel_l2, sd_hash, seeded_sd

torch.set_num_threads(8)
CHUNK = 256          # oracle/make_golden_stats.py draws the reference's samples in sample() calls of 256


def _shared_condition(z, n):
    one = O.synthetic_condition(1, int(z["size"]), int(z["size"]), seed=int(z["cond_seed"]))
    return {k: v.expand(n, *v.shape[1:]).contiguous() for k, v in one.items()}


def test_fixture_belongs_to_the_seeded_weights_and_is_self_consistent():
    z = load("stats_1k.npz")
    assert int(z["n"]) == 1024 and int(z["size"]) == 32 and int(z["timesteps"]) == 24
    assert str(z["weights_sha256"]) == sd_hash(seeded_sd())
    assert z["psd2d"].shape == (4, 32, 32) and z["radial"].shape == (4, 8)
    assert np.allclose(z["radial"].sum(axis=1), 1.0)
    # Parseval: the mean periodogram (per-patch DC removed) sums to the within-patch variance, which is below the pooled one
    within = z["psd2d"].astype(np.float64).sum(axis=(1, 2)) / (32 * 32)
    assert (within > 0).all() and (within <= z["var"] * 1.001).all()
    # the two independent reference sets agree with each other to the sampling noise the tolerances are stated against
    assert float(z["self_mean"].max()) < 0.02 and float(z["self_var"].max()) < 0.03 and float(z["self_radial_max"]) < 0.05


def test_oracle_reproduces_the_fixtures_first_samples():
    """Sample-level pin of the fixture: replaying the reference's RNG consumption (x_T = randn(shape), then one randn_like per
    noisy step, denoising_diffusion_pytorch.py:381,371) through the oracle chain gives the reference's first two samples."""
    z = load("stats_1k.npz")
    S, T = int(z["size"]), int(z["timesteps"])
    torch.manual_seed(1000)
    x_T = torch.randn(CHUNK, 4, S, S)
    zs = [torch.randn(CHUNK, 4, S, S) for _ in range(T - 1)]
    cond = _shared_condition(z, 2)
    with torch.no_grad():
        got = O.sample_chain(seeded_sd(), cond, x_T[:2], [n[:2] for n in zs], T=T)[-1]
    assert rel_l2(got, torch.from_numpy(z["first"])) < 1e-4     # conv kernels chosen per batch size may round differently


def test_oracle_statistics_agree_with_reference_fixture():
    """64 oracle chains with the oracle's own draws: coarse agreement (the sampling error of 64 patches dominates)."""
    z = load("stats_1k.npz")
    S, T, n = int(z["size"]), int(z["timesteps"]), 64
    g = torch.Generator().manual_seed(4321)
    x_T = torch.randn(n, 4, S, S, generator=g)
    zs = [torch.randn(n, 4, S, S, generator=g) for _ in range(T - 1)]
    with torch.no_grad():
        got = O.sample_chain(seeded_sd(), _shared_condition(z, n), x_T, zs, T=T)[-1]
    st = noise_stats(got)
    assert (np.abs(st["mean"] - z["mean"]) <= 0.08 * np.sqrt(z["var"])).all()
    assert (np.abs(st["var"] - z["var"]) / z["var"] <= 0.12).all()
    assert float((np.abs(st["radial"] - z["radial"]) / z["radial"]).max()) <= 0.15
