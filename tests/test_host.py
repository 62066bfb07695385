"""CPU: host-side mirror of the reference interface (no compute on the library without a GPU)."""
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from torch import nn

import noisediff_b200 as nd
from noisediff_b200 import _lib
from tests.util import ROOT, load, net_args, seeded_net


def test_library_loads_and_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 20 and "ndiff_forward" in names and "ndiff_chain_run" in names
    l = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/noisediff_b200.h but not exported"
    assert set(names) == set(_lib._SIGS), "ctypes signature table out of sync with the header"
    assert _lib.lib().ndiff_abi_version() == 1


def test_c_abi_from_plain_c(tmp_path):
    """The header is valid C99 (-pedantic) and a C program can link the library and read its errors."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", exe, "-L", libdir, "-lnoisediff_b200",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "multiple of 8 up to 64" in r.stdout and "training step runs at dim = 64" in r.stdout and "multiples of 8" in r.stdout


def test_abi_struct_layouts():
    assert ctypes.sizeof(_lib.Step) == 48 and ctypes.sizeof(_lib.Config) == 24


def test_engine_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        nd.Engine(dim=64, batch=1, height=64, width=64)
    net = seeded_net()
    cond = {"clean_img": torch.zeros(1, 4, 8, 8), "position": torch.zeros(1, 2, 8, 8),
            "iso_ratio_idx": torch.zeros(1, dtype=torch.long)}
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 4, 8, 8), torch.zeros(1, dtype=torch.long), cond)
    gd = nd.GaussianDiffusion(net, image_size=8, timesteps=4, beta_schedule="sigmoid2")
    with pytest.raises(RuntimeError, match="CUDA"):
        gd.sample(batch_size=1, condition=cond)


def test_net_attributes_and_state_dict_names():
    net = seeded_net()
    assert (net.channels, net.out_dim, net.self_condition, net.random_or_learned_sinusoidal_cond) == (4, 4, False, False)
    sd = net.state_dict()
    assert len(sd) == 416
    for k, shp in {"init_conv.weight": (64, 4, 7, 7), "iso_embed.weight": (100, 16), "time_mlp.1.weight": (256, 64),
                   "downs.0.0.mlp.1.weight": (128, 256), "downs.0.2.attn.to_q.weight": (128, 64),
                   "downs.0.2.ff.net.0.0.weight": (128, 64), "downs.0.2.ff.net.2.weight": (64, 128),
                   "downs.0.3.1.weight": (64, 256, 1, 1), "downs.3.3.weight": (512, 256, 3, 3),
                   "ups.0.0.res_conv.weight": (512, 768, 1, 1), "ups.0.3.1.weight": (256, 512, 3, 3),
                   "ups.3.3.weight": (64, 64, 3, 3), "pos_enc.weights.weight": (8, 2, 1, 1),
                   "pos_block1.mlp.1.weight": (128, 8, 1, 1), "shot_mlp3.fc2.weight": (4, 64, 1, 1),
                   "final_conv.weight": (4, 64, 1, 1)}.items():
        assert tuple(sd[k].shape) == shp, k
    # survives the reference's wrapping (models/modules.py:81) and strict reload (trainer_diffusion.py:333-349)
    wrapped = nn.DataParallel(net)
    clone = nd.NoiseDiffNet(net_args())
    clone.load_state_dict({("module." + k)[7:]: v for k, v in wrapped.module.state_dict().items()}, strict=True)
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 4, 12, 12), torch.zeros(1, dtype=torch.long), {})


def test_engine_cache_is_bounded_lru(monkeypatch):
    """A generation loop with ragged batch sizes must not pile up engines (each owns its activation pool)."""
    from noisediff_b200 import arch
    made, closed = [], []

    class FakeEngine:
        def __init__(self, *, dim, batch, height, width, device):
            self.key, self.weights_version = (batch, height, width), None
            made.append(self.key)

        def load_state_dict(self, sd):
            assert len(sd) == 416

        def close(self):
            closed.append(self.key)

    monkeypatch.setattr(arch._engine, "Engine", FakeEngine)
    net = nd.NoiseDiffNet(net_args())
    dev = torch.device("cuda", 0)
    for b in (1, 2, 3, 4):
        net.engine_for(b, 32, 32, dev)
    assert made == [(b, 32, 32) for b in (1, 2, 3, 4)] and closed == []
    net.engine_for(1, 32, 32, dev)                         # touch: batch 1 becomes the most recently used
    net.engine_for(5, 32, 32, dev)                         # fifth geometry evicts the least recently used (batch 2)
    assert closed == [(2, 32, 32)] and len(net._engines) == 4
    assert net.engine_for(1, 32, 32, dev).key == (1, 32, 32) and made.count((1, 32, 32)) == 1
    net.release_engines()
    assert len(net._engines) == 0 and len(closed) == 5


def test_drop_in_through_the_references_own_plugin_registry():
    """a new `models/archs/B200_arch.py` exporting the class is all the reference's registry needs
    (models/modules.py:20-41,86-92).  Drives the LIVE reference's define_G / init_net with `--net_name NoiseDiffNetB200`, then
    the strict checkpoint exchange of Trainer.load_networks (models/trainer_diffusion.py:333-349) in both directions."""
    import importlib
    import types
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not mounted")
    RefNet, RefGD = ref_shim.load()
    M = importlib.import_module("models.modules")
    plug = types.ModuleType("models.archs.B200_arch")          # what importing the new arch file yields
    plug.NoiseDiffNetB200 = nd.NoiseDiffNet
    M._arch_modules.append(plug)
    try:
        args = SimpleNamespace(net_name="NoiseDiffNetB200", gpu_ids=[], device=torch.device("cpu"), dist=False, **vars(net_args()))
        net = M.define_G(args)
        assert isinstance(net, nd.NoiseDiffNet)
        with pytest.raises(ValueError):
            M.define_G(SimpleNamespace(**{**vars(args), "net_name": "NoSuchNet"}))
    finally:
        M._arch_modules.remove(plug)
    torch.manual_seed(0)
    ref = RefNet(net_args())
    ckpt = {"module." + k: v for k, v in ref.state_dict().items()}           # a DataParallel-saved reference checkpoint
    net.load_state_dict({k[7:]: v for k, v in ckpt.items()}, strict=True)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), ref.state_dict().values()))
    ref.load_state_dict(net.state_dict(), strict=True)                       # and back: same 416 keys and shapes
    # the reference's GaussianDiffusion reads the same attributes off either network class (denoising_diffusion_pytorch.py:184-198)
    mine = nd.GaussianDiffusion(nn.DataParallel(net), image_size=64, timesteps=10, beta_schedule="sigmoid2")
    theirs = RefGD(nn.DataParallel(ref), image_size=64, timesteps=10, beta_schedule="sigmoid2", objective="pred_v")
    assert (mine.channels, mine.self_condition, mine.num_timesteps, mine.objective, mine.is_ddim_sampling) == \
        (theirs.channels, theirs.self_condition, theirs.num_timesteps, theirs.objective, theirs.is_ddim_sampling)
    assert sorted(k for k, _ in mine.named_buffers(recurse=False)) == sorted(k for k, _ in theirs.named_buffers(recurse=False))


@pytest.mark.parametrize("name", ["linear", "cosine", "sigmoid1", "sigmoid2", "sigmoid3"])
def test_schedule_buffers_match_reference(name):
    z = load("schedules.npz")
    for T in (1000, 50):
        gd = nd.GaussianDiffusion(seeded_net(), image_size=64, timesteps=T, beta_schedule=name)
        bufs = dict(gd.named_buffers(recurse=False))
        assert len(bufs) == 13
        for k, v in bufs.items():
            assert v.dtype == torch.float32 and np.array_equal(v.numpy(), z[f"{name}/{T}/{k}"]), (name, T, k)


def test_ctor_contract():
    net = seeded_net()
    with pytest.raises(ValueError):
        nd.GaussianDiffusion(net, image_size=64)                      # default 'sigmoid' raises as in the reference
    with pytest.raises(AssertionError):
        nd.GaussianDiffusion(net, image_size=64, beta_schedule="linear", objective="nope")
    gd = nd.GaussianDiffusion(nn.DataParallel(net), image_size=64, timesteps=100, sampling_timesteps=10,
                              beta_schedule="sigmoid2", ddim_sampling_eta=1.0)
    assert gd.is_ddim_sampling and gd.channels == 4 and gd.num_timesteps == 100
    assert gd.ddim_time_pairs()[0] == (99, 89) and gd.ddim_time_pairs()[-1][1] == -1
    assert nd.GaussianDiffusion(net, image_size=64, timesteps=10, beta_schedule="linear", auto_normalize=True) \
        .unnormalize(torch.zeros(1)).item() == 0.5


def test_step_tables_follow_the_buffers():
    gd = nd.GaussianDiffusion(seeded_net(), image_size=64, timesteps=1000, beta_schedule="sigmoid2", objective="pred_v")
    steps = gd.ddpm_steps()
    assert [s.t for s in steps[:3]] == [999, 998, 997] and steps[-1].t == 0 and len(steps) == 1000
    s = steps[500]
    t = s.t
    assert s.p == pytest.approx(float(gd.sqrt_alphas_cumprod[t])) and s.q == pytest.approx(-float(gd.sqrt_one_minus_alphas_cumprod[t]))
    assert s.a == pytest.approx(float(gd.posterior_mean_coef1[t])) and s.b == pytest.approx(float(gd.posterior_mean_coef2[t]))
    assert s.sigma == pytest.approx(float((0.5 * gd.posterior_log_variance_clipped[t]).exp())) and s.c == 0.0 and s.clip == 1
    assert steps[-1].sigma == 0.0                                     # no noise at t == 0 (ref :371)
    g2 = nd.GaussianDiffusion(seeded_net(), image_size=64, timesteps=50, sampling_timesteps=5, ddim_sampling_eta=0.5,
                              beta_schedule="sigmoid2")
    d = g2.ddim_steps()
    assert len(d) == 5 and d[-1].a == 1.0 and d[-1].sigma == 0.0 and d[0].b == 0.0 and d[0].sigma > 0.0
    g3 = nd.GaussianDiffusion(seeded_net(), image_size=64, timesteps=4, beta_schedule="linear", objective="pred_x0")
    assert (g3.ddpm_steps()[0].p, g3.ddpm_steps()[0].q) == (0.0, 1.0)


def test_elementwise_helpers_match_oracle():
    from oracle import noisediff_oracle as O
    gd = nd.GaussianDiffusion(seeded_net(), image_size=8, timesteps=20, beta_schedule="sigmoid2")
    tab = O.schedule_tables("sigmoid2", 20)
    g = torch.Generator().manual_seed(0)
    x, v, z = (torch.randn(2, 4, 8, 8, generator=g) for _ in range(3))
    t = torch.tensor([7, 7])
    x0 = gd.predict_start_from_v(x, t, v)
    assert torch.equal(x0, tab["sqrt_alphas_cumprod"][7] * x - tab["sqrt_one_minus_alphas_cumprod"][7] * v)
    ref, _ = O.ddpm_step(tab, "pred_v", x, 7, v, z)
    mean, _, logvar = gd.q_posterior(x0.clamp(-1, 1), x, t)
    assert torch.allclose(mean + (0.5 * logvar).exp() * z, ref, atol=0, rtol=0)
    assert torch.equal(gd.q_sample(x, t, z), tab["sqrt_alphas_cumprod"][7] * x + tab["sqrt_one_minus_alphas_cumprod"][7] * z)
    # training: gradients are not built — asking for them raises; the loss value is a no_grad/CUDA-only call
    trainable = nd.GaussianDiffusion(nd.NoiseDiffNet(net_args()), image_size=8, timesteps=20, beta_schedule="sigmoid2")
    with pytest.raises(NotImplementedError):
        trainable(torch.zeros(1, 4, 8, 8), {})
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        gd(torch.zeros(1, 4, 8, 8), {"clean_img": torch.zeros(1, 4, 8, 8), "position": torch.zeros(1, 2, 8, 8),
                                     "iso_ratio_idx": torch.zeros(1, dtype=torch.long)})


def test_fast_gelu_coefficients_track_exact_erf_gelu():
    """The CUDA kernels evaluate nn.GELU() (exact erf, ref Diffusion_arch.py:345,412) as x * sigmoid(x * P(x^2)) with the
    coefficients below (noisediff_b200/csrc/common.cuh::gelu_erf); pin their error against torch's erf GELU."""
    import re
    src = open(os.path.join(ROOT, "noisediff_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_erf(float x)"):]
    body = body[:body.index("}")]
    a2, a1, a0 = [float(v) for v in re.findall(r"(-?\d\.\d+e?-?\d*)f \* kL2e", body)]
    x = torch.linspace(-12, 12, 480001, dtype=torch.float64)
    x2 = (x * x).clamp(max=52.6)
    L2E = 1.4426950408889634
    p = ((a2 * x2 + a1) * x2 + a0) * 1.0              # = -(P) / ... : the kernel folds the minus sign and log2(e) into the constants
    approx = x / (1 + torch.exp(x * p))               # the kernel's ex2(x * p * log2 e) == exp(x * p)
    exact = 0.5 * x * (1 + torch.erf(x / 2 ** 0.5))
    assert float((approx - exact).abs().max()) < 4e-5


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the driver's reference arm): one JSON line on stdout with the contract's keys, no GPU needed."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
