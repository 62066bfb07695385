"""The shipped checkpoint's geometry (script.sh:10: --dim 48 --crop_size 512; channels 48/96/192/384, GroupNorm groups of
6/12/24/48, time_dim 192) on the CUDA library: the narrower model is embedded into the 64-channel kernels with zero-padded,
slot-permuted channels (engine.cu embed_params).  Checked against the reference's own dim-48 forward (tests/golden/next_rows.npz,
minted from the unmodified reference), layer by layer against the oracle, at 512 x 512 on the GPU-fp32 oracle, and with the
teacher-forced T = 1000 step gate."""
import copy

import pytest
import torch

import noisediff_b200 as nd
from noisediff_b200 import _lib
from oracle import noisediff_oracle as O
from tests.test_gpu_fullshape import fp32_oracle_on_gpu
from tests.test_gpu_net import TAPS
from tests.util import load, rel_l2, seeded_net, seeded_sd, sd_hash

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net48():
    return copy.deepcopy(seeded_net(dim=48, seed=0)).cuda()


def test_dim48_forward_matches_the_references_golden(net48):
    z = load("next_rows.npz")
    assert sd_hash(seeded_sd(dim=48)) == str(z["dim48/weights_sha256"])       # same weights the reference minted the vector with
    cond = {k: v.cuda() for k, v in O.synthetic_condition(1, 64, 64, seed=5).items()}
    out = net48(torch.from_numpy(z["dim48/x"]).cuda(), torch.from_numpy(z["dim48/t"]).cuda(), cond)
    err = rel_l2(out, torch.from_numpy(z["dim48/out"]))
    print(f"dim=48 forward vs the reference's golden: rel-L2 {err:.3e}")
    assert out.shape == (1, 4, 64, 64) and err < 2.5e-2, err


def test_dim48_layer_by_layer(net48):
    """Every named activation of the embedded network, mapped back from physical (padded, slot-permuted) to logical channels,
    against the oracle; the padding itself must be exactly zero."""
    sd = seeded_sd(dim=48)
    B, S = 2, 64
    cond = O.synthetic_condition(B, S, S, seed=21)
    cond["iso_ratio_idx"] = torch.tensor([3, 57])
    x = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(22))
    t = torch.tensor([812, 40])
    taps = {}
    ref = O.net_forward(sd, x, t, cond, taps=taps)
    eng = nd.Engine(dim=48, batch=B, height=S, width=S, flags=_lib.FLAG_KEEP_ACTIVATIONS)
    eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
    eng.set_condition(*(cond[k].cuda() for k in ("clean_img", "position", "iso_ratio_idx")))
    out = eng.forward(x.cuda(), t.cuda())
    torch.cuda.synchronize()
    bad = []
    for name in TAPS:
        if name == "shot_attn":
            continue                                   # lives only on chip inside the fused shot-branch chain
        phys = eng.debug_tensor(name).cpu()
        C = taps[name].shape[1]
        Cp = phys.shape[1]
        assert Cp == C * 64 // 48, (name, C, Cp)
        idx = torch.tensor([(c // (C // 8)) * (Cp // 8) + c % (C // 8) for c in range(C)])
        pad = torch.ones(Cp, dtype=torch.bool)
        pad[idx] = False
        assert float(phys[:, pad].abs().max()) == 0.0, f"{name}: padding channels are not zero"
        e = rel_l2(phys[:, idx], taps[name])
        print(f"{name:20s} {e:.3e}")
        if not e < 3e-2:
            bad.append((name, e))
    eng.close()
    assert not bad, bad
    assert rel_l2(out, ref) < 2.5e-2, rel_l2(out, ref)


def test_dim48_forward_at_the_shipped_crop_size_512(net48):
    sd_gpu = {k: v.cuda() for k, v in seeded_sd(dim=48).items()}
    cond = {k: v.cuda() for k, v in O.synthetic_condition(1, 512, 512, seed=23).items()}
    x = torch.randn(1, 4, 512, 512, generator=torch.Generator().manual_seed(24)).cuda()
    t = torch.tensor([431]).cuda()
    out = net48(x, t, cond)
    with fp32_oracle_on_gpu():
        ref = O.net_forward(sd_gpu, x, t, cond)
    err = rel_l2(out, ref)
    print(f"dim=48, 512 x 512 forward vs the GPU-fp32 oracle: rel-L2 {err:.3e}")
    net48.release_engines()
    assert out.shape == (1, 4, 512, 512) and err < 2.5e-2, err


def test_dim48_teacher_forced_steps_T1000(net48):
    sd = seeded_sd(dim=48)
    gd = nd.GaussianDiffusion(net48, image_size=64, timesteps=1000, beta_schedule="sigmoid2", objective="pred_v").cuda()
    tab = O.schedule_tables("sigmoid2", 1000)
    ts = [999, 500, 100, 10, 1, 0]
    all_steps = {s.t: s for s in gd.ddpm_steps()}
    steps = [all_steps[t] for t in ts]
    g = torch.Generator().manual_seed(3)
    cond = O.synthetic_condition(2, 64, 64, seed=6)
    x0 = torch.randn(2, 4, 64, 64, generator=g) * 0.05
    x_in = [tab["sqrt_alphas_cumprod"][t] * x0 + tab["sqrt_one_minus_alphas_cumprod"][t] * torch.randn(2, 4, 64, 64, generator=g)
            for t in ts]
    noises = torch.randn(len(ts), 2, 4, 64, 64, generator=g)
    eng = net48.engine_for(2, 64, 64, torch.device("cuda", 0))
    eng.set_condition(*(cond[k].cuda() for k in ("clean_img", "position", "iso_ratio_idx")))
    eng.chain_begin(steps, x_in[0].cuda(), 0)
    snaps = torch.empty((len(ts), 2, 4, 64, 64), device="cuda")
    eng.chain_run(len(ts), noises.cuda().contiguous(), torch.stack(x_in).cuda().contiguous(), snaps)
    torch.cuda.synchronize()
    worst = 0.0
    for i, t in enumerate(ts):
        out = O.net_forward(sd, x_in[i], torch.full((2,), t, dtype=torch.long), cond)
        ref, _ = O.ddpm_step(tab, "pred_v", x_in[i], t, out, noises[i])
        e = rel_l2(snaps[i], ref)
        print(f"dim=48 t={t:4d}  x_(t-1) rel-L2 {e:.3e}")
        worst = max(worst, e)
    assert worst <= 2e-3, worst


def test_dim48_short_free_running_chain_and_sampling_api(net48):
    gd = nd.GaussianDiffusion(net48, image_size=64, timesteps=8, beta_schedule="sigmoid2", objective="pred_v").cuda()
    sd = seeded_sd(dim=48)
    g = torch.Generator().manual_seed(11)
    cond = O.synthetic_condition(2, 64, 64, seed=12)
    x_T = torch.randn(2, 4, 64, 64, generator=g)
    noises = [torch.randn(2, 4, 64, 64, generator=g) for _ in range(8)]
    xs = O.sample_chain(sd, cond, x_T, noises, T=8)
    got = gd._run_chain(gd.ddpm_steps(), (2, 4, 64, 64), {k: v.cuda() for k, v in cond.items()}, x_T.cuda(), False,
                        noises=torch.stack(noises).cuda())
    err = rel_l2(got, xs[-1])
    print(f"dim=48 free-running T=8 final rel-L2 {err:.3e}")
    assert err <= 1e-2, err
    torch.manual_seed(4)
    a = gd.sample(batch_size=2, condition={k: v.cuda() for k, v in cond.items()})
    assert a.shape == (2, 4, 64, 64) and torch.isfinite(a).all()
