"""CPU: the oracle restatement (oracle/noisediff_oracle.py) against the golden vectors minted from the unmodified
reference (oracle/make_golden.py), and — when /root/reference is present — against the reference itself."""
import numpy as np
import pytest
import torch

from oracle import noisediff_oracle as O
from oracle import ref_shim
from tests.util import load, rel_l2, sd_hash, seeded_sd

torch.set_num_threads(8)


def _cond(z):
    return {"clean_img": torch.from_numpy(z["clean"]), "position": torch.from_numpy(z["position"]),
            "iso_ratio_idx": torch.from_numpy(z["iso"])}


def test_weights_are_the_reference_weights():
    z = load("fwd_64.npz")
    sd = seeded_sd()
    assert len(sd) == int(z["n_keys"]) == 416
    assert sum(v.numel() for v in sd.values()) == int(z["n_params"])
    assert sd_hash(sd) == str(z["weights_sha256"])


def test_forward_64_matches_reference_bit_exact():
    z = load("fwd_64.npz")
    out = O.net_forward(seeded_sd(), torch.from_numpy(z["x"]), torch.from_numpy(z["t"]), _cond(z))
    assert torch.equal(out, torch.from_numpy(z["out"]))


def test_ddpm_chain_matches_reference():
    z = load("chain_ddpm_T8.npz")
    xs = O.sample_chain(seeded_sd(), _cond(z), torch.from_numpy(z["x_T"]), list(torch.from_numpy(z["noises"])), T=8)
    got = torch.stack(xs, dim=1)
    ref = torch.from_numpy(z["xs"])
    assert got.shape == ref.shape == (2, 9, 4, 64, 64)
    assert rel_l2(got, ref) < 1e-6          # same ops; only scalar-tensor broadcasting order may differ by an ulp
    assert rel_l2(got[:, -1], ref[:, -1]) < 1e-6


def test_ddim_chain_matches_reference():
    z = load("chain_ddim_T50_S5.npz")
    zc = load("chain_ddpm_T8.npz")
    xs = O.sample_chain(seeded_sd(), _cond(zc), torch.from_numpy(z["x_T"]), list(torch.from_numpy(z["noises"])), T=50,
                        sampling_steps=5, eta=float(z["eta"]))
    ref = torch.from_numpy(z["xs"])
    assert rel_l2(torch.stack(xs, dim=1), ref) < 1e-6
    assert O.ddim_pairs(50, 5)[-1][1] == -1


@pytest.mark.parametrize("objective", ["pred_noise", "pred_x0"])
def test_other_objectives_match_reference(objective):
    z = load("chain_objectives_T4.npz")
    zc = load("chain_ddpm_T8.npz")
    xs = O.sample_chain(seeded_sd(), _cond(zc), torch.from_numpy(z["x_T"]),
                        list(torch.from_numpy(z[objective + "_noises"])), T=4, schedule="cosine", objective=objective)
    assert rel_l2(torch.stack(xs, dim=1), torch.from_numpy(z[objective])) < 1e-6


@pytest.mark.parametrize("objective", ["pred_v", "pred_noise", "pred_x0"])
def test_training_loss_value_matches_reference(objective):
    z = load("losses.npz")
    assert str(z["weights_sha256"]) == sd_hash(seeded_sd())
    tab = O.schedule_tables(str(z[objective + "/schedule"]), int(z[objective + "/T"]), objective)
    loss = O.p_losses(seeded_sd(), tab, objective, torch.from_numpy(z["x_start"]), torch.from_numpy(z[objective + "/t"]),
                      _cond(z), torch.from_numpy(z["noise"]))
    assert abs(float(loss) - float(z[objective + "/loss"])) <= 1e-6 * float(z[objective + "/loss"])


def test_next_row_dim48_forward_matches_reference():
    """SURVEY §8f N3 (the shipped --dim 48): the oracle and the host module tree are already generic in dim; the CUDA library is
    not (ndiff_engine_create rejects dim != 64) — this fixture is what its 48-channel kernels will be built against."""
    import noisediff_b200 as nd
    from tests.util import net_args
    z = load("next_rows.npz")
    torch.manual_seed(0)
    net = nd.NoiseDiffNet(net_args(48))
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    assert sd_hash(sd) == str(z["dim48/weights_sha256"]) and sum(v.numel() for v in sd.values()) == int(z["dim48/n_params"])
    out = O.net_forward(sd, torch.from_numpy(z["dim48/x"]), torch.from_numpy(z["dim48/t"]), O.synthetic_condition(1, 64, 64, seed=5))
    assert rel_l2(out, torch.from_numpy(z["dim48/out"])) < 1e-6


def test_next_row_training_step_matches_reference():
    """loss, every parameter's gradient norm, a few gradients in full, and the parameters after one Adam step."""
    z, zl = load("next_rows.npz"), load("losses.npz")
    sd = seeded_sd()
    tab = O.schedule_tables("sigmoid2", 1000, "pred_v")
    loss, grads = O.loss_gradients(sd, tab, "pred_v", torch.from_numpy(zl["x_start"]), torch.from_numpy(zl["pred_v/t"]), _cond(zl),
                                   torch.from_numpy(zl["noise"]))
    assert abs(float(loss) - float(z["train/loss"])) <= 1e-6 * float(z["train/loss"])
    names, norms = [str(n) for n in z["train/names"]], z["train/grad_norms"]
    assert len(names) == 416
    dead = [n for n, g in zip(names, norms) if g == 0.0]
    assert dead and all(("attn.to_q" in n or "attn.to_k" in n or ".norm1." in n) for n in dead)      # and nothing else is dead
    for n, g in zip(names, norms):
        mine = float(grads[n].double().norm())
        assert abs(mine - g) <= 2e-4 * g + 1e-9, (n, mine, g)
    small = [k[len("train/grad/"):] for k in z if k.startswith("train/grad/")]
    for k in small:
        assert rel_l2(grads[k], torch.from_numpy(z["train/grad/" + k])) < 1e-4, k
    params = {k: sd[k] for k in small}
    after = O.adam_step(params, {k: grads[k] for k in small}, {}, lr=float(z["train/lr"]))
    for k in small:
        ref = torch.from_numpy(z["train/after/" + k])
        assert not torch.equal(ref, sd[k]) and torch.allclose(after[k], ref, rtol=0, atol=2e-7), k


def test_schedule_tables_bit_exact():
    z = load("schedules.npz")
    for name in ("linear", "cosine", "sigmoid1", "sigmoid2", "sigmoid3"):
        for T in (1000, 50):
            tab = O.schedule_tables(name, T)
            for k, v in tab.items():
                assert np.array_equal(v.numpy(), z[f"{name}/{T}/{k}"]), (name, T, k)
    with pytest.raises(ValueError):
        O.beta_schedule("sigmoid", 10)       # the reference's ctor default string raises (ref :218)


def test_tile_grid_matches_survey():
    tiles = O.tile_origins(256)
    assert len(tiles) == 88 and tiles[0] == (0, 0) and tiles[-1] == (1872, 1168)
    assert len(O.tile_origins(512)) == 24
    pos = O.make_position(8, 8, x0=16, y0=32)
    assert pos.shape == (2, 8, 8) and abs(float(pos[0, 0, 0]) - 32 / 1423) < 1e-7 and abs(float(pos[1, 0, 0]) - 16 / 2127) < 1e-7


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_against_live_reference_forward_and_taps():
    net, gd = ref_shim.build(dim=64, image_size=32, timesteps=6)
    sd = {k: v.detach() for k, v in net.module.state_dict().items()}
    cond = O.synthetic_condition(1, 32, 32, seed=4)
    x = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(8))
    t = torch.tensor([2])
    with torch.no_grad():
        assert torch.equal(net(x, t, cond), O.net_forward(sd, x, t, cond))
    torch.manual_seed(11)
    with torch.no_grad():
        ref = gd.sample(batch_size=1, condition=cond, return_all_timesteps=True)
    torch.manual_seed(11)
    x_T = torch.randn(1, 4, 32, 32)
    zs = [torch.randn(1, 4, 32, 32) for _ in range(5)]
    mine = torch.stack(O.sample_chain(sd, cond, x_T, zs, T=6), dim=1)
    assert rel_l2(mine, ref) < 1e-6
