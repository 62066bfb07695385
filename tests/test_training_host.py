"""CPU: the host-side schedules of the training row (noisediff_b200/training.py) against the oracle's restatements and torch's
own scheduler, and the data-parallel gradient all-reduce over gloo with two ranks (the B200 path uses the same call over NCCL)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from noisediff_b200 import training
from oracle import noisediff_oracle as O


def test_cosine_lr_equals_torch_scheduler():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=500)
    for epoch in range(1, 40):
        opt.step()
        sch.step()                                   # the reference steps the scheduler at the top of every epoch (ref :152-154)
        want = opt.param_groups[0]["lr"]
        assert math.isclose(training.cosine_lr(1e-4, epoch, 500), want, rel_tol=1e-9)
        assert math.isclose(O.cosine_annealing_lr(1e-4, epoch, 500), want, rel_tol=1e-9)


@pytest.mark.parametrize("cfg", [dict(beta=0.995, update_after_step=500, update_every=20), dict(beta=0.9, update_after_step=3, update_every=2),
                                 dict(beta=0.99, update_after_step=0, update_every=1)])
def test_ema_schedule_equals_the_oracles_restatement(cfg):
    """EmaSchedule only decides copy / lerp weights; applied to tensors it must reproduce oracle.ema_update step for step."""
    g = torch.Generator().manual_seed(0)
    params = {"w": torch.randn(5, generator=g)}
    st, mine, sched = {}, {"w": params["w"].clone()}, training.EmaSchedule(**cfg)
    n = 700 if cfg["update_after_step"] == 500 else 40
    for i in range(n):
        params = {"w": params["w"] + 0.01 * torch.randn(5, generator=g)}          # the optimizer moved the weights
        O.ema_update(params, st, **cfg)
        for kind, w in sched.update():
            mine["w"] = params["w"].clone() if kind == "copy" else torch.lerp(mine["w"], params["w"], w)
        assert torch.equal(mine["w"], st["ema"]["w"]), i
    assert sched.step == st["step"] == n
    if cfg["update_after_step"] == 500:               # the reference's setting: first real average at update call 520
        assert not torch.equal(mine["w"], params["w"])


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)        # this rank's flat gradient buffer
        scale = training.allreduce_gradients(flat)
        out[rank] = (flat * scale).tolist()
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_over_two_ranks_averages():
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_rank_main, args=(2, 29533, out), nprocs=2, join=True)
        want = (torch.arange(10, dtype=torch.float32) * 1.5).tolist()     # mean of 1x and 2x
        assert out[0] == want and out[1] == want


def test_allreduce_is_a_noop_without_a_process_group():
    flat = torch.ones(4)
    assert training.allreduce_gradients(flat) == 1.0 and torch.equal(flat, torch.ones(4))
