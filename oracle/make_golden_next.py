"""Generates tests/golden/next_rows.npz: reference outputs for the two scope rows that come AFTER the sampling path
(SURVEY.md §8f), so that their CUDA kernels can be built oracle-first next round.  TEST INFRASTRUCTURE; build container only:

    python -m oracle.make_golden_next

  N3  the shipped network width: ``NoiseDiffNet(dim=48)`` forward (script.sh:10 ``--dim 48``), 1 x 4 x 64 x 64, seed-0 weights.
  N1  one training step of ``Trainer.train`` (models/trainer_diffusion.py:141-227) at dim = 64 on the inputs of losses.npz
      (pred_v, sigmoid2, T = 1000, per-sample t): ``loss.backward()`` then ``torch.optim.Adam(lr=1e-4, weight_decay=0).step()``
      (:92, train_diffusion.py:85-87).  Stored: the loss, the L2 norm of every parameter's gradient, a few small tensors'
      gradients and post-step values in full.  EMA (``ema_pytorch``, an external dependency absent here) is not exercised.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import noisediff_oracle as O   # noqa: E402
from oracle import ref_shim                # noqa: E402
from oracle.make_golden import sd_hash     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "next_rows.npz")
SMALL = ["final_conv.weight", "time_mlp.3.bias", "iso_embed.weight", "init_conv.bias", "shot_mlp3.fc2.weight",
         "downs.0.2.attn.to_v.weight", "mid_block1.block1.norm.weight"]


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    # ---- N3: dim = 48 forward -----------------------------------------------------------------------------------------
    net48, _ = ref_shim.build(dim=48, seed=0, image_size=64, timesteps=8)
    sd48 = {k: v.detach() for k, v in net48.module.state_dict().items()}
    cond = O.synthetic_condition(1, 64, 64, seed=5)
    x = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(6))
    t = torch.tensor([421])
    with torch.no_grad():
        v = net48(x, t, cond)
    out.update({"dim48/x": x.numpy(), "dim48/t": t.numpy(), "dim48/out": v.numpy(), "dim48/weights_sha256": sd_hash(sd48),
                "dim48/n_params": int(sum(p.numel() for p in sd48.values()))})
    print("dim48 forward", tuple(v.shape), float(v.std()), out["dim48/n_params"])

    # ---- N1: one training step ----------------------------------------------------------------------------------------
    z = np.load(os.path.join(ROOT, "tests", "golden", "losses.npz"))
    net, _ = ref_shim.build(dim=64, seed=0, image_size=64, timesteps=8)
    _, GD = ref_shim.load()
    assert sd_hash({k: v.detach() for k, v in net.module.state_dict().items()}) == str(z["weights_sha256"])
    net.train()
    gd = GD(net, image_size=64, timesteps=1000, beta_schedule="sigmoid2", objective="pred_v")
    cond = {"clean_img": torch.from_numpy(z["clean"]), "position": torch.from_numpy(z["position"]), "iso_ratio_idx": torch.from_numpy(z["iso"])}
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=0)
    opt.zero_grad()
    loss = gd.p_losses(torch.from_numpy(z["x_start"]), torch.from_numpy(z["pred_v/t"]), cond, noise=torch.from_numpy(z["noise"]).clone())
    loss.backward()
    names, norms = [], []
    for k, p in net.module.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    grads = {k: dict(net.module.named_parameters())[k].grad.clone() for k in SMALL}
    opt.step()
    after = {k: dict(net.module.named_parameters())[k].detach().clone() for k in SMALL}
    out.update({"train/loss": np.float32(loss.item()), "train/names": np.array(names), "train/grad_norms": np.array(norms),
                "train/lr": 1e-4})
    for k in SMALL:
        out["train/grad/" + k] = grads[k].numpy()
        out["train/after/" + k] = after[k].numpy()
    print("train loss", float(loss), "grad norm total", float(np.sqrt((np.array(norms) ** 2).sum())),
          "zero-grad params", [n for n, g in zip(names, norms) if g == 0.0][:8], "...")
    np.savez(OUT, **out)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
