"""fp32 CPU oracle for the NoiseDiff sampling hot path (TEST INFRASTRUCTURE, not product code).

A from-scratch *functional* restatement (plain ``torch.nn.functional`` calls driven by a state_dict, no
``nn.Module`` tree) of what the reference computes on the path
``GaussianDiffusion.p_sample_loop / ddim_sample -> NoiseDiffNet.forward``.  Every function cites the reference
lines it follows (paths relative to /root/reference).  Parity pin: ``oracle/make_golden.py`` imports the
*unmodified* reference in the build container, runs both on identical seeded weights / conditions / noise and
stores the reference outputs under ``tests/golden/``; ``tests/test_oracle.py`` re-checks this file against
those fixtures on any machine (the reference ships no tests or golden vectors of its own — SURVEY.md §4).

Arithmetic type: float32 everywhere (float64 only for the beta schedule, as in the reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------------------------------
# network pieces
# --------------------------------------------------------------------------------------------------------------

def _conv(sd: SD, name: str, x: Tensor, pad: int = 0) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad)


def _linear(sd: SD, name: str, x: Tensor, bias: bool = True) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"] if bias else None)


def block(sd: SD, pfx: str, x: Tensor, groups: int, scale_shift=None) -> Tensor:
    """models/archs/Diffusion_arch.py:128-144 — conv3x3(pad 1) -> GroupNorm(eps 1e-5) -> x*(scale+1)+shift -> SiLU.
    (ResnetBlock hard-codes ks=3,pd=1 for both of its Blocks, :154-155.)"""
    y = _conv(sd, pfx + ".proj", x, pad=1)
    y = F.group_norm(y, groups, sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"], eps=1e-5)
    if scale_shift is not None:
        sc, sh = scale_shift
        y = y * (sc + 1) + sh
    return F.silu(y)


def resnet_block(sd: SD, pfx: str, x: Tensor, temb: Tensor, groups: int = 8) -> Tensor:
    """Diffusion_arch.py:146-170 — time MLP (SiLU, Linear) -> (scale, shift) into block1 only; + res_conv(x)."""
    e = _linear(sd, pfx + ".mlp.1", F.silu(temb))[:, :, None, None]
    sc, sh = e.chunk(2, dim=1)
    h = block(sd, pfx + ".block1", x, groups, (sc, sh))
    h = block(sd, pfx + ".block2", h, groups)
    res = _conv(sd, pfx + ".res_conv", x) if (pfx + ".res_conv.weight") in sd else x
    return h + res


def resnet_block_pos(sd: SD, pfx: str, x: Tensor, pos_emb: Tensor, groups: int = 2) -> Tensor:
    """Diffusion_arch.py:173-196 (ResnetBlock2) — scale/shift are per-pixel maps conv1x1(SiLU(pos_emb))."""
    e = _conv(sd, pfx + ".mlp.1", F.silu(pos_emb))
    sc, sh = e.chunk(2, dim=1)
    h = block(sd, pfx + ".block1", x, groups, (sc, sh))
    h = block(sd, pfx + ".block2", h, groups)
    res = _conv(sd, pfx + ".res_conv", x) if (pfx + ".res_conv.weight") in sd else x
    return h + res


def cross_attention(sd: SD, pfx: str, x: Tensor, ctx: Tensor, heads: int = 4) -> Tensor:
    """Diffusion_arch.py:361-402 — full (un-collapsed) cross attention, kept literal so the algebraic collapse
    used by the CUDA path is itself checked against it."""
    b, n, _ = x.shape
    q = _linear(sd, pfx + ".to_q", x, bias=False)
    k = _linear(sd, pfx + ".to_k", ctx, bias=False)
    v = _linear(sd, pfx + ".to_v", ctx, bias=False)
    d = q.shape[-1] // heads
    split = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    att = sim.softmax(dim=-1)
    o = torch.einsum("bij,bjd->bid", att, v)
    o = o.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, heads * d)
    return _linear(sd, pfx + ".to_out.0", o)


def attn_block(sd: SD, pfx: str, x: Tensor, ctx: Tensor) -> Tensor:
    """Diffusion_arch.py:425-443 — tokens = pixels; attn(LN1)+x ; FF(LN2)+x ; proj_out(.) + x_in."""
    b, c, h, w = x.shape
    tok = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    n1 = F.layer_norm(tok, (c,), sd[pfx + ".norm1.weight"], sd[pfx + ".norm1.bias"], eps=1e-5)
    tok = cross_attention(sd, pfx + ".attn", n1, ctx) + tok
    n2 = F.layer_norm(tok, (c,), sd[pfx + ".norm2.weight"], sd[pfx + ".norm2.bias"], eps=1e-5)
    ff = _linear(sd, pfx + ".ff.net.2", F.gelu(_linear(sd, pfx + ".ff.net.0.0", n2)))   # :405-422, exact-erf GELU
    tok = ff + tok
    y = tok.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return _conv(sd, pfx + ".proj_out", y) + x


def mlp1x1(sd: SD, pfx: str, x: Tensor) -> Tensor:
    """Diffusion_arch.py:340-356 — conv1x1 -> GELU(erf) -> conv1x1 (dropout p=0)."""
    return _conv(sd, pfx + ".fc2", F.gelu(_conv(sd, pfx + ".fc1", x)))


def time_embedding(sd: SD, time: Tensor, dim: int) -> Tensor:
    """Diffusion_arch.py:94-107 + :502-507 — sinusoidal(dim, theta 1e4) -> Linear -> GELU -> Linear."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=time.device) * -(math.log(10000.0) / (half - 1)))
    ang = time.to(torch.float32)[:, None] * freq[None, :]
    emb = torch.cat((ang.sin(), ang.cos()), dim=-1)
    return _linear(sd, "time_mlp.3", F.gelu(_linear(sd, "time_mlp.1", emb)))


def space_to_depth(x: Tensor) -> Tensor:
    """Diffusion_arch.py:78-82 — 'b c (h p1) (w p2) -> b (c p1 p2) h w', p1=p2=2."""
    b, c, h, w = x.shape
    x = x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(b, c * 4, h // 2, w // 2)


def net_forward(sd: SD, x: Tensor, time: Tensor, condition: Dict[str, Tensor], dim: Optional[int] = None,
                taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """NoiseDiffNet.forward, Diffusion_arch.py:577-646.  ``taps`` (optional dict) receives named intermediate
    activations so individual CUDA kernels can be checked layer by layer."""
    if dim is None:
        dim = sd["init_conv.weight"].shape[0]
    rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
    assert x.shape[-1] % 8 == 0 and x.shape[-2] % 8 == 0                           # :578
    clean, position = condition["clean_img"], condition["position"]

    # positional condition (:584-585) — LearnedSinusoidalPosEmb :322-337 then Mlp
    w = _conv(sd, "pos_enc.weights", position)
    fr = w * 2 * math.pi
    pos = torch.cat((w, torch.cat((fr.sin(), fr.cos()), dim=1)), dim=1)
    pos_emb = mlp1x1(sd, "pos_mlp", pos)
    rec("pos_emb", pos_emb)
    # camera condition (:590-591)
    iso = F.embedding(condition["iso_ratio_idx"], sd["iso_embed.weight"])[:, None, :]
    # time (:595)
    temb = time_embedding(sd, time, dim)
    rec("temb", temb)

    # shot-noise branch (:598-604) — clean image first, x second
    s = mlp1x1(sd, "shot_mlp1", torch.cat([clean, x], dim=1))
    rs = s
    rec("shot_mlp1", s)
    s = attn_block(sd, "shot_attn", s, iso)
    rec("shot_attn", s)
    s = mlp1x1(sd, "shot_mlp2", s)
    rec("shot_mlp2", s)
    s = resnet_block(sd, "shot_time", s, temb, groups=2)
    s = s + rs
    rec("shot_time", s)
    shot_noise = mlp1x1(sd, "shot_mlp3", s)
    rec("shot_noise", shot_noise)

    # main U-Net (:606-643)
    y = _conv(sd, "init_conv", x, pad=3)
    r = y
    rec("init_conv", y)
    skips: List[Tensor] = []
    y = resnet_block_pos(sd, "pos_block1", y, pos_emb)
    rec("pos_block1", y)
    for i in range(4):
        y = resnet_block(sd, f"downs.{i}.0", y, temb); skips.append(y)
        rec(f"downs.{i}.0", y)
        y = resnet_block(sd, f"downs.{i}.1", y, temb); skips.append(y)
        rec(f"downs.{i}.1", y)
        y = attn_block(sd, f"downs.{i}.2", y, iso)
        rec(f"downs.{i}.2", y)
        if i < 3:
            y = _conv(sd, f"downs.{i}.3.1", space_to_depth(y))
        else:
            y = _conv(sd, f"downs.{i}.3", y, pad=1)
        rec(f"downs.{i}.3", y)
    y = resnet_block(sd, "mid_block1", y, temb)
    rec("mid_block1", y)
    y = resnet_block(sd, "mid_block2", y, temb)
    rec("mid_block2", y)
    for i in range(4):
        y = resnet_block(sd, f"ups.{i}.0", torch.cat((y, skips.pop()), dim=1), temb)
        rec(f"ups.{i}.0", y)
        y = resnet_block(sd, f"ups.{i}.1", torch.cat((y, skips.pop()), dim=1), temb)
        rec(f"ups.{i}.1", y)
        y = attn_block(sd, f"ups.{i}.2", y, iso)
        rec(f"ups.{i}.2", y)
        if i < 3:
            y = _conv(sd, f"ups.{i}.3.1", F.interpolate(y, scale_factor=2, mode="nearest"), pad=1)   # :72-76
        else:
            y = _conv(sd, f"ups.{i}.3", y, pad=1)
        rec(f"ups.{i}.3", y)
    y = resnet_block_pos(sd, "pos_block2", y, pos_emb)
    rec("pos_block2", y)
    y = resnet_block(sd, "final_res_block", torch.cat((y, r), dim=1), temb)
    rec("final_res_block", y)
    read_noise = _conv(sd, "final_conv", y)
    rec("read_noise", read_noise)
    return shot_noise + read_noise                                                  # :644


# --------------------------------------------------------------------------------------------------------------
# diffusion process
# --------------------------------------------------------------------------------------------------------------

def _sigmoid_schedule(T: int, start: float, end: float, tau: float) -> Tensor:
    """models/denoising_diffusion_pytorch.py:119-164 (the three sigmoid variants differ only in start/end/tau)."""
    t = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
    v0 = torch.tensor(start / tau).sigmoid()
    v1 = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v1) / (v1 - v0)
    ac = ac / ac[0]
    return torch.clip(1 - ac[1:] / ac[:-1], 0, 0.999)


def beta_schedule(name: str, T: int) -> Tensor:
    """denoising_diffusion_pytorch.py:96-164, dispatch :207-218 (float64)."""
    if name == "linear":
        s = 1000 / T
        return torch.linspace(s * 1e-4, s * 0.02, T, dtype=torch.float64)
    if name == "cosine":
        t = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
        ac = torch.cos((t + 0.008) / 1.008 * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - ac[1:] / ac[:-1], 0, 0.999)
    if name == "sigmoid1":
        return _sigmoid_schedule(T, -3, 3, 0.5)
    if name == "sigmoid2":
        return _sigmoid_schedule(T, -7, 3, 0.7)
    if name == "sigmoid3":
        return _sigmoid_schedule(T, -10, 3, 0.7)
    raise ValueError(f"unknown beta schedule {name}")


def schedule_tables(name: str, T: int, objective: str = "pred_v") -> Dict[str, Tensor]:
    """The 13 fp32 buffers of GaussianDiffusion.__init__, denoising_diffusion_pytorch.py:220-286."""
    betas = beta_schedule(name, T)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    acp = F.pad(ac[:-1], (1, 0), value=1.0)
    pv = betas * (1.0 - acp) / (1.0 - ac)
    snr = ac / (1 - ac)
    lw = {"pred_noise": snr / snr, "pred_x0": snr, "pred_v": snr / (snr + 1)}[objective]
    tab = dict(
        betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=acp,
        sqrt_alphas_cumprod=ac.sqrt(), sqrt_one_minus_alphas_cumprod=(1.0 - ac).sqrt(),
        log_one_minus_alphas_cumprod=(1.0 - ac).log(), sqrt_recip_alphas_cumprod=(1.0 / ac).sqrt(),
        sqrt_recipm1_alphas_cumprod=(1.0 / ac - 1).sqrt(), posterior_variance=pv,
        posterior_log_variance_clipped=pv.clamp(min=1e-20).log(),
        posterior_mean_coef1=betas * acp.sqrt() / (1.0 - ac),
        posterior_mean_coef2=(1.0 - acp) * alphas.sqrt() / (1.0 - ac), loss_weight=lw)
    return {k: v.to(torch.float32) for k, v in tab.items()}


def predictions(tab, objective: str, x: Tensor, t: int, out: Tensor, clip: bool) -> Tuple[Tensor, Tensor]:
    """model_predictions, denoising_diffusion_pytorch.py:331-354 -> (pred_noise, x_start)."""
    r1, r2 = tab["sqrt_recip_alphas_cumprod"][t], tab["sqrt_recipm1_alphas_cumprod"][t]
    mc = (lambda z: z.clamp(-1.0, 1.0)) if clip else (lambda z: z)
    if objective == "pred_noise":
        x0 = mc(r1 * x - r2 * out)                                                    # :298-302
        eps = (r1 * x - x0) / r2 if clip else out                                     # rederive only when clipping
    elif objective == "pred_x0":
        x0 = mc(out)
        eps = (r1 * x - x0) / r2
    elif objective == "pred_v":
        x0 = mc(tab["sqrt_alphas_cumprod"][t] * x - tab["sqrt_one_minus_alphas_cumprod"][t] * out)   # :316-320
        eps = (r1 * x - x0) / r2                                                      # :304-308
    else:
        raise ValueError(objective)
    return eps, x0


def ddpm_step(tab, objective: str, x: Tensor, t: int, out: Tensor, z: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """p_sample / p_mean_variance / q_posterior, denoising_diffusion_pytorch.py:322-329,356-373.
    ``z`` is the injected N(0,1) draw for this step (ignored at t == 0, as the reference uses 0 there)."""
    _, x0 = predictions(tab, objective, x, t, out, clip=False)
    x0 = x0.clamp(-1.0, 1.0)
    mean = tab["posterior_mean_coef1"][t] * x0 + tab["posterior_mean_coef2"][t] * x
    if t > 0:
        return mean + (0.5 * tab["posterior_log_variance_clipped"][t]).exp() * z, x0
    return mean + 0.0, x0


def p_losses(sd: SD, tab, objective: str, x_start: Tensor, t: Tensor, condition, noise: Tensor) -> Tensor:
    """Training loss value, denoising_diffusion_pytorch.py:481-531 (q_sample :473-479, predict_v :310-314); per-sample ``t``.
    The reference's offset noise is off by default (strength 0) and its pred_x0 prints are not restated."""
    def ex(name):
        return tab[name][t].reshape(-1, 1, 1, 1)
    x = ex("sqrt_alphas_cumprod") * x_start + ex("sqrt_one_minus_alphas_cumprod") * noise
    out = net_forward(sd, x, t, condition)
    if objective == "pred_noise":
        target = noise
    elif objective == "pred_x0":
        target = x_start
    elif objective == "pred_v":
        target = ex("sqrt_alphas_cumprod") * noise - ex("sqrt_one_minus_alphas_cumprod") * x_start
    else:
        raise ValueError(objective)
    loss = ((out - target) ** 2).flatten(1).mean(dim=1) * tab["loss_weight"][t]
    if objective == "pred_x0":
        return loss.mean() + (out.mean(dim=(2, 3)) - target.mean(dim=(2, 3))).abs().mean()
    return loss.mean()


def loss_gradients(sd: SD, tab, objective: str, x_start: Tensor, t: Tensor, condition, noise: Tensor) -> Tuple[Tensor, SD]:
    """``loss.backward()`` of the training step (models/trainer_diffusion.py:179-188): autograd through this restatement of
    p_losses.  Returns (loss, {name: dloss/dparam}); parameters the loss does not reach (the dead ``attn.to_q / to_k / norm1``
    of every AttnBlock — 1-token softmax, Diffusion_arch.py:361-402) get zeros, where the reference leaves ``.grad = None``."""
    leaf = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    with torch.enable_grad():
        loss = p_losses(leaf, tab, objective, x_start, t, condition, noise)
        loss.backward()
    return loss.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}


def adam_step(params: SD, grads: SD, state: Dict[str, Dict[str, Tensor]], lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999,
              eps: float = 1e-8, weight_decay: float = 0.0) -> SD:
    """One ``torch.optim.Adam.step()`` as Trainer.train runs it (models/trainer_diffusion.py:92,189; lr / weight_decay from
    train_diffusion.py:85-87), written out: m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2;
    p <- p - lr/(1-b1^k) * m / (sqrt(v)/sqrt(1-b2^k) + eps).  ``state`` carries (step, m, v) per parameter between calls."""
    out = {}
    for k, p in params.items():
        g = grads[k]
        if weight_decay:
            g = g + weight_decay * p
        st = state.setdefault(k, {"step": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["step"] += 1
        st["m"] = beta1 * st["m"] + (1 - beta1) * g
        st["v"] = beta2 * st["v"] + (1 - beta2) * g * g
        denom = st["v"].sqrt() / math.sqrt(1 - beta2 ** st["step"]) + eps
        out[k] = p - (lr / (1 - beta1 ** st["step"])) * st["m"] / denom
    return out


def ema_update(params: SD, st: Dict, beta: float = 0.995, update_after_step: int = 500, update_every: int = 20,
               inv_gamma: float = 1.0, power: float = 2.0 / 3.0, min_value: float = 0.0) -> None:
    """One ``ema.update()`` of the training loop (models/trainer_diffusion.py:63-69,191).  The class is the pip package
    ``ema_pytorch`` — NOT vendored in the reference tree and unpinned in install.sh:12, so this part of the oracle is restated
    from the package's published algorithm (ema-pytorch 0.2 - 0.7, ``EMA.update`` / ``get_current_decay`` /
    ``update_moving_average`` with the constructor defaults inv_gamma = 1, power = 2/3, min_value = 0) and is PARITY-UNPINNED:
    there is nothing in /root/reference to check it against.  ``st`` = {"step": int, "initted": bool, "ema": SD}, updated in place.

        step = self.step; self.step += 1
        if step % update_every != 0: return
        if step <= update_after_step: ema <- params; return
        if not initted: ema <- params; initted = True
        epoch = max(self.step - update_after_step - 1, 0)
        decay = 0 if epoch <= 0 else clamp(1 - (1 + epoch / inv_gamma) ** -power, min_value, beta)
        ema.lerp_(params, 1 - decay)
    """
    st.setdefault("step", 0)
    st.setdefault("initted", False)
    if "ema" not in st:
        st["ema"] = {k: v.clone() for k, v in params.items()}       # EMA.__init__ deep-copies the model
    step = st["step"]
    st["step"] += 1
    if step % update_every != 0:
        return
    if step <= update_after_step:
        st["ema"] = {k: v.clone() for k, v in params.items()}
        return
    if not st["initted"]:
        st["ema"] = {k: v.clone() for k, v in params.items()}
        st["initted"] = True
    epoch = max(st["step"] - update_after_step - 1, 0)
    decay = 0.0 if epoch <= 0 else min(max(1.0 - (1.0 + epoch / inv_gamma) ** -power, min_value), beta)
    for k, v in params.items():
        if v.is_floating_point():
            st["ema"][k] = torch.lerp(st["ema"][k], v, 1.0 - decay)


def cosine_annealing_lr(base_lr: float, epoch: int, t_max: int, eta_min: float = 0.0) -> float:
    """``CosineAnnealingLR(optimizer, T_max=max_iter)`` stepped once per epoch (models/trainer_diffusion.py:93,152-154), closed
    form of torch's recurrence."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / t_max)) / 2.0


def ddim_pairs(T: int, S: int) -> List[Tuple[int, int]]:
    """denoising_diffusion_pytorch.py:409-411."""
    times = torch.linspace(-1, T - 1, steps=S + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_step(tab, objective: str, x: Tensor, t: int, t_next: int, out: Tensor, z: Optional[Tensor], eta: float):
    """One iteration of ddim_sample, denoising_diffusion_pytorch.py:418-439."""
    eps, x0 = predictions(tab, objective, x, t, out, clip=True)
    if t_next < 0:
        return x0, x0
    a, an = tab["alphas_cumprod"][t], tab["alphas_cumprod"][t_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return x0 * an.sqrt() + c * eps + sigma * z, x0


def sample_chain(sd: SD, condition, x_T: Tensor, noises: Sequence[Tensor], *, T: int, schedule: str = "sigmoid2",
                 objective: str = "pred_v", sampling_steps: Optional[int] = None, eta: float = 0.0,
                 teacher: Optional[Sequence[Tensor]] = None) -> List[Tensor]:
    """Full reverse chain with an injected noise stream; returns [x_T, x_{T-1}, ..., x_0] (all timesteps).
    DDPM: p_sample_loop :375-402 (``noises[i]`` is the draw of the i-th loop iteration, t = T-1-i; the last
    one is unused).  DDIM when ``sampling_steps < T``: ddim_sample :404-444 (one draw per non-final pair).
    ``teacher``: if given, the network input at iteration i is teacher[i] instead of the running state
    (teacher-forced per-step parity)."""
    tab = schedule_tables(schedule, T, objective)
    b = x_T.shape[0]
    xs = [x_T]
    x = x_T
    if sampling_steps is None or sampling_steps >= T:
        for i, t in enumerate(reversed(range(T))):
            xin = teacher[i] if teacher is not None else x
            out = net_forward(sd, xin, torch.full((b,), t, dtype=torch.long, device=x_T.device), condition)
            x, _ = ddpm_step(tab, objective, xin, t, out, noises[i] if t > 0 else None)
            xs.append(x)
    else:
        for i, (t, tn) in enumerate(ddim_pairs(T, sampling_steps)):
            xin = teacher[i] if teacher is not None else x
            out = net_forward(sd, xin, torch.full((b,), t, dtype=torch.long, device=x_T.device), condition)
            x, _ = ddim_step(tab, objective, xin, t, tn, out, noises[i] if tn >= 0 else None, eta)
            xs.append(x)
    return xs


# --------------------------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench
# --------------------------------------------------------------------------------------------------------------

def make_position(H: int, W: int, x0: int = 0, y0: int = 0, full_h: int = 1424, full_w: int = 2128) -> Tensor:
    """utils/util.py:138-147 make_coord(rescale=True) cropped as dataloader/dataset.py:242-281 -> (2,H,W):
    ch0 = row/(full_h-1), ch1 = col/(full_w-1)."""
    rows = (torch.arange(y0, y0 + H).float() / (full_h - 1))[:, None].expand(H, W)
    cols = (torch.arange(x0, x0 + W).float() / (full_w - 1))[None, :].expand(H, W)
    return torch.stack((rows, cols), dim=0).contiguous()


def synthetic_condition(B: int, H: int, W: int, seed: int = 1, iso_idx: int = 24) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    clean = torch.rand((B, 4, H, W), generator=g) * 0.3
    origins = tile_origins(H)
    pos = torch.stack([make_position(H, W, *origins[i % len(origins)]) for i in range(B)])
    return {"clean_img": clean, "position": pos, "iso_ratio_idx": torch.full((B,), iso_idx, dtype=torch.long)}


def tile_origins(ps: int, full_h: int = 1424, full_w: int = 2128) -> List[Tuple[int, int]]:
    """dataloader/dataset.py:203-219 — (x, y) origins of the overlapping tile grid (step = ps - ps//4)."""
    step = ps - ps // 4
    def axis(n):
        v = list(range(0, n - ps + 1, step))
        if n - (v[-1] + ps) < ps:
            v.append(n - ps)
        return v
    return [(x, y) for y in axis(full_h) for x in axis(full_w)]
