"""``GaussianDiffusion`` — drop-in for the reference class (``models/denoising_diffusion_pytorch.py:167-542``).

Same constructor keywords, same 13 fp32 schedule buffers, same ``sample / p_sample_loop / ddim_sample / p_sample /
model_predictions / q_sample`` signatures.  ``sample()`` does not loop in Python over ~1.5k torch kernels per step:
it hands the per-step scalars (taken from the very same fp32 buffers) to the C-ABI engine, which replays one CUDA
graph per reverse step (network + posterior update fused).  Patches are independent, so a batch is processed in
micro-batches sized to keep producer->consumer activations resident in the B200's L2.

Host code here is plumbing (schedule tables, RNG stream, chunking); it never computes network arithmetic.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import List, Optional, Tuple

import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .arch import NoiseDiffNet

__all__ = ["GaussianDiffusion", "ModelPrediction"]

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start"])


def _draws(flag: bool):
    """ndiff_step.reserved[0] on the HOST side: this step consumes one (B, C, H, W) draw from the caller's generator in
    noise_source='torch' mode.  The library ignores the field (it multiplies whatever noise it is handed by sigma)."""
    return (C.c_int32 * 2)(1 if flag else 0, 0)


def _gather(a: torch.Tensor, t: torch.Tensor, ndim: int) -> torch.Tensor:     # ref extract :91-94
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


def _sigmoid_alphas_cumprod(T: int, start: float, end: float, tau: float) -> torch.Tensor:
    u = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
    lo, hi = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
    ac = (hi - ((u * (end - start) + start) / tau).sigmoid()) / (hi - lo)
    return ac / ac[0]


def make_betas(name: str, T: int, **kw) -> torch.Tensor:
    """float64 beta schedules, ref :96-164; unknown names (including the ctor default 'sigmoid') raise as in ref :218."""
    def from_ac(ac):
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    if name == "linear":
        k = 1000 / T
        return torch.linspace(k * 0.0001, k * 0.02, T, dtype=torch.float64)
    if name == "cosine":
        s = kw.get("s", 0.008)
        u = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
        ac = torch.cos((u + s) / (1 + s) * math.pi * 0.5) ** 2
        return from_ac(ac / ac[0])
    presets = {"sigmoid1": (-3, 3, 0.5), "sigmoid2": (-7, 3, 0.7), "sigmoid3": (-10, 3, 0.7)}
    if name in presets:
        s, e, tau = presets[name]
        return from_ac(_sigmoid_alphas_cumprod(T, kw.get("start", s), kw.get("end", e), kw.get("tau", tau)))
    raise ValueError(f"unknown beta schedule {name}")


class GaussianDiffusion(nn.Module):
    #: steps per engine call; also the granularity at which torch-RNG noise is pre-drawn
    chunk_steps = 25
    #: patches processed together by one engine (L2-residency knob)
    micro_batch = 64
    #: "torch": draw x_T and every z_t from torch's global CUDA generator in the reference's order (ref :381,:371);
    #: "philox": in-kernel counter-based Philox4x32-10, seeded from torch's generator
    noise_source = "torch"

    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, objective="pred_v",
                 beta_schedule="sigmoid", schedule_fn_kwargs=dict(), ddim_sampling_eta=0., auto_normalize=False,
                 offset_noise_strength=0., min_snr_gamma=5):
        super().__init__()
        net = model.module if isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel)) else model
        assert not (type(self) == GaussianDiffusion and net.channels != net.out_dim)
        assert not net.random_or_learned_sinusoidal_cond
        self.model = model
        self.channels = net.channels
        self.self_condition = net.self_condition
        self.image_size = image_size
        self.objective = objective
        assert objective in {"pred_noise", "pred_x0", "pred_v"}, \
            "objective must be either pred_noise (predict noise) or pred_x0 (predict image start) or pred_v (predict v)"

        betas = make_betas(beta_schedule, timesteps, **schedule_fn_kwargs)
        alphas = 1. - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = F.pad(ac[:-1], (1, 0), value=1.)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = ddim_sampling_eta

        post_var = betas * (1. - ac_prev) / (1. - ac)
        snr = ac / (1 - ac)
        tables = dict(
            betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
            sqrt_alphas_cumprod=torch.sqrt(ac), sqrt_one_minus_alphas_cumprod=torch.sqrt(1. - ac),
            log_one_minus_alphas_cumprod=torch.log(1. - ac), sqrt_recip_alphas_cumprod=torch.sqrt(1. / ac),
            sqrt_recipm1_alphas_cumprod=torch.sqrt(1. / ac - 1), posterior_variance=post_var,
            posterior_log_variance_clipped=torch.log(post_var.clamp(min=1e-20)),
            posterior_mean_coef1=betas * torch.sqrt(ac_prev) / (1. - ac),
            posterior_mean_coef2=(1. - ac_prev) * torch.sqrt(alphas) / (1. - ac),
            loss_weight={"pred_noise": snr / snr, "pred_x0": snr, "pred_v": snr / (snr + 1)}[objective])
        for k, v in tables.items():
            self.register_buffer(k, v.to(torch.float32))
        self.offset_noise_strength = offset_noise_strength
        self.normalize = (lambda x: x * 2 - 1) if auto_normalize else (lambda x, *a, **k: x)
        self.unnormalize = (lambda x: (x + 1) * 0.5) if auto_normalize else (lambda x, *a, **k: x)

    # ------------------------------------------------------------------------------------------------------------
    @property
    def device(self):
        return self.betas.device

    def _net(self) -> NoiseDiffNet:
        m = self.model
        if isinstance(m, (nn.DataParallel, nn.parallel.DistributedDataParallel)):
            m = m.module
        if not isinstance(m, NoiseDiffNet):
            raise TypeError("noisediff_b200.GaussianDiffusion.sample drives noisediff_b200.NoiseDiffNet only")
        return m

    # ---- elementwise helpers with the reference's formulas (ref :298-329) -------------------------------------------
    def predict_start_from_noise(self, x_t, t, noise):
        return _gather(self.sqrt_recip_alphas_cumprod, t, x_t.dim()) * x_t - \
            _gather(self.sqrt_recipm1_alphas_cumprod, t, x_t.dim()) * noise

    def predict_noise_from_start(self, x_t, t, x0):
        return (_gather(self.sqrt_recip_alphas_cumprod, t, x_t.dim()) * x_t - x0) / \
            _gather(self.sqrt_recipm1_alphas_cumprod, t, x_t.dim())

    def predict_v(self, x_start, t, noise):
        return _gather(self.sqrt_alphas_cumprod, t, x_start.dim()) * noise - \
            _gather(self.sqrt_one_minus_alphas_cumprod, t, x_start.dim()) * x_start

    def predict_start_from_v(self, x_t, t, v):
        return _gather(self.sqrt_alphas_cumprod, t, x_t.dim()) * x_t - \
            _gather(self.sqrt_one_minus_alphas_cumprod, t, x_t.dim()) * v

    def q_posterior(self, x_start, x_t, t):
        mean = _gather(self.posterior_mean_coef1, t, x_t.dim()) * x_start + \
            _gather(self.posterior_mean_coef2, t, x_t.dim()) * x_t
        return mean, _gather(self.posterior_variance, t, x_t.dim()), \
            _gather(self.posterior_log_variance_clipped, t, x_t.dim())

    def model_predictions(self, x, t, condition=None, clip_x_start=False, rederive_pred_noise=False):
        out = self.model(x, t, condition)
        clip = (lambda z: z.clamp(-1., 1.)) if clip_x_start else (lambda z: z)
        if self.objective == "pred_noise":
            eps = out
            x0 = clip(self.predict_start_from_noise(x, t, eps))
            if clip_x_start and rederive_pred_noise:
                eps = self.predict_noise_from_start(x, t, x0)
        elif self.objective == "pred_x0":
            x0 = clip(out)
            eps = self.predict_noise_from_start(x, t, x0)
        else:
            x0 = clip(self.predict_start_from_v(x, t, out))
            eps = self.predict_noise_from_start(x, t, x0)
        return ModelPrediction(eps, x0)

    def p_mean_variance(self, x, t, condition=None, clip_denoised=True):
        x0 = self.model_predictions(x, t, condition).pred_x_start
        if clip_denoised:
            x0 = x0.clamp(-1., 1.)
        mean, var, logvar = self.q_posterior(x_start=x0, x_t=x, t=t)
        return mean, var, logvar, x0

    @torch.inference_mode()
    def p_sample(self, x, t: int, condition=None):
        """One reverse step through the public per-step API (ref :366-373): network on the GPU library, the four
        elementwise lines in torch.  ``sample()`` does not go through here."""
        tt = torch.full((x.shape[0],), t, device=x.device, dtype=torch.long)
        mean, _, logvar, x0 = self.p_mean_variance(x=x, t=tt, condition=condition, clip_denoised=True)
        noise = torch.randn_like(x) if t > 0 else 0.
        return mean + (0.5 * logvar).exp() * noise, x0

    # ---- step tables for the engine ------------------------------------------------------------------------------
    def _xstart_coefs(self, t: int, tab) -> Tuple[float, float]:
        if self.objective == "pred_v":
            return float(tab["sqrt_alphas_cumprod"][t]), -float(tab["sqrt_one_minus_alphas_cumprod"][t])
        if self.objective == "pred_noise":
            return float(tab["sqrt_recip_alphas_cumprod"][t]), -float(tab["sqrt_recipm1_alphas_cumprod"][t])
        return 0.0, 1.0

    def _cpu_tables(self):
        return {k: v.detach().float().cpu() for k, v in self.named_buffers(recurse=False)}

    def ddpm_steps(self) -> List[_lib.Step]:
        """p_sample_loop order (ref :394): t = T-1 ... 0; sigma = exp(0.5*logvar_t), 0 at t == 0 (ref :371-372)."""
        tab = self._cpu_tables()
        out = []
        for t in reversed(range(self.num_timesteps)):
            p, q = self._xstart_coefs(t, tab)
            sigma = float((0.5 * tab["posterior_log_variance_clipped"][t]).exp()) if t > 0 else 0.0
            out.append(_lib.Step(t=t, p=p, q=q, a=float(tab["posterior_mean_coef1"][t]),
                                 b=float(tab["posterior_mean_coef2"][t]), c=0.0,
                                 r1=float(tab["sqrt_recip_alphas_cumprod"][t]),
                                 r2=float(tab["sqrt_recipm1_alphas_cumprod"][t]), sigma=sigma, clip=1,
                                 reserved=_draws(t > 0)))                 # ref :371: randn_like iff t > 0
        return out

    def ddim_time_pairs(self) -> List[Tuple[int, int]]:
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)   # ref :409-411
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def ddim_steps(self) -> List[_lib.Step]:
        tab = self._cpu_tables()
        eta = self.ddim_sampling_eta
        out = []
        for t, tn in self.ddim_time_pairs():
            p, q = self._xstart_coefs(t, tab)
            r1, r2 = float(tab["sqrt_recip_alphas_cumprod"][t]), float(tab["sqrt_recipm1_alphas_cumprod"][t])
            if tn < 0:                                                    # ref :422-425: img = x_start
                out.append(_lib.Step(t=t, p=p, q=q, a=1.0, b=0.0, c=0.0, r1=r1, r2=r2, sigma=0.0, clip=1, reserved=_draws(False)))
                continue
            alpha, alpha_next = tab["alphas_cumprod"][t], tab["alphas_cumprod"][tn]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()      # ref :430-431
            c = (1 - alpha_next - sigma ** 2).sqrt()
            out.append(_lib.Step(t=t, p=p, q=q, a=float(alpha_next.sqrt()), b=0.0, c=float(c), r1=r1, r2=r2,
                                 sigma=float(sigma), clip=1, reserved=_draws(True)))   # ref :433: every non-final pair draws, even at eta = 0
        return out

    # ---- the fast path -------------------------------------------------------------------------------------------
    def _run_chain(self, steps: List[_lib.Step], shape, condition, x_T: Optional[torch.Tensor], return_all: bool,
                   noises: Optional[torch.Tensor] = None, teacher: Optional[torch.Tensor] = None):
        """Executes `steps` for a batch of independent patches, micro-batch by micro-batch, chunk by chunk.
        noises / teacher (tests): [n_steps, B, 4, H, W] injected draws / teacher-forced inputs."""
        net = self._net()
        B, Cc, H, W = shape
        dev = self.device
        n_steps = len(steps)
        if B == 0:                      # the reference runs an empty batch through and returns an empty result
            return torch.empty((0, n_steps + 1, Cc, H, W) if return_all else (0, Cc, H, W), device=dev)
        mb = min(B, int(self.micro_batch))
        net.engine_for(mb, H, W, dev)       # raises off-GPU / without the library before any RNG is consumed: there is no CPU path
        groups = [(lo, min(lo + mb, B)) for lo in range(0, B, mb)]
        use_torch_rng = self.noise_source == "torch" and noises is None
        if self.noise_source not in ("torch", "philox"):
            raise ValueError("noise_source must be 'torch' or 'philox'")
        base_seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (self.noise_source == "philox") else 0

        if x_T is None and (use_torch_rng or noises is not None):
            x_T = torch.randn(shape, device=dev)
        state = [x_T[lo:hi].contiguous() if x_T is not None else None for lo, hi in groups]
        conds = [(condition["clean_img"][lo:hi], condition["position"][lo:hi], condition["iso_ratio_idx"][lo:hi])
                 for lo, hi in groups]
        snaps, xT_parts = [], [None] * len(groups)
        done = 0
        started = [False] * len(groups)
        while done < n_steps:
            n = min(int(self.chunk_steps), n_steps - done)
            chunk_noise = None
            if use_torch_rng:
                # the global generator must end up where the reference leaves it: DDPM draws randn_like for every t > 0
                # (ref :371), DDIM for every non-final pair whatever sigma is (ref :433, also at the default eta = 0)
                chunk_noise = torch.zeros((n, B, Cc, H, W), device=dev)
                for i in range(n):
                    if steps[done + i].reserved[0]:
                        torch.randn((B, Cc, H, W), device=dev, out=chunk_noise[i])
            elif noises is not None:
                chunk_noise = noises[done:done + n]
            snaps_full = torch.empty((n, B, Cc, H, W), device=dev) if return_all else None
            for gi, (lo, hi) in enumerate(groups):
                pad = mb - (hi - lo)                      # the last micro-batch may be short: pad by repetition
                def fit(t, dim=0):
                    if pad == 0:
                        return t.contiguous()
                    idx = [slice(None)] * t.dim()
                    idx[dim] = slice(0, 1)
                    rep = [1] * t.dim()
                    rep[dim] = pad
                    return torch.cat([t, t[tuple(idx)].repeat(*rep)], dim=dim).contiguous()
                eng = net.engine_for(mb, H, W, dev)
                c, p, i = conds[gi]
                if len(groups) > 1 or not started[gi]:     # one micro-batch owns its engine for the whole chain: set once
                    eng.set_condition(fit(c.to(dev)), fit(p.to(dev)), fit(i.to(dev)))
                if not started[gi]:
                    eng.chain_begin(steps, fit(state[gi]) if state[gi] is not None else None, base_seed + gi)
                    started[gi] = True
                    if return_all and x_T is None:        # x_T was drawn in the library (Philox): read it back for slot 0
                        xT_parts[gi] = eng.chain_read()[: hi - lo]
                elif len(groups) > 1:
                    eng.chain_seek(done, fit(state[gi]), base_seed + gi)
                nz = fit(chunk_noise[:, lo:hi], 1) if chunk_noise is not None else None
                tf = fit(teacher[done:done + n, lo:hi].to(dev), 1) if teacher is not None else None
                sn = torch.empty((n, mb, Cc, H, W), device=dev) if return_all else None
                eng.chain_run(n, nz, tf, sn)
                if len(groups) > 1 or done + n >= n_steps:
                    state[gi] = eng.chain_read()[: hi - lo]
                if return_all:
                    snaps_full[:, lo:hi] = sn[:, : hi - lo]
            if return_all:
                snaps.extend(snaps_full[i] for i in range(n))
            done += n
        if return_all:                                    # (B, n_steps + 1, C, H, W), slot 0 = x_T (ref :389-399)
            first = x_T if x_T is not None else torch.cat(xT_parts, dim=0)
            return torch.stack([first] + snaps, dim=1)
        return torch.cat(state, dim=0)

    @torch.inference_mode()
    def p_sample_loop(self, shape, condition=None, return_all_timesteps=False, preset_mean=None):
        x_T = None
        if preset_mean is not None:
            if self.noise_source == "torch":
                torch.randn(shape, device=self.device)        # drawn and discarded, as the reference does (ref :381-387)
            x_T = preset_mean
        ret = self._run_chain(self.ddpm_steps(), tuple(shape), condition, x_T, return_all_timesteps)
        return self.unnormalize(ret)

    @torch.inference_mode()
    def ddim_sample(self, shape, condition=None, return_all_timesteps=False, preset_mean=None):
        ret = self._run_chain(self.ddim_steps(), tuple(shape), condition, None, return_all_timesteps)
        return self.unnormalize(ret)

    @torch.inference_mode()
    def sample(self, batch_size=16, condition=None, return_all_timesteps=False, preset_mean=None):
        fn = self.ddim_sample if self.is_ddim_sampling else self.p_sample_loop
        return fn((batch_size, self.channels, self.image_size, self.image_size), condition=condition,
                  return_all_timesteps=return_all_timesteps, preset_mean=preset_mean)

    @torch.inference_mode()
    def interpolate(self, x1, x2, t=None, lam=0.5):
        raise NotImplementedError("interpolate() passes self_cond as the condition in the reference (ref :468-469), "
                                  "which NoiseDiffNet cannot consume; it is unreachable there too")

    # ---- training-side API: signatures kept; the loss is evaluated forward-only (backward is not built) ----
    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return _gather(self.sqrt_alphas_cumprod, t, x_start.dim()) * x_start + \
            _gather(self.sqrt_one_minus_alphas_cumprod, t, x_start.dim()) * noise

    def p_losses(self, x_start, t, condition=None, noise=None, offset_noise_strength=None):
        """Loss VALUE of the training objective (ref :481-531): q_sample -> network (per-sample t) -> weighted MSE against the
        objective's target.  Forward half of SURVEY.md §8f N1 only — validation / loss curves under ``torch.no_grad()``; the
        library has no backward kernels yet, so asking for gradients raises instead of silently returning a detached loss."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            raise NotImplementedError("diffusion training (forward+backward) is the next scope row (SURVEY.md §8f N1); "
                                      "the B200 library evaluates the loss only under torch.no_grad()")
        noise = torch.randn_like(x_start) if noise is None else noise
        ons = self.offset_noise_strength if offset_noise_strength is None else offset_noise_strength
        if ons > 0.:                                                                   # ref :490-492
            noise = noise + ons * torch.randn(x_start.shape[:2], device=self.device)[:, :, None, None]
        x = self.q_sample(x_start=x_start, t=t, noise=noise)
        out = self.model(x, t, condition)
        if self.objective == "pred_noise":
            target = noise
        elif self.objective == "pred_x0":
            target = x_start
        else:
            target = self.predict_v(x_start, t, noise)
        loss = F.mse_loss(out, target, reduction="none").flatten(1).mean(dim=1)        # ref :518-519
        loss = loss * self.loss_weight.gather(-1, t)
        if self.objective == "pred_x0":                                                # ref :522-526 (its prints are dropped)
            return loss.mean() + (out.mean(dim=(2, 3)) - target.mean(dim=(2, 3))).abs().mean()
        return loss.mean()

    def forward(self, img, condition, *args, **kwargs):
        b, c, h, w = img.shape
        assert h == self.image_size and w == self.image_size, f"height and width of image must be {self.image_size}"
        t = torch.randint(0, self.num_timesteps, (b,), device=img.device).long()
        return self.p_losses(self.normalize(img), t, condition, *args, **kwargs)
