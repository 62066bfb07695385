"""``NoiseDiffNet`` — drop-in for the reference network class (``models/archs/Diffusion_arch.py:447-646``).

Same constructor ``NoiseDiffNet(args)`` (reads ``args.dim, cond_dim, inp_dim, self_condition, normalize_condition``,
ref :453-470), same attributes the sampler reads (``channels, out_dim, self_condition,
random_or_learned_sinusoidal_cond``, ``models/denoising_diffusion_pytorch.py:184-198``), same 416-key
``state_dict`` (names, shapes and — because sub-modules are created in the reference's order with torch's
default initialisers — the same values for a given ``torch.manual_seed``), same
``forward(x, time, condition) -> (B, out_dim, H, W)``.

The module tree below only *holds parameters*; all arithmetic runs in the sm_100a CUDA library through the C ABI
(``noisediff_b200/csrc``).  There is no PyTorch/CPU fallback: calling ``forward`` without the library or off-GPU
raises.
"""
from __future__ import annotations

from functools import partial

import torch
from torch import nn

from . import engine as _engine

__all__ = ["NoiseDiffNet"]


class _Tag(nn.Module):
    """Parameter-free placeholder that keeps ``nn.Sequential`` indices equal to the reference's."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class _Block(nn.Module):                       # ref Block :128-144
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.proj = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm = nn.GroupNorm(groups, cout)


class _ResnetBlock(nn.Module):                 # ref ResnetBlock :146-170 / ResnetBlock2 :173-196
    def __init__(self, cin, cout, *, time_emb_dim=None, pos_emb_dim=None, groups=8):
        super().__init__()
        if time_emb_dim is not None:
            self.mlp = nn.Sequential(_Tag("SiLU"), nn.Linear(time_emb_dim, cout * 2))
        else:
            self.mlp = nn.Sequential(_Tag("SiLU"), nn.Conv2d(pos_emb_dim, cout * 2, 1))
        self.block1 = _Block(cin, cout, groups)
        self.block2 = _Block(cout, cout, groups)
        self.res_conv = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()
        self.groups = groups


class _CrossAttention(nn.Module):              # ref CrossAttention :361-402 (to_q/to_k are dead on this path)
    def __init__(self, qdim, cdim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.to_q = nn.Linear(qdim, inner, bias=False)
        self.to_k = nn.Linear(cdim, inner, bias=False)
        self.to_v = nn.Linear(cdim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, qdim), _Tag("Dropout(0)"))


class _FeedForward(nn.Module):                 # ref FeedForward :405-422
    def __init__(self, dim, mult=2):
        super().__init__()
        self.net = nn.Sequential(nn.Sequential(nn.Linear(dim, dim * mult), _Tag("GELU")), _Tag("Dropout(0)"),
                                 nn.Linear(dim * mult, dim))


class _AttnBlock(nn.Module):                   # ref AttnBlock :425-443
    def __init__(self, qdim, cdim, heads=4, dim_head=32):
        super().__init__()
        self.attn = _CrossAttention(qdim, cdim, heads, dim_head)
        self.norm1 = nn.LayerNorm(qdim)
        self.norm2 = nn.LayerNorm(qdim)
        self.ff = _FeedForward(qdim)
        self.proj_out = nn.Conv2d(qdim, qdim, 1)


class _Mlp(nn.Module):                         # ref Mlp :340-356
    def __init__(self, cin, hidden, cout):
        super().__init__()
        self.fc1 = nn.Conv2d(cin, hidden, 1)
        self.fc2 = nn.Conv2d(hidden, cout, 1)


class _PosEnc(nn.Module):                      # ref LearnedSinusoidalPosEmb :322-337
    def __init__(self, cin, hidden):
        super().__init__()
        self.weights = nn.Conv2d(cin, hidden, 1)


class NoiseDiffNet(nn.Module):
    #: engines kept alive per network (one per (device, batch, H, W) seen); each owns GiBs of activation buffers at the bench
    #: geometry, so a loop over ragged batch sizes must not accumulate them
    max_engines = 4

    def __init__(self, args):
        super().__init__()
        dim = int(args.dim)
        self.dim = dim
        self.cond_dim = getattr(args, "cond_dim", 4)
        self.channels = int(args.inp_dim)
        self.out_dim = self.channels
        self.self_condition = args.self_condition
        self.normalize_condition = args.normalize_condition
        self.random_or_learned_sinusoidal_cond = False
        if self.channels != 4:
            raise ValueError("the B200 path implements packed 4-channel Bayer input (inp_dim=4)")
        iso_dim, pos_dim, time_dim = 16, 8, dim * 4
        dims = [dim, dim, dim * 2, dim * 4, dim * 8]
        in_out = list(zip(dims[:-1], dims[1:]))
        rb = partial(_ResnetBlock, time_emb_dim=time_dim, groups=8)

        # creation order == reference order (ref :479-573) so default init consumes the RNG identically
        self.init_conv = nn.Conv2d(self.channels, dim, 7, padding=3)
        self.iso_embed = nn.Embedding(100, iso_dim)
        self.time_mlp = nn.Sequential(_Tag("SinusoidalPosEmb"), nn.Linear(dim, time_dim), _Tag("GELU"),
                                      nn.Linear(time_dim, time_dim))
        self.downs = nn.ModuleList()
        self.ups = nn.ModuleList()
        for i, (ci, co) in enumerate(in_out):
            last = i == len(in_out) - 1
            self.downs.append(nn.ModuleList([
                rb(ci, ci), rb(ci, ci), _AttnBlock(ci, iso_dim),
                nn.Conv2d(ci, co, 3, padding=1) if last else nn.Sequential(_Tag("space_to_depth"), nn.Conv2d(ci * 4, co, 1))]))
        mid = dims[-1]
        self.mid_block1 = rb(mid, mid)
        self.mid_block2 = rb(mid, mid)
        for i, (ci, co) in enumerate(reversed(in_out)):
            last = i == len(in_out) - 1
            self.ups.append(nn.ModuleList([
                rb(co + ci, co), rb(co + ci, co), _AttnBlock(co, iso_dim),
                nn.Conv2d(co, ci, 3, padding=1) if last else nn.Sequential(_Tag("nearest_x2"), nn.Conv2d(co, ci, 3, padding=1))]))
        self.final_res_block = rb(dim * 2, dim)
        self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        self.pos_enc = _PosEnc(2, pos_dim)
        self.pos_mlp = _Mlp(pos_dim * 3, pos_dim * 2, pos_dim)
        self.pos_block1 = _ResnetBlock(dim, dim, pos_emb_dim=pos_dim, groups=2)
        self.pos_block2 = _ResnetBlock(dim, dim, pos_emb_dim=pos_dim, groups=2)
        self.shot_mlp1 = _Mlp(8, dim, dim)
        self.shot_attn = _AttnBlock(dim, iso_dim)
        self.shot_mlp2 = _Mlp(dim, dim, dim)
        self.shot_time = _ResnetBlock(dim, dim, time_emb_dim=time_dim, groups=2)
        self.shot_mlp3 = _Mlp(dim, dim, 4)

        self._engines = {}          # (device index, B, H, W) -> engine.Engine, least recently used first
        self._weights_version = None

    @property
    def downsample_factor(self):
        return 2 ** (len(self.downs) - 1)

    # ---- engine management -----------------------------------------------------------------------------------
    def _param_version(self):
        """Identity of the current weights: storage address + in-place version counter of every parameter.  The parameter LIST is
        cached (the module-tree walk is what costs on the per-step ``p_sample`` API) and dropped whenever the module's tensors are
        replaced (``.to()/.cuda()/.half()`` go through ``_apply``; ``load_state_dict`` copies in place and bumps the versions)."""
        ps = self.__dict__.get("_param_list")
        if ps is None:
            ps = list(self.parameters())
            self.__dict__["_param_list"] = ps
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _apply(self, fn, *a, **kw):
        self.__dict__.pop("_param_list", None)
        return super()._apply(fn, *a, **kw)

    def engine_for(self, batch: int, height: int, width: int, device: torch.device) -> "_engine.Engine":
        """Returns the (cached) CUDA engine for this geometry with the current parameters uploaded."""
        if device.type != "cuda":
            raise RuntimeError("noisediff_b200.NoiseDiffNet runs only on CUDA (sm_100a); there is no CPU path")
        key = (device.index if device.index is not None else torch.cuda.current_device(), batch, height, width)
        ver = self._param_version()
        eng = self._engines.pop(key, None)
        if eng is None:
            while len(self._engines) >= max(int(self.max_engines), 1):
                self._engines.pop(next(iter(self._engines))).close()      # least recently used: frees its activation pool
            eng = _engine.Engine(dim=self.dim, batch=batch, height=height, width=width, device=key[0])
        self._engines[key] = eng                                          # (re)inserted last = most recently used
        if eng.weights_version != ver:
            eng.load_state_dict({k: v.detach() for k, v in self.state_dict().items()})
            eng.weights_version = ver
        return eng

    def release_engines(self):
        for e in self._engines.values():
            e.close()
        self._engines.clear()

    # ---- reference-facing call --------------------------------------------------------------------------------
    def forward(self, x, time, condition=None):
        f = self.downsample_factor
        assert all(d % f == 0 for d in x.shape[-2:]), \
            f"your input dimensions {tuple(x.shape[-2:])} need to be divisible by {f}, given the unet"
        if condition is None:
            raise TypeError("NoiseDiffNet.forward needs condition={'clean_img','position','iso_ratio_idx'}")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # a training call (autograd on, trainable weights) must not get a silently detached tensor back
            raise NotImplementedError("NoiseDiffNet.forward on the CUDA library is inference-only: call it under torch.no_grad() / "
                                      "inference_mode (the training step is SURVEY §8f N1)")
        B, C, H, W = x.shape
        eng = self.engine_for(B, H, W, x.device)
        eng.set_condition(condition["clean_img"], condition["position"], condition["iso_ratio_idx"])
        return eng.forward(x, time)
