"""GPU parity of the individual sm_100a kernels against plain fp32 torch ops on the same (bf16-rounded) inputs.
All calls go through the C ABI (include/noisediff_b200.h)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from noisediff_b200 import _lib
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16).float()


def _check(out_nhwc, ref_nchw, tol=1.5e-2):
    got = G.from_nhwc(out_nhwc)
    err = (got - ref_nchw).abs().max().item()
    scale = ref_nchw.abs().max().item()
    rel = ((got - ref_nchw).norm() / ref_nchw.norm()).item()
    assert rel < 4e-3 and err <= tol * max(scale, 1.0), f"rel-L2 {rel:.3e}, max abs {err:.3e} (scale {scale:.3e})"


# (B, H, W, Cin0, Cin1, Cout)
GEMM_CASES = [(2, 16, 16, 64, 0, 64), (1, 32, 32, 128, 0, 256), (2, 8, 8, 256, 128, 128), (1, 64, 64, 64, 64, 64),
              (3, 24, 40, 64, 0, 128), (1, 8, 8, 512, 0, 512),
              # many tiles per CTA: the weight slice stays resident in shared memory
              (4, 128, 128, 64, 0, 64), (2, 128, 128, 128, 0, 64), (2, 128, 128, 64, 64, 128), (8, 64, 64, 256, 0, 512)]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm_1x1(case):
    B, H, W, c0, c1, co = case
    x0 = _rand((B, c0, H, W), 1)
    x1 = _rand((B, c1, H, W), 2) if c1 else None
    w = _rand((co, c0 + c1, 1, 1), 3, 1.0 / math.sqrt(c0 + c1))
    bias = torch.randn(co, device="cuda")
    ref = F.conv2d(torch.cat([x0, x1], 1) if c1 else x0, w, bias)
    out = G.conv(G.MODE_DIRECT, G.to_nhwc_bf16(x0), G.pack_weight(w), co, src1=G.to_nhwc_bf16(x1) if c1 else None,
                 bias=bias)
    _check(out, ref)


def test_gemm_epilogue_gelu_residual_vector():
    B, H, W, c, co = 2, 16, 16, 128, 64
    x = _rand((B, c, H, W), 4)
    w = _rand((co, c, 1, 1), 5, 1.0 / math.sqrt(c))
    bias = torch.randn(co, device="cuda")
    res = _rand((B, co, H, W), 6)
    vec = torch.randn(B, 96, device="cuda")            # leading dimension larger than Cout on purpose
    ref = F.gelu(F.conv2d(x, w, bias)) + vec[:, :co, None, None] + res
    out = G.conv(G.MODE_DIRECT, G.to_nhwc_bf16(x), G.pack_weight(w), co, bias=bias, vec=vec, res=G.to_nhwc_bf16(res), act=1)
    _check(out, ref)


CONV3_CASES = [(8, 32, 32, 512, 256, 512), (8, 64, 64, 256, 0, 256), (2, 32, 32, 64, 0, 64), (1, 16, 16, 128, 64, 128), (1, 8, 8, 256, 0, 256), (2, 24, 40, 64, 64, 64),
               (1, 64, 64, 64, 0, 64), (1, 8, 8, 512, 256, 512), (4, 128, 128, 64, 0, 64), (2, 128, 128, 64, 64, 64),
               (1, 256, 256, 64, 0, 64), (1, 256, 256, 64, 64, 64)]


@pytest.mark.parametrize("mode", ["halo", "halo2", "direct"])
@pytest.mark.parametrize("case", CONV3_CASES)
def test_conv3x3(case, mode):
    B, H, W, c0, c1, co = case
    x0 = _rand((B, c0, H, W), 7)
    x1 = _rand((B, c1, H, W), 8) if c1 else None
    w = _rand((co, c0 + c1, 3, 3), 9, 1.0 / math.sqrt(9 * (c0 + c1)))
    bias = torch.randn(co, device="cuda")
    ref = F.conv2d(torch.cat([x0, x1], 1) if c1 else x0, w, bias, padding=1)
    kw = dict(src1=G.to_nhwc_bf16(x1) if c1 else None, bias=bias)
    if mode != "direct":
        out = G.conv(G.MODE_HALO2 if mode == "halo2" else G.MODE_HALO1, G.to_nhwc_bf16(x0), G.pack_weight(w), co, **kw)
    else:
        out = G.conv(G.MODE_DIRECT, G.to_nhwc_bf16(x0), G.pack_weight(w), co, taps=(3, 3), pad=(1, 1), **kw)
    _check(out, ref)


@pytest.mark.parametrize("tile_w", [8, 16, 32, 64])
def test_direct_conv_tile_shapes(tile_w):
    B, H, W, c, co = 1, 32, 64, 64, 128
    x = _rand((B, c, H, W), 10)
    w = _rand((co, c, 3, 3), 11, 1.0 / math.sqrt(9 * c))
    out = G.conv(G.MODE_DIRECT, G.to_nhwc_bf16(x), G.pack_weight(w), co, taps=(3, 3), pad=(1, 1), tile_w=tile_w)
    _check(out, F.conv2d(x, w, None, padding=1))


@pytest.mark.parametrize("case", [(2, 32, 32, 64, 64), (1, 16, 16, 128, 256), (1, 64, 64, 64, 128), (4, 256, 256, 64, 64)])
def test_downsample_space_to_depth(case):
    B, H, W, c, co = case                                 # H, W = input size
    x = _rand((B, c, H, W), 12)
    w = _rand((co, 4 * c, 1, 1), 13, 1.0 / math.sqrt(4 * c))
    bias = torch.randn(co, device="cuda")
    s2d = x.reshape(B, c, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 4 * c, H // 2, W // 2)
    ref = F.conv2d(s2d, w, bias)
    out = G.conv(G.MODE_S2D, G.to_nhwc_bf16(x), G.pack_weight(w, s2d=True), co, bias=bias, out_hw=(H // 2, W // 2))
    _check(out, ref)


@pytest.mark.parametrize("force_nt", [0, 64])
@pytest.mark.parametrize("case", [(2, 16, 16, 128, 64), (1, 32, 32, 512, 256), (2, 64, 64, 256, 128), (2, 128, 128, 128, 64),
                                  (1, 24, 40, 64, 64), (8, 32, 32, 512, 256)])
def test_upsample_conv3x3_phase_form(case, force_nt):
    """Upsample = nearest x2 then conv3x3 pad 1 (ref Diffusion_arch.py:72-76) as four 2x2 phase convolutions (kHaloUp)."""
    B, H, W, c, co = case                                 # H, W = low-resolution input size
    x = _rand((B, c, H, W), 30)
    w = _rand((co, c, 3, 3), 31, 1.0 / math.sqrt(9 * c))
    bias = torch.randn(co, device="cuda")
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, bias, padding=1)
    out = G.conv(G.MODE_HALO_UP, G.to_nhwc_bf16(x), G.pack_upconv_weight(w), co, bias=bias, alloc_hw=(2 * H, 2 * W),
                 force_nt=force_nt)
    _check(out, ref)


@pytest.mark.parametrize("case", [(2, 32, 32, 64, 8), (2, 16, 16, 128, 8), (1, 8, 8, 512, 8), (2, 32, 32, 64, 2),
                                  (1, 256, 256, 64, 2), (2, 256, 256, 64, 8), (4, 64, 64, 128, 8)])
def test_groupnorm_stats_and_apply(case):
    B, H, W, c, groups = case
    x = _rand((B, c, H, W), 14)
    w = _rand((c, c, 3, 3), 15, 1.0 / math.sqrt(9 * c))
    bias = torch.randn(c, device="cuda")
    gamma, beta = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
    ss = torch.randn(B, 3 * c, device="cuda") * 0.5
    res = _rand((B, c, H, W), 16)
    stats = torch.zeros(B, groups, 2, device="cuda", dtype=torch.int64)       # 2^-24 fixed point
    y = G.conv(G.MODE_HALO1, G.to_nhwc_bf16(x), G.pack_weight(w), c, bias=bias, stats=stats, groups=groups)
    stats3 = torch.zeros_like(stats)
    y3 = G.conv(G.MODE_HALO2, G.to_nhwc_bf16(x), G.pack_weight(w), c, bias=bias, stats=stats3, groups=groups)
    assert torch.equal(y3, y)
    assert ((stats - stats3).abs().double() <= 2 ** 24 * 1e-2 + 2e-6 * stats.abs().double()).all()   # fp32 partials
    stats2 = torch.zeros_like(stats)
    G.conv(G.MODE_DIRECT, G.to_nhwc_bf16(x), G.pack_weight(w), c, bias=bias, stats=stats2, groups=groups, tile_w=16,
           taps=(3, 3), pad=(1, 1))
    conv_ref = F.conv2d(x, w, bias, padding=1)
    grp = conv_ref.reshape(B, groups, -1)
    sums = stats.double() / 2 ** 24
    assert torch.allclose(sums[..., 0], grp.sum(-1).double(), rtol=2e-3, atol=2e-2 * math.sqrt(grp.shape[-1]))
    assert torch.allclose(sums[..., 1], (grp ** 2).sum(-1).double(), rtol=2e-3)
    assert ((stats - stats2).abs().double() <= 2 ** 24 * 1e-2 + 2e-6 * stats.abs().double()).all()   # tiling changes only fp32 partial rounding
    out = torch.empty_like(y)
    _lib.check(_lib.lib().ndiff_op_gn_apply(G.P(y), G.P(out), G.P(stats), G.P(gamma), G.P(beta), G.P(ss), 3 * c, c // 2,
                                            None, G.P(G.to_nhwc_bf16(res)), None, B, H * W, c, groups, G.stream()))
    torch.cuda.synchronize()
    sc, sh = ss[:, c // 2:c // 2 + c, None, None], ss[:, c // 2 + c:c // 2 + 2 * c, None, None]
    ref = F.silu(F.group_norm(conv_ref, groups, gamma, beta, eps=1e-5) * (sc + 1) + sh) + res
    _check(out, ref, tol=3e-2)


def test_groupnorm_apply_with_pixel_maps():
    B, H, W, c, groups = 2, 16, 16, 64, 2
    y = _rand((B, c, H, W), 17)
    maps = _rand((B, 2 * c, H, W), 18, 0.5)
    gamma, beta = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
    grp = y.reshape(B, groups, -1)
    stats = (torch.stack([grp.sum(-1), (grp ** 2).sum(-1)], dim=-1).double() * 2 ** 24).round().to(torch.int64).contiguous()
    yb = G.to_nhwc_bf16(y)
    out = torch.empty_like(yb)
    _lib.check(_lib.lib().ndiff_op_gn_apply(G.P(yb), G.P(out), G.P(stats), G.P(gamma), G.P(beta), None, 0, 0,
                                            G.P(G.to_nhwc_bf16(maps)), None, None, B, H * W, c, groups, G.stream()))
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(y, groups, gamma, beta, eps=1e-5) * (maps[:, :c] + 1) + maps[:, c:])
    _check(out, ref, tol=3e-2)


@pytest.mark.parametrize("c", [64, 128, 512])
def test_layernorm_with_sample_vector(c):
    B, H, W = 2, 16, 8
    x = _rand((B, c, H, W), 19)
    vec = torch.randn(B, c + 32, device="cuda")
    g, b = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
    xb = G.to_nhwc_bf16(x)
    out = torch.empty_like(xb)
    _lib.check(_lib.lib().ndiff_op_layernorm(G.P(xb), G.P(vec), c + 32, G.P(g), G.P(b), G.P(out), B, H * W, c, G.stream()))
    torch.cuda.synchronize()
    tok = x.permute(0, 2, 3, 1) + vec[:, None, None, :c]
    ref = F.layer_norm(tok, (c,), g, b, eps=1e-5).permute(0, 3, 1, 2)
    _check(out, ref, tol=3e-2)


def test_philox_normals_are_standard_normal_and_counter_based():
    n4 = 1 << 18
    a = torch.empty(n4 * 4, device="cuda")
    b = torch.empty(n4 * 4, device="cuda")
    lib = _lib.lib()
    _lib.check(lib.ndiff_op_philox_normal(G.P(a), n4, 42, 7, G.stream()))
    _lib.check(lib.ndiff_op_philox_normal(G.P(b), n4, 42, 7, G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(a, b)                                                 # deterministic in (seed, stream, index)
    _lib.check(lib.ndiff_op_philox_normal(G.P(b), n4, 42, 8, G.stream()))
    torch.cuda.synchronize()
    assert abs(float((a * b).mean())) < 5e-3                                 # streams are uncorrelated
    n = a.numel()
    assert abs(float(a.mean())) < 4 / math.sqrt(n) and abs(float(a.var()) - 1) < 6 * math.sqrt(2 / n)
    assert abs(float((a ** 3).mean())) < 0.02 and abs(float((a ** 4).mean()) - 3) < 0.05
    # Kolmogorov-Smirnov distance to the normal CDF on a subsample
    s = a[:: 64].double().sort().values
    cdf = 0.5 * (1 + torch.erf(s / math.sqrt(2)))
    emp = (torch.arange(1, s.numel() + 1, device="cuda", dtype=torch.float64)) / s.numel()
    assert float((cdf - emp).abs().max()) < 1.63 / math.sqrt(s.numel())      # 1% critical value
    assert float(a.abs().max()) < 7 and float((a.reshape(-1, 4)[:, 0] * a.reshape(-1, 4)[:, 1]).mean()) < 5e-3


# ---- fused per-pixel chains (pixel_chain.cu) ---------------------------------------------------------------------------
def _bf(t):
    return t.to(torch.bfloat16).float()


def _hf(t):
    return t.to(torch.float16).float()


def _blob(w, f16=False):
    """fp32 [N][K] -> 16-bit [ceil(K/64)][N][64] rows (the chain kernels' K-blocked weight layout); f16 rows (layers fed
    by an on-chip fp16 GELU output) are returned as their bit pattern in a bf16-typed tensor."""
    n, k = w.shape
    kb = (k + 63) // 64
    wp = torch.zeros(n, kb * 64, device=w.device)
    wp[:, :k] = w
    rows = wp.reshape(n, kb, 64).permute(1, 0, 2).reshape(kb * n, 64).contiguous()
    return rows.to(torch.float16).view(torch.bfloat16) if f16 else rows.to(torch.bfloat16)


def _attn_params(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, sc=1.0: torch.randn(*s, generator=g, device="cuda") * sc
    p = dict(ln_g=1 + 0.3 * rn(64), ln_b=0.2 * rn(64), W1=_bf(rn(128, 64, sc=0.125)), b1=0.3 * rn(128),
             W2=_bf(rn(64, 128, sc=0.09)), b2=0.3 * rn(64), Wp=_bf(rn(64, 64, sc=0.125)), bp=0.3 * rn(64))
    # packer convention (pixel_chain.cuh): LayerNorm's affine is folded into the first Linear
    p["W1f"] = _bf(p["W1"] * p["ln_g"][None, :])
    p["b1f"] = p["b1"] + p["W1"] @ p["ln_b"]
    # ... and (stand-alone AttnBlock chain only) ff.net.2 and proj_out are ONE stage: out = (Wp W2) h + Wp x + [Wp (b2 + c) + bp] + x
    p["Wfold"] = _hf(p["Wp"] @ p["W2"])
    blob = torch.cat([_blob(p["W1f"]), _blob(p["W2"], f16=True), _blob(p["Wp"])])             # unfolded layout (the shot test takes W1f from it)
    p["blob_attn"] = torch.cat([_blob(p["W1f"]), _blob(p["Wfold"], f16=True), _blob(p["Wp"])])
    fvec = torch.cat([torch.zeros(128, device="cuda"), p["b1f"], p["b2"], p["bp"]])
    return p, blob, fvec


def _attn_ref(tok, cv, p, folded=False):
    """tok: [npix, 64] fp32 (bf16-representable), cv: [npix, 64].  Rounds to bf16 where the kernel stores bf16."""
    y = tok + cv
    u = _bf(F.layer_norm(y, (64,), None, None, eps=1e-5))
    h = _hf(F.gelu(u @ p["W1f"].T + p["b1f"]))            # hidden activations stay on chip in fp16
    # the folded form is the unfolded block up to the bf16 rounding of the weights
    h_ref = F.gelu(F.layer_norm(y, (64,), p["ln_g"], p["ln_b"], eps=1e-5) @ p["W1"].T + p["b1"])
    assert ((h - h_ref).norm() / h_ref.norm()).item() < 1e-2
    if folded:      # what the stand-alone chain computes; equal to the two-stage form up to the rounding of z / of the folded weight
        out = h @ p["Wfold"].T + tok @ p["Wp"].T + ((p["b2"] + cv) @ p["Wp"].T + p["bp"]) + tok
        z = h @ p["W2"].T + p["b2"] + y
        two_stage = z @ p["Wp"].T + p["bp"] + tok
        assert ((out - two_stage).norm() / two_stage.norm()).item() < 3e-3
        return out
    z = _bf(h @ p["W2"].T + p["b2"] + y)
    return z @ p["Wp"].T + p["bp"] + tok


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("case", [(2, 16, 16), (3, 8, 8), (1, 8, 24), (4, 128, 128)])
def test_attn_chain(case):
    B, H, W = case
    npix, HW = B * H * W, H * W
    x = _rand((npix, 64), 31)
    cvec = torch.randn(B, 96, device="cuda")                    # leading dimension larger than C on purpose
    p, blob, fvec = _attn_params(32)
    out = torch.zeros((npix, 64), dtype=torch.bfloat16, device="cuda")
    xb = x.to(torch.bfloat16)
    cvec2 = torch.zeros_like(cvec)
    cvec2[:, :64] = (p["b2"] + cvec[:, :64]) @ p["Wp"].T + p["bp"]          # per-sample vector of the folded stage
    _lib.check(_lib.lib().ndiff_op_pixel_chain(0, npix, HW, G.P(xb), None, None, G.P(p["blob_attn"]), G.P(fvec), G.P(cvec), 96,
                                               G.P(cvec2), G.P(out), None, G.stream()))
    torch.cuda.synchronize()
    ref = _attn_ref(x, cvec[:, :64].repeat_interleave(HW, dim=0), p, folded=True)
    assert _rel(out.float(), ref) < 5e-3, _rel(out.float(), ref)
    assert (out.float() - ref).abs().max().item() < 3e-2 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize("case", [(2, 16, 16), (3, 8, 8), (2, 256, 256)])
def test_shot_chain(case):
    B, H, W = case
    npix, HW = B * H * W, H * W
    g = torch.Generator(device="cuda").manual_seed(41)
    clean = torch.rand((npix, 4), generator=g, device="cuda") * 0.3
    xt = torch.randn((npix, 4), generator=g, device="cuda")
    cvec = torch.randn(B, 64, device="cuda")
    p, ablob, afvec = _attn_params(42)
    rn = lambda *s, sc=1.0: torch.randn(*s, generator=g, device="cuda") * sc
    W0, b0 = _bf(rn(64, 8, sc=0.35)), 0.3 * rn(64)
    Wf, bf_ = _bf(rn(64, 64, sc=0.125)), 0.3 * rn(64)
    Wm1, bm1, Wm2, bm2 = _bf(rn(64, 64, sc=0.125)), 0.3 * rn(64), _bf(rn(64, 64, sc=0.125)), 0.3 * rn(64)
    # packer convention (pixel_chain.cuh): shot_attn.ff.net.2, shot_attn.proj_out (+ both residuals) and shot_mlp2.fc1 are ONE stage:
    #   fc1(Wp (W2 h + b2 + c + s1) + bp + s1) = (M W2) h + (M + Wm1) s1 + [M (b2 + c) + Wm1 bp + bm1],  M = Wm1 Wp
    # the W2 rows hold fp16(M W2), the Wp rows bf16(M + Wm1), the bracket arrives per sample as cvec2
    M = Wm1 @ p["Wp"]
    blob = torch.cat([_blob(W0), _blob(Wf, f16=True), ablob[:128], _blob(M @ p["W2"], f16=True), _blob(M + Wm1), _blob(Wm2, f16=True)])
    fvec = torch.cat([b0, bf_, afvec, torch.zeros(64, device="cuda"), bm2])
    assert blob.shape == (512, 64) and fvec.numel() == 640
    cvec2 = (p["b2"] + cvec) @ M.T + Wm1 @ p["bp"] + bm1
    out = torch.zeros((npix, 64), dtype=torch.bfloat16, device="cuda")
    out2 = torch.zeros_like(out)
    _lib.check(_lib.lib().ndiff_op_pixel_chain(1, npix, HW, None, G.P(clean), G.P(xt), G.P(blob), G.P(fvec), G.P(cvec), 64,
                                               G.P(cvec2), G.P(out), G.P(out2), G.stream()))
    torch.cuda.synchronize()
    a0 = _bf(torch.cat([clean, xt], dim=1))
    h0 = _hf(F.gelu(a0 @ W0.T + b0))
    s1 = _bf(h0 @ Wf.T + bf_)
    s2 = _bf(_attn_ref(s1, cvec.repeat_interleave(HW, dim=0), p))
    h5 = _hf(F.gelu(s2 @ Wm1.T + bm1))
    ref = h5 @ Wm2.T + bm2
    assert _rel(out2.float(), s1) < 4e-3, _rel(out2.float(), s1)
    assert _rel(out.float(), ref) < 8e-3, _rel(out.float(), ref)


# ---- fused forms of the ResnetBlock convolutions (conv_gemm.cu kHalo1R / XF) and the shot-branch tail ------------------------
def _stats_buf(B, G_):
    return torch.zeros((B, G_, 2), dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("case", [(2, 32, 32, 64, 64, 64), (1, 64, 64, 128, 64, 128), (2, 16, 24, 256, 128, 256), (2, 128, 128, 64, 64, 64),
                                  (8, 32, 32, 512, 256, 512)])
def test_conv3x3_with_fused_res_conv(case):
    """block1.proj (3x3) and res_conv (1x1) of a ResnetBlock on the same concatenated input in one launch (ref
    Diffusion_arch.py:157,163-169)."""
    B, H, W, c0, c1, co = case
    x0, x1 = _rand((B, c0, H, W), 50), _rand((B, c1, H, W), 51)
    w3 = _rand((co, c0 + c1, 3, 3), 52, 1.0 / math.sqrt(9 * (c0 + c1)))
    w1 = _rand((co, c0 + c1, 1, 1), 53, 1.0 / math.sqrt(c0 + c1))
    b3, b1 = torch.randn(co, device="cuda"), torch.randn(co, device="cuda")
    xin = torch.cat([x0, x1], 1)
    wp = torch.cat([w3.reshape(co, (c0 + c1) // 64, 64, 9), w1.reshape(co, (c0 + c1) // 64, 64, 1)], dim=3)
    wp = wp.permute(0, 1, 3, 2).contiguous().to(torch.bfloat16)               # [Cout][cblk][10][64]
    out = torch.empty((B, H, W, co), dtype=torch.bfloat16, device="cuda")
    out2 = torch.empty_like(out)
    stats = _stats_buf(B, 8)
    ex = _lib.ConvEx(out2.data_ptr(), b1.data_ptr(), None, None, None, None, 0, 0)
    import ctypes as C
    s0, s1 = G.to_nhwc_bf16(x0), G.to_nhwc_bf16(x1)          # (named: the raw pointers must outlive the call)
    _lib.check(_lib.lib().ndiff_op_conv_ex(6, B, H, W, G.P(s0), c0, G.P(s1), c1, G.P(wp), co,
                                           G.P(b3), G.P(stats), 8, G.P(out), C.byref(ex), G.stream()))
    torch.cuda.synchronize()
    ref3 = F.conv2d(xin, w3, b3, padding=1)
    _check(out, ref3)
    _check(out2, F.conv2d(xin, w1, b1))
    # the GroupNorm sums of the 3x3 output ride along as usual
    got = stats.double() / 2 ** 24
    o = G.from_nhwc(out).double().reshape(B, 8, -1)
    assert torch.allclose(got[:, :, 0], o.sum(-1), rtol=2e-3, atol=2.0)


@pytest.mark.parametrize("mode", [G.MODE_HALO1, G.MODE_HALO2])
@pytest.mark.parametrize("case", [(2, 32, 32, 64, 2, True), (3, 64, 64, 64, 8, True), (1, 32, 48, 128, 8, True), (2, 16, 16, 256, 8, False),
                                  (1, 128, 128, 64, 8, True)])
def test_conv3x3_with_groupnorm_apply_on_the_input(case, mode):
    """block1.norm (GroupNorm + scale/shift + SiLU) evaluated inside block2's conv must equal the stand-alone
    GroupNorm-apply kernel followed by the plain conv — bit for bit (same formulas, same bf16 rounding)."""
    B, H, W, c, groups, with_ss = case
    if mode == G.MODE_HALO2 and c != 64:
        pytest.skip("256-pixel tiles are planned for the 64-channel layers only")
    h = G.to_nhwc_bf16(_rand((B, c, H, W), 60, 1.7) + 0.4)
    gamma, beta = 1 + 0.2 * torch.randn(c, device="cuda"), 0.3 * torch.randn(c, device="cuda")
    ss = 0.3 * torch.randn(B, 2 * c + 32, device="cuda") if with_ss else None       # row pitch larger than 2C on purpose
    hf = h.float()
    stats = torch.stack([(hf.reshape(B, H * W, groups, c // groups).sum((1, 3)) * 2 ** 24).round(),
                         ((hf * hf).reshape(B, H * W, groups, c // groups).sum((1, 3)) * 2 ** 24).round()], dim=2).to(torch.int64)
    w = _rand((c, c, 3, 3), 61, 1.0 / math.sqrt(9 * c))
    bias = torch.randn(c, device="cuda")
    # separate pass, then plain conv
    hn = torch.empty_like(h)
    _lib.check(_lib.lib().ndiff_op_gn_apply(G.P(h), G.P(hn), G.P(stats), G.P(gamma), G.P(beta), G.P(ss), ss.shape[1] if with_ss else 0,
                                            0, None, None, None, B, H * W, c, groups, G.stream()))
    ref = G.conv(mode, hn, G.pack_weight(w), c, bias=bias)
    # fused
    import ctypes as C
    out = torch.empty((B, H, W, c), dtype=torch.bfloat16, device="cuda")
    ex = _lib.ConvEx(None, None, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ss.data_ptr() if with_ss else None,
                     ss.shape[1] if with_ss else 0, groups)
    wpk = G.pack_weight(w)
    _lib.check(_lib.lib().ndiff_op_conv_ex(mode, B, H, W, G.P(h), c, None, 0, G.P(wpk), c, G.P(bias), None, 0,
                                           G.P(out), C.byref(ex), G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, ref), (out.float() - ref.float()).abs().max().item()
    # ... and both follow the fp32 definition
    y = F.group_norm(G.from_nhwc(h), groups, gamma, beta, eps=1e-5)
    if with_ss:
        y = y * (ss[:, :c, None, None] + 1) + ss[:, c:2 * c, None, None]
    _check(out, F.conv2d(_bf(F.silu(y)), w, bias, padding=1), tol=2.5e-2)


@pytest.mark.parametrize("case", [(2, 16, 16), (3, 8, 8), (1, 8, 24), (2, 256, 256)])
def test_shot_tail_chain(case):
    B, H, W = case
    npix, HW = B * H * W, H * W
    g = torch.Generator(device="cuda").manual_seed(71)
    rn = lambda *s, sc=1.0: torch.randn(*s, generator=g, device="cuda") * sc
    h2 = _bf(rn(npix, 64, sc=1.5) + 0.3)
    r1, r2 = _bf(rn(npix, 64)), _bf(rn(npix, 64))
    gamma, beta = 1 + 0.2 * rn(64), 0.3 * rn(64)
    W1, b1 = _bf(rn(64, 64, sc=0.125)), 0.3 * rn(64)
    W2, b2 = _bf(rn(4, 64, sc=0.125)), 0.3 * rn(4)
    groups = 2
    hb = h2.reshape(B, HW, groups, 32)
    stats = torch.stack([(hb.sum((1, 3)) * 2 ** 24).round(), ((hb * hb).sum((1, 3)) * 2 ** 24).round()], dim=2).to(torch.int64)
    blob = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    blob[:64] = _blob(W1)
    blob[64:68] = _blob(W2, f16=True)
    fvec = torch.zeros(128, device="cuda")
    fvec[:64], fvec[64:68] = b1, b2
    out = torch.zeros((npix, 4), device="cuda")
    hb16, r1b, r2b = h2.to(torch.bfloat16), r1.to(torch.bfloat16), r2.to(torch.bfloat16)     # (named: pointers must outlive the call)
    _lib.check(_lib.lib().ndiff_op_tail_chain(npix, HW, G.P(hb16), G.P(r1b), G.P(r2b),
                                              G.P(blob), G.P(fvec), G.P(stats), G.P(gamma), G.P(beta), groups, G.P(out), G.stream()))
    torch.cuda.synchronize()
    y = F.group_norm(h2.reshape(B, HW, 64).permute(0, 2, 1), groups, gamma, beta, eps=1e-5).permute(0, 2, 1).reshape(npix, 64)
    y = _bf(F.silu(y) + r1 + r2)
    ref = _hf(F.gelu(y @ W1.T + b1)) @ _hf(W2).T + b2
    assert _rel(out, ref) < 6e-3, _rel(out, ref)


def test_shared_memory_operand_and_per_lane_store_forms_still_agree():
    """Two more default-on forms have an A/B switch that is read once per process: the chain kernels hand every thread-written A
    operand to the next GEMM through tensor memory (NDIFF_CHAIN_TS=0: through swizzled shared memory), and the XF conv kernels
    store their output through a staging block + TMA (NDIFF_NO_STAGED_STORE=1: two 32-byte sectors per lane).  The chain, tail and
    fused-GroupNorm-input parity cases of this file re-run in a child process with both switched off, so that the alternative
    instruction streams in the shipped library stay covered."""
    import subprocess
    import sys
    env = dict(os.environ, NDIFF_CHAIN_TS="0", NDIFF_NO_STAGED_STORE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.abspath(__file__), "-m", "gpu", "-k",
                        "test_attn_chain or test_shot_chain or tail_chain or groupnorm_apply_on_the_input"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_plain_mma_form_of_the_halo2_kernels_still_agrees():
    """The N = 64 kHalo2 kernels run the weight-stationary tcgen05.mma.ws form by default (collector-buffer reuse of the weight
    block across the two sub-tiles; conv_gemm.cu `WS`, verified and measured on B200 in round 2) — every halo2 parity case of
    this file exercises it.  NDIFF_NO_WS=1 (read once per process) selects the plain form: the same cases re-run in a child
    process so that both instruction streams stay covered."""
    import subprocess
    import sys
    env = dict(os.environ, NDIFF_NO_WS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.abspath(__file__), "-m", "gpu", "-k",
                        "(test_conv3x3 and halo2) or groupnorm_apply_on_the_input"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
