// noisediff_b200 — the diffusion TRAINING step (BASELINE configs[4]): forward with saved activations,
// backward, Adam and EMA on the B200, behind the C ABI (ndiff_trainer_*).
//
// What it replaces in the reference (paths relative to /root/reference):
//   GaussianDiffusion.p_losses -> NoiseDiffNet.forward (per-sample t)     models/denoising_diffusion_pytorch.py:481-531
//   loss.backward()                                                        models/trainer_diffusion.py:188
//   torch.optim.Adam.step()  (lr, weight_decay from train_diffusion.py)    models/trainer_diffusion.py:92,189
//   ema_pytorch.EMA.update() (the lerp; the schedule lives on the host)    models/trainer_diffusion.py:63-69,191
//
// Plan: the network is laid out layer by layer (the sampling path's cross-layer fusions hide the activations a backward pass
// needs): every Block is conv (tcgen05, GroupNorm sums in the epilogue) + an out-of-place GroupNorm-apply, every AttnBlock is
// LayerNorm + three 1x1 GEMMs + GELU.  Every forward op registers the closure that emits its backward launches; the backward
// list is emitted in reverse order at plan time, so "first writer overwrites, later writers accumulate" is decided statically
// per gradient buffer.  Convolution gradients run on the tensor cores: dgrad = the forward implicit-GEMM kernel on flipped /
// transposed weights (one launch per concatenated source, writing straight into that source's gradient), wgrad = wgrad_gemm.cu.
// Parameters, gradients, Adam moments and the EMA copy are single flat fp32 buffers (one NCCL all-reduce, one Adam launch).
#include "engine_internal.cuh"
#include "train_ops.cuh"
#include "wgrad_gemm.cuh"

namespace ndiff {
int xpad_pack_launch(const float* x_nhwc4, bf16* xpad, int H, int W, size_t npix, cudaStream_t s);     // engine.cu

namespace {

struct TT {                 // training tensor: activation and its gradient (NHWC bf16, same shape)
    bf16* p = nullptr; bf16* g = nullptr;
    int C = 0, H = 0, W = 0;
    bool g_set = false;     // plan-time: some backward op already writes g (later contributions accumulate)
};

// dst[(n * (K/64) + kb) * taps + tap][64]: GEMM weight [N rows][K] for the dgrad convolution of one concatenated source.
//   n = input channel (c_off + n of the forward weight), K = forward output channels, taps flipped (3x3: tap' = 8 - tap).
__global__ void pack_dgrad_weight_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin_total, int c_off,
                                         int Cs, int taps) {
    const size_t total = static_cast<size_t>(Cs) * Cout * taps;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int col = static_cast<int>(i % 64);
        size_t r = i / 64;
        const int tap = static_cast<int>(r % taps); r /= taps;
        const int kb = static_cast<int>(r % (Cout / 64));
        const int n = static_cast<int>(r / (Cout / 64));
        const int co = kb * 64 + col;
        dst[i] = __float2bfloat16_rn(w[(static_cast<size_t>(co) * Cin_total + c_off + n) * taps + (taps - 1 - tap)]);
    }
}
// space-to-depth conv (weight [Cout][4C], input channel = c * 4 + tap): dgrad GEMM rows n = tap * C + c, K = Cout
__global__ void pack_dgrad_s2d_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int C) {
    const size_t total = static_cast<size_t>(4) * C * Cout;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int col = static_cast<int>(i % 64);
        size_t r = i / 64;
        const int kb = static_cast<int>(r % (Cout / 64));
        const int n = static_cast<int>(r / (Cout / 64));
        const int tap = n / C, c = n % C, co = kb * 64 + col;
        dst[i] = __float2bfloat16_rn(w[static_cast<size_t>(co) * 4 * C + c * 4 + tap]);
    }
}
__global__ void i64_to_i32_kernel2(const long long* in, int* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<int>(in[i]);
}

}  // namespace
}  // namespace ndiff

using namespace ndiff;

struct ndiff_trainer {
    ndiff_engine* e = nullptr;
    int B = 0, H = 0, W = 0;
    bool built = false;
    std::vector<void*> owned;
    // flat fp32 buffers: [live parameters | dead parameters (attn.to_q / to_k / norm1: no gradient, never stepped)]
    float* flat_p = nullptr; float* flat_g = nullptr; float* flat_m = nullptr; float* flat_v = nullptr; float* flat_ema = nullptr;
    size_t n_flat = 0, n_live = 0;
    std::map<std::string, std::pair<size_t, size_t>> slot;     // name -> (offset, numel)
    int adam_steps = 0;
    // per-step buffers
    float* target = nullptr; float* w_b = nullptr; float* dss = nullptr; float* dcv = nullptr; float* st_saved = nullptr;
    float* dst_buf = nullptr; float* gn_acc = nullptr; double* loss = nullptr;
    bf16* dmap1 = nullptr; bf16* dmap2 = nullptr;
    struct DgradPack { std::string name; bf16* dst; int Cout, Cin_total, c_off, Cs, taps; bool s2d; };
    std::vector<DgradPack> dpacks;
    std::vector<std::function<int(cudaStream_t)>> fwd, bwd;
    std::vector<std::string> fwd_kind, bwd_kind;               // kernel family of every launch (per-family timing, ndiff_trainer_time)
    std::string kind = "other";
    void addf(std::function<int(cudaStream_t)> f) { fwd.push_back(std::move(f)); fwd_kind.push_back(kind); }
    void addb(std::function<int(cudaStream_t)> f) { bwd.push_back(std::move(f)); bwd_kind.push_back(kind); }
    std::vector<std::function<void()>> pending;                // emitters of backward launches, in forward order
    int err = 0;
    int stats_slot = 0;
    size_t act_bytes = 0;
    int steps_run = 0;
    cudaGraphExec_t step_exec = nullptr;      // time path + forward + backward of one step (captured on the second step)

    ~ndiff_trainer() {
        if (step_exec) cudaGraphExecDestroy(step_exec);
        for (void* p : owned) cudaFree(p);
        delete e;
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        NDIFF_CUDA_OK(cudaMalloc(&p, count * sizeof(T) ? count * sizeof(T) : 16));
        owned.push_back(p);
        *out = static_cast<T*>(p);
        return 0;
    }
    float* G(const std::string& name) { return flat_g + slot.at(name).first; }
};

namespace {

struct TB {      // plan builder
    ndiff_trainer* t;
    ndiff_engine* e;
    explicit TB(ndiff_trainer* tr) : t(tr), e(tr->e) {}

    TT make(int C, int H, int W) {
        TT a; a.C = C; a.H = H; a.W = W;
        const size_t n = static_cast<size_t>(t->B) * H * W * C;
        if (t->alloc(&a.p, n) || t->alloc(&a.g, n)) t->err = 1;
        t->act_bytes += 2 * n * sizeof(bf16);
        return a;
    }
    unsigned long long* next_stats() { return e->stats + static_cast<size_t>(t->stats_slot++) * t->B * 8 * 2; }
    size_t npix(const TT& a) const { return static_cast<size_t>(t->B) * a.H * a.W; }

    // dst.g (+)= src (same shape): gradient routing of residual adds
    void route(TT& dst, const bf16* src) {
        bf16* d = dst.g; const bool acc = dst.g_set; const int C = dst.C; const size_t n = npix(dst);
        t->kind = "route"; t->addb([=](cudaStream_t s) { return add_slice_launch(d, C, 0, src, C, 0, C, n, acc, s); });
        dst.g_set = true;
    }

    // ---- convolution / GEMM: out = conv(cat(s0, s1)) + bias (+ vec[b]) (+ res); optional GroupNorm sums of the output ----------
    // mode: kHalo1 (3x3 pad 1), kDirect (1x1), kS2D.  s0 / s1 / res are modified at backward-emission time (g_set flags).
    TT conv(const std::string& wname, int mode, TT* s0, TT* s1, int Cout, const float* vec, int vec_off, TT* res,
            unsigned long long* stats, int groups, bool input_grad = true) {
        const int Ho = mode == kS2D ? s0->H / 2 : s0->H, Wo = mode == kS2D ? s0->W / 2 : s0->W;
        TT out = make(Cout, Ho, Wo);
        if (t->err) return out;
        ConvGemmDesc d;
        d.mode = mode; d.B = t->B; d.H = Ho; d.W = Wo;
        d.src0 = s0->p; d.C0 = s0->C;
        if (s1) { d.src1 = s1->p; d.C1 = s1->C; }
        d.weight = e->packed.at(wname);
        d.Cout = Cout; d.bias = e->pf(wname + ".bias");
        if (vec) { d.vec = vec; d.vec_ld = e->cv_total; }
        if (res) { d.res = res->p; d.res_ld = res->C; }
        d.out = out.p; d.out_ld = Cout;
        d.stats = stats; d.groups = groups;
        // 256-pixel tiles + weight-stationary MMAs where the sampling path uses them (64 -> 64, single source; engine.cu Builder::conv)
        if (mode == kHalo1 && Ho >= 32 && Cout == 64 && s0->C == 64 && !s1) d.mode = kHalo2;
        auto plan = std::make_shared<ConvGemmPlan>();
        if (conv_gemm_plan(d, e->num_sms, plan.get())) { t->err = 1; return out; }
        t->kind = "conv_fwd"; t->addf([plan](cudaStream_t s) { return conv_gemm_launch(*plan, s); });

        // dgrad weight packs (one per concatenated source), refreshed after every optimizer step
        const int taps = mode == kHalo1 ? 9 : 1;
        const int Cin = s0->C + (s1 ? s1->C : 0);
        bf16* dw0 = nullptr; bf16* dw1 = nullptr;
        if (input_grad) {
            if (mode == kS2D) {
                if (t->alloc(&dw0, static_cast<size_t>(4) * s0->C * Cout)) { t->err = 1; return out; }
                t->dpacks.push_back({wname, dw0, Cout, 0, 0, s0->C, 1, true});
            } else {
                if (t->alloc(&dw0, static_cast<size_t>(s0->C) * Cout * taps)) { t->err = 1; return out; }
                t->dpacks.push_back({wname, dw0, Cout, Cin, 0, s0->C, taps, false});
                if (s1) {
                    if (t->alloc(&dw1, static_cast<size_t>(s1->C) * Cout * taps)) { t->err = 1; return out; }
                    t->dpacks.push_back({wname, dw1, Cout, Cin, s0->C, s1->C, taps, false});
                }
            }
        }
        ndiff_trainer* tr = t;
        TB* self = this;
        const TT o = out;
        t->pending.push_back([=]() {
            // (runs at plan time, in reverse forward order)  gradient of `out` is complete in o.g
            const size_t np = static_cast<size_t>(tr->B) * o.H * o.W;
            float* gb = tr->G(wname + ".bias");
            // (a conv followed by GroupNorm gets its bias gradient from the GroupNorm backward's second pass: GnBwdArgs::dbias)
            if (!stats) { tr->kind = "bias_grad"; tr->addb([=](cudaStream_t s) { return colsum_launch(o.g, o.C, 0, gb, 0, false, tr->B, o.H * o.W, o.C, s); }); }
            if (vec) {
                float* dcv = tr->dcv + vec_off;
                const int ld = tr->e->cv_total;
                tr->addb([=](cudaStream_t s) { return colsum_launch(o.g, o.C, 0, dcv, ld, true, tr->B, o.H * o.W, o.C, s); });
            }
            if (res) self->route(*res, o.g);
            {   // weight gradient
                WgradDesc wd;
                wd.mode = mode == kHalo1 ? kWg3x3 : (mode == kS2D ? kWgS2D : kWg1x1);
                wd.B = tr->B; wd.H = o.H; wd.W = o.W;
                wd.dy = o.g; wd.Cout = o.C;
                wd.src0 = s0->p; wd.C0 = s0->C;
                if (s1) { wd.src1 = s1->p; wd.C1 = s1->C; }
                wd.dw = tr->G(wname + ".weight");
                auto wp = std::make_shared<WgradPlan>();
                if (wgrad_gemm_plan(wd, tr->e->num_sms, wp.get())) { tr->err = 1; return; }
                tr->kind = "wgrad"; tr->addb([wp](cudaStream_t s) { return wgrad_gemm_launch(*wp, s); });
            }
            if (!input_grad) return;
            if (mode == kS2D) {
                // dX[b, 2h+p1, 2w+p2, c] = sum_co dY[b,h,w,co] W[co][c*4 + p1*2 + p2]: a 1x1 GEMM to [B,H,W,4C], then depth-to-space
                bf16* tmp = nullptr;
                if (tr->alloc(&tmp, np * 4 * s0->C)) { tr->err = 1; return; }
                ConvGemmDesc g;
                g.mode = kDirect; g.B = tr->B; g.H = o.H; g.W = o.W;
                g.src0 = o.g; g.C0 = o.C; g.weight = dw0; g.Cout = 4 * s0->C; g.out = tmp; g.out_ld = 4 * s0->C;
                auto gp = std::make_shared<ConvGemmPlan>();
                if (conv_gemm_plan(g, tr->e->num_sms, gp.get())) { tr->err = 1; return; }
                tr->kind = "dgrad"; tr->addb([gp](cudaStream_t s) { return conv_gemm_launch(*gp, s); });
                tr->kind = "route";
                bf16* dx = s0->g; const bool acc = s0->g_set; const int C = s0->C, hh = o.H, ww = o.W, B = tr->B;
                tr->addb([=](cudaStream_t s) { return depth_to_space_launch(tmp, dx, B, hh, ww, C, acc, s); });
                s0->g_set = true;
                return;
            }
            for (int si = 0; si < (s1 ? 2 : 1); ++si) {
                TT* src = si == 0 ? s0 : s1;
                ConvGemmDesc g;
                g.mode = mode; g.B = tr->B; g.H = o.H; g.W = o.W;       // kHalo1 (flipped taps) or kDirect (transposed 1x1)
                if (mode == kHalo1 && o.H >= 32 && o.C == 64 && src->C == 64) g.mode = kHalo2;
                g.src0 = o.g; g.C0 = o.C;
                g.weight = si == 0 ? dw0 : dw1;
                g.Cout = src->C; g.out = src->g; g.out_ld = src->C;
                if (src->g_set) { g.res = src->g; g.res_ld = src->C; }     // accumulate in place: each thread reads its own pixel first
                auto gp = std::make_shared<ConvGemmPlan>();
                if (conv_gemm_plan(g, tr->e->num_sms, gp.get())) { tr->err = 1; return; }
                tr->kind = "dgrad"; tr->addb([gp](cudaStream_t s) { return conv_gemm_launch(*gp, s); });
                src->g_set = true;
            }
        });
        return out;
    }

    // ---- GroupNorm + (scale+1)/shift + SiLU (+ residual adds), out of place: h stays for the backward pass ---------------
    TT gn(const std::string& nname, const std::string& conv_name, TT* h, unsigned long long* stats, int groups, int ss_off,
          const bf16* maps, bf16* dmaps, TT* r1, TT* r2) {
        TT out = make(h->C, h->H, h->W);
        if (t->err) return out;
        GnApplyArgs g{};
        g.x = h->p; g.out = out.p; g.stats = stats;
        g.gamma = e->pf(nname + ".weight"); g.beta = e->pf(nname + ".bias");
        if (ss_off >= 0) { g.ss = e->ss_cur; g.ss_ld = e->ss_total; g.ss_off = ss_off; }
        g.maps = maps;
        g.res1 = r1 ? r1->p : nullptr; g.res2 = r2 ? r2->p : nullptr;
        g.B = t->B; g.HW = h->H * h->W; g.C = h->C; g.G = groups; g.eps = 1e-5f; g.real_frac = 1.0f;
        t->kind = "gn_fwd"; t->addf([g](cudaStream_t s) { return gn_apply_launch(g, s); });
        ndiff_trainer* tr = t;
        TB* self = this;
        const TT o = out;
        t->pending.push_back([=]() {
            if (r1) self->route(*r1, o.g);
            if (r2) self->route(*r2, o.g);
            GnBwdArgs b{};
            b.h = h->p; b.dout = o.g; b.dh = h->g; b.stats = stats;
            b.gamma = g.gamma; b.beta = g.beta;
            b.ss = g.ss; b.ss_ld = g.ss_ld; b.ss_off = g.ss_off;
            b.maps = maps; b.dmaps = dmaps;
            b.acc = tr->gn_acc;
            b.dgamma = tr->G(nname + ".weight"); b.dbeta = tr->G(nname + ".bias");
            b.dbias = tr->G(conv_name + ".bias");
            if (ss_off >= 0) { b.dss = tr->dss; b.dss_ld = tr->e->ss_total; }
            b.B = tr->B; b.HW = g.HW; b.C = g.C; b.G = groups; b.eps = 1e-5f; b.real_frac = 1.0f;
            tr->kind = "gn_bwd"; tr->addb([b](cudaStream_t s) { return gn_backward_launch(b, s); });
            h->g_set = true;      // the raw conv output has exactly one consumer: written, not accumulated
        });
        return out;
    }

    TT gelu(TT* pre) {
        TT out = make(pre->C, pre->H, pre->W);
        if (t->err) return out;
        const size_t n = npix(*pre) * pre->C;
        const bf16* pp = pre->p; bf16* op = out.p;
        t->kind = "gelu"; t->addf([=](cudaStream_t s) { return gelu_forward_launch(pp, op, n, s); });
        ndiff_trainer* tr = t;
        const TT o = out;
        t->pending.push_back([=]() {
            bf16* dp = pre->g; const bf16* dy = o.g;
            tr->kind = "gelu"; tr->addb([=](cudaStream_t s) { return gelu_backward_launch(pp, dy, dp, n, s); });
            pre->g_set = true;
        });
        return out;
    }

    // ResnetBlock / ResnetBlock2 (Diffusion_arch.py:146-196)
    TT resblock(const std::string& n, TT* s0, TT* s1, int Cout, int groups, const bf16* maps, bf16* dmaps, TT* extra_res,
                std::vector<std::unique_ptr<TT>>& keep) {
        auto hold = [&](TT v) { keep.push_back(std::make_unique<TT>(v)); return keep.back().get(); };
        const int Cin = s0->C + (s1 ? s1->C : 0);
        unsigned long long* st1 = next_stats();
        TT* h1 = hold(conv(n + ".block1.proj", kHalo1, s0, s1, Cout, nullptr, 0, nullptr, st1, groups));
        TT* a1 = hold(gn(n + ".block1.norm", n + ".block1.proj", h1, st1, groups, maps ? -1 : e->ss_off.at(n), maps, dmaps, nullptr, nullptr));
        unsigned long long* st2 = next_stats();
        TT* h2 = hold(conv(n + ".block2.proj", kHalo1, a1, nullptr, Cout, nullptr, 0, nullptr, st2, groups));
        TT* r = s0;
        if (Cin != Cout) r = hold(conv(n + ".res_conv", kDirect, s0, s1, Cout, nullptr, 0, nullptr, nullptr, 0));
        return gn(n + ".block2.norm", n + ".block2.proj", h2, st2, groups, -1, nullptr, nullptr, r, extra_res);
    }

    // AttnBlock with the collapsed 1-token cross attention (Diffusion_arch.py:425-443):
    //   u = LN(x + c); z = ff2(GELU(ff1(u))) + c + x; out = proj(z) + x
    TT attn(const std::string& n, TT* x, std::vector<std::unique_ptr<TT>>& keep) {
        auto hold = [&](TT v) { keep.push_back(std::make_unique<TT>(v)); return keep.back().get(); };
        const int C = x->C, off = e->cv_off.at(n);
        const float* cv = e->cvec + off;
        TT* u = hold(make(C, x->H, x->W));
        if (t->err) return *u;
        {
            const bf16* xp = x->p; bf16* up = u->p;
            const float* g = e->pf(n + ".norm2.weight"); const float* bt = e->pf(n + ".norm2.bias");
            const int B = t->B, HW = x->H * x->W, ld = e->cv_total;
            t->kind = "layernorm"; t->addf([=](cudaStream_t s) { return layernorm_launch(xp, cv, ld, g, bt, up, B, HW, C, s, 1.0f); });
            ndiff_trainer* tr = t;
            TB* self = this;
            t->pending.push_back([=]() {
                bf16* du = u->g;
                float* dg = tr->G(n + ".norm2.weight"); float* db = tr->G(n + ".norm2.bias");
                float* dcv = tr->dcv + off;
                // dy overwrites du in place (each thread reads its du elements before it writes them)
                tr->kind = "layernorm"; tr->addb([=](cudaStream_t s) { return layernorm_backward_launch(xp, cv, ld, g, du, du, dg, db, B, HW, C, 1.0f, s); });
                tr->addb([=](cudaStream_t s) { return colsum_launch(du, C, 0, dcv, ld, true, B, HW, C, s); });
                self->route(*x, du);
            });
        }
        TT* hpre = hold(conv(n + ".ff.net.0.0", kDirect, u, nullptr, 2 * C, nullptr, 0, nullptr, nullptr, 0));
        TT* hh = hold(gelu(hpre));
        TT* z = hold(conv(n + ".ff.net.2", kDirect, hh, nullptr, C, cv, off, x, nullptr, 0));
        return conv(n + ".proj_out", kDirect, z, nullptr, C, nullptr, 0, x, nullptr, 0);
    }
};

int build_training_plan(ndiff_trainer* t) {
    ndiff_engine* e = t->e;
    const int dim = e->dim, B = t->B, H = t->H, W = t->W;
    const int d[5] = {dim, dim, dim * 2, dim * 4, dim * 8};
    const size_t npix = static_cast<size_t>(B) * H * W;
    TB b(t);
    std::vector<std::unique_ptr<TT>> keep;      // stable addresses: backward emitters hold TT pointers
    auto hold = [&](TT v) { keep.push_back(std::make_unique<TT>(v)); return keep.back().get(); };

    // ---- shot-noise branch (Diffusion_arch.py:598-604)
    TT* s0 = hold(b.make(dim, H, W));
    if (t->err) return 1;
    {
        const float* cl = e->clean; const float* x = e->x; bf16* o = s0->p;
        const float* w = e->pf("shot_mlp1.fc1.weight"); const float* bs = e->pf("shot_mlp1.fc1.bias");
        t->addf([=](cudaStream_t s) { return shot_in_launch(cl, x, w, bs, o, static_cast<int>(npix), dim, s); });
        t->pending.push_back([=]() {
            const bf16* ds0 = s0->g;
            float* dw = t->G("shot_mlp1.fc1.weight"); float* db = t->G("shot_mlp1.fc1.bias");
            t->addb([=](cudaStream_t s) { return shot_in_backward_launch(cl, x, w, bs, ds0, dw, db, npix, dim, s); });
        });
    }
    TT* s1 = hold(b.conv("shot_mlp1.fc2", kDirect, s0, nullptr, dim, nullptr, 0, nullptr, nullptr, 0));
    TT* s2 = hold(b.attn("shot_attn", s1, keep));
    TT* s3p = hold(b.conv("shot_mlp2.fc1", kDirect, s2, nullptr, dim, nullptr, 0, nullptr, nullptr, 0));
    TT* s3 = hold(b.gelu(s3p));
    TT* s4 = hold(b.conv("shot_mlp2.fc2", kDirect, s3, nullptr, dim, nullptr, 0, nullptr, nullptr, 0));
    TT* s5 = hold(b.resblock("shot_time", s4, nullptr, dim, 2, nullptr, nullptr, s1, keep));      // + r_s (ref :603)
    TT* s6p = hold(b.conv("shot_mlp3.fc1", kDirect, s5, nullptr, dim, nullptr, 0, nullptr, nullptr, 0));
    TT* s6 = hold(b.gelu(s6p));

    // ---- main U-Net (ref :606-643)
    TT* x0 = hold(b.make(dim, H, W));
    if (t->err) return 1;
    {
        ConvGemmDesc cd;
        cd.mode = kDirect; cd.B = B; cd.H = H; cd.W = W;
        cd.src0 = e->xpad; cd.C0 = 64;
        cd.taps_y = 4; cd.taps_x = 1; cd.pad_y = 0; cd.pad_x = 0; cd.tap_sy = 2;
        cd.custom_src0 = true;
        cd.cdim[0] = 64; cd.cdim[1] = static_cast<uint64_t>(W); cd.cdim[2] = static_cast<uint64_t>(H + 6); cd.cdim[3] = static_cast<uint64_t>(B);
        cd.cstride[0] = 16; cd.cstride[1] = static_cast<uint64_t>(W + 8) * 16; cd.cstride[2] = static_cast<uint64_t>(H + 6) * (W + 8) * 16;
        cd.weight = e->init_w_tc; cd.Cout = dim; cd.bias = e->pf("init_conv.bias");
        cd.out = x0->p; cd.out_ld = dim;
        auto plan = std::make_shared<ConvGemmPlan>();
        if (conv_gemm_plan(cd, e->num_sms, plan.get())) return 1;
        const float* xs = e->x; bf16* xp = e->xpad;
        t->addf([=](cudaStream_t s) { return xpad_pack_launch(xs, xp, H, W, npix, s); });
        t->addf([plan](cudaStream_t s) { return conv_gemm_launch(*plan, s); });
        t->pending.push_back([=]() {
            const bf16* dy = x0->g;
            float* dw = t->G("init_conv.weight"); float* db = t->G("init_conv.bias");
            t->addb([=](cudaStream_t s) { return colsum_launch(dy, dim, 0, db, 0, false, B, H * W, dim, s); });
            t->addb([=](cudaStream_t s) { return init_conv_wgrad_launch(xs, dy, dw, B, H, W, dim, s); });
        });
    }
    TT* cur = hold(b.resblock("pos_block1", x0, nullptr, dim, 2, e->map1, t->dmap1, nullptr, keep));
    std::vector<TT*> skips;
    for (int i = 0; i < 4; ++i) {
        const std::string p = "downs." + std::to_string(i);
        TT* a1 = hold(b.resblock(p + ".0", cur, nullptr, d[i], 8, nullptr, nullptr, nullptr, keep));
        skips.push_back(a1);
        TT* a2 = hold(b.resblock(p + ".1", a1, nullptr, d[i], 8, nullptr, nullptr, nullptr, keep));
        skips.push_back(a2);
        TT* a3 = hold(b.attn(p + ".2", a2, keep));
        if (i < 3) cur = hold(b.conv(p + ".3.1", kS2D, a3, nullptr, d[i + 1], nullptr, 0, nullptr, nullptr, 0));
        else cur = hold(b.conv(p + ".3", kHalo1, a3, nullptr, d[i + 1], nullptr, 0, nullptr, nullptr, 0));
    }
    {
        TT* m1 = hold(b.resblock("mid_block1", cur, nullptr, d[4], 8, nullptr, nullptr, nullptr, keep));
        cur = hold(b.resblock("mid_block2", m1, nullptr, d[4], 8, nullptr, nullptr, nullptr, keep));
    }
    for (int i = 0; i < 4; ++i) {
        const std::string p = "ups." + std::to_string(i);
        const int co = d[4 - i], ci = d[3 - i];
        TT* sk = skips.back(); skips.pop_back();
        TT* a1 = hold(b.resblock(p + ".0", cur, sk, co, 8, nullptr, nullptr, nullptr, keep));
        sk = skips.back(); skips.pop_back();
        TT* a2 = hold(b.resblock(p + ".1", a1, sk, co, 8, nullptr, nullptr, nullptr, keep));
        TT* a3 = hold(b.attn(p + ".2", a2, keep));
        if (i < 3) {
            // Upsample = nearest x2 (materialised here: its backward is a 2x2 sum) + conv3x3 (ref :72-76)
            TT* up = hold(b.make(co, a3->H * 2, a3->W * 2));
            if (t->err) return 1;
            const bf16* in = a3->p; bf16* o = up->p; const int hh = a3->H, ww = a3->W;
            t->addf([=](cudaStream_t s) { return upsample2x_launch(in, o, B, hh, ww, co, s); });
            t->pending.push_back([=]() {
                const bf16* dy = up->g; bf16* dx = a3->g; const bool acc = a3->g_set;
                t->addb([=](cudaStream_t s) { return upsample2x_backward_launch(dy, dx, B, hh, ww, co, acc, s); });
                a3->g_set = true;
            });
            cur = hold(b.conv(p + ".3.1", kHalo1, up, nullptr, ci, nullptr, 0, nullptr, nullptr, 0));
        } else {
            cur = hold(b.conv(p + ".3", kHalo1, a3, nullptr, ci, nullptr, 0, nullptr, nullptr, 0));
        }
    }
    TT* pb2 = hold(b.resblock("pos_block2", cur, nullptr, dim, 2, e->map2, t->dmap2, nullptr, keep));
    TT* fr = hold(b.resblock("final_res_block", pb2, x0, dim, 8, nullptr, nullptr, nullptr, keep));
    if (t->err) return 1;
    NDIFF_REQUIRE(t->stats_slot <= e->n_stats, "GroupNorm statistics arena too small");

    // ---- heads: v = final_conv(fr) + shot_mlp3.fc2(s6)  (ref :643-644) -------------------------------------------------------
    {
        FinalArgs f;
        memset(&f, 0, sizeof(f));
        f.xf = fr->p; f.sf = s6->p;
        f.wf = e->pf("final_conv.weight"); f.bfin = e->pf("final_conv.bias");
        f.ws = e->pf("shot_mlp3.fc2.weight"); f.bs = e->pf("shot_mlp3.fc2.bias");
        f.npix = static_cast<int>(npix); f.C = dim; f.HW = H * W; f.v_out = e->v_out;
        t->kind = "heads"; t->addf([f](cudaStream_t s) { return final_launch(f, s); });
    }
    // ---- backward list: loss + heads first, then every forward op's emitter in reverse order ---------------------------------
    {
        HeadsBwdArgs hb{};
        hb.xf = fr->p; hb.sf = s6->p;
        hb.wf = e->pf("final_conv.weight"); hb.ws = e->pf("shot_mlp3.fc2.weight");
        hb.v = reinterpret_cast<const float4*>(e->v_out); hb.target = reinterpret_cast<const float4*>(t->target); hb.w_b = t->w_b;
        hb.dxf = fr->g; hb.dsf = s6->g;
        hb.dwf = t->G("final_conv.weight"); hb.dbf = t->G("final_conv.bias");
        hb.dws = t->G("shot_mlp3.fc2.weight"); hb.dbs = t->G("shot_mlp3.fc2.bias");
        hb.loss = t->loss; hb.B = B; hb.HW = H * W; hb.C = dim;
        t->kind = "heads"; t->addb([hb](cudaStream_t s) { return heads_backward_launch(hb, s); });
        fr->g_set = true; s6->g_set = true;
    }
    for (auto it = t->pending.rbegin(); it != t->pending.rend(); ++it) {
        (*it)();
        if (t->err) return 1;
    }
    t->pending.clear();
    // ---- per-sample / small dense paths -----------------------------------------------------------------------------------------
    t->kind = "small_paths";
    const int td = e->dim_real * 4;
    for (const RbSpec& rb : resblocks(dim)) {
        if (rb.pos) continue;
        const int off = e->ss_off.at(rb.name), rows = 2 * rb.cout, ld = e->ss_total;
        const float* dss = t->dss + off; const float* st = e->st_buf;
        float* dw = t->G(rb.name + ".mlp.1.weight"); float* db = t->G(rb.name + ".mlp.1.bias");
        t->addb([=](cudaStream_t s) { return small_gemm_launch(true, false, rows, td, B, dss, ld, st, td, dw, td, true, s); });
        t->addb([=](cudaStream_t s) { return rowsum_f32_launch(dss, ld, B, rows, db, s); });
    }
    {
        const float* dss = t->dss; const float* ssw = e->ss_w; float* dst = t->dst_buf; const int ld = e->ss_total;
        t->addb([=](cudaStream_t s) { return small_gemm_launch(false, false, B, td, ld, dss, ld, ssw, td, dst, td, false, s); });
        const float* saved = t->st_saved;
        const float* w1 = e->pf("time_mlp.1.weight"); const float* w2 = e->pf("time_mlp.3.weight");
        float* dw1 = t->G("time_mlp.1.weight"); float* db1 = t->G("time_mlp.1.bias");
        float* dw2 = t->G("time_mlp.3.weight"); float* db2 = t->G("time_mlp.3.bias");
        const int dr = e->dim_real;
        t->addb([=](cudaStream_t s) { return time_mlp_backward_launch(dst, saved, B, dr, w1, w2, dw1, db1, dw2, db2, s); });
    }
    for (const AttnSpec& ab : attnblocks(dim)) {
        const float* emb = e->pf("iso_embed.weight");
        const float* wv = e->pf(ab.name + ".attn.to_v.weight"); const float* wo = e->pf(ab.name + ".attn.to_out.0.weight");
        float* demb = t->G("iso_embed.weight"); float* dwv = t->G(ab.name + ".attn.to_v.weight");
        float* dwo = t->G(ab.name + ".attn.to_out.0.weight"); float* dbo = t->G(ab.name + ".attn.to_out.0.bias");
        const float* dcv = t->dcv; const int ld = e->cv_total, off = e->cv_off.at(ab.name), C = ab.C;
        ndiff_trainer* tr = t;
        t->addb([=](cudaStream_t s) {
            return iso_vec_backward_launch(emb, reinterpret_cast<const long long*>(tr->e->iso_idx), wv, wo, dcv, ld, off, demb, dwv, dwo, dbo, B, C, s);
        });
    }
    {
        PosBwdArgs pa{};
        pa.fwd.position = e->position;
        pa.fwd.we = e->pf("pos_enc.weights.weight"); pa.fwd.be = e->pf("pos_enc.weights.bias");
        pa.fwd.w1 = e->pf("pos_mlp.fc1.weight"); pa.fwd.b1 = e->pf("pos_mlp.fc1.bias");
        pa.fwd.w2 = e->pf("pos_mlp.fc2.weight"); pa.fwd.b2 = e->pf("pos_mlp.fc2.bias");
        pa.fwd.wm1 = e->pf("pos_block1.mlp.1.weight"); pa.fwd.bm1 = e->pf("pos_block1.mlp.1.bias");
        pa.fwd.wm2 = e->pf("pos_block2.mlp.1.weight"); pa.fwd.bm2 = e->pf("pos_block2.mlp.1.bias");
        pa.fwd.B = B; pa.fwd.HW = H * W; pa.fwd.C = dim;
        pa.dmap1 = t->dmap1; pa.dmap2 = t->dmap2;
        pa.dwe = t->G("pos_enc.weights.weight"); pa.dbe = t->G("pos_enc.weights.bias");
        pa.dw1 = t->G("pos_mlp.fc1.weight"); pa.db1 = t->G("pos_mlp.fc1.bias");
        pa.dw2 = t->G("pos_mlp.fc2.weight"); pa.db2 = t->G("pos_mlp.fc2.bias");
        pa.dwm1 = t->G("pos_block1.mlp.1.weight"); pa.dbm1 = t->G("pos_block1.mlp.1.bias");
        pa.dwm2 = t->G("pos_block2.mlp.1.weight"); pa.dbm2 = t->G("pos_block2.mlp.1.bias");
        t->addb([pa](cudaStream_t s) { return pos_backward_launch(pa, s); });
    }
    // the TT objects only matter at plan time (pointers and flags were copied into the closures): `keep` may go
    t->built = true;
    return 0;
}

int repack_dgrad(ndiff_trainer* t, cudaStream_t s) {
    for (const ndiff_trainer::DgradPack& p : t->dpacks) {
        const float* w = t->e->pf(p.name + ".weight");
        if (p.s2d) pack_dgrad_s2d_kernel<<<256, 256, 0, s>>>(w, p.dst, p.Cout, p.Cs);
        else pack_dgrad_weight_kernel<<<256, 256, 0, s>>>(w, p.dst, p.Cout, p.Cin_total, p.c_off, p.Cs, p.taps);
        NDIFF_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

bool is_dead_param(const std::string& k) {       // never reached by the loss: 1-token softmax (Diffusion_arch.py:361-402)
    auto ends = [&](const char* suf) { const std::string s(suf); return k.size() >= s.size() && k.compare(k.size() - s.size(), s.size(), s) == 0; };
    return ends(".attn.to_q.weight") || ends(".attn.to_k.weight") || ends(".norm1.weight") || ends(".norm1.bias");
}

}  // namespace

// ================================================================================================================
// C ABI
// ================================================================================================================
extern "C" {

int32_t ndiff_trainer_create(const ndiff_config* cfg, ndiff_trainer** out) {
    NDIFF_REQUIRE(cfg && out, "null argument");
    NDIFF_REQUIRE(cfg->dim == 64, "the training step runs at dim = 64 (the embedded narrower widths are a sampling-path feature)");
    ndiff_engine* e = nullptr;
    if (ndiff_engine_create(cfg, &e)) return 1;
    e->skip_plan = true;
    std::unique_ptr<ndiff_trainer> t(new ndiff_trainer());
    t->e = e; t->B = cfg->batch; t->H = cfg->height; t->W = cfg->width;
    *out = t.release();
    return 0;
}

void ndiff_trainer_destroy(ndiff_trainer* t) {
    if (!t) return;
    DeviceGuard dev_guard(t->e->cfg.device);
    cudaDeviceSynchronize();
    delete t;
}

ndiff_engine* ndiff_trainer_engine(ndiff_trainer* t) { return t ? t->e : nullptr; }

int32_t ndiff_trainer_finalize(ndiff_trainer* t, void* stream) {
    NDIFF_REQUIRE(t && !t->built, "trainer already finalized (weights are updated in place by ndiff_trainer_adam_step)");
    ndiff_engine* e = t->e;
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    // ---- flatten: live parameters first, then the dead ones; 16-byte aligned slots
    std::vector<std::string> live, dead;
    for (auto& kv : e->params) (is_dead_param(kv.first) ? dead : live).push_back(kv.first);
    size_t at = 0;
    for (auto* lst : {&live, &dead}) {
        for (const std::string& k : *lst) {
            const size_t n = e->params.at(k).n;
            t->slot[k] = {at, n};
            at += (n + 3) / 4 * 4;
        }
        if (lst == &live) t->n_live = at;
    }
    t->n_flat = at;
    if (t->alloc(&t->flat_p, at) || t->alloc(&t->flat_g, at) || t->alloc(&t->flat_m, at) || t->alloc(&t->flat_v, at) ||
        t->alloc(&t->flat_ema, at)) return 1;
    NDIFF_CUDA_OK(cudaMemsetAsync(t->flat_p, 0, at * sizeof(float), s));
    NDIFF_CUDA_OK(cudaMemsetAsync(t->flat_g, 0, at * sizeof(float), s));
    NDIFF_CUDA_OK(cudaMemsetAsync(t->flat_m, 0, at * sizeof(float), s));
    NDIFF_CUDA_OK(cudaMemsetAsync(t->flat_v, 0, at * sizeof(float), s));
    for (auto& kv : e->params) {
        float* dst = t->flat_p + t->slot.at(kv.first).first;
        NDIFF_CUDA_OK(cudaMemcpyAsync(dst, kv.second.dev, kv.second.n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));
    for (auto& kv : e->params) {        // the engine now reads its fp32 parameters from the flat buffer (Adam updates them in place)
        e->release(kv.second.dev);
        kv.second.dev = t->flat_p + t->slot.at(kv.first).first;
    }
    NDIFF_CUDA_OK(cudaMemcpyAsync(t->flat_ema, t->flat_p, at * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (engine_finalize(e, s)) return 1;
    // ---- per-step buffers
    const size_t npix = static_cast<size_t>(t->B) * t->H * t->W;
    const int td = e->dim_real * 4;
    if (t->alloc(&t->target, npix * 4) || t->alloc(&t->w_b, t->B) || t->alloc(&t->dss, static_cast<size_t>(t->B) * e->ss_total) ||
        t->alloc(&t->dcv, static_cast<size_t>(t->B) * e->cv_total) || t->alloc(&t->st_saved, static_cast<size_t>(t->B) * (e->dim_real + 2 * td)) ||
        t->alloc(&t->dst_buf, static_cast<size_t>(t->B) * td) || t->alloc(&t->gn_acc, static_cast<size_t>(t->B) * (512 * 2 + 8 * 2)) ||
        t->alloc(&t->loss, 1) || t->alloc(&t->dmap1, npix * 2 * e->dim) || t->alloc(&t->dmap2, npix * 2 * e->dim)) return 1;
    if (e->t_buf_n < t->B) {
        e->release(e->t_buf); e->release(e->st_buf);
        if (e->alloc(&e->t_buf, t->B) || e->alloc(&e->st_buf, static_cast<size_t>(t->B) * td)) return 1;
        e->t_buf_n = t->B;
    }
    if (build_training_plan(t)) return 1;
    if (repack_dgrad(t, s)) return 1;
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int32_t ndiff_trainer_forward_backward(ndiff_trainer* t, const float* x_t_dev, const int64_t* time_dev, const float* target_dev,
                                       const float* loss_weight_dev, double* loss_host, void* stream) {
    NDIFF_REQUIRE(t && t->built, "trainer not finalized");
    ndiff_engine* e = t->e;
    NDIFF_REQUIRE(e->cond_set, "set the condition (ndiff_set_condition on ndiff_trainer_engine()) before the step");
    NDIFF_REQUIRE(x_t_dev && time_dev && target_dev && loss_weight_dev, "null argument");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    const int HW = t->H * t->W;
    const int td = e->dim_real * 4;
    // ---- inputs: x_t, target (fp32 NCHW -> NHWC4), per-sample t and loss weights
    if (nchw_to_nhwc4_launch(x_t_dev, e->x, t->B, HW, s)) return 1;
    if (nchw_to_nhwc4_launch(target_dev, t->target, t->B, HW, s)) return 1;
    NDIFF_CUDA_OK(cudaMemcpyAsync(t->w_b, loss_weight_dev, sizeof(float) * t->B, cudaMemcpyDeviceToDevice, s));
    i64_to_i32_kernel2<<<1, 256, 0, s>>>(reinterpret_cast<const long long*>(time_dev), e->t_buf, t->B);
    NDIFF_CUDA_OK(cudaGetLastError());
    g_use_pdl = false;
    // Everything below works on the trainer's own buffers: ~650 launches whose arguments never change -> one CUDA graph.  The
    // first step runs eagerly (one-time kernel attribute set-up happens there), the second is captured, later steps replay.
    auto run_lists = [&](cudaStream_t st) -> int {
        // ---- zero what is accumulated with atomics
        NDIFF_CUDA_OK(cudaMemsetAsync(t->flat_g, 0, t->n_flat * sizeof(float), st));
        NDIFF_CUDA_OK(cudaMemsetAsync(t->dcv, 0, static_cast<size_t>(t->B) * e->cv_total * sizeof(float), st));
        NDIFF_CUDA_OK(cudaMemsetAsync(t->loss, 0, sizeof(double), st));
        NDIFF_CUDA_OK(cudaMemsetAsync(e->stats, 0, e->stats_bytes, st));
        // ---- time path (per-sample t): st = SiLU(time_mlp(t)), every ResnetBlock's (scale, shift)
        if (time_mlp_train_launch(e->t_buf, t->B, e->dim_real, e->pf("time_mlp.1.weight"), e->pf("time_mlp.1.bias"),
                                  e->pf("time_mlp.3.weight"), e->pf("time_mlp.3.bias"), e->st_buf, t->st_saved, st)) return 1;
        if (rows_gemv_launch(e->ss_w, e->ss_b, e->st_buf, e->ss_cur, t->B, e->ss_total, td, st)) return 1;
        for (auto& f : t->fwd) if (f(st)) return 1;
        for (auto& f : t->bwd) if (f(st)) return 1;
        return 0;
    };
    const bool use_graph = (e->cfg.flags & NDIFF_FLAG_NO_GRAPH) == 0 && t->steps_run >= 1;
    if (use_graph && !t->step_exec) {
        NDIFF_CUDA_OK(cudaStreamSynchronize(s));
        cudaGraph_t graph = nullptr;
        NDIFF_CUDA_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = run_lists(e->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
        NDIFF_CUDA_OK(ce);
        NDIFF_CUDA_OK(cudaGraphInstantiate(&t->step_exec, graph, 0));
        cudaGraphDestroy(graph);
    }
    if (use_graph) NDIFF_CUDA_OK(cudaGraphLaunch(t->step_exec, s));
    else if (run_lists(s)) return 1;
    t->steps_run += 1;
    if (loss_host) {
        NDIFF_CUDA_OK(cudaMemcpyAsync(loss_host, t->loss, sizeof(double), cudaMemcpyDeviceToHost, s));
        NDIFF_CUDA_OK(cudaStreamSynchronize(s));
    }
    return 0;
}

int32_t ndiff_trainer_flat(ndiff_trainer* t, int32_t which, float** ptr, int64_t* n_live, int64_t* n_total) {
    NDIFF_REQUIRE(t && t->built && ptr, "trainer not finalized / null argument");
    float* tab[5] = {t->flat_p, t->flat_g, t->flat_m, t->flat_v, t->flat_ema};
    NDIFF_REQUIRE(which >= 0 && which < 5, "buffer index: 0 parameters, 1 gradients, 2 Adam m, 3 Adam v, 4 EMA");
    *ptr = tab[which];
    if (n_live) *n_live = static_cast<int64_t>(t->n_live);
    if (n_total) *n_total = static_cast<int64_t>(t->n_flat);
    return 0;
}

int32_t ndiff_trainer_slot(ndiff_trainer* t, const char* name, int64_t* offset, int64_t* numel) {
    NDIFF_REQUIRE(t && t->built && name && offset && numel, "trainer not finalized / null argument");
    auto it = t->slot.find(name);
    NDIFF_REQUIRE(it != t->slot.end(), std::string("no such parameter '") + name + "'");
    *offset = static_cast<int64_t>(it->second.first);
    *numel = static_cast<int64_t>(it->second.second);
    return 0;
}

int32_t ndiff_trainer_adam_step(ndiff_trainer* t, float lr, float beta1, float beta2, float eps, float weight_decay,
                                float grad_scale, void* stream) {
    NDIFF_REQUIRE(t && t->built, "trainer not finalized");
    ndiff_engine* e = t->e;
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    t->adam_steps += 1;
    if (adam_launch(t->flat_p, t->flat_g, t->flat_m, t->flat_v, t->n_live, lr, beta1, beta2, eps, weight_decay, t->adam_steps,
                    grad_scale, s)) return 1;
    // every derived weight form (bf16 GEMM packs, stacked heads, dgrad packs) follows the fp32 parameters; the condition-derived
    // buffers (maps, attention vectors) are stale until the next ndiff_set_condition
    if (engine_finalize(e, s)) return 1;
    e->cond_set = false;
    return repack_dgrad(t, s);
}

int32_t ndiff_trainer_ema_update(ndiff_trainer* t, float weight, void* stream) {
    NDIFF_REQUIRE(t && t->built, "trainer not finalized");
    DeviceGuard dev_guard(t->e->cfg.device);
    // weight = 1 - decay; weight >= 1 copies the parameters (ema_pytorch's copy_params_from_model_to_ema)
    if (weight >= 1.0f) {
        NDIFF_CUDA_OK(cudaMemcpyAsync(t->flat_ema, t->flat_p, t->n_flat * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
        return 0;
    }
    return ema_lerp_launch(t->flat_ema, t->flat_p, t->n_flat, weight, as_stream(stream));
}

int32_t ndiff_trainer_time(ndiff_trainer* t, char* out, int32_t cap, void* stream) {
    // Re-runs the forward and backward launch lists of the LAST step's inputs with a CUDA event around every launch and sums the
    // durations per kernel family; writes lines "fwd|bwd;family;launches;ms".  (Gradients accumulate once more: call it after
    // the measurements that matter.)
    NDIFF_REQUIRE(t && t->built && out && cap > 0, "trainer not finalized / null argument");
    DeviceGuard dev_guard(t->e->cfg.device);
    cudaStream_t s = as_stream(stream);
    std::map<std::string, std::pair<int, double>> acc;
    cudaEvent_t e0, e1;
    NDIFF_CUDA_OK(cudaEventCreate(&e0));
    NDIFF_CUDA_OK(cudaEventCreate(&e1));
    NDIFF_CUDA_OK(cudaMemsetAsync(t->e->stats, 0, t->e->stats_bytes, s));
    for (int pass = 0; pass < 2; ++pass) {
        auto& list = pass ? t->bwd : t->fwd;
        auto& kinds = pass ? t->bwd_kind : t->fwd_kind;
        for (size_t i = 0; i < list.size(); ++i) {
            NDIFF_CUDA_OK(cudaEventRecord(e0, s));
            if (list[i](s)) return 1;
            NDIFF_CUDA_OK(cudaEventRecord(e1, s));
            NDIFF_CUDA_OK(cudaEventSynchronize(e1));
            float ms = 0.f;
            NDIFF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
            auto& a = acc[std::string(pass ? "bwd;" : "fwd;") + kinds[i]];
            a.first += 1; a.second += ms;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    std::string txt;
    for (auto& kv : acc) txt += kv.first + ";" + std::to_string(kv.second.first) + ";" + std::to_string(kv.second.second) + "\n";
    strncpy(out, txt.c_str(), cap - 1);
    out[cap - 1] = 0;
    return 0;
}

int64_t ndiff_trainer_activation_bytes(const ndiff_trainer* t) { return t ? static_cast<int64_t>(t->act_bytes) : 0; }
int64_t ndiff_trainer_launches(const ndiff_trainer* t, int32_t backward) {
    return t ? static_cast<int64_t>(backward ? t->bwd.size() : t->fwd.size()) : 0;
}

// ---- single-operator entry points (parity tests of the backward kernels against torch autograd) ---------------------------------
int32_t ndiff_op_wgrad(int32_t mode, int32_t B, int32_t H, int32_t W, const void* dy, int32_t Cout, const void* src0, int32_t C0,
                       const void* src1, int32_t C1, float* dw, void* stream) {
    int dev = 0, sms = 0;
    NDIFF_CUDA_OK(cudaGetDevice(&dev));
    NDIFF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WgradDesc d;
    d.mode = mode; d.B = B; d.H = H; d.W = W;
    d.dy = static_cast<const bf16*>(dy); d.Cout = Cout;
    d.src0 = static_cast<const bf16*>(src0); d.C0 = C0;
    d.src1 = static_cast<const bf16*>(src1); d.C1 = C1;
    d.dw = dw;
    WgradPlan plan;
    if (wgrad_gemm_plan(d, sms, &plan)) return 1;
    return wgrad_gemm_launch(plan, as_stream(stream));
}

int32_t ndiff_op_gn_backward(const void* h, const void* dout, void* dh, const void* stats, const float* gamma, const float* beta,
                             const float* ss, int32_t ss_ld, int32_t ss_off, const void* maps, void* dmaps, float* dgamma,
                             float* dbeta, float* dss, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream) {
    float* acc = nullptr;
    NDIFF_CUDA_OK(cudaMalloc(&acc, static_cast<size_t>(B) * (C * 2 + G * 2) * sizeof(float)));
    GnBwdArgs a{};
    a.h = static_cast<const bf16*>(h); a.dout = static_cast<const bf16*>(dout); a.dh = static_cast<bf16*>(dh);
    a.stats = static_cast<const unsigned long long*>(stats); a.gamma = gamma; a.beta = beta;
    a.ss = ss; a.ss_ld = ss_ld; a.ss_off = ss_off;
    a.maps = static_cast<const bf16*>(maps); a.dmaps = static_cast<bf16*>(dmaps);
    a.acc = acc; a.dgamma = dgamma; a.dbeta = dbeta; a.dss = dss; a.dss_ld = ss_ld;
    a.B = B; a.HW = HW; a.C = C; a.G = G; a.eps = 1e-5f; a.real_frac = 1.0f;
    const int rc = gn_backward_launch(a, as_stream(stream));
    cudaStreamSynchronize(as_stream(stream));
    cudaFree(acc);
    return rc;
}

int32_t ndiff_op_layernorm_backward(const void* x, const float* vec, int32_t vec_ld, const float* g, const void* du, void* dy,
                                    float* dg, float* dbeta, int32_t B, int32_t HW, int32_t C, void* stream) {
    return layernorm_backward_launch(static_cast<const bf16*>(x), vec, vec_ld, g, static_cast<const bf16*>(du), static_cast<bf16*>(dy),
                                     dg, dbeta, B, HW, C, 1.0f, as_stream(stream));
}

}  // extern "C"
