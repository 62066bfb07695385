// noisediff_b200 — engine: weights, layer plan, CUDA-graph step, and the C ABI (include/noisediff_b200.h).
//
// The plan below is NoiseDiffNet.forward (reference models/archs/Diffusion_arch.py:577-646) re-expressed as a flat
// list of kernel launches over NHWC bf16 buffers:
//   * every Block.proj / res_conv / Downsample / AttnBlock.ff / proj_out / Mlp.fc is one tcgen05 implicit-GEMM launch;
//   * GroupNorm statistics come out of the conv epilogue, normalise + (scale+1)/shift + SiLU (+ residual adds) is one
//     pass; cat((x, skip)) is never materialised (two-source K loop);
//   * the 1-token cross attention collapses to a per-sample vector (softmax over one key == 1), the
//     positional branch and the time MLP are hoisted out of the per-pixel work;
//   * final_conv + shot_mlp3.fc2 + the DDPM/DDIM posterior update are one kernel; a whole step replays as one graph.
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "engine_internal.cuh"

namespace ndiff {

bool g_use_pdl = false;
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

namespace {

__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int Cin, int KH,
                                   int KW, int s2d) {
    // dst[co][cblk][tap][64]  <-  src[co][ci][ky][kx]   (s2d: src[co][c*4 + tap], KH = KW = 1 in storage)
    const int taps = s2d ? 4 : KH * KW;
    const size_t total = static_cast<size_t>(Cout) * Cin * taps;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cl = static_cast<int>(i % 64);
        size_t r = i / 64;
        const int tap = static_cast<int>(r % taps); r /= taps;
        const int cb = static_cast<int>(r % (Cin / 64));
        const int co = static_cast<int>(r / (Cin / 64));
        const int ci = cb * 64 + cl;
        const size_t s = s2d ? (static_cast<size_t>(co) * Cin * 4 + static_cast<size_t>(ci) * 4 + tap)
                             : ((static_cast<size_t>(co) * Cin + ci) * taps + tap);
        dst[i] = __float2bfloat16_rn(src[s]);
    }
}

// ResnetBlock block1 conv + res_conv sharing one pass over the input (conv_gemm.cuh kHalo1R):
// dst[co][cblk][10][64]: taps 0..8 = the 3x3 kernel, tap 9 = the 1x1 res_conv weight.
__global__ void pack_weight_res_kernel(const float* __restrict__ w3, const float* __restrict__ w1, bf16* __restrict__ dst,
                                       int Cout, int Cin) {
    const size_t total = static_cast<size_t>(Cout) * Cin * 10;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cl = static_cast<int>(i % 64);
        size_t r = i / 64;
        const int tap = static_cast<int>(r % 10); r /= 10;
        const int cb = static_cast<int>(r % (Cin / 64));
        const int co = static_cast<int>(r / (Cin / 64));
        const int ci = cb * 64 + cl;
        const size_t o = static_cast<size_t>(co) * Cin + ci;
        dst[i] = __float2bfloat16_rn(tap < 9 ? w3[o * 9 + tap] : w1[o]);
    }
}

// Upsample(nearest x2) + conv3x3 as four 2x2 phase convolutions (conv_gemm.cuh kHaloUp):
// dst[co][cblk][phase = py*2+px][tap = dy*2+dx][64]; rows {w0 | w1+w2} for py = 0 and {w0+w1 | w2} for py = 1, same in x.
__global__ void pack_upconv_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int Cin) {
    const size_t total = static_cast<size_t>(Cout) * Cin * 16;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int kk = static_cast<int>(i % 64);
        size_t r = i / 64;
        const int tap = static_cast<int>(r % 4); r /= 4;
        const int phase = static_cast<int>(r % 4); r /= 4;
        const int cb = static_cast<int>(r % (Cin / 64));
        const int co = static_cast<int>(r / (Cin / 64));
        const int ci = cb * 64 + kk, py = phase >> 1, px = phase & 1, dy = tap >> 1, dx = tap & 1;
        // taps of the 3x3 kernel that read low-res row y - 1 + py + dy: py=0: {0},{1,2}; py=1: {0,1},{2}
        const int ky0 = py == 0 ? (dy == 0 ? 0 : 1) : (dy == 0 ? 0 : 2), ky1 = py == 0 ? (dy == 0 ? 0 : 2) : (dy == 0 ? 1 : 2);
        const int kx0 = px == 0 ? (dx == 0 ? 0 : 1) : (dx == 0 ? 0 : 2), kx1 = px == 0 ? (dx == 0 ? 0 : 2) : (dx == 0 ? 1 : 2);
        const float* w = src + (static_cast<size_t>(co) * Cin + ci) * 9;
        float acc = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) acc += w[ky * 3 + kx];
        dst[i] = __float2bfloat16_rn(acc);
    }
}

// AttnBlock (C >= 128): ff.net.2 and proj_out meet without a nonlinearity (ref Diffusion_arch.py:439-443):
//   out = Wp (W2 h + b2 + c + x) + bp + x = (Wp W2) h + Wp x + [Wp (b2 + c) + bp] + x
// -> ONE two-source GEMM over [h | x] with the packed operand [Wp W2 | Wp] (folded in fp32, rounded to bf16 once), a per-sample
// vector and the residual x, instead of two GEMM launches with the C-channel tensor z written and re-read in between.
// dst[co][cblk][64]: cblk < 2C/64 from the fold, then C/64 blocks of Wp.
__global__ void pack_attn_ff2proj_kernel(const float* __restrict__ wp, const float* __restrict__ w2, bf16* __restrict__ dst, int C) {
    const int K = 3 * C;
    const size_t total = static_cast<size_t>(C) * K;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i / K), k = static_cast<int>(i % K);
        float acc;
        if (k < 2 * C) {
            acc = 0.f;
            for (int j = 0; j < C; ++j) acc = fmaf(wp[static_cast<size_t>(co) * C + j], w2[static_cast<size_t>(j) * 2 * C + k], acc);
        } else {
            acc = wp[static_cast<size_t>(co) * C + (k - 2 * C)];
        }
        dst[i] = __float2bfloat16_rn(acc);
    }
}
// out[N][K] = a[N][M] @ b[M][K]  (fp32, tiny: weight folds at pack time).  grid = N blocks, one thread per output column
__global__ void fold_matmul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int N, int M, int K) {
    const int n = blockIdx.x;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < M; ++j) acc = fmaf(a[n * M + j], b[j * K + k], acc);
        out[n * K + k] = acc;
    }
}
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}
// per-sample vector of the folded form: v2[b][c] = sum_k Wp[c][k] (b2[k] + cvec[b][k]) + bp[c]
__global__ void attn_vec2_kernel(const float* __restrict__ wp, const float* __restrict__ bp, const float* __restrict__ b2,
                                 const float* __restrict__ cvec, float* __restrict__ out, int ld, int off, int C) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = bp[c];
        for (int k = 0; k < C; ++k) acc = fmaf(wp[static_cast<size_t>(c) * C + k], b2[k] + cvec[static_cast<size_t>(b) * ld + off + k], acc);
        out[static_cast<size_t>(b) * ld + off + c] = acc;
    }
}

__global__ void init_conv_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout) {
    // dst[tap(49)][ci(4)][co]  <-  src[co][ci][7][7]
    const int total = Cout * 4 * 49;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout, ci = (i / Cout) % 4, tap = i / (Cout * 4);
        dst[i] = src[(co * 4 + ci) * 49 + tap];
    }
}

// init_conv on tensor cores: x_t is re-laid as bf16 [B][H+6][W+8][8] with a 3-pixel zero border, where element (r, c) holds
// the 4 channels of padded pixel (r, c) AND, in channels 4..7, those of the pixel one row below (r + 1, c).  One contiguous
// 128-byte window (8 pixels x 8 values) is then the 7 taps of TWO kernel rows -> one SWIZZLE_128B operand row, and the 7x7 conv
// is 4 row-pair GEMM taps (input rows y, y+2, y+4, y+6) instead of 7 single-row taps whose K blocks were half zeros.
__global__ void xpad_pack_kernel(const float4* __restrict__ x, uint2* __restrict__ xpad, int H, int W, size_t npix) {
    pdl_trigger();
    pdl_wait();
    const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const int xx = static_cast<int>(pix % W);
    const size_t r = pix / W;
    const int yy = static_cast<int>(r % H);
    const size_t b = r / H;
    const float4 v = x[pix];
    const uint2 o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    const size_t at = (b * (H + 6) + yy + 3) * (W + 8) + xx + 3;      // element index of padded pixel (yy + 3, xx + 3)
    xpad[at * 2] = o;                                                  // its own row: channels 0..3
    xpad[(at - (W + 8)) * 2 + 1] = o;                                  // the row above sees it as "one row below": channels 4..7
}

__global__ void init_conv_pack_tc_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout) {
    // dst[co][kp(4)][kxw(8)][8]  <-  src[co][ci(4)][7][7]: values 0..3 = kernel row 2 kp, values 4..7 = kernel row 2 kp + 1;
    // zero where the row is 7 or kxw == 7
    const int total = Cout * 4 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int v8 = i & 7, kxw = (i >> 3) & 7, kp = (i >> 6) & 3, co = i >> 8;
        const int ci = v8 & 3, ky = 2 * kp + (v8 >> 2);
        const float v = (ky < 7 && kxw < 7) ? src[((co * 4 + ci) * 7 + ky) * 7 + kxw] : 0.f;
        dst[i] = __float2bfloat16_rn(v);
    }
}

__global__ void embed_param_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc, int inner,
                                   const int* __restrict__ rmap, const int* __restrict__ cmap, int pC) {
    const size_t total = static_cast<size_t>(R) * Cc * inner;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % inner);
        const int c = static_cast<int>((i / inner) % Cc);
        const int r = static_cast<int>(i / (static_cast<size_t>(inner) * Cc));
        dst[(static_cast<size_t>(rmap[r]) * pC + cmap[c]) * inner + k] = src[i];
    }
}

__global__ void i64_to_i32_kernel(const long long* in, int* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<int>(in[i]);
}

__global__ void bf16_nhwc_to_f32_nchw_kernel(const bf16* __restrict__ in, float* __restrict__ out, int C, int HW,
                                             size_t total) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const size_t p = i / C;
        const size_t b = p / HW, hw = p % HW;
        out[(b * C + c) * HW + hw] = __bfloat162float(in[i]);
    }
}

// step prologue: publish this step's scalars + time vectors; optional teacher forcing of the state
__global__ void zero_u64_kernel(unsigned long long* p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0ull;
}

__global__ void __launch_bounds__(256) chain_step_begin_kernel(ChainState* chain, const StepParams* table,
                                                               const float* __restrict__ ss_table, int ss_len,
                                                               float* __restrict__ ss_cur, int B, int HW,
                                                               float4* __restrict__ x, unsigned long long* stats,
                                                               int n_stats_words) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_stats_words; i += gridDim.x * blockDim.x) stats[i] = 0ull;
    const int step = chain->step;
    if (blockIdx.x == 0 && threadIdx.x == 0) chain->cur = table[step];
    const float* row = ss_table + static_cast<size_t>(step) * ss_len;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < ss_len * B; i += nth) ss_cur[i] = row[i % ss_len];
    const float* teacher = chain->teacher;
    if (teacher) {
        const size_t npix = static_cast<size_t>(B) * HW;
        const float* tp = teacher + static_cast<size_t>(step - chain->base_step) * npix * 4;
        for (size_t pix = tid; pix < npix; pix += nth) {
            const size_t b = pix / HW, hw = pix % HW;
            const float* p = tp + b * 4 * HW + hw;
            x[pix] = make_float4(p[0], p[HW], p[2 * static_cast<size_t>(HW)], p[3 * static_cast<size_t>(HW)]);
        }
    }
}

}  // namespace
}  // namespace ndiff

namespace ndiff {

std::vector<RbSpec> resblocks(int dim) {
    std::vector<RbSpec> v;
    const int d[5] = {dim, dim, dim * 2, dim * 4, dim * 8};
    v.push_back({"shot_time", dim, dim, 2, false, dim, 0});
    v.push_back({"pos_block1", dim, dim, 2, true, dim, 0});
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j) v.push_back({"downs." + std::to_string(i) + "." + std::to_string(j), d[i], d[i], 8, false, d[i], 0});
    v.push_back({"mid_block1", d[4], d[4], 8, false, d[4], 0});
    v.push_back({"mid_block2", d[4], d[4], 8, false, d[4], 0});
    for (int i = 0; i < 4; ++i) {
        const int co = d[4 - i], ci = d[3 - i];
        for (int j = 0; j < 2; ++j) v.push_back({"ups." + std::to_string(i) + "." + std::to_string(j), co + ci, co, 8, false, co, ci});
    }
    v.push_back({"pos_block2", dim, dim, 2, true, dim, 0});
    v.push_back({"final_res_block", dim * 2, dim, 8, false, dim, dim});
    return v;
}

std::vector<AttnSpec> attnblocks(int dim) {
    std::vector<AttnSpec> v;
    const int d[5] = {dim, dim, dim * 2, dim * 4, dim * 8};
    v.push_back({"shot_attn", dim});
    for (int i = 0; i < 4; ++i) v.push_back({"downs." + std::to_string(i) + ".2", d[i]});
    for (int i = 0; i < 4; ++i) v.push_back({"ups." + std::to_string(i) + ".2", d[4 - i]});
    return v;
}
}  // namespace ndiff

using namespace ndiff;

namespace {

int check_shape(ndiff_engine* e, const std::string& name, std::initializer_list<int64_t> want) {
    const Param* p = e->param(name);
    NDIFF_REQUIRE(p != nullptr, "missing state_dict entry '" + name + "'");
    std::vector<int64_t> w(want);
    bool ok = p->shape.size() == w.size();
    for (size_t i = 0; ok && i < w.size(); ++i) ok = p->shape[i] == w[i];
    NDIFF_REQUIRE(ok, "state_dict entry '" + name + "' has an unexpected shape");
    return 0;
}

int pack_conv(ndiff_engine* e, const std::string& name, int Cout, int Cin, int K, bool s2d, cudaStream_t s) {
    if (s2d) { if (check_shape(e, name + ".weight", {Cout, Cin * 4, 1, 1})) return 1; }
    else if (check_shape(e, name + ".weight", {Cout, Cin, K, K})) return 1;
    if (check_shape(e, name + ".bias", {Cout})) return 1;
    const int taps = s2d ? 4 : K * K;
    bf16* dst = e->packed.count(name) ? e->packed[name] : nullptr;
    if (!dst && e->alloc(&dst, static_cast<size_t>(Cout) * Cin * taps)) return 1;
    pack_weight_kernel<<<256, 256, 0, s>>>(e->pf(name + ".weight"), dst, Cout, Cin, K, K, s2d ? 1 : 0);
    NDIFF_CUDA_OK(cudaGetLastError());
    e->packed[name] = dst;
    return 0;
}
int pack_linear(ndiff_engine* e, const std::string& name, int N, int K, cudaStream_t s) {
    if (check_shape(e, name + ".weight", {N, K})) return 1;
    if (check_shape(e, name + ".bias", {N})) return 1;
    bf16* dst = e->packed.count(name) ? e->packed[name] : nullptr;
    if (!dst && e->alloc(&dst, static_cast<size_t>(N) * K)) return 1;
    pack_weight_kernel<<<256, 256, 0, s>>>(e->pf(name + ".weight"), dst, N, K, 1, 1, 0);
    NDIFF_CUDA_OK(cudaGetLastError());
    e->packed[name] = dst;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Models narrower than the kernels' 64-channel base (the shipped checkpoint: --dim 48, script.sh:10; channels 48/96/192/384,
// GroupNorm groups of 6/12/24/48): every tensor with C live channels is laid out in C' = C * 64 / dim PHYSICAL channels as eight
// slots of (C/8 live + zero padding) — slot = GroupNorm group of the 8-group blocks, and four consecutive slots = one group of
// the 2-group blocks, so group boundaries stay powers of two in physical channels.  Weights are embedded once per load into
// zero-initialised physical tensors (rows = permuted output channels, columns = permuted input channels of each concatenated
// source); padding rows / biases / gammas are 0, so padded channels are exactly 0 everywhere and the network is, physically, the
// dim = 64 network.  Only the element COUNT of the normalisations changes (real_frac).  Cost: the padded MACs are executed.
// ------------------------------------------------------------------------------------------------------------
struct EmbedSpec {
    std::string name;
    std::vector<int64_t> lshape, pshape;      // logical (state_dict) and physical shapes
    std::vector<int> rmap, cmap;              // logical row / column -> physical row / column
};

struct EmbedBuilder {
    int dim, base;                             // live and physical base width
    std::vector<EmbedSpec> specs;
    int P(int C) const { return C / dim * base; }                       // physical width of a C-channel tensor
    std::vector<int> ident(int n) const { std::vector<int> m(n); for (int i = 0; i < n; ++i) m[i] = i; return m; }
    std::vector<int> perm(int C) const {      // eight slots of C/8 live channels inside P(C)/8 physical ones
        std::vector<int> m(C);
        const int rs = C / 8, ps = P(C) / 8;
        for (int c = 0; c < C; ++c) m[c] = (c / rs) * ps + c % rs;
        return m;
    }
    std::vector<int> cat(int c0, int c1) const {   // [x | skip]: each source in its own layout, skip after ALL of x's physical channels
        std::vector<int> m = perm(c0);
        if (c1 > 0) for (int v : perm(c1)) m.push_back(P(c0) + v);
        return m;
    }
    std::vector<int> scale_shift(int C) const {    // [scale C | shift C]
        std::vector<int> m = perm(C);
        for (int v : perm(C)) m.push_back(P(C) + v);
        return m;
    }
    std::vector<int> s2d(int C) const {            // input channel c * 4 + p1 * 2 + p2 (Diffusion_arch.py:78-82)
        std::vector<int> m(4 * C), pm = perm(C);
        for (int c = 0; c < C; ++c) for (int t = 0; t < 4; ++t) m[c * 4 + t] = pm[c] * 4 + t;
        return m;
    }
    void mat(const std::string& name, std::vector<int> rm, int pR, std::vector<int> cm, int pC, int kh = 0, int kw = 0) {
        EmbedSpec sp;
        sp.name = name;
        sp.lshape = {static_cast<int64_t>(rm.size()), static_cast<int64_t>(cm.size())};
        sp.pshape = {pR, pC};
        if (kh > 0) { sp.lshape.push_back(kh); sp.lshape.push_back(kw); sp.pshape.push_back(kh); sp.pshape.push_back(kw); }
        sp.rmap = std::move(rm); sp.cmap = std::move(cm);
        specs.push_back(std::move(sp));
    }
    void vec(const std::string& name, std::vector<int> rm, int pR) {
        EmbedSpec sp;
        sp.name = name;
        sp.lshape = {static_cast<int64_t>(rm.size())};
        sp.pshape = {pR};
        sp.rmap = std::move(rm); sp.cmap = {0};
        specs.push_back(std::move(sp));
    }
    void conv(const std::string& n, int co, std::vector<int> cm, int pCin, int k) {
        mat(n + ".weight", perm(co), P(co), std::move(cm), pCin, k, k);
        vec(n + ".bias", perm(co), P(co));
    }
};

int embed_params(ndiff_engine* e, cudaStream_t s);

int finalize(ndiff_engine* e, cudaStream_t s) {
    if (e->dim_real != e->dim && embed_params(e, s)) return 1;
    const int dim = e->dim, td = e->dim_real * 4;      // channel counts are physical; the time vector keeps its logical width
    const int d[5] = {dim, dim, dim * 2, dim * 4, dim * 8};
    // --- convolution / GEMM weights -> bf16 [Cout][cblk][tap][64]
    for (const RbSpec& rb : resblocks(dim)) {
        if (pack_conv(e, rb.name + ".block1.proj", rb.cout, rb.cin, 3, false, s)) return 1;
        if (pack_conv(e, rb.name + ".block2.proj", rb.cout, rb.cout, 3, false, s)) return 1;
        if (rb.cin != rb.cout && pack_conv(e, rb.name + ".res_conv", rb.cout, rb.cin, 1, false, s)) return 1;
        if (rb.cin != rb.cout) {   // the same two layers as ONE operand for the fused kHalo1R form
            const std::string key = rb.name + ".block1.proj#res";
            bf16* dst = e->packed.count(key) ? e->packed[key] : nullptr;
            if (!dst && e->alloc(&dst, static_cast<size_t>(rb.cout) * rb.cin * 10)) return 1;
            pack_weight_res_kernel<<<256, 256, 0, s>>>(e->pf(rb.name + ".block1.proj.weight"), e->pf(rb.name + ".res_conv.weight"),
                                                       dst, rb.cout, rb.cin);
            NDIFF_CUDA_OK(cudaGetLastError());
            e->packed[key] = dst;
        }
        for (const char* blk : {".block1.norm", ".block2.norm"}) {
            if (check_shape(e, rb.name + blk + ".weight", {rb.cout})) return 1;
            if (check_shape(e, rb.name + blk + ".bias", {rb.cout})) return 1;
        }
    }
    for (const AttnSpec& ab : attnblocks(dim)) {
        if (pack_linear(e, ab.name + ".ff.net.0.0", 2 * ab.C, ab.C, s)) return 1;
        if (pack_linear(e, ab.name + ".ff.net.2", ab.C, 2 * ab.C, s)) return 1;
        if (pack_conv(e, ab.name + ".proj_out", ab.C, ab.C, 1, false, s)) return 1;
        if (check_shape(e, ab.name + ".attn.to_v.weight", {128, 16})) return 1;
        if (check_shape(e, ab.name + ".attn.to_out.0.weight", {ab.C, 128})) return 1;
        if (check_shape(e, ab.name + ".attn.to_out.0.bias", {ab.C})) return 1;
        if (check_shape(e, ab.name + ".norm2.weight", {ab.C})) return 1;
        if (check_shape(e, ab.name + ".norm2.bias", {ab.C})) return 1;
    }
    for (const AttnSpec& ab : attnblocks(dim)) {
        if (ab.C < 128) continue;
        const std::string key = ab.name + "#ff2proj";
        bf16* dst = e->packed.count(key) ? e->packed[key] : nullptr;
        if (!dst && e->alloc(&dst, static_cast<size_t>(ab.C) * 3 * ab.C)) return 1;
        pack_attn_ff2proj_kernel<<<256, 256, 0, s>>>(e->pf(ab.name + ".proj_out.weight"), e->pf(ab.name + ".ff.net.2.weight"), dst, ab.C);
        NDIFF_CUDA_OK(cudaGetLastError());
        e->packed[key] = dst;
    }
    // --- fused per-pixel chains (pixel_chain.cuh): K-blocked weight blobs + fp32 parameter blocks
    // fold_proj: the stand-alone AttnBlock chain (W2 slot = Wp W2, Wp slot = Wp); otherwise the caller fills both slots
    auto pack_attn_chain = [&](const std::string& n, bf16* w, float* f, bool fold_proj) -> int {
        // AttnBlock.norm2's affine folded into ff.net.0.0 (exact algebra; the fold runs in fp32 before the bf16 rounding)
        if (!e->fold_tmp && e->alloc(&e->fold_tmp, 128 * 64)) return 1;
        if (fold_layernorm_launch(e->pf(n + ".ff.net.0.0.weight"), e->pf(n + ".ff.net.0.0.bias"), e->pf(n + ".norm2.weight"),
                                  e->pf(n + ".norm2.bias"), e->fold_tmp, f + 128, 128, 64, s)) return 1;
        if (pack_chain_weight_launch(e->fold_tmp, w, 128, 64, false, s)) return 1;
        if (fold_proj) {      // stand-alone AttnBlock chain: ff.net.2's slot holds Wp W2 (folded in fp32, rounded to fp16 once)
            fold_matmul_kernel<<<64, 128, 0, s>>>(e->pf(n + ".proj_out.weight"), e->pf(n + ".ff.net.2.weight"), e->fold_tmp, 64, 64, 128);
            NDIFF_CUDA_OK(cudaGetLastError());
            if (pack_chain_weight_launch(e->fold_tmp, w + 128 * 64, 64, 128, true, s)) return 1;
            if (pack_chain_weight_launch(e->pf(n + ".proj_out.weight"), w + 256 * 64, 64, 64, false, s)) return 1;
        }
        NDIFF_CUDA_OK(cudaMemsetAsync(f, 0, sizeof(float) * 128, s));                      // reserved
        NDIFF_CUDA_OK(cudaMemcpyAsync(f + 256, e->pf(n + ".ff.net.2.bias"), sizeof(float) * 64, cudaMemcpyDeviceToDevice, s));
        NDIFF_CUDA_OK(cudaMemcpyAsync(f + 320, e->pf(n + ".proj_out.bias"), sizeof(float) * 64, cudaMemcpyDeviceToDevice, s));
        return 0;
    };
    if (dim == 64) {
        for (const AttnSpec& ab : attnblocks(dim)) {
            if (ab.C != 64 || ab.name == "shot_attn") continue;
            if (!e->chain_w.count(ab.name)) {
                if (e->alloc(&e->chain_w[ab.name], static_cast<size_t>(kChainAttnRows) * 64)) return 1;
                if (e->alloc(&e->chain_f[ab.name], kChainAttnFloats)) return 1;
            }
            if (pack_attn_chain(ab.name, e->chain_w[ab.name], e->chain_f[ab.name], true)) return 1;
        }
        if (!e->chain_w.count("shot")) {
            if (e->alloc(&e->chain_w["shot"], static_cast<size_t>(kChainShotRows) * 64)) return 1;
            if (e->alloc(&e->chain_f["shot"], kChainShotFloats)) return 1;
        }
        bf16* w = e->chain_w["shot"];
        float* f = e->chain_f["shot"];
        if (pack_chain_weight_launch(e->pf("shot_mlp1.fc1.weight"), w, 64, 8, false, s)) return 1;
        if (pack_chain_weight_launch(e->pf("shot_mlp1.fc2.weight"), w + 64 * 64, 64, 64, true, s)) return 1;
        if (pack_attn_chain("shot_attn", w + 128 * 64, f + 128, false)) return 1;
        // shot_attn.ff.net.2 + proj_out + shot_mlp2.fc1 as one stage (pixel_chain.cuh): with M = Wm1 Wp (kept in fp32 for the
        // per-sample vector, f[512..] = Wm1 bp + bm1), rows 256..383 = fp16(M W2) applied to the hidden layer, rows 384..447 =
        // bf16(M + Wm1) applied to s1
        if (!e->shot_fold && e->alloc(&e->shot_fold, 64 * 64)) return 1;
        if (fold_linear_launch(e->pf("shot_mlp2.fc1.weight"), e->pf("shot_attn.proj_out.weight"), e->pf("shot_attn.proj_out.bias"),
                               e->pf("shot_mlp2.fc1.bias"), e->shot_fold, f + 512, 64, s)) return 1;
        fold_matmul_kernel<<<64, 128, 0, s>>>(e->shot_fold, e->pf("shot_attn.ff.net.2.weight"), e->fold_tmp, 64, 64, 128);
        NDIFF_CUDA_OK(cudaGetLastError());
        if (pack_chain_weight_launch(e->fold_tmp, w + 256 * 64, 64, 128, true, s)) return 1;
        add_vec_kernel<<<16, 256, 0, s>>>(e->shot_fold, e->pf("shot_mlp2.fc1.weight"), e->fold_tmp, 64 * 64);
        NDIFF_CUDA_OK(cudaGetLastError());
        if (pack_chain_weight_launch(e->fold_tmp, w + 384 * 64, 64, 64, false, s)) return 1;
        if (pack_chain_weight_launch(e->pf("shot_mlp2.fc2.weight"), w + 448 * 64, 64, 64, true, s)) return 1;
        NDIFF_CUDA_OK(cudaMemcpyAsync(f, e->pf("shot_mlp1.fc1.bias"), sizeof(float) * 64, cudaMemcpyDeviceToDevice, s));
        NDIFF_CUDA_OK(cudaMemcpyAsync(f + 64, e->pf("shot_mlp1.fc2.bias"), sizeof(float) * 64, cudaMemcpyDeviceToDevice, s));
        NDIFF_CUDA_OK(cudaMemcpyAsync(f + 576, e->pf("shot_mlp2.fc2.bias"), sizeof(float) * 64, cudaMemcpyDeviceToDevice, s));
        // shot-branch tail: fc1 (bf16 rows) | fc2 (fp16, 4 live rows of 16) | zero padding
        if (!e->chain_w.count("tail")) {
            if (e->alloc(&e->chain_w["tail"], static_cast<size_t>(kTailRows) * 64)) return 1;
            if (e->alloc(&e->chain_f["tail"], kTailFloats)) return 1;
        }
        bf16* tw = e->chain_w["tail"];
        float* tf = e->chain_f["tail"];
        NDIFF_CUDA_OK(cudaMemsetAsync(tw, 0, sizeof(bf16) * kTailRows * 64, s));
        NDIFF_CUDA_OK(cudaMemsetAsync(tf, 0, sizeof(float) * kTailFloats, s));
        if (pack_chain_weight_launch(e->pf("shot_mlp3.fc1.weight"), tw, 64, 64, false, s)) return 1;
        if (pack_chain_weight_launch(e->pf("shot_mlp3.fc2.weight"), tw + 64 * 64, 4, 64, true, s)) return 1;
        NDIFF_CUDA_OK(cudaMemcpyAsync(tf, e->pf("shot_mlp3.fc1.bias"), sizeof(float) * 64, cudaMemcpyDeviceToDevice, s));
        NDIFF_CUDA_OK(cudaMemcpyAsync(tf + 64, e->pf("shot_mlp3.fc2.bias"), sizeof(float) * 4, cudaMemcpyDeviceToDevice, s));
    }
    for (int i = 0; i < 3; ++i) {
        if (pack_conv(e, "downs." + std::to_string(i) + ".3.1", d[i + 1], d[i], 1, true, s)) return 1;
        if (pack_conv(e, "ups." + std::to_string(i) + ".3.1", d[3 - i], d[4 - i], 3, false, s)) return 1;
        {   // the same layer in the fused nearest-x2 form (four 2x2 phase kernels)
            const std::string n = "ups." + std::to_string(i) + ".3.1";
            const int co = d[3 - i], ci = d[4 - i];
            bf16* dst = e->packed.count(n + "#up") ? e->packed[n + "#up"] : nullptr;
            if (!dst && e->alloc(&dst, static_cast<size_t>(co) * ci * 16)) return 1;
            pack_upconv_weight_kernel<<<256, 256, 0, s>>>(e->pf(n + ".weight"), dst, co, ci);
            NDIFF_CUDA_OK(cudaGetLastError());
            e->packed[n + "#up"] = dst;
        }
    }
    if (pack_conv(e, "downs.3.3", d[4], d[3], 3, false, s)) return 1;
    if (pack_conv(e, "ups.3.3", d[0], d[1], 3, false, s)) return 1;
    if (pack_conv(e, "shot_mlp1.fc2", dim, dim, 1, false, s)) return 1;
    if (pack_conv(e, "shot_mlp2.fc1", dim, dim, 1, false, s)) return 1;
    if (pack_conv(e, "shot_mlp2.fc2", dim, dim, 1, false, s)) return 1;
    if (pack_conv(e, "shot_mlp3.fc1", dim, dim, 1, false, s)) return 1;
    // --- small fp32 layers used as-is
    if (check_shape(e, "shot_mlp1.fc1.weight", {dim, 8, 1, 1}) || check_shape(e, "shot_mlp1.fc1.bias", {dim})) return 1;
    if (check_shape(e, "shot_mlp3.fc2.weight", {4, dim, 1, 1}) || check_shape(e, "shot_mlp3.fc2.bias", {4})) return 1;
    if (check_shape(e, "final_conv.weight", {4, dim, 1, 1}) || check_shape(e, "final_conv.bias", {4})) return 1;
    if (check_shape(e, "init_conv.weight", {dim, 4, 7, 7}) || check_shape(e, "init_conv.bias", {dim})) return 1;
    if (check_shape(e, "iso_embed.weight", {100, 16})) return 1;
    if (check_shape(e, "time_mlp.1.weight", {td, e->dim_real}) || check_shape(e, "time_mlp.1.bias", {td})) return 1;
    if (check_shape(e, "time_mlp.3.weight", {td, td}) || check_shape(e, "time_mlp.3.bias", {td})) return 1;
    if (check_shape(e, "pos_enc.weights.weight", {8, 2, 1, 1}) || check_shape(e, "pos_enc.weights.bias", {8})) return 1;
    if (check_shape(e, "pos_mlp.fc1.weight", {16, 24, 1, 1}) || check_shape(e, "pos_mlp.fc1.bias", {16})) return 1;
    if (check_shape(e, "pos_mlp.fc2.weight", {8, 16, 1, 1}) || check_shape(e, "pos_mlp.fc2.bias", {8})) return 1;
    for (const char* pb : {"pos_block1", "pos_block2"}) {
        if (check_shape(e, std::string(pb) + ".mlp.1.weight", {2 * dim, 8, 1, 1})) return 1;
        if (check_shape(e, std::string(pb) + ".mlp.1.bias", {2 * dim})) return 1;
    }
    if (!e->init_w && e->alloc(&e->init_w, static_cast<size_t>(dim) * 4 * 49)) return 1;
    init_conv_pack_kernel<<<64, 256, 0, s>>>(e->pf("init_conv.weight"), e->init_w, dim);
    NDIFF_CUDA_OK(cudaGetLastError());
    if (!e->init_w_tc && e->alloc(&e->init_w_tc, static_cast<size_t>(dim) * 7 * 64)) return 1;
    init_conv_pack_tc_kernel<<<64, 256, 0, s>>>(e->pf("init_conv.weight"), e->init_w_tc, dim);
    NDIFF_CUDA_OK(cudaGetLastError());
    // --- stacked time-MLP heads: rows [scale C | shift C] per ResnetBlock, in plan order
    e->ss_total = 0;
    e->ss_off.clear();
    for (const RbSpec& rb : resblocks(dim)) {
        if (rb.pos) continue;
        if (check_shape(e, rb.name + ".mlp.1.weight", {2 * rb.cout, td})) return 1;
        if (check_shape(e, rb.name + ".mlp.1.bias", {2 * rb.cout})) return 1;
        e->ss_off[rb.name] = e->ss_total;
        e->ss_total += 2 * rb.cout;
    }
    if (!e->ss_w) {
        if (e->alloc(&e->ss_w, static_cast<size_t>(e->ss_total) * td)) return 1;
        if (e->alloc(&e->ss_b, e->ss_total)) return 1;
        if (e->alloc(&e->ss_cur, static_cast<size_t>(e->B) * e->ss_total)) return 1;
    }
    for (const RbSpec& rb : resblocks(dim)) {
        if (rb.pos) continue;
        const int off = e->ss_off[rb.name];
        NDIFF_CUDA_OK(cudaMemcpyAsync(e->ss_w + static_cast<size_t>(off) * td, e->pf(rb.name + ".mlp.1.weight"),
                                      sizeof(float) * 2 * rb.cout * td, cudaMemcpyDeviceToDevice, s));
        NDIFF_CUDA_OK(cudaMemcpyAsync(e->ss_b + off, e->pf(rb.name + ".mlp.1.bias"), sizeof(float) * 2 * rb.cout,
                                      cudaMemcpyDeviceToDevice, s));
    }
    e->cv_total = 0;
    e->cv_off.clear();
    for (const AttnSpec& ab : attnblocks(dim)) { e->cv_off[ab.name] = e->cv_total; e->cv_total += ab.C; }
    if (!e->cvec && e->alloc(&e->cvec, static_cast<size_t>(e->B) * e->cv_total)) return 1;
    if (!e->cvec2 && e->alloc(&e->cvec2, static_cast<size_t>(e->B) * e->cv_total)) return 1;
    e->finalized = true;
    return 0;
}

int embed_params(ndiff_engine* e, cudaStream_t s) {
    EmbedBuilder b{e->dim_real, e->dim, {}};
    const int dim = e->dim_real, td = dim * 4;
    const int d[5] = {dim, dim, dim * 2, dim * 4, dim * 8};
    b.mat("init_conv.weight", b.perm(dim), b.P(dim), b.ident(4), 4, 7, 7);
    b.vec("init_conv.bias", b.perm(dim), b.P(dim));
    for (const RbSpec& rb : resblocks(dim)) {
        b.conv(rb.name + ".block1.proj", rb.cout, b.cat(rb.c0, rb.c1), b.P(rb.c0) + (rb.c1 ? b.P(rb.c1) : 0), 3);
        b.conv(rb.name + ".block2.proj", rb.cout, b.perm(rb.cout), b.P(rb.cout), 3);
        if (rb.cin != rb.cout) b.conv(rb.name + ".res_conv", rb.cout, b.cat(rb.c0, rb.c1), b.P(rb.c0) + (rb.c1 ? b.P(rb.c1) : 0), 1);
        for (const char* blk : {".block1.norm", ".block2.norm"}) {
            b.vec(rb.name + blk + ".weight", b.perm(rb.cout), b.P(rb.cout));
            b.vec(rb.name + blk + ".bias", b.perm(rb.cout), b.P(rb.cout));
        }
        // (scale, shift) head: Linear(time_dim, 2C) on the LOGICAL time vector, or conv1x1(8 -> 2C) of pos_emb
        if (rb.pos) b.mat(rb.name + ".mlp.1.weight", b.scale_shift(rb.cout), 2 * b.P(rb.cout), b.ident(8), 8, 1, 1);
        else b.mat(rb.name + ".mlp.1.weight", b.scale_shift(rb.cout), 2 * b.P(rb.cout), b.ident(td), td);
        b.vec(rb.name + ".mlp.1.bias", b.scale_shift(rb.cout), 2 * b.P(rb.cout));
    }
    for (const AttnSpec& ab : attnblocks(dim)) {
        const int C = ab.C, pC = b.P(C);
        b.mat(ab.name + ".ff.net.0.0.weight", b.ident(2 * C), 2 * pC, b.perm(C), pC);      // hidden units: first 2C of 2C'
        b.vec(ab.name + ".ff.net.0.0.bias", b.ident(2 * C), 2 * pC);
        b.mat(ab.name + ".ff.net.2.weight", b.perm(C), pC, b.ident(2 * C), 2 * pC);
        b.vec(ab.name + ".ff.net.2.bias", b.perm(C), pC);
        b.conv(ab.name + ".proj_out", C, b.perm(C), pC, 1);
        b.mat(ab.name + ".attn.to_out.0.weight", b.perm(C), pC, b.ident(128), 128);
        b.vec(ab.name + ".attn.to_out.0.bias", b.perm(C), pC);
        b.vec(ab.name + ".norm2.weight", b.perm(C), pC);
        b.vec(ab.name + ".norm2.bias", b.perm(C), pC);
    }
    for (int i = 0; i < 3; ++i) {
        b.conv("downs." + std::to_string(i) + ".3.1", d[i + 1], b.s2d(d[i]), 4 * b.P(d[i]), 1);
        b.conv("ups." + std::to_string(i) + ".3.1", d[3 - i], b.perm(d[4 - i]), b.P(d[4 - i]), 3);
    }
    b.conv("downs.3.3", d[4], b.perm(d[3]), b.P(d[3]), 3);
    b.conv("ups.3.3", d[0], b.perm(d[1]), b.P(d[1]), 3);
    b.conv("shot_mlp1.fc1", dim, b.ident(8), 8, 1);
    for (const char* n : {"shot_mlp1.fc2", "shot_mlp2.fc1", "shot_mlp2.fc2", "shot_mlp3.fc1"}) b.conv(n, dim, b.perm(dim), b.P(dim), 1);
    b.mat("shot_mlp3.fc2.weight", b.ident(4), 4, b.perm(dim), b.P(dim), 1, 1);
    b.mat("final_conv.weight", b.ident(4), 4, b.perm(dim), b.P(dim), 1, 1);

    // validate the logical shapes, allocate (once, zero-filled: the padding is never written) and scatter
    size_t n_maps = 0;
    for (const EmbedSpec& sp : b.specs) n_maps += sp.rmap.size() + sp.cmap.size();
    std::vector<int> host(n_maps);
    if (e->emb_maps_n != n_maps) {
        e->release(e->emb_maps);
        e->emb_maps = nullptr;
        if (e->alloc(&e->emb_maps, n_maps)) return 1;
        e->emb_maps_n = n_maps;
    }
    size_t at = 0;
    std::vector<std::pair<size_t, size_t>> offs;
    for (const EmbedSpec& sp : b.specs) {
        offs.push_back({at, at + sp.rmap.size()});
        std::copy(sp.rmap.begin(), sp.rmap.end(), host.begin() + at); at += sp.rmap.size();
        std::copy(sp.cmap.begin(), sp.cmap.end(), host.begin() + at); at += sp.cmap.size();
    }
    NDIFF_CUDA_OK(cudaMemcpyAsync(e->emb_maps, host.data(), n_maps * sizeof(int), cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));      // `host` is a local
    for (size_t i = 0; i < b.specs.size(); ++i) {
        const EmbedSpec& sp = b.specs[i];
        auto it = e->params.find(sp.name);
        NDIFF_REQUIRE(it != e->params.end(), "missing state_dict entry '" + sp.name + "'");
        NDIFF_REQUIRE(it->second.shape == sp.lshape, "state_dict entry '" + sp.name + "' has an unexpected shape for dim = " +
                                                         std::to_string(e->dim_real));
        Param& pp = e->phys[sp.name];
        size_t n = 1;
        for (int64_t v : sp.pshape) n *= static_cast<size_t>(v);
        if (!pp.dev) {
            if (e->alloc(&pp.dev, n)) return 1;
            NDIFF_CUDA_OK(cudaMemsetAsync(pp.dev, 0, n * sizeof(float), s));
            pp.n = n;
            pp.shape = sp.pshape;
        }
        const int R = static_cast<int>(sp.rmap.size()), Cc = static_cast<int>(sp.cmap.size());
        const int inner = sp.lshape.size() == 4 ? static_cast<int>(sp.lshape[2] * sp.lshape[3]) : 1;
        const int pC = sp.pshape.size() >= 2 ? static_cast<int>(sp.pshape[1]) : 1;
        embed_param_kernel<<<64, 256, 0, s>>>(it->second.dev, pp.dev, R, Cc, inner, e->emb_maps + offs[i].first,
                                              e->emb_maps + offs[i].second, pC);
        NDIFF_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------------------------
struct DeferredNorm {   // a ResnetBlock whose last GroupNorm-apply (+ residual) is left to the consuming kernel
    const unsigned long long* stats = nullptr;
    Act res{};
    int groups = 0;
    std::string norm;
};

struct XfSpec {      // GroupNorm-apply of a conv's INPUT (ConvGemmDesc::xf_*)
    const unsigned long long* stats = nullptr;
    const float* gamma = nullptr; const float* beta = nullptr;
    const float* ss = nullptr; int ss_ld = 0;
    int groups = 0;
};

struct Builder {
    ndiff_engine* e;
    int err = 0;
    int stats_slot = 0;
    bool direct3;
    bool fused;       // per-pixel 1x1 chains run as one kernel (pixel_chain.cuh)
    bool xf_ok;       // block1.norm folded into block2's conv (conv_gemm.cu XF kernels)

    explicit Builder(ndiff_engine* eng)
        : e(eng), direct3((eng->cfg.flags & NDIFF_FLAG_CONV_DIRECT) != 0),
          fused((eng->cfg.flags & NDIFF_FLAG_UNFUSED) == 0 && eng->dim == 64),
          xf_ok((eng->cfg.flags & (NDIFF_FLAG_UNFUSED | NDIFF_FLAG_NO_XF | NDIFF_FLAG_CONV_DIRECT)) == 0) {}

    Act make(int C, int H, int W) {
        Act a; a.C = C; a.H = H; a.W = W;
        a.p = e->pool_get(static_cast<size_t>(e->B) * H * W * C);
        if (!a.p) err = 1;
        return a;
    }
    void drop(const Act& a) { e->pool_put(a.p); }
    void name(const std::string& n, const Act& a) { e->named[n] = a; }
    unsigned long long* next_stats() { return e->stats + static_cast<size_t>(stats_slot++) * e->B * 8 * 2; }

    // generic conv / GEMM launch -> new activation
    Act conv(const std::string& wname, int mode, const Act& s0, const Act* s1, int Cout, int act, const float* vec,
             int vec_ld, const Act* res, unsigned long long* stats, int groups, const Act* out2 = nullptr,
             const std::string& res_name = "", const XfSpec* xf = nullptr) {
        const int Ho = mode == kS2D ? s0.H / 2 : s0.H, Wo = mode == kS2D ? s0.W / 2 : s0.W;
        Act out = make(Cout, Ho, Wo);
        if (err) return out;
        ConvGemmDesc d;
        d.mode = mode;
        if (mode == kHalo1 && direct3) { d.mode = kDirect; d.taps_y = 3; d.taps_x = 3; d.pad_y = 1; d.pad_x = 1; }
        // 256-pixel CTA tiles where the whole weight slice stays resident next to the larger halo stages (measured: +4 %
        // on the 64 -> 64 layers; streamed-weight shapes are faster with 128-pixel tiles and deeper rings)
        else if (mode == kHalo1 && Ho >= 32 && Cout == 64 && s0.C == 64 && !s1 && !(e->cfg.flags & NDIFF_FLAG_HALO1)) d.mode = kHalo2;
        d.B = e->B; d.H = Ho; d.W = Wo;
        d.src0 = s0.p; d.C0 = s0.C;
        if (s1) { d.src1 = s1->p; d.C1 = s1->C; }
        d.weight = e->packed.at(mode == kHalo1R ? wname + "#res" : wname);
        if (mode == kHalo1R) { d.bias2 = e->pf(res_name + ".bias"); d.out2 = out2->p; d.out2_ld = Cout; }
        d.Cout = Cout;
        d.bias = e->pf(wname + ".bias");
        d.vec = vec; d.vec_ld = vec_ld;
        if (res) { d.res = res->p; d.res_ld = res->C; }
        d.out = out.p; d.out_ld = Cout;
        d.act = act;
        d.stats = stats; d.groups = groups;
        if (xf) {
            d.xf_stats = xf->stats; d.xf_gamma = xf->gamma; d.xf_beta = xf->beta; d.xf_ss = xf->ss; d.xf_ss_ld = xf->ss_ld;
            d.xf_groups = xf->groups; d.xf_real_frac = e->real_frac;
        }
        auto plan = std::make_shared<ConvGemmPlan>();
        if (conv_gemm_plan(d, e->num_sms, plan.get())) { err = 1; return out; }
        const int taps = mode == kHalo1 ? 9 : (mode == kHalo1R ? 10 : (mode == kS2D ? 4 : 1));
        Op op;
        op.name = wname;
        op.flops = 2.0 * e->B * Ho * Wo * Cout * static_cast<double>(taps) * (s0.C + (s1 ? s1->C : 0));
        {
            const double opix = static_cast<double>(e->B) * Ho * Wo, ipix = mode == kS2D ? 4.0 * opix : opix;
            op.bytes = 2.0 * (ipix * (s0.C + (s1 ? s1->C : 0)) + opix * Cout * (1 + (res ? 1 : 0) + (mode == kHalo1R ? 1 : 0)));
        }
        op.fn = [plan](cudaStream_t st) { return conv_gemm_launch(*plan, st); };
        e->net_ops.push_back(op);
        e->conv_flops += op.flops;
        return out;
    }

    void gn(const std::string& nname, const Act& xio, unsigned long long* stats, int groups, int ss_off, const bf16* maps,
            const Act* r1, const Act* r2) {
        GnApplyArgs g{};
        g.x = xio.p; g.out = xio.p; g.stats = stats;
        g.gamma = e->pf(nname + ".weight"); g.beta = e->pf(nname + ".bias");
        if (ss_off >= 0) { g.ss = e->ss_cur; g.ss_ld = e->ss_total; g.ss_off = ss_off; }
        g.maps = maps;
        g.res1 = r1 ? r1->p : nullptr; g.res2 = r2 ? r2->p : nullptr;
        g.B = e->B; g.HW = xio.H * xio.W; g.C = xio.C; g.G = groups; g.eps = 1e-5f; g.real_frac = e->real_frac;
        Op op; op.name = nname;
        op.bytes = 2.0 * e->B * g.HW * xio.C * (2 + (maps ? 2 : 0) + (r1 ? 1 : 0) + (r2 ? 1 : 0));
        op.fn = [g](cudaStream_t st) { return gn_apply_launch(g, st); };
        e->net_ops.push_back(op);
    }

    // ResnetBlock / ResnetBlock2 (ref :146-196): returns block output; consumes nothing
    Act resblock(const std::string& n, const Act& s0, const Act* s1, int Cout, int groups, const bf16* maps,
                 const Act* extra_res, DeferredNorm* defer = nullptr) {
        const int Cin = s0.C + (s1 ? s1->C : 0);
        unsigned long long* st1 = next_stats();
        // res_conv (1x1 on the same concatenated input) rides along with block1's 3x3 conv as a tenth weight block
        // (measured: a win where the activations dominate the traffic — C_out <= 128, i.e. the 128^2 and 256^2 levels; the
        // 32^2 / 64^2 levels stream their weights and lose more to the smaller weight stages than the 1x1 pass costs)
        const bool fuse_res = Cin != Cout && Cout <= 128 && fused && !direct3;
        Act rfused;
        if (fuse_res) rfused = make(Cout, s0.H, s0.W);
        Act h = fuse_res ? conv(n + ".block1.proj", kHalo1R, s0, s1, Cout, kActNone, nullptr, 0, nullptr, st1, groups, &rfused,
                                n + ".res_conv")
                         : conv(n + ".block1.proj", kHalo1, s0, s1, Cout, kActNone, nullptr, 0, nullptr, st1, groups);
        // block1.norm (GroupNorm + time scale/shift + SiLU) has ONE consumer, block2's conv: it is evaluated inside that conv,
        // on the activation tiles in shared memory, instead of as a read-modify-write pass over HBM.  (ResnetBlock2's per-pixel
        // scale/shift maps stay a separate pass.)
        // (measured: a win at the 128^2 and 256^2 levels; below that the tensors are L2-sized, the separate pass costs 15-50 us
        // and the in-kernel transform, which competes with the MMA for shared-memory bandwidth, costs more)
        const bool fuse_norm1 = xf_ok && !maps && s0.H * s0.W >= (e->H * e->W) / 4;
        XfSpec xs;
        if (fuse_norm1) {
            xs.stats = st1; xs.gamma = e->pf(n + ".block1.norm.weight"); xs.beta = e->pf(n + ".block1.norm.bias");
            xs.ss = e->ss_cur + e->ss_off.at(n); xs.ss_ld = e->ss_total; xs.groups = groups;
        } else {
            gn(n + ".block1.norm", h, st1, groups, maps ? -1 : e->ss_off.at(n), maps, nullptr, nullptr);
        }
        unsigned long long* st2 = next_stats();
        Act h2 = conv(n + ".block2.proj", kHalo1, h, nullptr, Cout, kActNone, nullptr, 0, nullptr, st2, groups, nullptr, "",
                      fuse_norm1 ? &xs : nullptr);
        drop(h);
        if (defer) {
            // the consumer (heads kernel / shot-branch tail) applies block2.norm + the residual itself; the residual stays live
            Act r = s0;
            if (fuse_res) r = rfused;
            else if (Cin != Cout) r = conv(n + ".res_conv", kDirect, s0, s1, Cout, kActNone, nullptr, 0, nullptr, nullptr, 0);
            defer->stats = st2; defer->res = r; defer->groups = groups; defer->norm = n + ".block2.norm";
            return h2;
        }
        if (Cin != Cout) {
            Act r = fuse_res ? rfused : conv(n + ".res_conv", kDirect, s0, s1, Cout, kActNone, nullptr, 0, nullptr, nullptr, 0);
            gn(n + ".block2.norm", h2, st2, groups, -1, nullptr, &r, extra_res);
            drop(r);
        } else {
            gn(n + ".block2.norm", h2, st2, groups, -1, nullptr, &s0, extra_res);
        }
        name(n, h2);
        return h2;
    }

    // AttnBlock with the collapsed 1-token cross attention (ref :425-443)
    Act attn(const std::string& n, const Act& xin) {
        const int C = xin.C;
        const float* cv = e->cvec + e->cv_off.at(n);
        if (fused && C == 64) {
            Act o = make(C, xin.H, xin.W);
            if (err) return o;
            ChainDesc d;
            d.prog = kProgAttn;
            d.npix = e->B * xin.H * xin.W; d.HW = xin.H * xin.W;
            d.x = xin.p; d.weights = e->chain_w.at(n); d.fvec = e->chain_f.at(n);
            d.cvec = cv; d.cvec_ld = e->cv_total; d.real_frac = e->real_frac;
            d.cvec2 = e->cvec2 + e->cv_off.at(n);
            d.out = o.p;
            auto plan = std::make_shared<ChainPlan>();
            if (pixel_chain_plan(d, e->num_sms, plan.get())) { err = 1; return o; }
            Op op; op.name = n + "(fused chain)";
            op.flops = 2.0 * d.npix * (64.0 * 128 + 128.0 * 64 + 64.0 * 64);
            op.bytes = 2.0 * d.npix * 64 * 2;
            op.chain = true;
            op.fn = [plan](cudaStream_t st) { return pixel_chain_launch(*plan, st); };
            e->net_ops.push_back(op);
            name(n, o);
            return o;
        }
        Act u = make(C, xin.H, xin.W);
        if (err) return u;
        {
            const bf16* xp = xin.p; bf16* up = u.p;
            const float* g = e->pf(n + ".norm2.weight"); const float* bt = e->pf(n + ".norm2.bias");
            const int B = e->B, HW = xin.H * xin.W, ld = e->cv_total;
            Op op; op.name = n + ".norm2";
            op.bytes = 2.0 * B * HW * C * 2;
            const float rf = e->real_frac;
            op.fn = [=](cudaStream_t st) { return layernorm_launch(xp, cv, ld, g, bt, up, B, HW, C, st, rf); };
            e->net_ops.push_back(op);
        }
        Act hh = conv(n + ".ff.net.0.0", kDirect, u, nullptr, 2 * C, kActGelu, nullptr, 0, nullptr, nullptr, 0);
        drop(u);
        if (fused) {
            // ff.net.2 and proj_out folded into one two-source GEMM over [h | x] (pack_attn_ff2proj_kernel): z never exists
            Act o = make(C, xin.H, xin.W);
            if (err) return o;
            ConvGemmDesc d;
            d.mode = kDirect; d.B = e->B; d.H = xin.H; d.W = xin.W;
            d.src0 = hh.p; d.C0 = 2 * C; d.src1 = xin.p; d.C1 = C;
            d.weight = e->packed.at(n + "#ff2proj"); d.Cout = C;
            d.vec = e->cvec2 + e->cv_off.at(n); d.vec_ld = e->cv_total;
            d.res = xin.p; d.res_ld = C;
            d.out = o.p; d.out_ld = C;
            auto plan = std::make_shared<ConvGemmPlan>();
            if (conv_gemm_plan(d, e->num_sms, plan.get())) { err = 1; return o; }
            Op op; op.name = n + ".ff.net.2+proj_out";
            op.flops = 2.0 * e->B * xin.H * xin.W * C * 3.0 * C;
            op.bytes = 2.0 * e->B * xin.H * xin.W * (2.0 * C + C + C);
            op.fn = [plan](cudaStream_t st) { return conv_gemm_launch(*plan, st); };
            e->net_ops.push_back(op);
            e->conv_flops += op.flops;
            drop(hh);
            name(n, o);
            return o;
        }
        Act z = conv(n + ".ff.net.2", kDirect, hh, nullptr, C, kActNone, cv, e->cv_total, &xin, nullptr, 0);
        drop(hh);
        Act o = conv(n + ".proj_out", kDirect, z, nullptr, C, kActNone, nullptr, 0, &xin, nullptr, 0);
        drop(z);
        name(n, o);
        return o;
    }

    Act gemm(const std::string& w, const Act& s, int Cout, int act) {
        return conv(w, kDirect, s, nullptr, Cout, act, nullptr, 0, nullptr, nullptr, 0);
    }
};

int build_plan(ndiff_engine* e) {
    const int dim = e->dim, B = e->B, H = e->H, W = e->W;
    const int d[5] = {dim, dim, dim * 2, dim * 4, dim * 8};
    e->net_ops.clear();
    e->named.clear();
    e->conv_flops = 0.0;
    Builder b(e);
    const int npix = B * H * W;

    // ---- shot-noise branch (ref :598-604)
    Act s1, s4;
    if (b.fused) {
        s1 = b.make(dim, H, W);
        s4 = b.make(dim, H, W);
        if (b.err) return 1;
        ChainDesc d;
        d.prog = kProgShot;
        d.npix = npix; d.HW = H * W;
        d.weights = e->chain_w.at("shot"); d.fvec = e->chain_f.at("shot");
        d.cvec = e->cvec + e->cv_off.at("shot_attn"); d.cvec_ld = e->cv_total; d.real_frac = e->real_frac;
        d.cvec2 = e->cvec2 + e->cv_off.at("shot_attn");
        d.clean = e->clean; d.xt = e->x;
        d.out = s4.p; d.out2 = s1.p;
        auto plan = std::make_shared<ChainPlan>();
        if (pixel_chain_plan(d, e->num_sms, plan.get())) return 1;
        Op op; op.name = "shot_mlp1+shot_attn+shot_mlp2(fused chain)";
        op.flops = 2.0 * npix * (8.0 * 64 + 64.0 * 64 * 4 + 64.0 * 128 * 2);
        op.bytes = static_cast<double>(npix) * (2 * 16 + 2 * 64 * 2);      // clean + x_t (fp32 x 4 each) in, s1 and s4 out
        op.chain = true;
        op.fn = [plan](cudaStream_t st) { return pixel_chain_launch(*plan, st); };
        e->net_ops.push_back(op);
        b.name("shot_mlp1", s1);
        b.name("shot_mlp2", s4);
    } else {
        Act s0 = b.make(dim, H, W);
        if (b.err) return 1;
        {
            Op op; op.name = "shot_mlp1.fc1";
            const float* cl = e->clean; const float* x = e->x; bf16* o = s0.p;
            const float* w = e->pf("shot_mlp1.fc1.weight"); const float* bs = e->pf("shot_mlp1.fc1.bias");
            op.fn = [=](cudaStream_t st) { return shot_in_launch(cl, x, w, bs, o, npix, dim, st); };
            e->net_ops.push_back(op);
        }
        s1 = b.gemm("shot_mlp1.fc2", s0, dim, kActNone);
        b.drop(s0);
        b.name("shot_mlp1", s1);
        Act s2 = b.attn("shot_attn", s1);
        Act s3 = b.gemm("shot_mlp2.fc1", s2, dim, kActGelu);
        b.drop(s2);
        s4 = b.gemm("shot_mlp2.fc2", s3, dim, kActNone);
        b.drop(s3);
        b.name("shot_mlp2", s4);
    }
    // fused tail: block2.norm + residuals + shot_mlp3 (fc1, GELU, fc2) in one per-pixel kernel that leaves the 4-channel
    // shot-noise image (debug runs that keep every activation use the layer-by-layer form)
    e->tail_fused = b.fused && !e->keep_all;
    if (e->tail_fused) {
        DeferredNorm dn;
        Act h2 = b.resblock("shot_time", s4, nullptr, dim, 2, nullptr, &s1, &dn);
        if (b.err) return 1;
        if (!e->sn && e->alloc(&e->sn, static_cast<size_t>(npix) * 4)) return 1;
        TailDesc td;
        td.npix = npix; td.HW = H * W;
        td.h2 = h2.p; td.r1 = dn.res.p; td.r2 = s1.p;
        td.weights = e->chain_w.at("tail"); td.fvec = e->chain_f.at("tail");
        td.stats = dn.stats; td.gamma = e->pf(dn.norm + ".weight"); td.beta = e->pf(dn.norm + ".bias"); td.groups = dn.groups; td.real_frac = e->real_frac;
        td.out = e->sn;
        auto plan = std::make_shared<TailPlan>();
        if (tail_chain_plan(td, e->num_sms, plan.get())) return 1;
        Op op; op.name = "shot_time.block2.norm+shot_mlp3(fused chain)";
        op.flops = 2.0 * npix * (64.0 * 64 + 64.0 * 4);
        op.bytes = static_cast<double>(npix) * (3 * 64 * 2 + 16);
        op.chain = true;
        op.fn = [plan](cudaStream_t st) { return tail_chain_launch(*plan, st); };
        e->net_ops.push_back(op);
        b.drop(h2); b.drop(s4); b.drop(s1);
        e->sf = Act{};
    } else {
        Act s5 = b.resblock("shot_time", s4, nullptr, dim, 2, nullptr, &s1);   // + r (ref :603) folded into the apply pass
        b.drop(s4); b.drop(s1);
        Act s6 = b.gemm("shot_mlp3.fc1", s5, dim, kActGelu);
        b.drop(s5);
        e->sf = s6;
    }

    // ---- main U-Net (ref :606-643)
    Act x0 = b.make(dim, H, W);
    if (b.err) return 1;
    {
        // tensor-core path: pack x_t to the padded bf16 layout, then a 7-tap "direct" GEMM whose A rows are overlapping
        // 128-byte windows (tensor-map stride 16 B < row length 128 B); CUDA-core kernel only if the driver rejects the map
        ConvGemmDesc d;
        d.mode = kDirect; d.B = B; d.H = H; d.W = W;
        d.src0 = e->xpad; d.C0 = 64;
        d.taps_y = 4; d.taps_x = 1; d.pad_y = 0; d.pad_x = 0; d.tap_sy = 2;      // four row-pair taps: input rows y, y+2, y+4, y+6
        d.custom_src0 = true;
        d.toeplitz = W % 128 == 0 && (e->cfg.flags & NDIFF_FLAG_INIT_WINDOWS) == 0;      // one landed row per tap instead of 128-byte windows
        d.cdim[0] = d.toeplitz ? 8 : 64; d.cdim[1] = static_cast<uint64_t>(d.toeplitz ? W + 8 : W);
        d.cdim[2] = static_cast<uint64_t>(H + 6); d.cdim[3] = static_cast<uint64_t>(B);
        d.cstride[0] = 16; d.cstride[1] = static_cast<uint64_t>(W + 8) * 16; d.cstride[2] = static_cast<uint64_t>(H + 6) * (W + 8) * 16;
        d.weight = e->init_w_tc; d.Cout = dim; d.bias = e->pf("init_conv.bias");
        d.out = x0.p; d.out_ld = dim;
        auto plan = std::make_shared<ConvGemmPlan>();
        bool tc_ok = (e->cfg.flags & NDIFF_FLAG_INIT_SIMT) == 0 && conv_gemm_plan(d, e->num_sms, plan.get()) == 0;
        if (!tc_ok && d.toeplitz && (e->cfg.flags & NDIFF_FLAG_INIT_SIMT) == 0) {      // e.g. too few tiles for resident weights
            d.toeplitz = false;
            d.cdim[0] = 64; d.cdim[1] = static_cast<uint64_t>(W);
            tc_ok = conv_gemm_plan(d, e->num_sms, plan.get()) == 0;
        }
        if (tc_ok) {
            Op pk; pk.name = "init_conv.pack";
            pk.bytes = static_cast<double>(npix) * (16 + 16);
            const float4* xs = reinterpret_cast<const float4*>(e->x); uint2* xp = reinterpret_cast<uint2*>(e->xpad);
            pk.fn = [=](cudaStream_t st) {
                NDIFF_CUDA_OK(launch_pdl(xpad_pack_kernel, dim3((npix + 255) / 256), dim3(256), 0, st, xs, xp, H, W,
                                         static_cast<size_t>(npix)));
                return 0;
            };
            e->net_ops.push_back(pk);
            Op op; op.name = d.toeplitz ? "init_conv" : "init_conv(windows)";
            op.flops = 2.0 * npix * dim * 196.0;
            op.bytes = static_cast<double>(npix) * (16 + 2.0 * dim);
            op.fn = [plan](cudaStream_t st) { return conv_gemm_launch(*plan, st); };
            e->net_ops.push_back(op);
            e->conv_flops += op.flops;
        } else {
            Op op; op.name = "init_conv(simt)";
            const float* x = e->x; const float* w = e->init_w; const float* bs = e->pf("init_conv.bias"); bf16* o = x0.p;
            op.fn = [=](cudaStream_t st) { return init_conv7_launch(x, w, bs, o, B, H, W, dim, st); };
            e->net_ops.push_back(op);
        }
    }
    b.name("init_conv", x0);
    Act cur = b.resblock("pos_block1", x0, nullptr, dim, 2, e->map1, nullptr);
    std::vector<Act> skips;
    for (int i = 0; i < 4; ++i) {
        const std::string p = "downs." + std::to_string(i);
        Act a1 = b.resblock(p + ".0", cur, nullptr, d[i], 8, nullptr, nullptr);
        if (i > 0 || true) b.drop(cur);
        skips.push_back(a1);
        Act a2 = b.resblock(p + ".1", a1, nullptr, d[i], 8, nullptr, nullptr);
        skips.push_back(a2);
        Act a3 = b.attn(p + ".2", a2);
        if (i < 3) {
            cur = b.conv(p + ".3.1", kS2D, a3, nullptr, d[i + 1], kActNone, nullptr, 0, nullptr, nullptr, 0);
        } else {
            cur = b.conv(p + ".3", kHalo1, a3, nullptr, d[i + 1], kActNone, nullptr, 0, nullptr, nullptr, 0);
        }
        b.drop(a3);
        b.name(p + ".3", cur);
    }
    {
        Act m1 = b.resblock("mid_block1", cur, nullptr, d[4], 8, nullptr, nullptr);
        b.drop(cur);
        Act m2 = b.resblock("mid_block2", m1, nullptr, d[4], 8, nullptr, nullptr);
        b.drop(m1);
        cur = m2;
    }
    for (int i = 0; i < 4; ++i) {
        const std::string p = "ups." + std::to_string(i);
        const int co = d[4 - i], ci = d[3 - i];
        Act sk = skips.back(); skips.pop_back();
        Act a1 = b.resblock(p + ".0", cur, &sk, co, 8, nullptr, nullptr);
        b.drop(cur); b.drop(sk);
        sk = skips.back(); skips.pop_back();
        Act a2 = b.resblock(p + ".1", a1, &sk, co, 8, nullptr, nullptr);
        b.drop(a1); b.drop(sk);
        Act a3 = b.attn(p + ".2", a2);
        b.drop(a2);
        if (i < 3 && b.fused) {
            // Upsample(nearest x2) + conv3x3 in one kernel: four 2x2 phase convolutions on the low-resolution tensor
            Act out = b.make(ci, a3.H * 2, a3.W * 2);
            if (b.err) return 1;
            ConvGemmDesc d;
            d.mode = kHaloUp; d.B = B; d.H = a3.H; d.W = a3.W;
            d.src0 = a3.p; d.C0 = a3.C;
            d.weight = e->packed.at(p + ".3.1#up"); d.Cout = ci; d.bias = e->pf(p + ".3.1.bias");
            d.out = out.p; d.out_ld = ci;
            auto plan = std::make_shared<ConvGemmPlan>();
            if (conv_gemm_plan(d, e->num_sms, plan.get())) return 1;
            Op op; op.name = p + ".3.1";
            op.flops = 2.0 * B * (4.0 * a3.H * a3.W) * ci * 4.0 * a3.C;      // executed: 4 taps per output pixel
            op.bytes = 2.0 * B * a3.H * a3.W * (a3.C + 4.0 * ci);
            op.fn = [plan](cudaStream_t st) { return conv_gemm_launch(*plan, st); };
            e->net_ops.push_back(op);
            e->conv_flops += op.flops;
            b.drop(a3);
            cur = out;
        } else if (i < 3) {
            Act up = b.make(co, a3.H * 2, a3.W * 2);
            if (b.err) return 1;
            {
                Op op; op.name = p + ".3.0(nearest x2)";
                const bf16* in = a3.p; bf16* o = up.p; const int hh = a3.H, ww = a3.W;
                op.fn = [=](cudaStream_t st) { return upsample2x_launch(in, o, B, hh, ww, co, st); };
                e->net_ops.push_back(op);
            }
            b.drop(a3);
            cur = b.conv(p + ".3.1", kHalo1, up, nullptr, ci, kActNone, nullptr, 0, nullptr, nullptr, 0);
            b.drop(up);
        } else {
            cur = b.conv(p + ".3", kHalo1, a3, nullptr, ci, kActNone, nullptr, 0, nullptr, nullptr, 0);
            b.drop(a3);
        }
        b.name(p + ".3", cur);
    }
    Act pb2 = b.resblock("pos_block2", cur, nullptr, dim, 2, e->map2, nullptr);
    b.drop(cur);
    // fused tail: the heads kernel applies final_res_block.block2.norm on the fly (debug runs that keep every activation
    // materialise it instead, so the layer-by-layer parity test still sees the block output)
    const bool fuse_tail = b.fused && !e->keep_all;
    e->xf_stats = nullptr;
    DeferredNorm fdn;
    Act fr = b.resblock("final_res_block", pb2, &x0, dim, 8, nullptr, nullptr, fuse_tail ? &fdn : nullptr);
    if (fuse_tail) { e->xf_stats = fdn.stats; e->xf_res = fdn.res; e->xf_groups = fdn.groups; e->xf_norm = fdn.norm; }
    b.drop(pb2); b.drop(x0);
    e->xf = fr;
    if (b.err) return 1;
    NDIFF_REQUIRE(b.stats_slot <= e->n_stats, "GroupNorm statistics arena too small");
    e->plan_built = true;
    return 0;
}

int run_net(ndiff_engine* e, cudaStream_t s, bool zero_stats = true) {
    g_use_pdl = (e->cfg.flags & NDIFF_FLAG_PDL) != 0;      // per engine: every launch of this engine's plan goes through here
    // GroupNorm sums are accumulated with atomics: clear the arena first (a kernel, not a memset node, so that the first
    // layer can hang off it with a programmatic edge)
    if (zero_stats) {
        zero_u64_kernel<<<32, 256, 0, s>>>(e->stats, static_cast<int>(e->stats_bytes / sizeof(unsigned long long)));
        NDIFF_CUDA_OK(cudaGetLastError());
    }
    for (Op& op : e->net_ops)
        if (op.fn(s)) return 1;
    return 0;
}

int final_args(ndiff_engine* e, bool chain, FinalArgs* f) {
    memset(f, 0, sizeof(*f));
    f->xf = e->xf.p; f->sf = e->sf.p;
    if (e->tail_fused) { f->sf = nullptr; f->sn = reinterpret_cast<const float4*>(e->sn); }
    f->wf = e->pf("final_conv.weight"); f->bfin = e->pf("final_conv.bias");
    f->ws = e->pf("shot_mlp3.fc2.weight"); f->bs = e->pf("shot_mlp3.fc2.bias");
    f->npix = e->B * e->H * e->W; f->C = e->dim; f->HW = e->H * e->W;
    if (chain) { f->chain = e->chain; f->x = e->x; } else { f->v_out = e->v_out; }
    if (e->xf_stats) {
        f->gn_stats = e->xf_stats; f->gn_res = e->xf_res.p; f->gn_G = e->xf_groups; f->gn_eps = 1e-5f; f->gn_real_frac = e->real_frac;
        f->gn_gamma = e->pf(e->xf_norm + ".weight"); f->gn_beta = e->pf(e->xf_norm + ".bias");
    }
    return 0;
}

int run_step(ndiff_engine* e, cudaStream_t s) {
    chain_step_begin_kernel<<<32, 256, 0, s>>>(e->chain, e->step_table, e->ss_table, e->ss_total, e->ss_cur, e->B,
                                               e->H * e->W, reinterpret_cast<float4*>(e->x), e->stats,
                                               static_cast<int>(e->stats_bytes / sizeof(unsigned long long)));
    NDIFF_CUDA_OK(cudaGetLastError());
    if (run_net(e, s, false)) return 1;
    FinalArgs f;
    final_args(e, true, &f);
    return final_launch(f, s);
}

int capture(ndiff_engine* e, bool step, cudaGraphExec_t* exec, cudaStream_t user) {
    // the caller's stream may be the legacy default stream, which cannot be captured: record on our own stream
    NDIFF_CUDA_OK(cudaStreamSynchronize(user));
    cudaStream_t s = e->cap_stream;
    cudaGraph_t graph = nullptr;
    NDIFF_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc;
    if (step) rc = run_step(e, s);
    else {
        rc = run_net(e, s);
        if (!rc) { FinalArgs f; final_args(e, false, &f); rc = final_launch(f, s); }
    }
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
    NDIFF_CUDA_OK(ce);
    NDIFF_CUDA_OK(cudaGraphInstantiate(exec, graph, 0));
    cudaGraphDestroy(graph);
    return 0;
}

}  // namespace

// ================================================================================================================
// C ABI
// ================================================================================================================
namespace {
int ensure_time_bufs(ndiff_engine* e, int n) {
    if (n > e->t_buf_n) {
        e->release(e->t_buf); e->release(e->st_buf);
        e->t_buf = nullptr; e->st_buf = nullptr;
        if (e->alloc(&e->t_buf, n)) return 1;
        if (e->alloc(&e->st_buf, static_cast<size_t>(n) * e->dim_real * 4)) return 1;
        e->t_buf_n = n;
    }
    return 0;
}
}  // namespace

namespace ndiff {
int xpad_pack_launch(const float* x_nhwc4, bf16* xpad, int H, int W, size_t npix, cudaStream_t s) {
    NDIFF_CUDA_OK(launch_pdl(xpad_pack_kernel, dim3(static_cast<unsigned>((npix + 255) / 256)), dim3(256), 0, s,
                             reinterpret_cast<const float4*>(x_nhwc4), reinterpret_cast<uint2*>(xpad), H, W, npix));
    return 0;
}
int engine_finalize(ndiff_engine* e, cudaStream_t s) {
    if (finalize(e, s)) return 1;
    if (!e->plan_built && !e->skip_plan && build_plan(e)) return 1;
    return 0;
}
}  // namespace ndiff

extern "C" {

int32_t ndiff_abi_version(void) { return NDIFF_ABI_VERSION; }
const char* ndiff_last_error(void) { return get_error(); }

int32_t ndiff_engine_create(const ndiff_config* cfg, ndiff_engine** out) {
    NDIFF_REQUIRE(cfg && out, "null argument");
    NDIFF_REQUIRE(cfg->dim >= 8 && cfg->dim <= 64 && cfg->dim % 8 == 0,
                  "dim must be a multiple of 8 up to 64 (64 runs natively; smaller widths such as the shipped checkpoint's 48 are "
                  "embedded in the 64-channel kernels with zero padding)");
    NDIFF_REQUIRE(cfg->batch >= 1 && cfg->batch <= 256, "batch must be in [1, 256]");
    NDIFF_REQUIRE(cfg->height % 8 == 0 && cfg->width % 8 == 0 && cfg->height >= 8 && cfg->width >= 8,
                  "height/width must be multiples of 8 (Diffusion_arch.py:578)");
    int ndev = 0;
    NDIFF_CUDA_OK(cudaGetDeviceCount(&ndev));
    NDIFF_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "no such CUDA device");
    DeviceGuard dev_guard(cfg->device);
    cudaDeviceProp prop;
    NDIFF_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    NDIFF_REQUIRE(prop.major == 10, "noisediff_b200 needs an sm_100a GPU (B200); found sm_" + std::to_string(prop.major) +
                                        std::to_string(prop.minor));
    if (conv_gemm_init() || pointwise_init() || pixel_chain_init()) return 1;
    std::unique_ptr<ndiff_engine> e(new ndiff_engine());
    NDIFF_CUDA_OK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    e->cfg = *cfg;
    e->num_sms = prop.multiProcessorCount;
    e->B = cfg->batch; e->H = cfg->height; e->W = cfg->width;
    e->dim_real = cfg->dim; e->dim = 64; e->real_frac = static_cast<float>(cfg->dim) / 64.0f;
    e->keep_all = (cfg->flags & NDIFF_FLAG_KEEP_ACTS) != 0;
    const size_t npix = static_cast<size_t>(e->B) * e->H * e->W;
    if (e->alloc(&e->clean, npix * 4) || e->alloc(&e->x, npix * 4) || e->alloc(&e->v_out, npix * 4)) return 1;
    if (e->alloc(&e->map1, npix * 2 * e->dim) || e->alloc(&e->map2, npix * 2 * e->dim)) return 1;
    if (e->alloc(&e->pos_emb, npix * 8)) return 1;
    if (e->alloc(&e->position, npix * 2) || e->alloc(&e->iso_idx, static_cast<size_t>(e->B))) return 1;
    {
        const size_t n = static_cast<size_t>(e->B) * (e->H + 6) * (e->W + 8) * 8;
        if (e->alloc(&e->xpad, n)) return 1;
        NDIFF_CUDA_OK(cudaMemset(e->xpad, 0, n * sizeof(bf16)));
    }
    e->n_stats = 64;
    e->stats_bytes = static_cast<size_t>(e->n_stats) * e->B * 8 * 2 * sizeof(unsigned long long);
    if (e->alloc(&e->stats, e->stats_bytes / sizeof(unsigned long long))) return 1;
    if (e->alloc(&e->chain, 1)) return 1;
    NDIFF_CUDA_OK(cudaMemset(e->chain, 0, sizeof(ChainState)));
    *out = e.release();
    return 0;
}

void ndiff_engine_destroy(ndiff_engine* e) {
    if (!e) return;
    DeviceGuard dev_guard(e->cfg.device);
    cudaDeviceSynchronize();
    delete e;
}

int32_t ndiff_load_param_async(ndiff_engine* e, const char* name, const float* data, int32_t ndim, const int64_t* shape,
                               void* stream) {
    NDIFF_REQUIRE(e && name && data && ndim >= 0 && ndim <= 4, "bad argument");
    DeviceGuard dev_guard(e->cfg.device);
    Param& p = e->params[name];
    size_t n = 1;
    std::vector<int64_t> shp(shape, shape + ndim);
    for (int64_t v : shp) n *= static_cast<size_t>(v);
    if (p.dev == nullptr || p.n != n) {
        NDIFF_REQUIRE(!e->plan_built, std::string("parameter '") + name + "' changed size after the layer plan was built");
        e->release(p.dev);
        p.dev = nullptr;
        if (e->alloc(&p.dev, n)) return 1;
        p.n = n;
    }
    p.shape = shp;
    // ordered on the caller's stream: behind whatever produced `data` there (an optimizer step), ahead of the repack that
    // ndiff_finalize_params enqueues on the same stream, and behind earlier graph replays on it
    NDIFF_CUDA_OK(cudaMemcpyAsync(p.dev, data, n * sizeof(float), cudaMemcpyDefault, as_stream(stream)));
    e->finalized = false;
    e->cond_set = false;      // map1 / map2 / cvec were derived from the old weights: the condition must be set again
    return 0;
}

int32_t ndiff_load_param(ndiff_engine* e, const char* name, const float* data, int32_t ndim, const int64_t* shape) {
    // stream-less form: fence against everything in flight on the device, copy, and return with the copy complete
    NDIFF_REQUIRE(e, "null engine");
    DeviceGuard dev_guard(e->cfg.device);
    NDIFF_CUDA_OK(cudaDeviceSynchronize());
    if (ndiff_load_param_async(e, name, data, ndim, shape, nullptr)) return 1;
    NDIFF_CUDA_OK(cudaStreamSynchronize(nullptr));
    return 0;
}

int32_t ndiff_finalize_params(ndiff_engine* e, void* stream) {
    NDIFF_REQUIRE(e, "null engine");
    DeviceGuard dev_guard(e->cfg.device);
    // Packed buffers and fp32 parameter storage keep their addresses across reloads, so the layer plan and the
    // captured graphs stay valid; only the first call builds them.
    if (engine_finalize(e, as_stream(stream))) return 1;
    NDIFF_CUDA_OK(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

int32_t ndiff_set_condition(ndiff_engine* e, const float* clean_dev, const float* position_dev,
                            const int64_t* iso_idx_dev, void* stream) {
    NDIFF_REQUIRE(e && e->finalized, "engine has no finalized weights");
    NDIFF_REQUIRE(clean_dev && position_dev && iso_idx_dev, "null condition tensor");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    const int HW = e->H * e->W;
    if (nchw_to_nhwc4_launch(clean_dev, e->clean, e->B, HW, s)) return 1;
    NDIFF_CUDA_OK(cudaMemcpyAsync(e->position, position_dev, sizeof(float) * e->B * 2 * HW, cudaMemcpyDeviceToDevice, s));
    NDIFF_CUDA_OK(cudaMemcpyAsync(e->iso_idx, iso_idx_dev, sizeof(long long) * e->B, cudaMemcpyDeviceToDevice, s));
    PosArgs pa{};
    pa.position = position_dev;
    pa.we = e->pf("pos_enc.weights.weight"); pa.be = e->pf("pos_enc.weights.bias");
    pa.w1 = e->pf("pos_mlp.fc1.weight"); pa.b1 = e->pf("pos_mlp.fc1.bias");
    pa.w2 = e->pf("pos_mlp.fc2.weight"); pa.b2 = e->pf("pos_mlp.fc2.bias");
    pa.wm1 = e->pf("pos_block1.mlp.1.weight"); pa.bm1 = e->pf("pos_block1.mlp.1.bias");
    pa.wm2 = e->pf("pos_block2.mlp.1.weight"); pa.bm2 = e->pf("pos_block2.mlp.1.bias");
    pa.map1 = e->map1; pa.map2 = e->map2; pa.pos_emb = e->pos_emb;
    pa.B = e->B; pa.HW = HW; pa.C = e->dim;
    if (pos_maps_launch(pa, s)) return 1;
    for (const AttnSpec& ab : attnblocks(e->dim)) {
        if (iso_vec_launch(e->pf("iso_embed.weight"), reinterpret_cast<const long long*>(iso_idx_dev),
                           e->pf(ab.name + ".attn.to_v.weight"), e->pf(ab.name + ".attn.to_out.0.weight"),
                           e->pf(ab.name + ".attn.to_out.0.bias"), e->cvec, e->cv_total, e->cv_off.at(ab.name), e->B,
                           ab.C, s))
            return 1;
        // every AttnBlock runs its last linear layers as one folded stage with a per-sample vector: Wp (b2 + c) + bp, and for
        // the shot branch (which continues into shot_mlp2.fc1) Wm1 Wp (b2 + c) + Wm1 bp + bm1 (packer: shot_fold, f[512..])
        const bool shot = ab.name == "shot_attn";
        if (shot && !e->chain_f.count("shot")) continue;      // the unfused plan (dim != 64 layouts) has no folded shot stage
        attn_vec2_kernel<<<e->B, 256, 0, s>>>(shot ? e->shot_fold : e->pf(ab.name + ".proj_out.weight"),
                                              shot ? e->chain_f.at("shot") + 512 : e->pf(ab.name + ".proj_out.bias"),
                                              e->pf(ab.name + ".ff.net.2.bias"), e->cvec, e->cvec2, e->cv_total, e->cv_off.at(ab.name), ab.C);
        NDIFF_CUDA_OK(cudaGetLastError());
    }
    e->cond_set = true;
    return 0;
}

int32_t ndiff_forward(ndiff_engine* e, const float* x_dev, const int64_t* time_dev, float* out_dev, void* stream) {
    NDIFF_REQUIRE(e && e->finalized && e->cond_set, "engine needs weights and a condition before forward");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    const int HW = e->H * e->W;
    if (ensure_time_bufs(e, e->B)) return 1;
    if (nchw_to_nhwc4_launch(x_dev, e->x, e->B, HW, s)) return 1;
    i64_to_i32_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const long long*>(time_dev), e->t_buf, e->B);
    NDIFF_CUDA_OK(cudaGetLastError());
    if (time_mlp_launch(e->t_buf, 1, e->B, e->dim_real, e->pf("time_mlp.1.weight"), e->pf("time_mlp.1.bias"),
                        e->pf("time_mlp.3.weight"), e->pf("time_mlp.3.bias"), e->st_buf, s)) return 1;
    if (rows_gemv_launch(e->ss_w, e->ss_b, e->st_buf, e->ss_cur, e->B, e->ss_total, e->dim_real * 4, s)) return 1;
    if (e->cfg.flags & NDIFF_FLAG_NO_GRAPH) {
        if (run_net(e, s)) return 1;
        FinalArgs f; final_args(e, false, &f);
        if (final_launch(f, s)) return 1;
    } else {
        if (!e->fwd_exec && capture(e, false, &e->fwd_exec, s)) return 1;
        NDIFF_CUDA_OK(cudaGraphLaunch(e->fwd_exec, s));
    }
    return nhwc4_to_nchw_launch(e->v_out, out_dev, e->B, HW, s);
}

int32_t ndiff_chain_begin(ndiff_engine* e, const ndiff_step* steps_host, int32_t n_steps, const float* x_init_dev,
                          uint64_t seed, void* stream) {
    NDIFF_REQUIRE(e && e->finalized && e->cond_set, "engine needs weights and a condition before sampling");
    NDIFF_REQUIRE(steps_host && n_steps > 0, "empty step table");
    static_assert(sizeof(ndiff_step) == sizeof(StepParams), "ABI step layout");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    if (n_steps > e->ss_table_rows) {
        // The captured step graph holds the two table pointers by value (chain_step_begin_kernel's arguments): a longer chain
        // after a shorter one on the same engine must re-capture, or every replay would index the old, shorter tables.
        if (e->step_exec) {
            NDIFF_CUDA_OK(cudaDeviceSynchronize());
            NDIFF_CUDA_OK(cudaGraphExecDestroy(e->step_exec));
            e->step_exec = nullptr;
        }
        e->release(e->step_table); e->release(e->ss_table);
        e->step_table = nullptr; e->ss_table = nullptr;
        if (e->alloc(&e->step_table, n_steps)) return 1;
        if (e->alloc(&e->ss_table, static_cast<size_t>(n_steps) * e->ss_total)) return 1;
        e->ss_table_rows = n_steps;
    }
    if (ensure_time_bufs(e, n_steps)) return 1;
    std::vector<int> ts(n_steps);
    for (int i = 0; i < n_steps; ++i) ts[i] = steps_host[i].t;
    NDIFF_CUDA_OK(cudaMemcpyAsync(e->step_table, steps_host, sizeof(StepParams) * n_steps, cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaMemcpyAsync(e->t_buf, ts.data(), sizeof(int) * n_steps, cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));   // ts / steps_host may be pageable stack memory
    if (time_mlp_launch(e->t_buf, 1, n_steps, e->dim_real, e->pf("time_mlp.1.weight"), e->pf("time_mlp.1.bias"),
                        e->pf("time_mlp.3.weight"), e->pf("time_mlp.3.bias"), e->st_buf, s)) return 1;
    if (rows_gemv_launch(e->ss_w, e->ss_b, e->st_buf, e->ss_table, n_steps, e->ss_total, e->dim_real * 4, s)) return 1;
    ChainState cs;
    memset(&cs, 0, sizeof(cs));
    cs.step = 0; cs.n_steps = n_steps; cs.seed = seed;
    NDIFF_CUDA_OK(cudaMemcpyAsync(e->chain, &cs, sizeof(cs), cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));
    const int HW = e->H * e->W;
    if (x_init_dev) { if (nchw_to_nhwc4_launch(x_init_dev, e->x, e->B, HW, s)) return 1; }
    else if (philox_normal_launch(e->x, static_cast<size_t>(e->B) * HW, seed, 0ull, s)) return 1;
    e->n_steps = n_steps;
    e->steps_done = 0;
    return 0;
}

int32_t ndiff_chain_run(ndiff_engine* e, int32_t n, const float* noise_dev, const float* teacher_dev,
                        float* snapshots_dev, void* stream) {
    NDIFF_REQUIRE(e && e->n_steps > 0, "ndiff_chain_begin has not been called");
    NDIFF_REQUIRE(n > 0 && e->steps_done + n <= e->n_steps, "step count exceeds the chain length");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    struct { int base; int pad; const float* noise; const float* teacher; float* snap; } io = {e->steps_done, 0, noise_dev,
                                                                                               teacher_dev, snapshots_dev};
    static_assert(offsetof(ChainState, base_step) + sizeof(io) <= sizeof(ChainState), "ChainState io block");
    NDIFF_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(e->chain) + offsetof(ChainState, base_step), &io, sizeof(io),
                                  cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));   // `io` lives on this stack frame
    if (e->cfg.flags & NDIFF_FLAG_NO_GRAPH) {
        for (int i = 0; i < n; ++i)
            if (run_step(e, s)) return 1;
    } else {
        if (!e->step_exec && capture(e, true, &e->step_exec, s)) return 1;
        for (int i = 0; i < n; ++i) NDIFF_CUDA_OK(cudaGraphLaunch(e->step_exec, s));
    }
    e->steps_done += n;
    return 0;
}

int32_t ndiff_chain_read(ndiff_engine* e, float* out_dev, void* stream) {
    NDIFF_REQUIRE(e && out_dev, "null argument");
    DeviceGuard dev_guard(e->cfg.device);
    return nhwc4_to_nchw_launch(e->x, out_dev, e->B, e->H * e->W, as_stream(stream));
}

int32_t ndiff_chain_seek(ndiff_engine* e, int32_t step, const float* x_dev, uint64_t seed, void* stream) {
    NDIFF_REQUIRE(e && e->n_steps > 0 && x_dev, "ndiff_chain_begin has not been called / null state");
    NDIFF_REQUIRE(step >= 0 && step <= e->n_steps, "step out of range");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    const unsigned long long sd = seed;
    NDIFF_CUDA_OK(cudaMemcpyAsync(&e->chain->step, &step, sizeof(int), cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaMemcpyAsync(&e->chain->seed, &sd, sizeof(sd), cudaMemcpyHostToDevice, s));
    NDIFF_CUDA_OK(cudaStreamSynchronize(s));   // `step` / `sd` are stack variables
    e->steps_done = step;
    return nchw_to_nhwc4_launch(x_dev, e->x, e->B, e->H * e->W, s);
}

int32_t ndiff_sample_host(ndiff_engine* e, const float* clean_host, const float* position_host,
                          const int64_t* iso_idx_host, const ndiff_step* steps_host, int32_t n_steps, uint64_t seed,
                          float* out_host) {
    NDIFF_REQUIRE(e && clean_host && position_host && iso_idx_host && out_host, "null argument");
    DeviceGuard dev_guard(e->cfg.device);
    const size_t npix = static_cast<size_t>(e->B) * e->H * e->W;
    float* d_clean = nullptr; float* d_pos = nullptr; float* d_out = nullptr; long long* d_iso = nullptr;
    NDIFF_CUDA_OK(cudaMalloc(&d_clean, npix * 4 * sizeof(float)));
    NDIFF_CUDA_OK(cudaMalloc(&d_pos, npix * 2 * sizeof(float)));
    NDIFF_CUDA_OK(cudaMalloc(&d_out, npix * 4 * sizeof(float)));
    NDIFF_CUDA_OK(cudaMalloc(&d_iso, e->B * sizeof(long long)));
    int rc = 0;
    cudaStream_t s = nullptr;
    do {
        if (cudaMemcpyAsync(d_clean, clean_host, npix * 4 * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d_pos, position_host, npix * 2 * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(d_iso, iso_idx_host, e->B * sizeof(long long), cudaMemcpyHostToDevice, s) != cudaSuccess) {
            set_error("host->device copy failed");
            rc = 1;
            break;
        }
        if ((rc = ndiff_set_condition(e, d_clean, d_pos, reinterpret_cast<const int64_t*>(d_iso), s))) break;
        if ((rc = ndiff_chain_begin(e, steps_host, n_steps, nullptr, seed, s))) break;
        if ((rc = ndiff_chain_run(e, n_steps, nullptr, nullptr, nullptr, s))) break;
        if ((rc = ndiff_chain_read(e, d_out, s))) break;
        if (cudaMemcpyAsync(out_host, d_out, npix * 4 * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            set_error(std::string("device->host copy / chain execution failed: ") + cudaGetErrorString(cudaGetLastError()));
            rc = 1;
        }
    } while (0);
    cudaFree(d_clean); cudaFree(d_pos); cudaFree(d_out); cudaFree(d_iso);
    return rc;
}

int32_t ndiff_debug_tensor(ndiff_engine* e, const char* name, float* out_dev_nchw, int64_t* shape4, void* stream) {
    NDIFF_REQUIRE(e && name && shape4, "null argument");
    DeviceGuard dev_guard(e->cfg.device);
    const std::string n(name);
    if (n == "pos_emb") {
        shape4[0] = e->B; shape4[1] = e->H; shape4[2] = e->W; shape4[3] = 8;   // NHWC fp32, copied verbatim
        if (out_dev_nchw)
            NDIFF_CUDA_OK(cudaMemcpyAsync(out_dev_nchw, e->pos_emb, sizeof(float) * e->B * e->H * e->W * 8,
                                          cudaMemcpyDeviceToDevice, as_stream(stream)));
        return 0;
    }
    if (n == "ss_cur") {
        shape4[0] = e->B; shape4[1] = e->ss_total; shape4[2] = 1; shape4[3] = 1;
        if (out_dev_nchw)
            NDIFF_CUDA_OK(cudaMemcpyAsync(out_dev_nchw, e->ss_cur, sizeof(float) * e->B * e->ss_total,
                                          cudaMemcpyDeviceToDevice, as_stream(stream)));
        return 0;
    }
    if (n == "cvec") {
        shape4[0] = e->B; shape4[1] = e->cv_total; shape4[2] = 1; shape4[3] = 1;
        if (out_dev_nchw)
            NDIFF_CUDA_OK(cudaMemcpyAsync(out_dev_nchw, e->cvec, sizeof(float) * e->B * e->cv_total,
                                          cudaMemcpyDeviceToDevice, as_stream(stream)));
        return 0;
    }
    auto it = e->named.find(n);
    NDIFF_REQUIRE(it != e->named.end(), "no such debug tensor '" + n + "'");
    const Act& a = it->second;
    shape4[0] = e->B; shape4[1] = a.C; shape4[2] = a.H; shape4[3] = a.W;
    if (out_dev_nchw) {
        const size_t total = static_cast<size_t>(e->B) * a.C * a.H * a.W;
        bf16_nhwc_to_f32_nchw_kernel<<<1024, 256, 0, as_stream(stream)>>>(a.p, out_dev_nchw, a.C, a.H * a.W, total);
        NDIFF_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int64_t ndiff_launches_per_step(const ndiff_engine* e) {
    if (!e) return 0;
    int64_t n = 0;
    for (const Op& op : e->net_ops) n += op.launches;
    return n + 3;   // step prologue + fused heads/update + step counter
}

double ndiff_conv_flops_per_step(const ndiff_engine* e) { return e ? e->conv_flops : 0.0; }

int32_t ndiff_time_layers(ndiff_engine* e, int32_t iters, float* ms_out, char* names_out, int32_t names_cap,
                          int32_t* n_out, void* stream) {
    // Per-launch durations measured INSIDE the step: the whole step (prologue, every layer in plan order, fused heads + update)
    // is enqueued back to back on `stream` with a CUDA event between consecutive ops, `iters` times; the figure reported per op
    // is the MEDIAN over the iterations.  Every kernel therefore runs behind its real predecessor, with the cache state and the
    // clocks of a long-running chain — not in isolation.  Once a chain has begun the iterations are real chain steps (they
    // advance the chain), otherwise they are forward evaluations.
    NDIFF_REQUIRE(e && e->plan_built && e->cond_set, "engine not ready");
    DeviceGuard dev_guard(e->cfg.device);
    cudaStream_t s = as_stream(stream);
    g_use_pdl = false;
    const int n_net = static_cast<int>(e->net_ops.size());
    const int n = n_net + 2;     // step prologue + layers + the fused heads / posterior-update kernel
    if (n_out) *n_out = n;
    if (!ms_out) return 0;
    const bool chain = e->n_steps > 0;
    NDIFF_REQUIRE(iters >= 1, "iters must be positive");
    NDIFF_REQUIRE(!chain || e->steps_done + iters + 1 <= e->n_steps, "not enough chain steps left to time (need iters + 1)");
    if (chain) {      // library-drawn noise, no teacher forcing, no snapshots for the timed steps (same io block as ndiff_chain_run)
        struct { int base; int pad; const float* noise; const float* teacher; float* snap; } io = {e->steps_done, 0, nullptr, nullptr, nullptr};
        NDIFF_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(e->chain) + offsetof(ChainState, base_step), &io, sizeof(io),
                                      cudaMemcpyHostToDevice, s));
        NDIFF_CUDA_OK(cudaStreamSynchronize(s));
    }
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& x : ev) NDIFF_CUDA_OK(cudaEventCreate(&x));
    FinalArgs f;
    final_args(e, chain, &f);
    std::vector<std::vector<float>> samples(n);
    for (int it = -1; it < iters; ++it) {      // iteration -1 is an untimed warm-up
        NDIFF_CUDA_OK(cudaEventRecord(ev[0], s));
        if (chain) {
            chain_step_begin_kernel<<<32, 256, 0, s>>>(e->chain, e->step_table, e->ss_table, e->ss_total, e->ss_cur, e->B,
                                                       e->H * e->W, reinterpret_cast<float4*>(e->x), e->stats,
                                                       static_cast<int>(e->stats_bytes / sizeof(unsigned long long)));
        } else {
            zero_u64_kernel<<<32, 256, 0, s>>>(e->stats, static_cast<int>(e->stats_bytes / sizeof(unsigned long long)));
        }
        NDIFF_CUDA_OK(cudaGetLastError());
        NDIFF_CUDA_OK(cudaEventRecord(ev[1], s));
        for (int i = 0; i < n_net; ++i) {
            if (e->net_ops[i].fn(s)) return 1;
            NDIFF_CUDA_OK(cudaEventRecord(ev[i + 2], s));
        }
        if (final_launch(f, s)) return 1;
        NDIFF_CUDA_OK(cudaEventRecord(ev[n], s));
        NDIFF_CUDA_OK(cudaEventSynchronize(ev[n]));
        if (chain) e->steps_done += 1;
        if (it < 0) continue;
        for (int i = 0; i < n; ++i) {
            float ms = 0.f;
            NDIFF_CUDA_OK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            samples[i].push_back(ms);
        }
    }
    for (auto& x : ev) cudaEventDestroy(x);
    std::string names;
    for (int i = 0; i < n; ++i) {
        std::sort(samples[i].begin(), samples[i].end());
        ms_out[i] = samples[i][samples[i].size() / 2];
        const Op* op = (i >= 1 && i <= n_net) ? &e->net_ops[i - 1] : nullptr;
        const double npix = static_cast<double>(e->B) * e->H * e->W;
        const std::string nm = i == 0 ? "step prologue" : (op ? op->name : "final(heads+update)");
        const double fl = op ? op->flops : (i == n - 1 ? 2.0 * npix * e->dim * 4 : 0.0);
        // heads + update: raw block2 output + residual (bf16 x dim each) in, shot image in, state read + written (+ v / noise)
        const double by = op ? op->bytes : (i == n - 1 ? npix * (2.0 * 2 * e->dim + 16 + 32) : 0.0);
        names += nm + ";" + std::to_string(fl) + ";" + std::to_string(by) + "\n";
    }
    if (names_out && names_cap > 0) {
        strncpy(names_out, names.c_str(), names_cap - 1);
        names_out[names_cap - 1] = 0;
    }
    return 0;
}

// ---- single-operator entry points ---------------------------------------------------------------------------
static int op_conv_plan(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                        int32_t C1, int32_t taps_y, int32_t taps_x, int32_t pad_y, int32_t pad_x, const void* weight_packed,
                        int32_t Cout, const float* bias, const float* vec, int32_t vec_ld, const void* res, int32_t act,
                        void* stats, int32_t groups, void* out, int32_t force_nt, int32_t tile_w, ConvGemmPlan* plan,
                        const ndiff_conv_ex* ex = nullptr) {
    int dev = 0;
    NDIFF_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    NDIFF_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    NDIFF_REQUIRE(prop.major == 10, "noisediff_b200 needs an sm_100a GPU (B200)");
    ConvGemmDesc d;
    d.mode = mode; d.B = B; d.H = H; d.W = W;
    d.src0 = static_cast<const bf16*>(src0); d.C0 = C0;
    d.src1 = static_cast<const bf16*>(src1); d.C1 = C1;
    d.taps_y = taps_y; d.taps_x = taps_x; d.pad_y = pad_y; d.pad_x = pad_x;
    d.weight = static_cast<const bf16*>(weight_packed); d.Cout = Cout;
    d.bias = bias; d.vec = vec; d.vec_ld = vec_ld;
    d.res = static_cast<const bf16*>(res); d.res_ld = Cout;
    d.out = static_cast<bf16*>(out); d.out_ld = Cout;
    d.act = act; d.stats = static_cast<unsigned long long*>(stats); d.groups = groups;
    d.force_nt = force_nt; d.TW = tile_w;
    if (ex) {
        d.out2 = static_cast<bf16*>(ex->out2); d.out2_ld = Cout; d.bias2 = ex->bias2;
        d.xf_stats = static_cast<const unsigned long long*>(ex->xf_stats); d.xf_gamma = ex->xf_gamma; d.xf_beta = ex->xf_beta;
        d.xf_ss = ex->xf_ss; d.xf_ss_ld = ex->xf_ss_ld; d.xf_groups = ex->xf_groups;
    }
    return conv_gemm_plan(d, prop.multiProcessorCount, plan);
}

int32_t ndiff_op_conv_ex(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                         int32_t C1, const void* weight_packed, int32_t Cout, const float* bias, void* stats, int32_t groups,
                         void* out, const ndiff_conv_ex* ex, void* stream) {
    NDIFF_REQUIRE(ex != nullptr, "null extension block");
    ConvGemmPlan plan;
    if (op_conv_plan(mode, B, H, W, src0, C0, src1, C1, 3, 3, 1, 1, weight_packed, Cout, bias, nullptr, 0, nullptr, 0, stats,
                     groups, out, 0, 0, &plan, ex)) return 1;
    return conv_gemm_launch(plan, as_stream(stream));
}

int32_t ndiff_op_tail_chain(int32_t npix, int32_t HW, const void* h2, const void* r1, const void* r2, const void* weights_blob,
                            const float* fvec, const void* stats, const float* gamma, const float* beta, int32_t groups,
                            float* out_npix4, void* stream) {
    int dev = 0;
    NDIFF_CUDA_OK(cudaGetDevice(&dev));
    int sms = 0;
    NDIFF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TailDesc d;
    d.npix = npix; d.HW = HW;
    d.h2 = static_cast<const bf16*>(h2); d.r1 = static_cast<const bf16*>(r1); d.r2 = static_cast<const bf16*>(r2);
    d.weights = static_cast<const bf16*>(weights_blob); d.fvec = fvec;
    d.stats = static_cast<const unsigned long long*>(stats); d.gamma = gamma; d.beta = beta; d.groups = groups;
    d.out = out_npix4;
    TailPlan plan;
    if (tail_chain_plan(d, sms, &plan)) return 1;
    return tail_chain_launch(plan, as_stream(stream));
}

int32_t ndiff_op_conv(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                      int32_t C1, int32_t taps_y, int32_t taps_x, int32_t pad_y, int32_t pad_x, const void* weight_packed,
                      int32_t Cout, const float* bias, const float* vec, int32_t vec_ld, const void* res, int32_t act,
                      void* stats, int32_t groups, void* out, int32_t force_nt, int32_t tile_w, void* stream) {
    ConvGemmPlan plan;
    if (op_conv_plan(mode, B, H, W, src0, C0, src1, C1, taps_y, taps_x, pad_y, pad_x, weight_packed, Cout, bias, vec, vec_ld,
                     res, act, stats, groups, out, force_nt, tile_w, &plan)) return 1;
    return conv_gemm_launch(plan, as_stream(stream));
}

int32_t ndiff_op_conv_time(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                           int32_t C1, int32_t taps_y, int32_t taps_x, int32_t pad_y, int32_t pad_x,
                           const void* weight_packed, int32_t Cout, const float* bias, void* stats, int32_t groups, void* out,
                           int32_t force_nt, int32_t tile_w, int32_t iters, float* ms_per_launch, void* stream) {
    ConvGemmPlan plan;
    if (op_conv_plan(mode, B, H, W, src0, C0, src1, C1, taps_y, taps_x, pad_y, pad_x, weight_packed, Cout, bias, nullptr, 0,
                     nullptr, 0, stats, groups, out, force_nt, tile_w, &plan)) return 1;
    cudaStream_t s = as_stream(stream);
    cudaEvent_t e0, e1;
    NDIFF_CUDA_OK(cudaEventCreate(&e0));
    NDIFF_CUDA_OK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i)
        if (conv_gemm_launch(plan, s)) return 1;
    NDIFF_CUDA_OK(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i)
        if (conv_gemm_launch(plan, s)) return 1;
    NDIFF_CUDA_OK(cudaEventRecord(e1, s));
    NDIFF_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    NDIFF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / static_cast<float>(iters > 0 ? iters : 1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int32_t ndiff_op_gn_apply(const void* x, void* out, const void* stats, const float* gamma, const float* beta,
                          const float* ss, int32_t ss_ld, int32_t ss_off, const void* maps, const void* res1,
                          const void* res2, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream) {
    GnApplyArgs g{};
    g.x = static_cast<const bf16*>(x); g.out = static_cast<bf16*>(out);
    g.stats = static_cast<const unsigned long long*>(stats); g.gamma = gamma; g.beta = beta;
    g.ss = ss; g.ss_ld = ss_ld; g.ss_off = ss_off;
    g.maps = static_cast<const bf16*>(maps);
    g.res1 = static_cast<const bf16*>(res1); g.res2 = static_cast<const bf16*>(res2);
    g.B = B; g.HW = HW; g.C = C; g.G = G; g.eps = 1e-5f;
    return gn_apply_launch(g, as_stream(stream));
}

int32_t ndiff_op_layernorm(const void* x, const float* vec, int32_t vec_ld, const float* g, const float* beta, void* out,
                           int32_t B, int32_t HW, int32_t C, void* stream) {
    return layernorm_launch(static_cast<const bf16*>(x), vec, vec_ld, g, beta, static_cast<bf16*>(out), B, HW, C,
                            as_stream(stream));
}

int32_t ndiff_compose_noisy(const float* noise_dev, const float* clean_dev, float* noisy_out_dev, float* clean_out_dev, int64_t n,
                            void* stream) {
    NDIFF_REQUIRE(n >= 0 && n % 4 == 0, "compose: element count must be a multiple of 4");
    return compose_noisy_launch(noise_dev, clean_dev, noisy_out_dev, clean_out_dev, static_cast<size_t>(n / 4), as_stream(stream));
}

int32_t ndiff_op_philox_normal(float* out, int64_t n4, uint64_t seed, uint64_t stream_id, void* stream) {
    return philox_normal_launch(out, static_cast<size_t>(n4), seed, stream_id, as_stream(stream));
}

int32_t ndiff_op_pixel_chain(int32_t prog, int32_t npix, int32_t HW, const void* x, const float* clean_nhwc4,
                             const float* xt_nhwc4, const void* weights_blob, const float* fvec, const float* cvec,
                             int32_t cvec_ld, const float* cvec2, void* out, void* out2, void* stream) {
    int dev = 0;
    NDIFF_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    NDIFF_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    NDIFF_REQUIRE(prop.major == 10, "noisediff_b200 needs an sm_100a GPU (B200)");
    ChainDesc d;
    d.prog = prog; d.npix = npix; d.HW = HW;
    d.x = static_cast<const bf16*>(x); d.clean = clean_nhwc4; d.xt = xt_nhwc4;
    d.weights = static_cast<const bf16*>(weights_blob); d.fvec = fvec; d.cvec = cvec; d.cvec_ld = cvec_ld; d.cvec2 = cvec2;
    d.out = static_cast<bf16*>(out); d.out2 = static_cast<bf16*>(out2);
    ChainPlan plan;
    if (pixel_chain_plan(d, prop.multiProcessorCount, &plan)) return 1;
    return pixel_chain_launch(plan, as_stream(stream));
}

}  // extern "C"
