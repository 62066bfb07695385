// noisediff_b200 — HBM-bound kernels around the tensor-core convolutions (sm_100a).
// All activations are NHWC bf16 (16-byte vectors of 8 channels); the chain state x_t is NHWC fp32 (one float4 / pixel).
#include "pointwise.cuh"

namespace ndiff {

static int g_num_sms = 148;

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 t;
    t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
    u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm apply.  grid = (blocks per sample, B).  A thread owns ONE 16-byte channel vector (8 channels, always inside
// one group because groups are >= 8 channels wide) for its whole life, so the folded affine
//     y = x * A + B,  A = rstd*gamma*(scale+1),  B = (beta - mean*rstd*gamma)*(scale+1) + shift
// lives in 16 registers and the streaming loop is: 16-B loads (kGnUnroll pixels in flight per stream) -> fma -> SiLU ->
// (+ residuals) -> 16-B store.  Statistics are the 2^-24 fixed-point sums written by the conv epilogue.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGnThreads = 256;
constexpr int kGnUnroll = 4;

template <bool kMaps, int kRes>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnApplyArgs a) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int C = a.C, cv = C >> 3, gs = C / a.G;
    const int cvi = threadIdx.x % cv;                  // this thread's channel vector (blockDim.x % cv == 0)
    const int c0 = cvi * 8;
    float A[8], Bc[8];
    {
        const int g = c0 / gs;
        const double inv_n = 1.0 / (static_cast<double>(a.HW) * gs * (a.real_frac > 0.f ? a.real_frac : 1.0f));
        const double s = static_cast<double>(static_cast<long long>(a.stats[(b * a.G + g) * 2])) * (1.0 / 16777216.0);
        const double ss = static_cast<double>(static_cast<long long>(a.stats[(b * a.G + g) * 2 + 1])) * (1.0 / 16777216.0);
        const double meand = s * inv_n;
        const float mean = static_cast<float>(meand);
        const float var = fmaxf(static_cast<float>(ss * inv_n - meand * meand), 0.f);
        const float rstd = rsqrtf(var + a.eps);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float Aj = rstd * __ldg(a.gamma + c0 + j);
            float Bj = __ldg(a.beta + c0 + j) - mean * Aj;
            if (a.ss) {
                const float sc = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + c0 + j] + 1.0f;
                const float sh = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + C + c0 + j];
                Aj *= sc;
                Bj = Bj * sc + sh;
            }
            A[j] = Aj; Bc[j] = Bj;
        }
    }
    const int ppb = kGnThreads / cv;                                   // pixels per block per pass
    const int pstride = gridDim.x * ppb;
    const size_t base = static_cast<size_t>(b) * a.HW * cv + cvi;      // vector index of (b, pixel 0, cvi)
    const uint4* xin = reinterpret_cast<const uint4*>(a.x) + base;
    uint4* xout = reinterpret_cast<uint4*>(a.out) + base;
    const uint4* r1 = kRes >= 1 ? reinterpret_cast<const uint4*>(a.res1) + base : nullptr;
    const uint4* r2 = kRes >= 2 ? reinterpret_cast<const uint4*>(a.res2) + base : nullptr;
    const uint4* mp = kMaps ? reinterpret_cast<const uint4*>(a.maps) + (base - cvi) * 2 : nullptr;   // (b, pixel 0, 0) of [2C]
    for (int p0 = blockIdx.x * ppb + threadIdx.x / cv; p0 < a.HW; p0 += pstride * kGnUnroll) {
        uint4 xv[kGnUnroll], rv1[kGnUnroll], rv2[kGnUnroll], ms[kGnUnroll], mh[kGnUnroll];
#pragma unroll
        for (int i = 0; i < kGnUnroll; ++i) {
            const int p = p0 + i * pstride;
            if (p < a.HW) {
                const size_t o = static_cast<size_t>(p) * cv;
                xv[i] = __ldg(xin + o);
                if (kRes >= 1) rv1[i] = __ldg(r1 + o);
                if (kRes >= 2) rv2[i] = __ldg(r2 + o);
                if (kMaps) { ms[i] = __ldg(mp + o * 2 + cvi); mh[i] = __ldg(mp + o * 2 + cv + cvi); }
            }
        }
#pragma unroll
        for (int i = 0; i < kGnUnroll; ++i) {
            const int p = p0 + i * pstride;
            if (p >= a.HW) break;
            float f[8];
            unpack8(xv[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], A[j], Bc[j]);
            if (kMaps) {
                float sc[8], sh[8];
                unpack8(ms[i], sc);
                unpack8(mh[i], sh);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j] + 1.0f, sh[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = silu(f[j]);
            if (kRes >= 1) {
                float r[8];
                unpack8(rv1[i], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += r[j];
            }
            if (kRes >= 2) {
                float r[8];
                unpack8(rv2[i], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += r[j];
            }
            xout[static_cast<size_t>(p) * cv] = pack8(f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over channels of (x + vec[b]).  L = min(32, C/8) lanes share one pixel, each lane owns VPL 16-byte vectors.
// ---------------------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ vec,
                                                        int vec_ld, const float* __restrict__ g,
                                                        const float* __restrict__ beta, bf16* __restrict__ out,
                                                        int HW, int C, int L, size_t npix_, float real_frac) {
    pdl_trigger();
    pdl_wait();
    // 32-bit pixel arithmetic (the launcher checks npix < 2^31): the kernel is close to instruction bound — ~70 instructions per
    // 16-byte vector is what keeps it on the HBM roofline — and a 64-bit division per pixel alone cost that much
    const unsigned npix = static_cast<unsigned>(npix_);
    const unsigned lane = threadIdx.x & 31;
    const unsigned sub = lane % L, slot = lane / L, ppw = 32 / L;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned cv = C >> 3;
    float gg[VPL][8], bb[VPL][8];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int c = (sub + j * L) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) { gg[j][k] = g[c + k]; bb[j][k] = beta[c + k]; }
    }
    // zero-padded channel layouts: x + vec is exactly 0 on the padding, so the sum is that of the live channels; the centred
    // squares pick up n_pad * mean^2 from the padding, which is taken out again
    const float inv_c = 1.0f / (static_cast<float>(C) * real_frac);
    const float n_pad = static_cast<float>(C) * (1.0f - real_frac);
    // kLnU pixels per lane group and iteration: all 16-byte loads are issued before the first reduction, so the shuffles and
    // the dependent arithmetic of one pixel overlap the memory latency of the others
    constexpr int kLnU = VPL == 1 ? 4 : 2;
    const uint4* __restrict__ x4 = reinterpret_cast<const uint4*>(x);
    uint4* __restrict__ o4 = reinterpret_cast<uint4*>(out);
    for (unsigned p0 = warp * ppw * kLnU; p0 < npix; p0 += nwarps * ppw * kLnU) {
        float f[kLnU][VPL][8];
        unsigned pixs[kLnU];
        uint4 raw[kLnU][VPL];
#pragma unroll
        for (int u = 0; u < kLnU; ++u) {
            pixs[u] = p0 + u * ppw + slot;
            const unsigned pp = pixs[u] < npix ? pixs[u] : npix - 1;
#pragma unroll
            for (int j = 0; j < VPL; ++j) raw[u][j] = __ldg(x4 + static_cast<size_t>(pp) * cv + sub + j * L);
        }
#pragma unroll
        for (int u = 0; u < kLnU; ++u) {
            const unsigned pp = pixs[u] < npix ? pixs[u] : npix - 1;
            const float* vp = vec + static_cast<size_t>(pp / static_cast<unsigned>(HW)) * vec_ld;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int cvi = sub + j * L;
                unpack8(raw[u][j], f[u][j]);
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(vp + cvi * 8));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(vp + cvi * 8 + 4));
                f[u][j][0] += v0.x; f[u][j][1] += v0.y; f[u][j][2] += v0.z; f[u][j][3] += v0.w;
                f[u][j][4] += v1.x; f[u][j][5] += v1.y; f[u][j][6] += v1.z; f[u][j][7] += v1.w;
#pragma unroll
                for (int k = 0; k < 8; ++k) sum += f[u][j][k];
            }
            for (int o = L >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * inv_c;
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < VPL; ++j)
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float d = f[u][j][k] - mean; sq += d * d; }
            for (int o = L >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(fmaxf(sq - n_pad * mean * mean, 0.f) * inv_c + 1e-5f);
            if (pixs[u] < npix) {
#pragma unroll
                for (int j = 0; j < VPL; ++j) {
                    float o8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) o8[k] = (f[u][j][k] - mean) * rstd * gg[j][k] + bb[j][k];
                    o4[static_cast<size_t>(pixs[u]) * cv + sub + j * L] = pack8(o8);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shot_mlp1.fc1: cat[clean, x] (8) -> C, GELU.  thread = (pixel, 8 output channels)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) shot_in_kernel(const float4* __restrict__ clean, const float4* __restrict__ x,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      bf16* __restrict__ out, size_t npix, int C) {
    // weights regrouped per 8-channel output group as [group][k][j] with a 68-float group pitch: the 8 lanes that
    // share a pixel read float4s from disjoint bank quads (no conflicts), lanes of other pixels broadcast
    extern __shared__ float sw[];
    const int cv = C >> 3;
    float* sbias = sw + cv * 68;
    for (int i = threadIdx.x; i < C * 8; i += blockDim.x) {
        const int c = i >> 3, k = i & 7;
        sw[(c >> 3) * 68 + k * 8 + (c & 7)] = w[i];
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) sbias[i] = bias[i];
    __syncthreads();
    const size_t total = npix * cv;
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
         v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t pix = v / cv;
        const int cg = static_cast<int>(v % cv);
        const float4 a = __ldg(clean + pix), bq = __ldg(x + pix);
        const float in[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = sbias[cg * 8 + j];
        const float4* wg = reinterpret_cast<const float4*>(sw + cg * 68);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 w0 = wg[k * 2], w1 = wg[k * 2 + 1];
            f[0] += w0.x * in[k]; f[1] += w0.y * in[k]; f[2] += w0.z * in[k]; f[3] += w0.w * in[k];
            f[4] += w1.x * in[k]; f[5] += w1.y * in[k]; f[6] += w1.z * in[k]; f[7] += w1.w * in[k];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
        reinterpret_cast<uint4*>(out)[v] = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// init_conv 7x7 pad 3, 4 -> C(=64), fp32.  block = 16x16 pixels; thread = 1 pixel x 64 outputs
// ---------------------------------------------------------------------------------------------------------------
constexpr int kIcTile = 16, kIcHalo = kIcTile + 6;

__global__ void __launch_bounds__(256) init_conv7_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, bf16* __restrict__ out, int H,
                                                         int W) {
    extern __shared__ float4 sm4[];
    float4* sx = sm4;                                               // [22*22]
    float4* swt = sm4 + kIcHalo * kIcHalo;                          // [49][4 ci][16 float4 of co]
    const int b = blockIdx.z, y0 = blockIdx.y * kIcTile, x0 = blockIdx.x * kIcTile;
    for (int i = threadIdx.x; i < 49 * 4 * 16; i += 256) swt[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    for (int i = threadIdx.x; i < kIcHalo * kIcHalo; i += 256) {
        const int yy = y0 + i / kIcHalo - 3, xx = x0 + i % kIcHalo - 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(x + (static_cast<size_t>(b) * H + yy) * W + xx);
        sx[i] = v;
    }
    __syncthreads();
    const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = __ldg(bias + c);
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll 1
        for (int kx = 0; kx < 7; ++kx) {
            const float4 xi = sx[(ly + ky) * kIcHalo + lx + kx];
            const float4* wt = swt + (ky * 7 + kx) * 64;
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
                const float4 w0 = wt[c4], w1 = wt[16 + c4], w2 = wt[32 + c4], w3 = wt[48 + c4];
                acc[c4 * 4 + 0] += xi.x * w0.x + xi.y * w1.x + xi.z * w2.x + xi.w * w3.x;
                acc[c4 * 4 + 1] += xi.x * w0.y + xi.y * w1.y + xi.z * w2.y + xi.w * w3.y;
                acc[c4 * 4 + 2] += xi.x * w0.z + xi.y * w1.z + xi.z * w2.z + xi.w * w3.z;
                acc[c4 * 4 + 3] += xi.x * w0.w + xi.y * w1.w + xi.z * w2.w + xi.w * w3.w;
            }
        }
    }
    const int yy = y0 + ly, xx = x0 + lx;
    if (yy < H && xx < W) {
        uint4* op = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * H + yy) * W + xx) * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint4 u;
            u.x = pack_bf16(acc[j * 8 + 0], acc[j * 8 + 1]); u.y = pack_bf16(acc[j * 8 + 2], acc[j * 8 + 3]);
            u.z = pack_bf16(acc[j * 8 + 4], acc[j * 8 + 5]); u.w = pack_bf16(acc[j * 8 + 6], acc[j * 8 + 7]);
            op[j] = u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H,
                                                         int W, int cv, size_t total) {
    // H, W = INPUT size; out is [B, 2H, 2W, C]
    pdl_trigger();
    pdl_wait();
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
         v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(v % cv);
        size_t p = v / cv;
        const int ox = static_cast<int>(p % (2 * W)); p /= (2 * W);
        const int oy = static_cast<int>(p % (2 * H));
        const size_t b = p / (2 * H);
        out[v] = __ldg(in + ((b * H + (oy >> 1)) * W + (ox >> 1)) * cv + c);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 -> 4 standard normals (Box-Muller)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long stream_id,
                                                 unsigned long long idx) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                             static_cast<uint32_t>(stream_id), static_cast<uint32_t>(stream_id >> 32)),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    // (0,1] uniforms with 32 random bits (bottom bits lost to fp32 rounding; never 0)
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = (static_cast<float>(r.x) + 1.0f) * k, u1 = static_cast<float>(r.y) * k;
    const float u2 = (static_cast<float>(r.z) + 1.0f) * k, u3 = static_cast<float>(r.w) * k;
    const float ra = sqrtf(-2.0f * logf(fminf(u0, 1.0f))), rb = sqrtf(-2.0f * logf(fminf(u2, 1.0f)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float4* out, size_t n4, unsigned long long seed,
                                                            unsigned long long stream_id) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = philox_normal4(seed, stream_id, i);
}

// ---------------------------------------------------------------------------------------------------------------
// final 1x1 heads + posterior update.  8 lanes per pixel, each lane 8 channels of both feature maps.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_kernel(const FinalArgs a) {
    extern __shared__ float sw[];  // wf[4][C], ws[4][C]
    const int C = a.C;
    pdl_trigger();
    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) { sw[i] = a.wf[i]; sw[4 * C + i] = a.ws[i]; }
    __syncthreads();
    pdl_wait();
    const int lanes = C >> 3;                                  // lanes cooperating on one pixel (8 for C = 64)
    const int sub = threadIdx.x % lanes;
    const float bsum[4] = {a.bs[0] + a.bfin[0], a.bs[1] + a.bfin[1], a.bs[2] + a.bfin[2], a.bs[3] + a.bfin[3]};
    const size_t ppb = blockDim.x / lanes;                     // pixels per block per pass
    const size_t npix = static_cast<size_t>(a.npix);
    const size_t passes = (npix + ppb * gridDim.x - 1) / (ppb * gridDim.x);
    StepParams sp{};
    int step = 0, rel = 0;
    const float* noise = nullptr;
    float* snap = nullptr;
    unsigned long long seed = 0ull;
    if (a.chain) {
        sp = a.chain->cur; step = a.chain->step; rel = step - a.chain->base_step;   // index into this run's noise / snapshot arrays
        noise = a.chain->noise; snap = a.chain->snap; seed = a.chain->seed;
    }
    for (size_t it = 0; it < passes; ++it) {
        const size_t pix = (it * gridDim.x + blockIdx.x) * ppb + threadIdx.x / lanes;
        const bool live = pix < npix;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            float f[8], g[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(a.xf) + pix * lanes + sub), f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(a.sf) + pix * lanes + sub), g);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += f[j] * sw[k * C + sub * 8 + j] + g[j] * sw[4 * C + k * C + sub * 8 + j];
                o[k] = acc;
            }
        }
        for (int off = lanes >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] += __shfl_xor_sync(0xffffffffu, o[k], off);
        }
        if (!live || sub != 0) continue;
        // reference order: shot_noise (= fc2 out + bias) + read_noise (= final_conv out + bias)
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = o[k] + bsum[k];
        if (a.v_out) reinterpret_cast<float4*>(a.v_out)[pix] = make_float4(v[0], v[1], v[2], v[3]);
        if (!a.chain) continue;

        const size_t HW = a.HW, bimg = pix / HW, hw = pix % HW;
        const size_t nB = npix / HW;
        const size_t plane0 = ((static_cast<size_t>(rel) * nB + bimg) * 4) * HW + hw;   // NCHW offset of channel 0
        const float4 xi4 = reinterpret_cast<const float4*>(a.x)[pix];
        const float xi[4] = {xi4.x, xi4.y, xi4.z, xi4.w};
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (sp.sigma != 0.f) {
            if (noise) {
#pragma unroll
                for (int k = 0; k < 4; ++k) z[k] = __ldg(noise + plane0 + k * HW);
            } else {
                const float4 z4 = philox_normal4(seed, static_cast<unsigned long long>(step) + 1ull, pix);
                z[0] = z4.x; z[1] = z4.y; z[2] = z4.z; z[3] = z4.w;
            }
        }
        float xn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // separate roundings (no FMA contraction) to follow the reference's elementwise torch ops
            float x0 = __fadd_rn(__fmul_rn(sp.p, xi[k]), __fmul_rn(sp.q, v[k]));
            if (sp.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
            float m = __fadd_rn(__fmul_rn(sp.a, x0), __fmul_rn(sp.b, xi[k]));
            if (sp.c != 0.f) {
                const float eps = __fdiv_rn(__fadd_rn(__fmul_rn(sp.r1, xi[k]), -x0), sp.r2);
                m = __fadd_rn(m, __fmul_rn(sp.c, eps));
            }
            xn[k] = sp.sigma != 0.f ? __fadd_rn(m, __fmul_rn(sp.sigma, z[k])) : m;
        }
        reinterpret_cast<float4*>(a.x)[pix] = make_float4(xn[0], xn[1], xn[2], xn[3]);
        if (snap) {
#pragma unroll
            for (int k = 0; k < 4; ++k) snap[plane0 + k * HW] = xn[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dim = 64 fast path of the fused heads + posterior update.  A warp owns 32 consecutive pixels (never straddling a sample:
// HW % 32 == 0).  Phase 1: 8 passes of 4 pixels x 8 lanes, all 16-byte loads of the 32 pixels issued up front, each lane
// dots its channel octet against head weights held in registers; a reduce-scatter butterfly (28 shuffles) leaves lane
// (grp, sub) with the four head outputs of pixel 4*sub + grp.  Phase 2: ALL 32 lanes run Philox / Box-Muller and the
// posterior arithmetic for one pixel each (the generic kernel below does that with 1 lane in 8).
// kGn: the xf operand is SiLU(GroupNorm(xf)) + gn_res evaluated on the fly — final_res_block.block2.norm never
// materialises (ref Diffusion_arch.py:135-170,640-643).
// ---------------------------------------------------------------------------------------------------------------
template <bool kGn, bool kSf>
__global__ void __launch_bounds__(256, 2) final64_kernel(const FinalArgs a) {
    __shared__ __align__(16) float sw[2][4][64];      // final_conv / shot_mlp3.fc2 head weights
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sw[0][i >> 6][i & 63] = __ldg(a.wf + i);
        sw[1][i >> 6][i & 63] = kSf ? __ldg(a.ws + i) : 0.f;
    }
    __syncthreads();
    // kSf == false: the shot head arrives precomputed in a.sn (bias included)
    const float bsum[4] = {(kSf ? a.bs[0] : 0.f) + a.bfin[0], (kSf ? a.bs[1] : 0.f) + a.bfin[1],
                           (kSf ? a.bs[2] : 0.f) + a.bfin[2], (kSf ? a.bs[3] : 0.f) + a.bfin[3]};
    pdl_wait();
    StepParams sp{};
    int step = 0, rel = 0;
    const float* noise = nullptr;
    float* snap = nullptr;
    unsigned long long seed = 0ull;
    if (a.chain) {
        sp = a.chain->cur; step = a.chain->step; rel = step - a.chain->base_step;
        noise = a.chain->noise; snap = a.chain->snap; seed = a.chain->seed;
    }
    const size_t HW = a.HW, npix = static_cast<size_t>(a.npix), nB = npix / HW;
    const size_t n_chunks = npix >> 5;
    const size_t n_warps = static_cast<size_t>(gridDim.x) * (blockDim.x >> 5);
    const size_t gw = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t c_begin = n_chunks * gw / n_warps, c_end = n_chunks * (gw + 1) / n_warps;   // contiguous range per warp
    float A[8], Bc[8];          // folded GroupNorm affine, halved: SiLU(y) = h + h tanh(h), h = y / 2 = fma(x, A, B)
    long long cur_b = -1;
    const uint4* xfv = reinterpret_cast<const uint4*>(a.xf);
    const uint4* sfv = reinterpret_cast<const uint4*>(a.sf);
    const uint4* rsv = reinterpret_cast<const uint4*>(a.gn_res);
    for (size_t ch = c_begin; ch < c_end; ++ch) {
        const size_t pix0 = ch << 5;
        const size_t bimg = pix0 / HW;
        if (kGn && static_cast<long long>(bimg) != cur_b) {
            cur_b = static_cast<long long>(bimg);
            const int gs = 64 / a.gn_G, g = (sub * 8) / gs;
            const double inv_n = 1.0 / (static_cast<double>(a.HW) * gs * (a.gn_real_frac > 0.f ? a.gn_real_frac : 1.0f));
            const double s = static_cast<double>(static_cast<long long>(a.gn_stats[(bimg * a.gn_G + g) * 2])) * (1.0 / 16777216.0);
            const double ss = static_cast<double>(static_cast<long long>(a.gn_stats[(bimg * a.gn_G + g) * 2 + 1])) * (1.0 / 16777216.0);
            const double meand = s * inv_n;
            const float mean = static_cast<float>(meand);
            const float var = fmaxf(static_cast<float>(ss * inv_n - meand * meand), 0.f);
            const float rstd = rsqrtf(var + a.gn_eps);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float Aj = rstd * __ldg(a.gn_gamma + sub * 8 + j);
                const float Bj = __ldg(a.gn_beta + sub * 8 + j) - mean * Aj;
                A[j] = 0.5f * Aj; Bc[j] = 0.5f * Bj;
            }
        }
        // ---- phase 1: per-lane partial dot products of 8 x 4 pixels, in two batches of four passes ------------------------
        float o[8][4];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            uint4 fv[4], gv[4], rv[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const size_t idx = (pix0 + (hb * 4 + ii) * 4 + grp) * 8 + sub;
                fv[ii] = __ldg(xfv + idx);
                if (kSf) gv[ii] = __ldg(sfv + idx);
                if (kGn && rsv) rv[ii] = __ldg(rsv + idx);
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                float f[8], g[8];
                unpack8(fv[ii], f);
                if (kSf) unpack8(gv[ii], g);
                if (kGn) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float h = fmaf(f[j], A[j], Bc[j]);
                        f[j] = fmaf(h, tanh_approx(h), h);
                    }
                    if (rsv) {
                        float r[8];
                        unpack8(rv[ii], r);
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] += r[j];
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&sw[0][k][sub * 8]), w1 = *reinterpret_cast<const float4*>(&sw[0][k][sub * 8 + 4]);
                    float acc = f[0] * w0.x;
                    acc = fmaf(f[1], w0.y, acc); acc = fmaf(f[2], w0.z, acc); acc = fmaf(f[3], w0.w, acc);
                    acc = fmaf(f[4], w1.x, acc); acc = fmaf(f[5], w1.y, acc); acc = fmaf(f[6], w1.z, acc); acc = fmaf(f[7], w1.w, acc);
                    if (kSf) {
                        const float4 v0 = *reinterpret_cast<const float4*>(&sw[1][k][sub * 8]), v1 = *reinterpret_cast<const float4*>(&sw[1][k][sub * 8 + 4]);
                        acc = fmaf(g[0], v0.x, acc); acc = fmaf(g[1], v0.y, acc); acc = fmaf(g[2], v0.z, acc); acc = fmaf(g[3], v0.w, acc);
                        acc = fmaf(g[4], v1.x, acc); acc = fmaf(g[5], v1.y, acc); acc = fmaf(g[6], v1.z, acc); acc = fmaf(g[7], v1.w, acc);
                    }
                    o[hb * 4 + ii][k] = acc;
                }
            }
        }
        // ---- reduce-scatter over the 8 lanes of a pixel group: lane `sub` ends with pass `sub` -------------------------
        float t4[4][4], t2[2][4], v[4];
        const bool hi = (sub & 4) != 0, mid = (sub & 2) != 0, lo = (sub & 1) != 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float keep = hi ? o[j + 4][k] : o[j][k], send = hi ? o[j][k] : o[j + 4][k];
                t4[j][k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float keep = mid ? t4[j + 2][k] : t4[j][k], send = mid ? t4[j][k] : t4[j + 2][k];
                t2[j][k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float keep = lo ? t2[1][k] : t2[0][k], send = lo ? t2[0][k] : t2[1][k];
            // reference order: shot_noise (= fc2 out + bias) + read_noise (= final_conv out + bias)
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1) + bsum[k];
        }
        // ---- phase 2: one pixel per lane -------------------------------------------------------------------------------
        const size_t pix = pix0 + sub * 4 + grp;
        if (!kSf) {         // reference order: shot_noise + read_noise
            const float4 s4 = __ldg(a.sn + pix);
            v[0] = s4.x + v[0]; v[1] = s4.y + v[1]; v[2] = s4.z + v[2]; v[3] = s4.w + v[3];
        }
        if (a.v_out) reinterpret_cast<float4*>(a.v_out)[pix] = make_float4(v[0], v[1], v[2], v[3]);
        if (!a.chain) continue;
        const size_t hw = pix - bimg * HW;
        const size_t plane0 = ((static_cast<size_t>(rel) * nB + bimg) * 4) * HW + hw;   // NCHW offset of channel 0
        const float4 xi4 = reinterpret_cast<const float4*>(a.x)[pix];
        const float xi[4] = {xi4.x, xi4.y, xi4.z, xi4.w};
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (sp.sigma != 0.f) {
            if (noise) {
#pragma unroll
                for (int k = 0; k < 4; ++k) z[k] = __ldg(noise + plane0 + k * HW);
            } else {
                const float4 z4 = philox_normal4(seed, static_cast<unsigned long long>(step) + 1ull, pix);
                z[0] = z4.x; z[1] = z4.y; z[2] = z4.z; z[3] = z4.w;
            }
        }
        float xn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // separate roundings (no FMA contraction) to follow the reference's elementwise torch ops
            float x0 = __fadd_rn(__fmul_rn(sp.p, xi[k]), __fmul_rn(sp.q, v[k]));
            if (sp.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
            float m = __fadd_rn(__fmul_rn(sp.a, x0), __fmul_rn(sp.b, xi[k]));
            if (sp.c != 0.f) {
                const float eps = __fdiv_rn(__fadd_rn(__fmul_rn(sp.r1, xi[k]), -x0), sp.r2);
                m = __fadd_rn(m, __fmul_rn(sp.c, eps));
            }
            xn[k] = sp.sigma != 0.f ? __fadd_rn(m, __fmul_rn(sp.sigma, z[k])) : m;
        }
        reinterpret_cast<float4*>(a.x)[pix] = make_float4(xn[0], xn[1], xn[2], xn[3]);
        if (snap) {
#pragma unroll
            for (int k = 0; k < 4; ++k) snap[plane0 + k * HW] = xn[k];
        }
    }
}

__global__ void chain_advance_kernel(ChainState* chain) {
    pdl_wait();
    chain->step += 1;
}

// ---------------------------------------------------------------------------------------------------------------
// time path
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_mlp_kernel(const int* __restrict__ t, int t_stride, int dim,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       float* __restrict__ st_out) {
    __shared__ float emb[128], h[512];
    const int n = blockIdx.x, td = dim * 4, half = dim / 2;
    const float tv = static_cast<float>(t[n * t_stride]);
    if (threadIdx.x < half) {
        const float f = expf(static_cast<float>(threadIdx.x) * -(logf(10000.0f) / static_cast<float>(half - 1)));
        emb[threadIdx.x] = sinf(tv * f);
        emb[half + threadIdx.x] = cosf(tv * f);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < td; o += blockDim.x) {
        float acc = b1[o];
        for (int k = 0; k < dim; ++k) acc += w1[o * dim + k] * emb[k];
        h[o] = gelu_exact(acc);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < td; o += blockDim.x) {
        float acc = b2[o];
        for (int k = 0; k < td; ++k) acc += w2[o * td + k] * h[k];
        st_out[static_cast<size_t>(n) * td + o] = acc / (1.0f + expf(-acc));   // SiLU feeding every ResnetBlock.mlp
    }
}

__global__ void __launch_bounds__(256) rows_gemv_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                        const float* __restrict__ in, float* __restrict__ out, int rows,
                                                        int K) {
    const int n = blockIdx.y;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* wr = W + static_cast<size_t>(row) * K;
    const float* x = in + static_cast<size_t>(n) * K;
    float acc = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
        const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + k));
        acc += w4.x * x4.x + w4.y * x4.y + w4.z * x4.z + w4.w * x4.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[static_cast<size_t>(n) * rows + row] = acc + bias[row];
}

__global__ void __launch_bounds__(128) iso_vec_kernel(const float* __restrict__ emb_table,
                                                      const long long* __restrict__ idx, const float* __restrict__ wv,
                                                      const float* __restrict__ wo, const float* __restrict__ bo,
                                                      float* __restrict__ out, int out_ld, int out_off, int C) {
    __shared__ float v[128];
    const int b = blockIdx.x;
    const float* e = emb_table + idx[b] * 16;
    {
        float acc = 0.f;
        for (int k = 0; k < 16; ++k) acc += wv[threadIdx.x * 16 + k] * e[k];
        v[threadIdx.x] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 128) {
        float acc = bo[c];
        for (int k = 0; k < 128; ++k) acc += wo[c * 128 + k] * v[k];
        out[static_cast<size_t>(b) * out_ld + out_off + c] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// positional path (runs once per condition)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pos_maps_kernel(const PosArgs a) {
    extern __shared__ float sw[];
    const int C2 = 2 * a.C;
    float* wm1 = sw;                 // [2C][8]
    float* wm2 = wm1 + C2 * 8;
    float* bm1 = wm2 + C2 * 8;       // [2C]
    float* bm2 = bm1 + C2;
    float* small = bm2 + C2;         // we[16] be[8] w1[384] b1[16] w2[128] b2[8]
    for (int i = threadIdx.x; i < C2 * 8; i += blockDim.x) { wm1[i] = a.wm1[i]; wm2[i] = a.wm2[i]; }
    for (int i = threadIdx.x; i < C2; i += blockDim.x) { bm1[i] = a.bm1[i]; bm2[i] = a.bm2[i]; }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) small[i] = a.we[i];
    for (int i = threadIdx.x; i < 8; i += blockDim.x) small[16 + i] = a.be[i];
    for (int i = threadIdx.x; i < 384; i += blockDim.x) small[24 + i] = a.w1[i];
    for (int i = threadIdx.x; i < 16; i += blockDim.x) small[408 + i] = a.b1[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) small[424 + i] = a.w2[i];
    for (int i = threadIdx.x; i < 8; i += blockDim.x) small[552 + i] = a.b2[i];
    __syncthreads();
    const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= static_cast<size_t>(a.B) * a.HW) return;
    const size_t b = pix / a.HW, hw = pix % a.HW;
    const float p0 = a.position[(b * 2 + 0) * a.HW + hw], p1 = a.position[(b * 2 + 1) * a.HW + hw];
    float feat[24];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float w = small[j * 2] * p0 + small[j * 2 + 1] * p1 + small[16 + j];
        const float fr = w * 2.0f * 3.14159265358979323846f;
        feat[j] = w; feat[8 + j] = sinf(fr); feat[16 + j] = cosf(fr);
    }
    float h[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float acc = small[408 + o];
#pragma unroll
        for (int k = 0; k < 24; ++k) acc += small[24 + o * 24 + k] * feat[k];
        h[o] = gelu_exact(acc);
    }
    float pe[8], ps[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float acc = small[552 + o];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += small[424 + o * 16 + k] * h[k];
        pe[o] = acc;
        ps[o] = acc / (1.0f + expf(-acc));
    }
    if (a.pos_emb) {
#pragma unroll
        for (int o = 0; o < 8; ++o) a.pos_emb[pix * 8 + o] = pe[o];
    }
    for (int which = 0; which < 2; ++which) {
        const float* wm = which ? wm2 : wm1;
        const float* bm = which ? bm2 : bm1;
        uint4* mp = reinterpret_cast<uint4*>((which ? a.map2 : a.map1) + pix * C2);
        for (int c0 = 0; c0 < C2; c0 += 8) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float acc = bm[c0 + j];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += wm[(c0 + j) * 8 + k] * ps[k];
                f[j] = acc;
            }
            mp[c0 >> 3] = pack8(f);
        }
    }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc4_kernel(const float* __restrict__ in, float4* __restrict__ out,
                                                            int HW, size_t npix) {
    const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const size_t b = pix / HW, hw = pix % HW;
    const float* p = in + b * 4 * HW + hw;
    out[pix] = make_float4(p[0], p[HW], p[2 * static_cast<size_t>(HW)], p[3 * static_cast<size_t>(HW)]);
}
__global__ void __launch_bounds__(256) nhwc4_to_nchw_kernel(const float4* __restrict__ in, float* __restrict__ out,
                                                            int HW, size_t npix) {
    const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const size_t b = pix / HW, hw = pix % HW;
    const float4 v = in[pix];
    float* p = out + b * 4 * HW + hw;
    p[0] = v.x; p[HW] = v.y; p[2 * static_cast<size_t>(HW)] = v.z; p[3 * static_cast<size_t>(HW)] = v.w;
}

__global__ void __launch_bounds__(256) compose_noisy_kernel(const float4* __restrict__ noise, const float4* __restrict__ clean,
                                                            float4* __restrict__ noisy, float4* __restrict__ clean_out, size_t n4) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 z = __ldg(noise + i), c = __ldg(clean + i);
        auto one = [](float zz, float cc) { return fminf(fmaxf(fminf(fmaxf(zz, -1.0f), 1.0f) + cc, 0.0f), 1.0f); };
        noisy[i] = make_float4(one(z.x, c.x), one(z.y, c.y), one(z.z, c.z), one(z.w, c.w));
        if (clean_out) clean_out[i] = make_float4(fminf(fmaxf(c.x, 0.f), 1.f), fminf(fmaxf(c.y, 0.f), 1.f), fminf(fmaxf(c.z, 0.f), 1.f),
                                                  fminf(fmaxf(c.w, 0.f), 1.f));
    }
}

inline int blocks_for(size_t n, int per_block, int cap = 1 << 20) {
    size_t b = (n + per_block - 1) / per_block;
    return static_cast<int>(b < static_cast<size_t>(cap) ? (b ? b : 1) : cap);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
static int g_gn_occ[2][3] = {{0, 0, 0}, {0, 0, 0}};
static int g_final_occ[2][2] = {{0, 0}, {0, 0}};      // resident blocks per SM of final64_kernel<kGn, kSf>   // resident blocks per SM of each gn_apply variant (pointwise_init)

int gn_apply_launch(const GnApplyArgs& a, cudaStream_t s) {
    NDIFF_REQUIRE(a.C % 64 == 0 && a.C <= 512 && a.C % a.G == 0 && (a.C / a.G) % 8 == 0,
                  "GroupNorm apply: channels must be a multiple of 64 (<= 512) in groups of >= 8");
    NDIFF_REQUIRE(!(a.res2 && !a.res1), "GroupNorm apply: res2 without res1");
    const int nres = a.res2 ? 2 : (a.res1 ? 1 : 0);
    const int cv = a.C / 8, ppb = kGnThreads / cv;
    const int want = (a.HW + ppb * kGnUnroll - 1) / (ppb * kGnUnroll);
    // exactly one wave: (resident blocks per SM) x SMs blocks over the whole batch, each looping over its share
    const int occ = g_gn_occ[a.maps ? 1 : 0][nres] > 0 ? g_gn_occ[a.maps ? 1 : 0][nres] : 2;
    const int cap = (g_num_sms * occ) / a.B > 0 ? (g_num_sms * occ) / a.B : 1;
    dim3 grid(want < cap ? want : cap, a.B);
    if (a.maps) {
        if (nres == 0) NDIFF_CUDA_OK(launch_pdl(gn_apply_kernel<true, 0>, grid, dim3(kGnThreads), 0, s, a));
        else if (nres == 1) NDIFF_CUDA_OK(launch_pdl(gn_apply_kernel<true, 1>, grid, dim3(kGnThreads), 0, s, a));
        else NDIFF_CUDA_OK(launch_pdl(gn_apply_kernel<true, 2>, grid, dim3(kGnThreads), 0, s, a));
    } else {
        if (nres == 0) NDIFF_CUDA_OK(launch_pdl(gn_apply_kernel<false, 0>, grid, dim3(kGnThreads), 0, s, a));
        else if (nres == 1) NDIFF_CUDA_OK(launch_pdl(gn_apply_kernel<false, 1>, grid, dim3(kGnThreads), 0, s, a));
        else NDIFF_CUDA_OK(launch_pdl(gn_apply_kernel<false, 2>, grid, dim3(kGnThreads), 0, s, a));
    }
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int layernorm_launch(const bf16* x, const float* vec, int vec_ld, const float* g, const float* beta, bf16* out, int B,
                     int HW, int C, cudaStream_t s, float real_frac) {
    NDIFF_REQUIRE(real_frac > 0.f && real_frac <= 1.f, "LayerNorm: live channel fraction must be in (0, 1]");
    NDIFF_REQUIRE(static_cast<long long>(B) * HW < (1ll << 31), "LayerNorm: 32-bit pixel indices");
    NDIFF_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, "LayerNorm: C must be 64, 128, 256 or 512");
    NDIFF_REQUIRE(vec_ld % 4 == 0, "LayerNorm: per-sample vector stride must keep 16-byte alignment");
    const size_t npix = static_cast<size_t>(B) * HW;
    const int L = C / 8 > 32 ? 32 : C / 8;
    const int ppw = 32 / L;
    const int grid = blocks_for(npix, 8 * ppw * 4, 148 * 8);      // (a grid-stride loop: any grid covers the tensor)
    if (C == 512)
        NDIFF_CUDA_OK(launch_pdl(layernorm_kernel<2>, dim3(grid), dim3(256), 0, s, x, vec, vec_ld, g, beta, out, HW, C, L, npix, real_frac));
    else
        NDIFF_CUDA_OK(launch_pdl(layernorm_kernel<1>, dim3(grid), dim3(256), 0, s, x, vec, vec_ld, g, beta, out, HW, C, L, npix, real_frac));
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int shot_in_launch(const float* clean, const float* x, const float* w, const float* bias, bf16* out, int npix, int C,
                   cudaStream_t s) {
    const size_t total = static_cast<size_t>(npix) * (C / 8);
    shot_in_kernel<<<blocks_for(total, 256 * 4, 148 * 8), 256, ((C / 8) * 68 + C) * sizeof(float), s>>>(
        reinterpret_cast<const float4*>(clean), reinterpret_cast<const float4*>(x), w, bias, out, npix, C);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pointwise_init() {
    {
        int dev = 0;
        NDIFF_CUDA_OK(cudaGetDevice(&dev));
        NDIFF_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_gn_occ[0][0], gn_apply_kernel<false, 0>, kGnThreads, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_gn_occ[0][1], gn_apply_kernel<false, 1>, kGnThreads, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_gn_occ[0][2], gn_apply_kernel<false, 2>, kGnThreads, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_gn_occ[1][0], gn_apply_kernel<true, 0>, kGnThreads, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_gn_occ[1][1], gn_apply_kernel<true, 1>, kGnThreads, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_gn_occ[1][2], gn_apply_kernel<true, 2>, kGnThreads, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_final_occ[0][0], final64_kernel<false, false>, 256, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_final_occ[0][1], final64_kernel<false, true>, 256, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_final_occ[1][0], final64_kernel<true, false>, 256, 0));
        NDIFF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_final_occ[1][1], final64_kernel<true, true>, 256, 0));
    }
    const int smem = (kIcHalo * kIcHalo + 49 * 4 * 16) * sizeof(float4);
    NDIFF_CUDA_OK(cudaFuncSetAttribute(init_conv7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return 0;
}

int init_conv7_launch(const float* x, const float* w, const float* bias, bf16* out, int B, int H, int W, int C,
                      cudaStream_t s) {
    NDIFF_REQUIRE(C == 64, "init_conv kernel is specialised for dim = 64");
    const int smem = (kIcHalo * kIcHalo + 49 * 4 * 16) * sizeof(float4);
    dim3 grid((W + kIcTile - 1) / kIcTile, (H + kIcTile - 1) / kIcTile, B);
    init_conv7_kernel<<<grid, 256, smem, s>>>(reinterpret_cast<const float4*>(x), w, bias, out, H, W);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int upsample2x_launch(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s) {
    const int cv = C / 8;
    const size_t total = static_cast<size_t>(B) * 4 * H * W * cv;
    NDIFF_CUDA_OK(launch_pdl(upsample2x_kernel, dim3(blocks_for(total, 256 * 4, 148 * 16)), dim3(256), 0, s,
                             reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), H, W, cv, total));
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int final_launch(const FinalArgs& a, cudaStream_t s) {
    NDIFF_REQUIRE(a.C == 64 || a.C == 128 || a.C == 256, "final heads: C/8 must be a power of two <= 32");
    if (a.C == 64 && a.HW % 32 == 0) {
        // one wave of 256-thread blocks, every warp walks a contiguous range of 32-pixel chunks
        const int occ_q = g_final_occ[a.gn_stats ? 1 : 0][a.sn ? 0 : 1];
        const int occ = occ_q > 0 ? occ_q : 1;
        const int blocks = blocks_for(static_cast<size_t>(a.npix) / 32, 8, g_num_sms * occ);
        NDIFF_REQUIRE(a.sn || (a.sf && a.ws && a.bs), "final heads: neither a shot feature map nor a precomputed shot head");
        if (a.gn_stats) {
            NDIFF_REQUIRE(a.gn_gamma && a.gn_beta && a.gn_G > 0 && 64 % a.gn_G == 0 && (64 / a.gn_G) % 8 == 0,
                          "final heads: bad fused GroupNorm arguments");
            if (a.sn) NDIFF_CUDA_OK(launch_pdl(final64_kernel<true, false>, dim3(blocks), dim3(256), 0, s, a));
            else NDIFF_CUDA_OK(launch_pdl(final64_kernel<true, true>, dim3(blocks), dim3(256), 0, s, a));
        } else {
            if (a.sn) NDIFF_CUDA_OK(launch_pdl(final64_kernel<false, false>, dim3(blocks), dim3(256), 0, s, a));
            else NDIFF_CUDA_OK(launch_pdl(final64_kernel<false, true>, dim3(blocks), dim3(256), 0, s, a));
        }
    } else {
        NDIFF_REQUIRE(!a.gn_stats && !a.sn, "final heads: the fused forms need dim = 64");
        const size_t threads = static_cast<size_t>(a.npix) * (a.C / 8);
        NDIFF_CUDA_OK(launch_pdl(final_kernel, dim3(blocks_for(threads, 256, g_num_sms * 4)), dim3(256), 8 * a.C * sizeof(float), s, a));
    }
    NDIFF_CUDA_OK(cudaGetLastError());
    if (a.chain) {
        NDIFF_CUDA_OK(launch_pdl(chain_advance_kernel, dim3(1), dim3(1), 0, s, a.chain));
        NDIFF_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int time_mlp_launch(const int* t, int t_stride, int n, int dim, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* st_out, cudaStream_t s) {
    NDIFF_REQUIRE(dim <= 128 && dim % 2 == 0, "time embedding: dim must be even and <= 128");
    time_mlp_kernel<<<n, 256, 0, s>>>(t, t_stride, dim, w1, b1, w2, b2, st_out);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int rows_gemv_launch(const float* W, const float* bias, const float* in, float* out, int n, int rows, int K,
                     cudaStream_t s) {
    NDIFF_REQUIRE(K % 4 == 0, "gemv: K must be a multiple of 4 (float4 rows)");
    dim3 grid((rows + 7) / 8, n);
    rows_gemv_kernel<<<grid, 256, 0, s>>>(W, bias, in, out, rows, K);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int iso_vec_launch(const float* emb_table, const long long* idx, const float* wv, const float* wo, const float* bo,
                   float* out, int out_ld, int out_off, int B, int C, cudaStream_t s) {
    iso_vec_kernel<<<B, 128, 0, s>>>(emb_table, idx, wv, wo, bo, out, out_ld, out_off, C);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pos_maps_launch(const PosArgs& a, cudaStream_t s) {
    const int C2 = 2 * a.C;
    const int smem = (2 * C2 * 8 + 2 * C2 + 560) * sizeof(float);
    const size_t npix = static_cast<size_t>(a.B) * a.HW;
    pos_maps_kernel<<<blocks_for(npix, 128), 128, smem, s>>>(a);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int nchw_to_nhwc4_launch(const float* in, float* out, int B, int HW, cudaStream_t s) {
    const size_t npix = static_cast<size_t>(B) * HW;
    nchw_to_nhwc4_kernel<<<blocks_for(npix, 256), 256, 0, s>>>(in, reinterpret_cast<float4*>(out), HW, npix);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int nhwc4_to_nchw_launch(const float* in, float* out, int B, int HW, cudaStream_t s) {
    const size_t npix = static_cast<size_t>(B) * HW;
    nhwc4_to_nchw_kernel<<<blocks_for(npix, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(in), out, HW, npix);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int compose_noisy_launch(const float* noise, const float* clean, float* noisy_out, float* clean_out, size_t n4, cudaStream_t s) {
    NDIFF_REQUIRE(noise && clean && noisy_out, "compose: null argument");
    NDIFF_REQUIRE(((reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(clean) | reinterpret_cast<uintptr_t>(noisy_out) |
                    reinterpret_cast<uintptr_t>(clean_out)) & 15) == 0, "compose: buffers must be 16-byte aligned");
    compose_noisy_kernel<<<blocks_for(n4, 256 * 4, 148 * 16), 256, 0, s>>>(reinterpret_cast<const float4*>(noise), reinterpret_cast<const float4*>(clean),
                                                                          reinterpret_cast<float4*>(noisy_out), reinterpret_cast<float4*>(clean_out), n4);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int philox_normal_launch(float* out, size_t n4, unsigned long long seed, unsigned long long stream_id, cudaStream_t s) {
    philox_normal_kernel<<<blocks_for(n4, 256 * 4, 148 * 16), 256, 0, s>>>(reinterpret_cast<float4*>(out), n4, seed,
                                                                           stream_id);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ndiff
