// noisediff_b200 — backward / optimizer kernels of the diffusion training step (sm_100a).  See train_ops.cuh.
// These are the HBM-bound halves of the backward pass (normalisation / activation / routing kernels and the small dense
// paths); the convolution gradients run on the tensor cores (dgrad = conv_gemm with flipped weights, wgrad = wgrad_gemm.cu).
#include "train_ops.cuh"

namespace ndiff {

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
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float y) {          // d/dy [y * sigmoid(y)] = s (1 + y (1 - s)), s = 0.5 + 0.5 tanh(y / 2)
    const float s = fmaf(0.5f, tanh_approx(0.5f * y), 0.5f);   // one MUFU op (the exp + reciprocal form needs two)
    return s * fmaf(y, 1.0f - s, 1.0f);
}
__device__ __forceinline__ float gelu_grad(float x) {          // d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    return cdf + x * 0.39894228040143267794f * __expf(-0.5f * x * x);
}
inline int blocks_for(size_t n, int per_block, int cap = 1 << 20) {
    size_t b = (n + per_block - 1) / per_block;
    return static_cast<int>(b < static_cast<size_t>(cap) ? (b ? b : 1) : cap);
}

constexpr int kT = 256;

// ---------------------------------------------------------------------------------------------------------------
// column sums.  grid = (blocks per sample, B); a thread owns one 8-channel vector, lanes with the same vector walk pixels
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) colsum_kernel(const bf16* __restrict__ x, int x_ld, int x_off, float* __restrict__ out,
                                                    int out_ld, int per_sample, int HW, int C) {
    __shared__ float red[kT * 8];
    const int b = blockIdx.y, cv = C >> 3, cvi = threadIdx.x % cv, lane_p = threadIdx.x / cv, ppb = kT / cv;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bf16* base = x + static_cast<size_t>(b) * HW * x_ld + x_off + cvi * 8;
    for (int p = blockIdx.x * ppb + lane_p; p < HW; p += gridDim.x * ppb) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(p) * x_ld)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kT) {
        float t = 0.f;
        for (int l = 0; l < ppb; ++l) t += red[(l * cv + (c >> 3)) * 8 + (c & 7)];
        atomicAdd(out + (per_sample ? static_cast<size_t>(b) * out_ld : 0) + c, t);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm backward.  Shared prologue: per-thread folded coefficients of its 8 channels.
//   xhat = h * rs + nm (rs = rstd, nm = -mean * rstd);  y1 = xhat * gamma + beta;  y2 = y1 * (sc + 1) + sh
// ---------------------------------------------------------------------------------------------------------------
struct GnCoef { float rs, nm; };
__device__ __forceinline__ GnCoef gn_coef(const GnBwdArgs& a, int b, int g, int gs) {
    const double inv_n = 1.0 / (static_cast<double>(a.HW) * gs * (a.real_frac > 0.f ? a.real_frac : 1.0f));
    const double s = static_cast<double>(static_cast<long long>(a.stats[(b * a.G + g) * 2])) * (1.0 / 16777216.0);
    const double ss = static_cast<double>(static_cast<long long>(a.stats[(b * a.G + g) * 2 + 1])) * (1.0 / 16777216.0);
    const double meand = s * inv_n;
    const float var = fmaxf(static_cast<float>(ss * inv_n - meand * meand), 0.f);
    GnCoef c;
    c.rs = rsqrtf(var + a.eps);
    c.nm = -static_cast<float>(meand) * c.rs;
    return c;
}

template <bool kMaps>
__global__ void __launch_bounds__(kT) gn_bwd_stats_kernel(const GnBwdArgs a) {
    __shared__ float red[kT * 16];
    const int b = blockIdx.y, C = a.C, cv = C >> 3, gs = C / a.G;
    const int cvi = threadIdx.x % cv, lane_p = threadIdx.x / cv, ppb = kT / cv, c0 = cvi * 8;
    const GnCoef k = gn_coef(a, b, c0 / gs, gs);
    float gam[8], bet[8], sc1[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gam[j] = __ldg(a.gamma + c0 + j); bet[j] = __ldg(a.beta + c0 + j);
        sc1[j] = 1.0f; sh[j] = 0.f;
        if (!kMaps && a.ss) {
            sc1[j] = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + c0 + j] + 1.0f;
            sh[j] = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + C + c0 + j];
        }
    }
    float u1[8], u2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { u1[j] = 0.f; u2[j] = 0.f; }
    const size_t base = static_cast<size_t>(b) * a.HW * cv + cvi;
    const uint4* hin = reinterpret_cast<const uint4*>(a.h) + base;
    const uint4* din = reinterpret_cast<const uint4*>(a.dout) + base;
    const uint4* mp = kMaps ? reinterpret_cast<const uint4*>(a.maps) + (base - cvi) * 2 : nullptr;
    uint4* dmp = kMaps ? reinterpret_cast<uint4*>(a.dmaps) + (base - cvi) * 2 : nullptr;
    // two pixels per iteration, all 16-byte loads issued before the arithmetic (the SiLU-gradient chain is long)
    const int pstride = gridDim.x * ppb;
    for (int p0 = blockIdx.x * ppb + lane_p; p0 < a.HW; p0 += 2 * pstride) {
        uint4 hq[2], dq[2], msq[2], mhq[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = p0 + i * pstride;
            if (p < a.HW) {
                const size_t o = static_cast<size_t>(p) * cv;
                hq[i] = __ldg(hin + o); dq[i] = __ldg(din + o);
                if (kMaps) { msq[i] = __ldg(mp + o * 2 + cvi); mhq[i] = __ldg(mp + o * 2 + cv + cvi); }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = p0 + i * pstride;
            if (p >= a.HW) break;
            const size_t o = static_cast<size_t>(p) * cv;
            float hv[8], dv[8], ms[8], mh[8];
            unpack8(hq[i], hv);
            unpack8(dq[i], dv);
            if (kMaps) { unpack8(msq[i], ms); unpack8(mhq[i], mh); }
            float dsc[8], dsh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = fmaf(hv[j], k.rs, k.nm);
                const float y1 = fmaf(xh, gam[j], bet[j]);
                const float s1 = kMaps ? ms[j] + 1.0f : sc1[j];
                const float y2 = fmaf(y1, s1, kMaps ? mh[j] : sh[j]);
                const float dy2 = dv[j] * silu_grad(y2);
                const float dy1 = kMaps ? dy2 * s1 : dy2;        // per-sample scale is applied after the pixel sum
                u1[j] += dy1;
                u2[j] = fmaf(dy1, xh, u2[j]);
                if (kMaps) { dsc[j] = dy2 * y1; dsh[j] = dy2; }
            }
            if (kMaps) { dmp[o * 2 + cvi] = pack8(dsc); dmp[o * 2 + cv + cvi] = pack8(dsh); }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = u1[j]; red[threadIdx.x * 16 + 8 + j] = u2[j]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kT) {
        float t1 = 0.f, t2 = 0.f;
        for (int l = 0; l < ppb; ++l) {
            t1 += red[(l * cv + (c >> 3)) * 16 + (c & 7)];
            t2 += red[(l * cv + (c >> 3)) * 16 + 8 + (c & 7)];
        }
        float* dst = a.acc + (static_cast<size_t>(b) * C + c) * 2;
        atomicAdd(dst, t1);
        atomicAdd(dst + 1, t2);
    }
}

// parameter gradients + group sums from acc.  One block per sample; thread = channel.
// acc holds (U1, U2) = sums of dy2 (vector scale/shift: scale applied here) or of dy1 (maps).
// gsum[b][g][2] = sum_{c in g} gamma_c * (sc_c + 1) * (U1, U2)   (= sum dxhat, sum dxhat * xhat over the group)
template <bool kMaps>
__global__ void gn_bwd_param_kernel(const GnBwdArgs a, float* __restrict__ gsum) {
    const int b = blockIdx.x, C = a.C, gs = C / a.G;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float u1 = a.acc[(static_cast<size_t>(b) * C + c) * 2], u2 = a.acc[(static_cast<size_t>(b) * C + c) * 2 + 1];
        const float gam = a.gamma[c], bet = a.beta[c];
        float s1 = 1.0f;
        if (!kMaps && a.ss) s1 = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + c] + 1.0f;
        atomicAdd(a.dbeta + c, s1 * u1);
        atomicAdd(a.dgamma + c, s1 * u2);
        if (!kMaps && a.dss) {
            a.dss[static_cast<size_t>(b) * a.dss_ld + a.ss_off + c] = fmaf(gam, u2, bet * u1);      // d scale = sum dy2 * y1
            a.dss[static_cast<size_t>(b) * a.dss_ld + a.ss_off + C + c] = u1;                          // d shift = sum dy2
        }
        atomicAdd(gsum + (static_cast<size_t>(b) * a.G + c / gs) * 2, gam * s1 * u1);
        atomicAdd(gsum + (static_cast<size_t>(b) * a.G + c / gs) * 2 + 1, gam * s1 * u2);
    }
}

template <bool kMaps>
__global__ void __launch_bounds__(kT) gn_bwd_apply_kernel(const GnBwdArgs a, const float* __restrict__ gsum) {
    __shared__ float red[kT * 8];
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int b = blockIdx.y, C = a.C, cv = C >> 3, gs = C / a.G;
    const int cvi = threadIdx.x % cv, lane_p = threadIdx.x / cv, ppb = kT / cv, c0 = cvi * 8, g = c0 / gs;
    const GnCoef k = gn_coef(a, b, g, gs);
    const float inv_n = 1.0f / (static_cast<float>(a.HW) * gs * (a.real_frac > 0.f ? a.real_frac : 1.0f));
    const float m1 = gsum[(static_cast<size_t>(b) * a.G + g) * 2] * inv_n, m2 = gsum[(static_cast<size_t>(b) * a.G + g) * 2 + 1] * inv_n;
    float gam[8], bet[8], sc1[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gam[j] = __ldg(a.gamma + c0 + j); bet[j] = __ldg(a.beta + c0 + j);
        sc1[j] = 1.0f; sh[j] = 0.f;
        if (!kMaps && a.ss) {
            sc1[j] = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + c0 + j] + 1.0f;
            sh[j] = a.ss[static_cast<size_t>(b) * a.ss_ld + a.ss_off + C + c0 + j];
        }
    }
    const size_t base = static_cast<size_t>(b) * a.HW * cv + cvi;
    const uint4* hin = reinterpret_cast<const uint4*>(a.h) + base;
    const uint4* din = reinterpret_cast<const uint4*>(a.dout) + base;
    uint4* dh = reinterpret_cast<uint4*>(a.dh) + base;
    const uint4* mp = kMaps ? reinterpret_cast<const uint4*>(a.maps) + (base - cvi) * 2 : nullptr;
    const int pstride = gridDim.x * ppb;
    for (int p0 = blockIdx.x * ppb + lane_p; p0 < a.HW; p0 += 2 * pstride) {
        uint4 hq[2], dq[2], msq[2], mhq[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = p0 + i * pstride;
            if (p < a.HW) {
                const size_t o = static_cast<size_t>(p) * cv;
                hq[i] = __ldg(hin + o); dq[i] = __ldg(din + o);
                if (kMaps) { msq[i] = __ldg(mp + o * 2 + cvi); mhq[i] = __ldg(mp + o * 2 + cv + cvi); }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p = p0 + i * pstride;
            if (p >= a.HW) break;
            float hv[8], dv[8], ms[8], mh[8], r[8];
            unpack8(hq[i], hv);
            unpack8(dq[i], dv);
            if (kMaps) { unpack8(msq[i], ms); unpack8(mhq[i], mh); }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = fmaf(hv[j], k.rs, k.nm);
                const float y1 = fmaf(xh, gam[j], bet[j]);
                const float s1 = kMaps ? ms[j] + 1.0f : sc1[j];
                const float y2 = fmaf(y1, s1, kMaps ? mh[j] : sh[j]);
                const float dxh = dv[j] * silu_grad(y2) * s1 * gam[j];
                r[j] = k.rs * (dxh - m1 - xh * m2);
                bsum[j] += r[j];
            }
            dh[static_cast<size_t>(p) * cv] = pack8(r);
        }
    }
    if (a.dbias) {      // h = conv + bias: the bias gradient is the column sum of dh
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = bsum[j];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += kT) {
            float t = 0.f;
            for (int l = 0; l < ppb; ++l) t += red[(l * cv + (c >> 3)) * 8 + (c & 7)];
            atomicAdd(a.dbias + c, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward.  Same lane layout as layernorm_kernel: L lanes share a pixel, each lane owns VPL 8-channel vectors.
// ---------------------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(kT) layernorm_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ vec, int vec_ld,
                                                           const float* __restrict__ g, const bf16* __restrict__ du,
                                                           bf16* __restrict__ dy_out, float* __restrict__ dg,
                                                           float* __restrict__ dbeta, int HW, int C, int L, size_t npix,
                                                           float real_frac) {
    extern __shared__ float sacc[];      // [2][C]: dg, dbeta of this block
    for (int i = threadIdx.x; i < 2 * C; i += kT) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sub = lane % L, slot = lane / L, ppw = 32 / L;
    const size_t warp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const int cv = C >> 3;
    float gg[VPL][8], ag[VPL][8], ab[VPL][8];
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) { gg[j][k] = g[(sub + j * L) * 8 + k]; ag[j][k] = 0.f; ab[j][k] = 0.f; }
    const float inv_c = 1.0f / (static_cast<float>(C) * real_frac);
    const float n_pad = static_cast<float>(C) * (1.0f - real_frac);
    for (size_t p0 = warp * ppw; p0 < npix; p0 += nwarps * ppw) {
        const size_t pix = p0 + slot;
        const bool live = pix < npix;
        const size_t pp = live ? pix : npix - 1;
        const float* vp = vec + (pp / HW) * static_cast<size_t>(vec_ld);
        float f[VPL][8], d[VPL][8];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int cvi = sub + j * L;
            unpack8(__ldg(reinterpret_cast<const uint4*>(x) + pp * cv + cvi), f[j]);
            unpack8(__ldg(reinterpret_cast<const uint4*>(du) + pp * cv + cvi), d[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) { f[j][k] += __ldg(vp + cvi * 8 + k); sum += f[j][k]; }
        }
        for (int o = L >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * inv_c;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float t = f[j][k] - mean; sq += t * t; }
        for (int o = L >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(fmaxf(sq - n_pad * mean * mean, 0.f) * inv_c + 1e-5f);
        // dxhat = du * g (zero on padded channels: their g is 0);  dy = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = (f[j][k] - mean) * rstd;
                const float dxh = d[j][k] * gg[j][k];
                s1 += dxh; s2 = fmaf(dxh, xh, s2);
                if (live) { ag[j][k] = fmaf(d[j][k], xh, ag[j][k]); ab[j][k] += d[j][k]; }
                f[j][k] = xh; d[j][k] = dxh;
            }
        for (int o = L >> 1; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        s1 *= inv_c; s2 *= inv_c;
        if (live) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                float r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = (gg[j][k] != 0.f || real_frac >= 1.0f) ? rstd * (d[j][k] - s1 - f[j][k] * s2) : 0.f;
                reinterpret_cast<uint4*>(dy_out)[pix * cv + sub + j * L] = pack8(r);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            atomicAdd(&sacc[(sub + j * L) * 8 + k], ag[j][k]);
            atomicAdd(&sacc[C + (sub + j * L) * 8 + k], ab[j][k]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += kT) { atomicAdd(dg + i, sacc[i]); atomicAdd(dbeta + i, sacc[C + i]); }
}

__global__ void __launch_bounds__(kT) gelu_bwd_kernel(const uint4* __restrict__ pre, const uint4* __restrict__ dy,
                                                      uint4* __restrict__ dpre, size_t nv) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float a[8], d[8];
        unpack8(__ldg(pre + i), a);
        unpack8(__ldg(dy + i), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= gelu_grad(a[j]);
        dpre[i] = pack8(d);
    }
}
__global__ void __launch_bounds__(kT) gelu_fwd_kernel(const uint4* __restrict__ pre, uint4* __restrict__ y, size_t nv) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float a[8];
        unpack8(__ldg(pre + i), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = gelu_exact(a[j]);
        y[i] = pack8(a);
    }
}

__global__ void __launch_bounds__(kT) add_slice_kernel(bf16* __restrict__ dst, int dst_ld, int dst_off, const bf16* __restrict__ src,
                                                       int src_ld, int src_off, int cv, size_t total, int accumulate) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t p = i / cv;
        const int c = static_cast<int>(i % cv) * 8;
        uint4* dp = reinterpret_cast<uint4*>(dst + p * dst_ld + dst_off + c);
        float s[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(src + p * src_ld + src_off + c)), s);
        if (accumulate) {
            float d[8];
            unpack8(*dp, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += d[j];
        }
        *dp = pack8(s);
    }
}

__global__ void __launch_bounds__(kT) upsample2x_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int H, int W,
                                                            int cv, size_t total, int accumulate) {
    // H, W = low-resolution size; dy is [B, 2H, 2W, C]
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(v % cv);
        size_t p = v / cv;
        const int xx = static_cast<int>(p % W); p /= W;
        const int yy = static_cast<int>(p % H);
        const size_t b = p / H;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float f[8];
                unpack8(__ldg(dy + ((b * 2 * H + 2 * yy + i) * (2 * W) + 2 * xx + j) * cv + c), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        if (accumulate) {
            float d[8];
            unpack8(dx[v], d);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += d[k];
        }
        dx[v] = pack8(acc);
    }
}

__global__ void __launch_bounds__(kT) depth_to_space_kernel(const uint4* __restrict__ t, uint4* __restrict__ dx, int H, int W,
                                                            int cv, size_t total, int accumulate) {
    // t: [B, H, W, 4, C] (tap = p1 * 2 + p2 major, then channel);  dx: [B, 2H, 2W, C];  total = B * 4HW * cv
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(v % cv);
        size_t p = v / cv;
        const int ox = static_cast<int>(p % (2 * W)); p /= (2 * W);
        const int oy = static_cast<int>(p % (2 * H));
        const size_t b = p / (2 * H);
        const int tap = (oy & 1) * 2 + (ox & 1);
        float f[8];
        unpack8(__ldg(t + (((b * H + (oy >> 1)) * W + (ox >> 1)) * 4 + tap) * cv + c), f);
        if (accumulate) {
            float d[8];
            unpack8(dx[v], d);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += d[k];
        }
        dx[v] = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// heads + loss.  8 lanes per pixel (C = 64: lane = channel octet).  Parameter gradients: registers -> shared -> global.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) heads_bwd_kernel(const HeadsBwdArgs a) {
    __shared__ float sw[2][4][64];
    __shared__ float sacc[2][4][64];
    __shared__ float sb[4];
    __shared__ double sloss;
    for (int i = threadIdx.x; i < 256; i += kT) {
        sw[0][i >> 6][i & 63] = a.wf[i]; sw[1][i >> 6][i & 63] = a.ws[i];
        sacc[0][i >> 6][i & 63] = 0.f; sacc[1][i >> 6][i & 63] = 0.f;
    }
    if (threadIdx.x < 4) sb[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) sloss = 0.0;
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const size_t npix = static_cast<size_t>(a.B) * a.HW;
    const float scale = 2.0f / (4.0f * static_cast<float>(a.HW) * static_cast<float>(a.B));
    float gwf[4][8], gws[4][8], gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) { gwf[k][j] = 0.f; gws[k][j] = 0.f; }
    double lacc = 0.0;
    for (size_t pix = (static_cast<size_t>(blockIdx.x) * kT + threadIdx.x) >> 3; pix < npix; pix += (static_cast<size_t>(gridDim.x) * kT) >> 3) {
        const float4 v = __ldg(a.v + pix), t = __ldg(a.target + pix);
        const float wb = __ldg(a.w_b + pix / a.HW);
        const float e[4] = {v.x - t.x, v.y - t.y, v.z - t.z, v.w - t.w};
        float dv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dv[k] = scale * wb * e[k];
        if (sub == 0) {
            lacc += static_cast<double>(wb) * (static_cast<double>(e[0]) * e[0] + static_cast<double>(e[1]) * e[1] +
                                               static_cast<double>(e[2]) * e[2] + static_cast<double>(e[3]) * e[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) gb[k] += dv[k];
        }
        float xf[8], sf[8], dxf[8], dsf[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(a.xf) + pix * 8 + sub), xf);
        unpack8(__ldg(reinterpret_cast<const uint4*>(a.sf) + pix * 8 + sub), sf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s0 = fmaf(dv[k], sw[0][k][sub * 8 + j], s0);
                s1 = fmaf(dv[k], sw[1][k][sub * 8 + j], s1);
                gwf[k][j] = fmaf(dv[k], xf[j], gwf[k][j]);
                gws[k][j] = fmaf(dv[k], sf[j], gws[k][j]);
            }
            dxf[j] = s0; dsf[j] = s1;
        }
        reinterpret_cast<uint4*>(a.dxf)[pix * 8 + sub] = pack8(dxf);
        reinterpret_cast<uint4*>(a.dsf)[pix * 8 + sub] = pack8(dsf);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sacc[0][k][sub * 8 + j], gwf[k][j]);
            atomicAdd(&sacc[1][k][sub * 8 + j], gws[k][j]);
        }
        if (sub == 0) atomicAdd(&sb[k], gb[k]);
    }
    if (sub == 0) atomicAdd(&sloss, lacc);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kT) {
        atomicAdd(a.dwf + i, sacc[0][i >> 6][i & 63]);
        atomicAdd(a.dws + i, sacc[1][i >> 6][i & 63]);
    }
    if (threadIdx.x < 4) { atomicAdd(a.dbf + threadIdx.x, sb[threadIdx.x]); atomicAdd(a.dbs + threadIdx.x, sb[threadIdx.x]); }
    if (threadIdx.x == 0) atomicAdd(a.loss, sloss / (4.0 * a.HW * a.B));
}

// ---------------------------------------------------------------------------------------------------------------
// shot_mlp1.fc1 parameter gradients.  thread = (pixel lane, 8-channel group); pre-activation recomputed.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) shot_in_bwd_kernel(const float4* __restrict__ clean, const float4* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const bf16* __restrict__ ds0, float* __restrict__ dw, float* __restrict__ db,
                                                         size_t npix, int C) {
    extern __shared__ float sm[];       // w [C][8] | bias [C] | acc [C][9]
    float* sw = sm; float* sbias = sm + C * 8; float* acc = sbias + C;
    for (int i = threadIdx.x; i < C * 8; i += kT) sw[i] = w[i];
    for (int i = threadIdx.x; i < C; i += kT) sbias[i] = bias[i];
    for (int i = threadIdx.x; i < C * 9; i += kT) acc[i] = 0.f;
    __syncthreads();
    const int cv = C >> 3, cg = threadIdx.x % cv;
    float gw[8][8], gb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { gb[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) gw[j][k] = 0.f; }
    const size_t ppb = kT / cv;
    for (size_t pix = static_cast<size_t>(blockIdx.x) * ppb + threadIdx.x / cv; pix < npix; pix += static_cast<size_t>(gridDim.x) * ppb) {
        const float4 a4 = __ldg(clean + pix), b4 = __ldg(x + pix);
        const float in[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
        float d[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(ds0) + pix * cv + cg), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg * 8 + j;
            float pre = sbias[c];
#pragma unroll
            for (int k = 0; k < 8; ++k) pre = fmaf(sw[c * 8 + k], in[k], pre);
            const float dp = d[j] * gelu_grad(pre);
            gb[j] += dp;
#pragma unroll
            for (int k = 0; k < 8; ++k) gw[j][k] = fmaf(dp, in[k], gw[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        atomicAdd(&acc[c * 9 + 8], gb[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&acc[c * 9 + k], gw[j][k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 8; i += kT) atomicAdd(dw + i, acc[(i >> 3) * 9 + (i & 7)]);
    for (int i = threadIdx.x; i < C; i += kT) atomicAdd(db + i, acc[i * 9 + 8]);
}

// ---------------------------------------------------------------------------------------------------------------
// init_conv 7x7 weight gradient.  Persistent blocks over 16x16 pixel tiles; thread = (output channel, group of taps).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kIwTile = 16, kIwHalo = kIwTile + 6;
__global__ void __launch_bounds__(kT) init_conv_wgrad_kernel(const float4* __restrict__ x, const bf16* __restrict__ dy,
                                                             float* __restrict__ dw, int B, int H, int W) {
    __shared__ float4 sx[kIwHalo * kIwHalo];
    __shared__ __align__(16) bf16 sdy[kIwTile * kIwTile][64 + 8];      // +8: rows start in different banks
    const int co = threadIdx.x & 63, tg = threadIdx.x >> 6;              // tap group tg handles taps tg, tg + 4, ... (13 / 12 taps)
    float acc[13][4];
#pragma unroll
    for (int i = 0; i < 13; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
    const int tiles_x = (W + kIwTile - 1) / kIwTile, tiles_y = (H + kIwTile - 1) / kIwTile;
    const int n_tiles = B * tiles_y * tiles_x;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int x0 = tx * kIwTile, y0 = ty * kIwTile;
        __syncthreads();
        for (int i = threadIdx.x; i < kIwHalo * kIwHalo; i += kT) {
            const int yy = y0 + i / kIwHalo - 3, xx = x0 + i % kIwHalo - 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(x + (static_cast<size_t>(b) * H + yy) * W + xx);
            sx[i] = v;
        }
        for (int i = threadIdx.x; i < kIwTile * kIwTile * 8; i += kT) {
            const int p = i >> 3, j = i & 7;
            const int yy = y0 + p / kIwTile, xx = x0 + p % kIwTile;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (yy < H && xx < W) v = __ldg(reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(b) * H + yy) * W + xx) * 64) + j);
            *reinterpret_cast<uint4*>(&sdy[p][j * 8]) = v;
        }
        __syncthreads();
        for (int p = 0; p < kIwTile * kIwTile; ++p) {
            const float g = __bfloat162float(sdy[p][co]);
            const int py = p / kIwTile, px = p % kIwTile;
#pragma unroll
            for (int i = 0; i < 13; ++i) {
                const int tap = tg + 4 * i;
                if (tap < 49) {
                    const float4 xi = sx[(py + tap / 7) * kIwHalo + px + tap % 7];
                    acc[i][0] = fmaf(g, xi.x, acc[i][0]); acc[i][1] = fmaf(g, xi.y, acc[i][1]);
                    acc[i][2] = fmaf(g, xi.z, acc[i][2]); acc[i][3] = fmaf(g, xi.w, acc[i][3]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int tap = tg + 4 * i;
        if (tap < 49) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) atomicAdd(dw + (co * 4 + ci) * 49 + tap, acc[i][ci]);     // [co][ci][7][7]
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// time path
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) time_mlp_train_kernel(const int* __restrict__ t, int dim, const float* __restrict__ w1,
                                                            const float* __restrict__ b1, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, float* __restrict__ st_out,
                                                            float* __restrict__ saved) {
    __shared__ float emb[128], h[512];
    const int n = blockIdx.x, td = dim * 4, half = dim / 2;
    float* sv = saved + static_cast<size_t>(n) * (dim + 2 * td);
    const float tv = static_cast<float>(t[n]);
    if (threadIdx.x < half) {
        const float f = expf(static_cast<float>(threadIdx.x) * -(logf(10000.0f) / static_cast<float>(half - 1)));
        emb[threadIdx.x] = sinf(tv * f);
        emb[half + threadIdx.x] = cosf(tv * f);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < dim; k += kT) sv[k] = emb[k];
    for (int o = threadIdx.x; o < td; o += kT) {
        float acc = b1[o];
        for (int k = 0; k < dim; ++k) acc += w1[o * dim + k] * emb[k];
        sv[dim + o] = acc;
        h[o] = gelu_exact(acc);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < td; o += kT) {
        float acc = b2[o];
        for (int k = 0; k < td; ++k) acc += w2[o * td + k] * h[k];
        sv[dim + td + o] = acc;
        st_out[static_cast<size_t>(n) * td + o] = acc / (1.0f + expf(-acc));
    }
}

// one block per sample: dst [td] -> da2 -> (dw2, db2, dh) -> da1 -> (dw1, db1)
__global__ void __launch_bounds__(kT) time_mlp_bwd_kernel(const float* __restrict__ dst, const float* __restrict__ saved, int dim,
                                                          const float* __restrict__ w1, const float* __restrict__ w2,
                                                          float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                                          float* __restrict__ db2) {
    __shared__ float emb[128], hh[512], da2[512], da1[512];
    const int n = blockIdx.x, td = dim * 4;
    const float* sv = saved + static_cast<size_t>(n) * (dim + 2 * td);
    for (int k = threadIdx.x; k < dim; k += kT) emb[k] = sv[k];
    for (int o = threadIdx.x; o < td; o += kT) {
        hh[o] = gelu_exact(sv[dim + o]);
        da2[o] = dst[static_cast<size_t>(n) * td + o] * silu_grad(sv[dim + td + o]);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < td; o += kT) atomicAdd(db2 + o, da2[o]);
    for (int i = threadIdx.x; i < td * td; i += kT) atomicAdd(dw2 + i, da2[i / td] * hh[i % td]);
    for (int k = threadIdx.x; k < td; k += kT) {
        float acc = 0.f;
        for (int o = 0; o < td; ++o) acc = fmaf(w2[o * td + k], da2[o], acc);
        da1[k] = acc * gelu_grad(sv[dim + k]);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < td; o += kT) atomicAdd(db1 + o, da1[o]);
    for (int i = threadIdx.x; i < td * dim; i += kT) atomicAdd(dw1 + i, da1[i / dim] * emb[i % dim]);
}

__global__ void __launch_bounds__(kT) small_gemm_kernel(int tA, int tB, int M, int N, int K, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ Bm, int ldb, float* __restrict__ C, int ldc,
                                                        int accumulate, int kchunk) {
    // blockIdx.y = K split (partials meet in atomics; C was zeroed by the launcher unless it accumulates anyway)
    const size_t total = static_cast<size_t>(M) * N;
    const int k0 = blockIdx.y * kchunk, k1 = k0 + kchunk < K ? k0 + kchunk : K;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int m = static_cast<int>(i / N), n = static_cast<int>(i % N);
        float acc = 0.f;
        for (int k = k0; k < k1; ++k) {
            const float av = tA ? A[static_cast<size_t>(k) * lda + m] : A[static_cast<size_t>(m) * lda + k];
            const float bv = tB ? Bm[static_cast<size_t>(n) * ldb + k] : Bm[static_cast<size_t>(k) * ldb + n];
            acc = fmaf(av, bv, acc);
        }
        float* c = C + static_cast<size_t>(m) * ldc + n;
        if (gridDim.y > 1) atomicAdd(c, acc);
        else *c = accumulate ? *c + acc : acc;
    }
}

__global__ void rowsum_f32_kernel(const float* __restrict__ in, int ld, int n, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += in[static_cast<size_t>(i) * ld + c];
    out[c] += acc;
}

// one block (128 threads) per sample: c = wo @ (wv @ e) + bo
__global__ void __launch_bounds__(128) iso_vec_bwd_kernel(const float* __restrict__ emb_table, const long long* __restrict__ idx,
                                                          const float* __restrict__ wv, const float* __restrict__ wo,
                                                          const float* __restrict__ dc, int dc_ld, int dc_off,
                                                          float* __restrict__ demb, float* __restrict__ dwv, float* __restrict__ dwo,
                                                          float* __restrict__ dbo, int C) {
    __shared__ float v[128], dv[128], e[16], sdc[512];
    const int b = blockIdx.x;
    const long long row = idx[b];
    if (threadIdx.x < 16) e[threadIdx.x] = emb_table[row * 16 + threadIdx.x];
    for (int c = threadIdx.x; c < C; c += 128) sdc[c] = dc[static_cast<size_t>(b) * dc_ld + dc_off + c];
    __syncthreads();
    {
        float acc = 0.f;
        for (int k = 0; k < 16; ++k) acc += wv[threadIdx.x * 16 + k] * e[k];
        v[threadIdx.x] = acc;
        float d = 0.f;
        for (int c = 0; c < C; ++c) d = fmaf(wo[c * 128 + threadIdx.x], sdc[c], d);
        dv[threadIdx.x] = d;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 128) atomicAdd(dbo + c, sdc[c]);
    for (int i = threadIdx.x; i < C * 128; i += 128) atomicAdd(dwo + i, sdc[i >> 7] * v[i & 127]);
    for (int i = threadIdx.x; i < 128 * 16; i += 128) atomicAdd(dwv + i, dv[i >> 4] * e[i & 15]);
    if (threadIdx.x < 16) {
        float de = 0.f;
        for (int k = 0; k < 128; ++k) de = fmaf(wv[k * 16 + threadIdx.x], dv[k], de);
        atomicAdd(demb + row * 16 + threadIdx.x, de);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// positional path backward.  Block = 256 threads working on chunks of 256 pixels, four phases per chunk:
//   1  (thread = pixel)        forward recompute; ps = SiLU(pos_emb) -> shared memory
//   2  (thread = map channel)  2 x 2C = 256 channels of the two ResnetBlock2 heads: dwm[c][k] += dmap[p][c] * ps[p][k] in registers
//   3a (thread = pixel)        dps = Wm^T dmap, then back through SiLU / fc2 / GELU / fc1 / sin-cos features: the per-pixel
//                              vectors every parameter gradient is an outer product of go to shared memory (row V[p])
//   3b (thread = parameter)    each of the 560 small parameters sums its products V[p][ia] * V[p][ib] over the chunk's pixels in
//                              a register that lives for the whole kernel (a first version reduced every product over the warp
//                              with shuffles: 6.9 ms at B = 32; the parameter-major form needs two shared loads per product)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPosSmall = 16 + 8 + 384 + 16 + 128 + 8;      // we, be, w1, b1, w2, b2 (offsets 0, 16, 24, 408, 424, 552)
constexpr int kPosRow = 75;                                  // V row: dpe 8 | hact 16 | da 16 | feat 24 | dw 8 | p0 | p1 (+1: odd pitch)
__device__ __forceinline__ void pos_param_operands(int q, int& ia, int& ib) {
    if (q < 16) { ia = 64 + (q >> 1); ib = 72 + (q & 1); }                       // pos_enc weight [8][2]: dw[j] * p{0,1}
    else if (q < 24) { ia = 64 + (q - 16); ib = -1; }                            // pos_enc bias
    else if (q < 408) { const int i = q - 24; ia = 24 + i / 24; ib = 40 + i % 24; }   // fc1 weight [16][24]: da[o] * feat[k]
    else if (q < 424) { ia = 24 + (q - 408); ib = -1; }                          // fc1 bias
    else if (q < 552) { const int i = q - 424; ia = i / 16; ib = 8 + i % 16; }    // fc2 weight [8][16]: dpe[o] * hact[k]
    else { ia = q - 552; ib = -1; }                                              // fc2 bias
}
__global__ void __launch_bounds__(kT) pos_bwd_kernel(const PosBwdArgs a) {
    extern __shared__ float sm[];
    const int C2 = 2 * a.fwd.C;              // map channels per block (128)
    float* wm = sm;                          // [2][C2][8]
    float* small = wm + 2 * C2 * 8;          // forward copies of the small weights (layout of pos_maps_kernel)
    float* ps_s = small + kPosSmall;         // [256][8]
    float* V = ps_s + kT * 8;                // [256][kPosRow]
    for (int i = threadIdx.x; i < C2 * 8; i += kT) { wm[i] = a.fwd.wm1[i]; wm[C2 * 8 + i] = a.fwd.wm2[i]; }
    for (int i = threadIdx.x; i < 16; i += kT) small[i] = a.fwd.we[i];
    for (int i = threadIdx.x; i < 8; i += kT) small[16 + i] = a.fwd.be[i];
    for (int i = threadIdx.x; i < 384; i += kT) small[24 + i] = a.fwd.w1[i];
    for (int i = threadIdx.x; i < 16; i += kT) small[408 + i] = a.fwd.b1[i];
    for (int i = threadIdx.x; i < 128; i += kT) small[424 + i] = a.fwd.w2[i];
    for (int i = threadIdx.x; i < 8; i += kT) small[552 + i] = a.fwd.b2[i];
    __syncthreads();
    const size_t npix = static_cast<size_t>(a.fwd.B) * a.fwd.HW;
    // phase-2 ownership: thread -> (which block, map channel); requires 2 * C2 == 256
    const int which = threadIdx.x / C2, mc = threadIdx.x % C2;
    const bf16* dmap_mine = which ? a.dmap2 : a.dmap1;
    float gwm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gbm = 0.f;
    // phase-3b ownership: parameters threadIdx.x, + 256, + 512
    int pia[3], pib[3];
    float pacc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int q = threadIdx.x + r * kT;
        pia[r] = 0; pib[r] = -1;
        if (q < kPosSmall) pos_param_operands(q, pia[r], pib[r]);
    }
    for (size_t base = static_cast<size_t>(blockIdx.x) * kT; base < npix; base += static_cast<size_t>(gridDim.x) * kT) {
        const size_t pix = base + threadIdx.x;
        const bool live = pix < npix;
        const size_t n_here = npix - base < kT ? npix - base : kT;
        // ---- phase 1: forward recompute for my pixel
        float feat[24], pre1[16], hact[16], pe[8], p0, p1;
        {
            const size_t pp = live ? pix : npix - 1;
            const size_t b = pp / a.fwd.HW, hw = pp % a.fwd.HW;
            p0 = a.fwd.position[(b * 2 + 0) * a.fwd.HW + hw]; p1 = a.fwd.position[(b * 2 + 1) * a.fwd.HW + hw];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float w = small[j * 2] * p0 + small[j * 2 + 1] * p1 + small[16 + j];
                const float fr = w * 2.0f * 3.14159265358979323846f;
                feat[j] = w; feat[8 + j] = sinf(fr); feat[16 + j] = cosf(fr);
            }
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                float acc = small[408 + o];
#pragma unroll
                for (int k = 0; k < 24; ++k) acc += small[24 + o * 24 + k] * feat[k];
                pre1[o] = acc; hact[o] = gelu_exact(acc);
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float acc = small[552 + o];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc += small[424 + o * 16 + k] * hact[k];
                pe[o] = acc;
                ps_s[threadIdx.x * 8 + o] = live ? acc / (1.0f + expf(-acc)) : 0.f;
            }
        }
        __syncthreads();
        // ---- phase 2: my map channel against the chunk's pixels
        for (size_t q = 0; q < n_here; ++q) {
            const float g = __bfloat162float(dmap_mine[(base + q) * C2 + mc]);
            gbm += g;
#pragma unroll
            for (int k = 0; k < 8; ++k) gwm[k] = fmaf(g, ps_s[q * 8 + k], gwm[k]);
        }
        // ---- phase 3a: my pixel: dps = Wm1^T dmap1 + Wm2^T dmap2, then back through the small MLP; row V[p] to shared memory
        {
            float dps[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) {
                for (int w = 0; w < 2; ++w) {
                    const uint4* row = reinterpret_cast<const uint4*>((w ? a.dmap2 : a.dmap1) + pix * C2);
                    for (int c8 = 0; c8 < C2 / 8; ++c8) {
                        float g[8];
                        unpack8(__ldg(row + c8), g);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int k = 0; k < 8; ++k) dps[k] = fmaf(g[j], wm[(w * C2 + c8 * 8 + j) * 8 + k], dps[k]);
                    }
                }
            }
            float* row = V + threadIdx.x * kPosRow;
            float dpe[8], dh[16], dfeat[24];
#pragma unroll
            for (int o = 0; o < 8; ++o) { dpe[o] = live ? dps[o] * silu_grad(pe[o]) : 0.f; row[o] = dpe[o]; }
#pragma unroll
            for (int k = 0; k < 16; ++k) { dh[k] = 0.f; row[8 + k] = hact[k]; }
#pragma unroll
            for (int o = 0; o < 8; ++o)
#pragma unroll
                for (int k = 0; k < 16; ++k) dh[k] = fmaf(small[424 + o * 16 + k], dpe[o], dh[k]);
#pragma unroll
            for (int k = 0; k < 24; ++k) { dfeat[k] = 0.f; row[40 + k] = feat[k]; }
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const float da = dh[o] * gelu_grad(pre1[o]);
                row[24 + o] = da;
#pragma unroll
                for (int k = 0; k < 24; ++k) dfeat[k] = fmaf(small[24 + o * 24 + k], da, dfeat[k]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)      // feat = (w, sin(2 pi w), cos(2 pi w))
                row[64 + j] = dfeat[j] + 6.28318530717958647692f * (dfeat[8 + j] * feat[16 + j] - dfeat[16 + j] * feat[8 + j]);
            row[72] = p0; row[73] = p1;
        }
        __syncthreads();
        // ---- phase 3b: my parameters against the chunk's rows (rows of dead pixels carry zero gradients)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (threadIdx.x + r * kT < kPosSmall) {
                float acc = 0.f;
                if (pib[r] >= 0) { for (size_t q = 0; q < n_here; ++q) acc = fmaf(V[q * kPosRow + pia[r]], V[q * kPosRow + pib[r]], acc); }
                else { for (size_t q = 0; q < n_here; ++q) acc += V[q * kPosRow + pia[r]]; }
                pacc[r] += acc;
            }
        }
        __syncthreads();
    }
    {
        float* dwm = which ? a.dwm2 : a.dwm1;
        float* dbm = which ? a.dbm2 : a.dbm1;
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dwm + mc * 8 + k, gwm[k]);
        atomicAdd(dbm + mc, gbm);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int q = threadIdx.x + r * kT;
        if (q >= kPosSmall) continue;
        float* dst = q < 16 ? a.dwe + q : (q < 24 ? a.dbe + (q - 16) : (q < 408 ? a.dw1 + (q - 24) : (q < 424 ? a.db1 + (q - 408)
                     : (q < 552 ? a.dw2 + (q - 424) : a.db2 + (q - 552)))));
        atomicAdd(dst, pacc[r]);
    }
}

__global__ void __launch_bounds__(kT) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, size_t n, float lr, float beta1, float beta2, float eps,
                                                  float wd, float bc1, float bc2_sqrt, float grad_scale) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
        m[i] = mi; v[i] = vi;
        // torch.optim.Adam: denom = sqrt(v) / sqrt(1 - beta2^t) + eps;  p -= (lr / (1 - beta1^t)) * m / denom
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

__global__ void __launch_bounds__(kT) ema_lerp_kernel(float* __restrict__ ema, const float* __restrict__ p, size_t n, float w) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float e = ema[i];
        ema[i] = fmaf(w, p[i] - e, e);      // torch lerp_: start + weight * (end - start)
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int colsum_launch(const bf16* x, int x_ld, int x_off, float* out, int out_ld, bool per_sample, int B, int HW, int C, cudaStream_t s) {
    NDIFF_REQUIRE(C % 8 == 0 && pow2(C / 8) && C / 8 <= kT && x_ld % 8 == 0 && x_off % 8 == 0, "colsum: channel count must be 8 * 2^k (<= 2048)");
    const int ppb = kT / (C / 8);
    int gx = (HW + ppb * 8 - 1) / (ppb * 8);
    const int cap = (148 * 8) / (B > 0 ? B : 1) > 0 ? (148 * 8) / B : 1;
    if (gx > cap) gx = cap;
    colsum_kernel<<<dim3(gx, B), kT, 0, s>>>(x, x_ld, x_off, out, out_ld, per_sample ? 1 : 0, HW, C);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int gn_backward_launch(const GnBwdArgs& a, cudaStream_t s) {
    NDIFF_REQUIRE(a.C % 64 == 0 && a.C <= 512 && pow2(a.C / 8) && a.C % a.G == 0 && (a.C / a.G) % 8 == 0,
                  "GroupNorm backward: channels must be 64 * 2^k (<= 512) in groups of >= 8");
    NDIFF_REQUIRE(a.h && a.dout && a.dh && a.stats && a.gamma && a.beta && a.acc && a.dgamma && a.dbeta, "GroupNorm backward: null argument");
    NDIFF_REQUIRE(!a.maps || a.dmaps, "GroupNorm backward: map gradients need an output");
    const size_t acc_floats = static_cast<size_t>(a.B) * a.C * 2, gsum_floats = static_cast<size_t>(a.B) * a.G * 2;
    NDIFF_CUDA_OK(cudaMemsetAsync(a.acc, 0, (acc_floats + gsum_floats) * sizeof(float), s));     // acc | gsum (caller sizes acc for both)
    float* gsum = a.acc + acc_floats;
    const int ppb = kT / (a.C / 8);
    int gx = (a.HW + ppb * 4 - 1) / (ppb * 4);
    const int cap = (148 * 6) / a.B > 0 ? (148 * 6) / a.B : 1;
    if (gx > cap) gx = cap;
    dim3 grid(gx, a.B);
    if (a.maps) {
        gn_bwd_stats_kernel<true><<<grid, kT, 0, s>>>(a);
        gn_bwd_param_kernel<true><<<a.B, 256, 0, s>>>(a, gsum);
        gn_bwd_apply_kernel<true><<<grid, kT, 0, s>>>(a, gsum);
    } else {
        gn_bwd_stats_kernel<false><<<grid, kT, 0, s>>>(a);
        gn_bwd_param_kernel<false><<<a.B, 256, 0, s>>>(a, gsum);
        gn_bwd_apply_kernel<false><<<grid, kT, 0, s>>>(a, gsum);
    }
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int layernorm_backward_launch(const bf16* x, const float* vec, int vec_ld, const float* g, const bf16* du, bf16* dy_out, float* dg,
                              float* dbeta, int B, int HW, int C, float real_frac, cudaStream_t s) {
    NDIFF_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512, "LayerNorm backward: C must be 64, 128, 256 or 512");
    NDIFF_REQUIRE(real_frac > 0.f && real_frac <= 1.f, "LayerNorm backward: live channel fraction must be in (0, 1]");
    const size_t npix = static_cast<size_t>(B) * HW;
    const int L = C / 8 > 32 ? 32 : C / 8;
    const int grid = blocks_for(npix, 8 * (32 / L) * 4, 148 * 6);
    const size_t smem = 2 * C * sizeof(float);
    if (C == 512) layernorm_bwd_kernel<2><<<grid, kT, smem, s>>>(x, vec, vec_ld, g, du, dy_out, dg, dbeta, HW, C, L, npix, real_frac);
    else layernorm_bwd_kernel<1><<<grid, kT, smem, s>>>(x, vec, vec_ld, g, du, dy_out, dg, dbeta, HW, C, L, npix, real_frac);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int gelu_backward_launch(const bf16* pre, const bf16* dy, bf16* dpre, size_t n, cudaStream_t s) {
    NDIFF_REQUIRE(n % 8 == 0, "GELU backward: element count must be a multiple of 8");
    gelu_bwd_kernel<<<blocks_for(n / 8, kT * 4, 148 * 16), kT, 0, s>>>(reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(dy),
                                                                      reinterpret_cast<uint4*>(dpre), n / 8);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int gelu_forward_launch(const bf16* pre, bf16* y, size_t n, cudaStream_t s) {
    NDIFF_REQUIRE(n % 8 == 0, "GELU: element count must be a multiple of 8");
    gelu_fwd_kernel<<<blocks_for(n / 8, kT * 4, 148 * 16), kT, 0, s>>>(reinterpret_cast<const uint4*>(pre), reinterpret_cast<uint4*>(y), n / 8);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int add_slice_launch(bf16* dst, int dst_ld, int dst_off, const bf16* src, int src_ld, int src_off, int C, size_t npix, bool accumulate,
                     cudaStream_t s) {
    NDIFF_REQUIRE(C % 8 == 0 && dst_ld % 8 == 0 && dst_off % 8 == 0 && src_ld % 8 == 0 && src_off % 8 == 0, "add: 16-byte channel alignment");
    const size_t total = npix * (C / 8);
    add_slice_kernel<<<blocks_for(total, kT * 4, 148 * 16), kT, 0, s>>>(dst, dst_ld, dst_off, src, src_ld, src_off, C / 8, total, accumulate ? 1 : 0);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int upsample2x_backward_launch(const bf16* dy, bf16* dx, int B, int H, int W, int C, bool accumulate, cudaStream_t s) {
    const int cv = C / 8;
    const size_t total = static_cast<size_t>(B) * H * W * cv;
    upsample2x_bwd_kernel<<<blocks_for(total, kT * 2, 148 * 16), kT, 0, s>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), H, W,
                                                                            cv, total, accumulate ? 1 : 0);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int depth_to_space_launch(const bf16* t, bf16* dx, int B, int H, int W, int C, bool accumulate, cudaStream_t s) {
    const int cv = C / 8;
    const size_t total = static_cast<size_t>(B) * 4 * H * W * cv;
    depth_to_space_kernel<<<blocks_for(total, kT * 4, 148 * 16), kT, 0, s>>>(reinterpret_cast<const uint4*>(t), reinterpret_cast<uint4*>(dx), H, W,
                                                                            cv, total, accumulate ? 1 : 0);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int heads_backward_launch(const HeadsBwdArgs& a, cudaStream_t s) {
    NDIFF_REQUIRE(a.C == 64, "heads backward: the final feature maps have 64 (physical) channels");
    const size_t npix = static_cast<size_t>(a.B) * a.HW;
    heads_bwd_kernel<<<blocks_for(npix, 32 * 8, 148 * 4), kT, 0, s>>>(a);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int shot_in_backward_launch(const float* clean, const float* x, const float* w, const float* bias, const bf16* ds0, float* dw, float* db,
                            size_t npix, int C, cudaStream_t s) {
    NDIFF_REQUIRE(C % 8 == 0 && kT % (C / 8) == 0, "shot_mlp1.fc1 backward: bad channel count");
    const size_t smem = static_cast<size_t>(C) * (8 + 1 + 9) * sizeof(float);
    shot_in_bwd_kernel<<<blocks_for(npix, (kT / (C / 8)) * 16, 148 * 4), kT, smem, s>>>(reinterpret_cast<const float4*>(clean),
                                                                                       reinterpret_cast<const float4*>(x), w, bias, ds0, dw, db, npix, C);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int init_conv_wgrad_launch(const float* x_nhwc4, const bf16* dy, float* dw, int B, int H, int W, int C, cudaStream_t s) {
    NDIFF_REQUIRE(C == 64, "init_conv weight gradient: 64 (physical) output channels");
    const int n_tiles = B * ((H + kIwTile - 1) / kIwTile) * ((W + kIwTile - 1) / kIwTile);
    init_conv_wgrad_kernel<<<n_tiles < 148 * 2 ? n_tiles : 148 * 2, kT, 0, s>>>(reinterpret_cast<const float4*>(x_nhwc4), dy, dw, B, H, W);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int time_mlp_train_launch(const int* t, int n, int dim, const float* w1, const float* b1, const float* w2, const float* b2, float* st_out,
                          float* saved, cudaStream_t s) {
    NDIFF_REQUIRE(dim <= 128 && dim % 2 == 0, "time embedding: dim must be even and <= 128");
    time_mlp_train_kernel<<<n, kT, 0, s>>>(t, dim, w1, b1, w2, b2, st_out, saved);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int time_mlp_backward_launch(const float* dst, const float* saved, int n, int dim, const float* w1, const float* w2, float* dw1, float* db1,
                             float* dw2, float* db2, cudaStream_t s) {
    NDIFF_REQUIRE(dim <= 128 && dim % 2 == 0, "time embedding: dim must be even and <= 128");
    time_mlp_bwd_kernel<<<n, kT, 0, s>>>(dst, saved, dim, w1, w2, dw1, db1, dw2, db2);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int small_gemm_launch(bool tA, bool tB, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                      bool accumulate, cudaStream_t s) {
    // long contractions with few outputs (d time-vector = d(scale, shift) @ W: 32 x 256 outputs, K = 8192) are split over K
    const int kchunk = K >= 2048 ? 256 : K;
    const int splits = (K + kchunk - 1) / kchunk;
    if (splits > 1 && !accumulate) {
        NDIFF_REQUIRE(ldc == N, "small GEMM: split-K overwrite needs a dense C");
        NDIFF_CUDA_OK(cudaMemsetAsync(C, 0, static_cast<size_t>(M) * N * sizeof(float), s));
    }
    dim3 grid(blocks_for(static_cast<size_t>(M) * N, kT, 148 * 16), splits);
    small_gemm_kernel<<<grid, kT, 0, s>>>(tA ? 1 : 0, tB ? 1 : 0, M, N, K, A, lda, Bm, ldb, C, ldc, accumulate ? 1 : 0, kchunk);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int rowsum_f32_launch(const float* in, int ld, int n, int C, float* out, cudaStream_t s) {
    rowsum_f32_kernel<<<(C + 255) / 256, 256, 0, s>>>(in, ld, n, C, out);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int iso_vec_backward_launch(const float* emb_table, const long long* idx, const float* wv, const float* wo, const float* dc, int dc_ld,
                            int dc_off, float* demb, float* dwv, float* dwo, float* dbo, int B, int C, cudaStream_t s) {
    NDIFF_REQUIRE(C <= 512, "iso vector backward: C <= 512");
    iso_vec_bwd_kernel<<<B, 128, 0, s>>>(emb_table, idx, wv, wo, dc, dc_ld, dc_off, demb, dwv, dwo, dbo, C);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pos_backward_launch(const PosBwdArgs& a, cudaStream_t s) {
    NDIFF_REQUIRE(a.fwd.C == 64, "positional backward: the ResnetBlock2 maps have 2 x 64 (physical) channels");
    const int C2 = 2 * a.fwd.C;
    const size_t smem = (2 * C2 * 8 + kPosSmall + kT * 8 + kT * kPosRow) * sizeof(float);
    static bool opted = false;
    if (!opted) {
        NDIFF_CUDA_OK(cudaFuncSetAttribute(pos_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        opted = true;
    }
    const size_t npix = static_cast<size_t>(a.fwd.B) * a.fwd.HW;
    pos_bwd_kernel<<<blocks_for(npix, kT * 4, 148 * 2), kT, smem, s>>>(a);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int adam_launch(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                int step, float grad_scale, cudaStream_t s) {
    NDIFF_REQUIRE(step >= 1, "Adam: step counts from 1");
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
    adam_kernel<<<blocks_for(n, kT * 4, 148 * 16), kT, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}
int ema_lerp_launch(float* ema, const float* p, size_t n, float weight, cudaStream_t s) {
    ema_lerp_kernel<<<blocks_for(n, kT * 4, 148 * 16), kT, 0, s>>>(ema, p, n, weight);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ndiff
