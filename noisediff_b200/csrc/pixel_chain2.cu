// noisediff_b200 — fused per-pixel MLP chains, TWO threads per pixel (sm_100a).  Same programs, blobs and arguments as
// pixel_chain.cu (see pixel_chain.cuh); different work split.
//
// pixel_chain.cu runs one thread per pixel: 3 warpgroups = 12 warps per SM, three per scheduler, and every stage is a long
// dependent chain per thread (TMEM load -> bias -> GELU -> pack -> shared store) — measured 33 % issue utilisation, the
// kernels are latency bound.  Here a 128-pixel tile belongs to a GROUP of 256 threads: warp w of the group reads TMEM lane
// quarter w % 4 (the hardware's rule) and column half w / 4, so each thread handles half of every row — half the registers,
// half the chain length per stage — and three groups put 24 warps on the SM (six per scheduler) within the same shared
// memory and TMEM budget.  LayerNorm's two moments cross the two halves through shared memory.
#include "pixel_chain.cuh"
#include "conv_gemm.cuh"

#include <cstring>
#include <mutex>

namespace ndiff {

namespace {

constexpr int kTile = 128;
constexpr int kBlk = kTile * 128;
constexpr int kNG = 3;                     // groups (tiles in flight) per CTA
constexpr int kGT = 256;                   // threads per group
constexpr int kWgBytes = 3 * kBlk;         // X | A0 | A1
constexpr int kThreads2 = kNG * kGT;       // 768 -> at most 80 registers per thread
constexpr int kTmemCols = 512;

struct Chain2Tail {
    uint64_t bar_w, bar_x[kNG], bar_mma[kNG];
    uint32_t tmem_base;
    uint32_t pad_;
    float fvec[kChainShotFloats];
    alignas(16) float ctab[kNG][2][2][64];      // per group / sample slot: [0] c, [1] shot: b2 + c, attn: Wp (b2 + c) + bp
    alignas(8) float2 ln[kNG][2][kTile];        // LayerNorm partial moments of the two column halves
};

__device__ __forceinline__ uint32_t swz(uint32_t blk, int r, int j) { return blk + r * 128 + ((j ^ (r & 7)) << 4); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

template <int N, bool F16 = false>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_blk, uint32_t w_blk, int kblocks, int k16, bool accumulate = false) {
    constexpr uint32_t idesc = F16 ? umma_idesc_f16(128, N) : umma_idesc_bf16(128, N);
    constexpr uint32_t hi = umma_desc_hi(1024);
    bool first = !accumulate;
    for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t a_lo = umma_desc_lo(a_blk + kb * kBlk), b_lo = umma_desc_lo(w_blk + kb * N * 128);
        for (int k = 0; k < k16; ++k) {
            if (first) umma_bf16_lohi<false>(d_tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc);
            else umma_bf16_lohi<true>(d_tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc);
            first = false;
        }
    }
}

// 16 accumulator columns (+ bias) -> GELU -> fp16 operand: chunks j0, j0 + 1 of row r
__device__ __forceinline__ void gelu16_to_f16(uint32_t blk, int r, int j0, const uint32_t (&raw)[16], const float* bias) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        uint4 u;
        u.x = gelu_f16x2(__uint_as_float(raw[jj * 8 + 0]) + bias[jj * 8 + 0], __uint_as_float(raw[jj * 8 + 1]) + bias[jj * 8 + 1]);
        u.y = gelu_f16x2(__uint_as_float(raw[jj * 8 + 2]) + bias[jj * 8 + 2], __uint_as_float(raw[jj * 8 + 3]) + bias[jj * 8 + 3]);
        u.z = gelu_f16x2(__uint_as_float(raw[jj * 8 + 4]) + bias[jj * 8 + 4], __uint_as_float(raw[jj * 8 + 5]) + bias[jj * 8 + 5]);
        u.w = gelu_f16x2(__uint_as_float(raw[jj * 8 + 6]) + bias[jj * 8 + 6], __uint_as_float(raw[jj * 8 + 7]) + bias[jj * 8 + 7]);
        sts128(swz(blk, r, j0 + jj), u);
    }
}
__device__ __forceinline__ void store16_bf16(uint32_t blk, int r, int j0, const float (&v)[16]) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        uint4 u;
        u.x = pack_bf16(v[jj * 8 + 0], v[jj * 8 + 1]); u.y = pack_bf16(v[jj * 8 + 2], v[jj * 8 + 3]);
        u.z = pack_bf16(v[jj * 8 + 4], v[jj * 8 + 5]); u.w = pack_bf16(v[jj * 8 + 6], v[jj * 8 + 7]);
        sts128(swz(blk, r, j0 + jj), u);
    }
}

template <int PROG>
__global__ void __launch_bounds__(kThreads2, 1) pixel_chain2_kernel(const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr bool kShot = PROG == kProgShot;
    constexpr int kWRows = kShot ? kChainShotRows : kChainAttnRows;
    constexpr int kNF = kShot ? kChainShotFloats : kChainAttnFloats;
    constexpr int kWBytes = kWRows * 128;
    Chain2Tail* tail = reinterpret_cast<Chain2Tail*>(smem + kWBytes + kNG * kWgBytes);

    const int tid = threadIdx.x, g = tid / kGT, tg = tid % kGT, wi = tg >> 5, q = wi & 3, hc = wi >> 2;
    const int r = q * 32 + (tid & 31);          // pixel row of the tile = TMEM lane
    const int c0 = hc * 32;                     // this thread's 32 of the 64 channels
    const uint32_t sW = smem_u32(smem);
    const uint32_t sX = sW + kWBytes + g * kWgBytes, sA0 = sX + kBlk, sA1 = sA0 + kBlk;
    const uint32_t sWattn = sW + (kShot ? 128 * 128 : 0);
    const uint32_t sW1 = sWattn, sW2 = sW1 + 128 * 128, sWp = sW2 + 128 * 128, sWm2 = sWp + 128 * 128;
    const float* fA = tail->fvec + (kShot ? 128 : 0);
    const float* f_b1 = fA + 128, *f_b2 = fA + 256, *f_bm1 = fA + 384, *f_bm2 = fA + 448;
    (void)f_b2; (void)f_bm1; (void)f_bm2; (void)sWm2;
    const uint32_t bar_w = smem_u32(&tail->bar_w), bar_x = smem_u32(&tail->bar_x[g]), bar_mma = smem_u32(&tail->bar_mma[g]);

    if (tid == 0) {
        tma_prefetch_desc(&a.tmW);
        tma_prefetch_desc(&a.tmOut);
        if (kShot) tma_prefetch_desc(&a.tmOut2); else tma_prefetch_desc(&a.tmX);
        mbar_init(&tail->bar_w, 1);
        for (int i = 0; i < kNG; ++i) { mbar_init(&tail->bar_x[i], 1); mbar_init(&tail->bar_mma[i], 1); }
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc<kTmemCols>(&tail->tmem_base);
    for (int i = tid; i < kNF; i += kThreads2) tail->fvec[i] = __ldg(a.fvec + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tail->tmem_base + g * 128;
    const uint32_t tmem_rd = tmem_d + (static_cast<uint32_t>(q * 32) << 16);

    if (tid == 0) {
        mbar_expect_tx(bar_w, kWBytes);
        for (int i = 0; i < kWRows / 64; ++i) tma_load_2d(sW + i * 64 * 128, &a.tmW, bar_w, 0, i * 64);
    }
    pdl_trigger();
    pdl_wait();
    const int tile0 = blockIdx.x * kNG + g, tile_step = gridDim.x * kNG;
    if (!kShot && tg == 0 && tile0 < a.n_tiles) {
        mbar_expect_tx(bar_x, kBlk);
        tma_load_2d(sX, &a.tmX, bar_x, 0, tile0 * kTile);
    }
    uint32_t xph = 0, mph = 0;
    bool w_ready = false;
    int tab_b0 = -1, tab_b1 = -1;

#define NDIFF_STAGE2(ISSUE)                                                  \
    do {                                                                     \
        fence_proxy_async();                                                 \
        tc_fence_before();                                                   \
        named_bar_sync(1 + g, kGT);                                          \
        if (tg == 0) {                                                       \
            if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }           \
            tc_fence_after();                                                \
            ISSUE;                                                           \
            umma_commit(bar_mma);                                            \
        }                                                                    \
        mbar_wait(bar_mma, mph);                                             \
        mph ^= 1;                                                            \
        tc_fence_after();                                                    \
    } while (0)

    for (int tile = tile0; tile < a.n_tiles; tile += tile_step) {
        const int p = tile * kTile + r;
        const bool live = p < a.npix;
        const int pc = live ? p : a.npix - 1;
        uint32_t xr[16];                      // my 32 channels of the attention-block input, packed bf16
        const int b_first = (tile * kTile) / a.HW;
        {
            const int last = tile * kTile + kTile - 1;
            const int b_last = (last < a.npix ? last : a.npix - 1) / a.HW;
            if (b_first != tab_b0 || b_last != tab_b1) {      // (uniform over the group)
                tab_b0 = b_first; tab_b1 = b_last;
                if (tg < 128) {
                    const int slot = tg >> 6, j = tg & 63;
                    const size_t co = static_cast<size_t>(slot ? b_last : b_first) * a.cvec_ld + j;
                    const float cj = __ldg(a.cvec + co);
                    tail->ctab[g][slot][0][j] = cj;
                    tail->ctab[g][slot][1][j] = kShot ? cj + f_b2[j] : __ldg(a.cvec2 + co);
                }
                named_bar_sync(1 + g, kGT);
            }
        }
        const float* ct = &tail->ctab[g][(pc / a.HW) != b_first ? 1 : 0][0][0];      // c at ct[j], second vector at ct[64 + j]

        if (tg == 0) tma_store_wait_read();       // X (shot) / A1 are sources of the previous tile's TMA stores

        if constexpr (kShot) {
            // ---- shot_mlp1.fc1 on cat[clean, x_t] (ref :598): one 16-byte operand chunk per pixel, written by the hc == 0 thread
            if (hc == 0) {
                float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f), x4 = c4;
                if (live) { c4 = __ldg(a.clean + p); x4 = a.x[p]; }
                uint4 u;
                u.x = pack_bf16(c4.x, c4.y); u.y = pack_bf16(c4.z, c4.w); u.z = pack_bf16(x4.x, x4.y); u.w = pack_bf16(x4.z, x4.w);
                sts128(swz(sA0, r, 0), u);
                sts128(swz(sA0, r, 1), make_uint4(0u, 0u, 0u, 0u));
            }
            NDIFF_STAGE2(issue_gemm<64>(tmem_d, sA0, sW, 1, 1));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[16];
                tmem_ld16(tmem_rd + c0 + h * 16, raw);
                tmem_ld_wait();
                gelu16_to_f16(sA0, r, hc * 4 + h * 2, raw, tail->fvec + c0 + h * 16);
            }
            // ---- shot_mlp1.fc2 -> s1 (the branch's residual r_s, ref :599): kept in registers and staged in the X slot
            NDIFF_STAGE2((issue_gemm<64, true>(tmem_d, sA0, sW + 64 * 128, 1, 4)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[16];
                tmem_ld16(tmem_rd + c0 + h * 16, raw);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 2)
                    xr[h * 8 + j / 2] = pack_bf16(__uint_as_float(raw[j]) + tail->fvec[64 + c0 + h * 16 + j],
                                                  __uint_as_float(raw[j + 1]) + tail->fvec[64 + c0 + h * 16 + j + 1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) sts128(swz(sX, r, hc * 4 + j), make_uint4(xr[j * 4], xr[j * 4 + 1], xr[j * 4 + 2], xr[j * 4 + 3]));
        } else {
            mbar_wait(bar_x, xph);
            xph ^= 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 u = lds128(swz(sX, r, hc * 4 + j));
                xr[j * 4] = u.x; xr[j * 4 + 1] = u.y; xr[j * 4 + 2] = u.z; xr[j * 4 + 3] = u.w;
            }
        }

        // ---- y = x + c; LayerNorm over the 64 channels: each half sums its 32, the halves meet in shared memory ----------------
        {
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(ct + c0 + j);
                const float2 f0 = unpack_bf16(xr[j / 2]), f1 = unpack_bf16(xr[j / 2 + 1]);
                const float y0 = f0.x + c4.x, y1 = f0.y + c4.y, y2 = f1.x + c4.z, y3 = f1.y + c4.w;
                sum += (y0 + y1) + (y2 + y3);
                sq = fmaf(y0, y0, sq); sq = fmaf(y1, y1, sq); sq = fmaf(y2, y2, sq); sq = fmaf(y3, y3, sq);
            }
            tail->ln[g][hc][r] = make_float2(sum, sq);
            named_bar_sync(1 + g, kGT);
            const float2 o = tail->ln[g][hc ^ 1][r];
            sum += o.x; sq += o.y;
            const float mean = sum * a.inv_c;
            const float rstd = rsqrtf(fmaxf(sq * a.inv_c - mean * mean, 0.f) + 1e-5f);
            const float nb = -mean * rstd;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 c4 = *reinterpret_cast<const float4*>(ct + c0 + h * 16 + j);
                    const float2 f0 = unpack_bf16(xr[(h * 16 + j) / 2]), f1 = unpack_bf16(xr[(h * 16 + j) / 2 + 1]);
                    v[j] = fmaf(f0.x + c4.x, rstd, nb); v[j + 1] = fmaf(f0.y + c4.y, rstd, nb);
                    v[j + 2] = fmaf(f1.x + c4.z, rstd, nb); v[j + 3] = fmaf(f1.y + c4.w, rstd, nb);
                }
                store16_bf16(sA0, r, hc * 4 + h * 2, v);
            }
        }
        // ---- FeedForward.net.0: Linear(C, 2C) + GELU; hidden K block hc -> operand block A0 / A1 (this thread: 64 of the 128 columns)
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1 + g, kGT);
        if (tg == 0) {
            if (kShot) {                       // s1 staged in the X slot by every thread before this barrier
                tma_store_2d(&a.tmOut2, sX, 0, tile * kTile);
                tma_store_commit();
            }
            if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
            tc_fence_after();
            issue_gemm<128>(tmem_d, sA0, sW1, 1, 4);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mph);
        mph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint32_t raw[16];
            tmem_ld16(tmem_rd + hc * 64 + h * 16, raw);
            tmem_ld_wait();
            gelu16_to_f16(sA0 + hc * kBlk, r, h * 2, raw, f_b1 + hc * 64 + h * 16);
        }
        if constexpr (!kShot) {
            // ---- ff.net.2 and proj_out as ONE stage: out = (Wp W2) h + Wp x + [Wp (b2 + c) + bp] + x   (see pixel_chain.cu)
            NDIFF_STAGE2((issue_gemm<64, true>(tmem_d, sA0, sW2, 2, 4), issue_gemm<64>(tmem_d, sX, sWp, 1, 4, true)));
            if (tg == 0 && tile + tile_step < a.n_tiles) {
                mbar_expect_tx(bar_x, kBlk);
                tma_load_2d(sX, &a.tmX, bar_x, 0, (tile + tile_step) * kTile);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[16];
                tmem_ld16(tmem_rd + c0 + h * 16, raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 c4 = *reinterpret_cast<const float4*>(ct + 64 + c0 + h * 16 + j);
                    const float2 f0 = unpack_bf16(xr[(h * 16 + j) / 2]), f1 = unpack_bf16(xr[(h * 16 + j) / 2 + 1]);
                    v[j] = __uint_as_float(raw[j]) + (f0.x + c4.x);
                    v[j + 1] = __uint_as_float(raw[j + 1]) + (f0.y + c4.y);
                    v[j + 2] = __uint_as_float(raw[j + 2]) + (f1.x + c4.z);
                    v[j + 3] = __uint_as_float(raw[j + 3]) + (f1.y + c4.w);
                }
                store16_bf16(sA1, r, hc * 4 + h * 2, v);
            }
        } else {
            // ---- FeedForward.net.2: z = ff + (b2 + c) + s1 -> A0 (bf16) ---------------------------------------------------------
            NDIFF_STAGE2((issue_gemm<64, true>(tmem_d, sA0, sW2, 2, 4)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[16];
                tmem_ld16(tmem_rd + c0 + h * 16, raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 c4 = *reinterpret_cast<const float4*>(ct + 64 + c0 + h * 16 + j);      // b2 + c
                    const float2 f0 = unpack_bf16(xr[(h * 16 + j) / 2]), f1 = unpack_bf16(xr[(h * 16 + j) / 2 + 1]);
                    v[j] = __uint_as_float(raw[j]) + (f0.x + c4.x);
                    v[j + 1] = __uint_as_float(raw[j + 1]) + (f0.y + c4.y);
                    v[j + 2] = __uint_as_float(raw[j + 2]) + (f1.x + c4.z);
                    v[j + 3] = __uint_as_float(raw[j + 3]) + (f1.y + c4.w);
                }
                store16_bf16(sA0, r, hc * 4 + h * 2, v);
            }
            // ---- proj_out + x_in folded into shot_mlp2.fc1: ONE K = 128 GEMM over [s1 | z] (X slot, A0); GELU; fc2 (ref :441-443, :601)
            NDIFF_STAGE2(issue_gemm<64>(tmem_d, sX, sWp, 2, 4));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[16];
                tmem_ld16(tmem_rd + c0 + h * 16, raw);
                tmem_ld_wait();
                gelu16_to_f16(sA0, r, hc * 4 + h * 2, raw, f_bm1 + c0 + h * 16);
            }
            NDIFF_STAGE2((issue_gemm<64, true>(tmem_d, sA0, sWm2, 1, 4)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[16];
                tmem_ld16(tmem_rd + c0 + h * 16, raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + f_bm2[c0 + h * 16 + j];
                store16_bf16(sA1, r, hc * 4 + h * 2, v);
            }
        }
        // ---- output tile: staged in A1, stored by TMA (rows beyond npix are clipped by the tensor map) -------------------------
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1 + g, kGT);
        if (tg == 0) {
            tma_store_2d(&a.tmOut, sA1, 0, tile * kTile);
            tma_store_commit();
        }
    }
#undef NDIFF_STAGE2
    if (tg == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tail->tmem_base);
    }
}

int chain2_smem_bytes(int prog) {
    const int rows = prog == kProgShot ? kChainShotRows : kChainAttnRows;
    return 1024 + rows * 128 + kNG * kWgBytes + static_cast<int>(sizeof(Chain2Tail));
}

}  // namespace

int pixel_chain2_smem_bytes(int prog) { return chain2_smem_bytes(prog); }

int pixel_chain2_init() {
    NDIFF_CUDA_OK(cudaFuncSetAttribute(pixel_chain2_kernel<kProgAttn>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain2_smem_bytes(kProgAttn)));
    NDIFF_CUDA_OK(cudaFuncSetAttribute(pixel_chain2_kernel<kProgShot>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain2_smem_bytes(kProgShot)));
    return 0;
}

int pixel_chain2_launch(const ChainPlan& plan, cudaStream_t stream) {
    {
        static std::once_flag once;
        static int init_rc = 0;
        std::call_once(once, [] { init_rc = pixel_chain2_init(); });
        if (init_rc) return 1;
    }
    const int smem = chain2_smem_bytes(plan.prog);
    NDIFF_REQUIRE(smem <= 227 * 1024, "pixel chain (two threads per pixel): shared-memory budget exceeded");
    // grid: one CTA per kNG tiles, at most one per SM (same tile walk as the one-thread-per-pixel kernel: plan.grid)
    if (plan.prog == kProgShot)
        NDIFF_CUDA_OK(launch_pdl(pixel_chain2_kernel<kProgShot>, dim3(plan.grid), dim3(kThreads2), smem, stream, plan.args));
    else
        NDIFF_CUDA_OK(launch_pdl(pixel_chain2_kernel<kProgAttn>, dim3(plan.grid), dim3(kThreads2), smem, stream, plan.args));
    return 0;
}

}  // namespace ndiff
