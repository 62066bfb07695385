// noisediff_b200 — tcgen05 implicit-GEMM convolution over NHWC bf16 activations (sm_100a).
//
// GEMM view (SURVEY.md §8a "3x3 conv shape census"): M = pixels, N = C_out, K = taps * C_in.
//   A (M x K)  : activations, K-major.  One smem row = one pixel's 64-channel block = 128 B -> SWIZZLE_128B rows.
//                TMA boxes {64 ch, TW, rows, 1 image} land a TH x TW pixel tile directly in that layout; a conv tap is
//                a shifted window, so out-of-image pixels are the TMA's zero fill (pad = 1 for free).
//   B (N x K)  : weights, repacked once to bf16 [C_out][cblk][tap][64], K-major, 2-D TMA boxes {64, NT}.
//   D (M x N)  : fp32 accumulator in TMEM, 128 lanes (pixels) x NT columns, double-buffered across tiles.
// Persistent CTA, 8 warps: w0 = TMA producer, w1 = MMA issuer (one elected lane issues tcgen05.mma),
// w2 = TMEM allocator, w4..7 = epilogue (tcgen05.ld -> bias / per-sample vector / residual / GELU / GroupNorm
// partial sums -> bf16 NHWC stores).  Replaces the reference's aten::conv2d / aten::linear calls inside
// Block.proj, res_conv, Downsample, AttnBlock.{ff,proj_out} and Mlp.fc* (models/archs/Diffusion_arch.py:128-443).
#include "conv_gemm.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

namespace ndiff {

namespace {

constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;                    // two warps per TMEM lane quarter, each takes every other 32-column chunk
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kXfWarp0 = kEpiWarp0 + kEpiWarps;  // XF kernels: four more warps transform the activation stages in place
constexpr int kXfExtra = 4;                      // warps added by the XF kernels
constexpr int kXfWarps = kXfExtra + 2;           // ... plus warps 2 and 3, idle after the TMEM allocation.  (Measured: ten transform
                                                 // warps with four epilogue warps are SLOWER — the transform shares the shared-memory
                                                 // pipe with the MMA's operand reads, which is what bounds the 64-channel layers.)
constexpr int kThreadsXf = (kXfWarp0 + kXfExtra) * 32;

struct SmemTail {  // lives after the operand rings
    uint64_t fullA[8], emptyA[8], fullB[16], emptyB[16], xfA[8];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad_;
    alignas(16) float bias[128];   // this CTA's N-tile slice of the bias (zeros when the layer has none)
    alignas(16) float bias2[128];  // kHalo1R: bias slice of the fused 1x1 residual convolution
    alignas(16) float2 coef[512];  // XF: folded GroupNorm affine (A, B) per input channel of the current sample
};

// Tap geometry of the single-copy halo mode (compile-time: TW = 8 so that one 8-row UMMA group == one output row).
constexpr int kHaloTW = 8, kHaloTH = 16;
constexpr int kHaloPitch = (kHaloTW + 2) * 128;                          // bytes between halo rows
constexpr int kHaloCopy = (kHaloTH + 2) * kHaloPitch;                    // 23040 B landed by one TMA box
constexpr int kHaloStage = (kHaloCopy + 1023) / 1024 * 1024;
constexpr int kHaloCopy2 = (2 * kHaloTH + 2) * kHaloPitch;                // 43520 B: halo of two stacked sub-tiles
constexpr int kHaloStage2 = (kHaloCopy2 + 1023) / 1024 * 1024;
constexpr int kSubStep = (kHaloTH * kHaloPitch) >> 4;                    // descriptor offset of the second sub-tile
// taps per streamed weight stage in halo mode
__host__ __device__ constexpr int halo_btaps(int nt, int sub) { return (sub == 2 && nt == 128) ? 1 : 3; }
constexpr int kToepPitch = 2304;                                         // Toeplitz operand: bytes between the tap rows of a stage (>= 136 x 16)
constexpr int kResBTaps = 2;                                             // kHalo1R: 10 weight blocks per channel block, two per stage
constexpr int kUpBTaps = 4;                                              // kHaloUp: one stage = the 4 taps of one phase

// XF: the input is the RAW output of the previous conv and the GroupNorm-apply pass that used to sit between the two
// (y = SiLU(x * A_c + B_c), A/B folding mean, rstd, gamma, beta and the time scale/shift; Block.forward, ref :135-144) runs
// here, in place on each activation stage between the TMA landing and the MMA: warps 2, 3, 12..15 rewrite the halo box in shared
// memory (out-of-image pixels stay the TMA's zero fill = the conv's zero padding of the NORMALISED tensor), then hand the
// stage to the MMA warp through xfA.  Same fp32 formulas and bf16 rounding as gn_apply_kernel, so results are bit-identical.
// WS (the N = 64 kHalo2 kernels, default): the two 128-pixel sub-tiles of a tile multiply the SAME weight block, and the
// 64-output-channel layers are bound by the shared-memory reads of their operands (4 KB of A + 2 KB of B per 32-cycle MMA).
// tcgen05.mma.ws keeps B in a collector buffer, so sub-tile 1 reuses what sub-tile 0 loaded: 6 -> 5 KB per MMA, the same saving
// as a cta_group::2 pair without the cluster.  Verified on B200 in round 2: same lane = row accumulator layout at M = 128
// (bit-compatible results, tests/test_gpu_ops.py), -6 % on the plain 64 -> 64 layers, -2 % on the XF ones.
template <int NT, int MODE, bool RES, bool XF = false, bool WS = false>
__global__ void __launch_bounds__(XF ? kThreadsXf : kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmArgs a) {
    static_assert(!WS || MODE == kHalo2, "the weight-stationary form exists for the two-sub-tile (kHalo2) kernels only");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem is only guaranteed 16-B aligned by the ABI; SWIZZLE_128B wants 1024
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __builtin_assume(__isShared(smem));      // the manual alignment hides the address space: without the hint table reads are generic LD.E
    constexpr bool kHalo = MODE == kHalo1 || MODE == kHalo2 || MODE == kHaloUp || MODE == kHalo1R;
    constexpr bool kUp = MODE == kHaloUp;
    constexpr bool kRes2 = MODE == kHalo1R;                      // second accumulator: 1x1 conv of the same input (centre tap)
    constexpr int kSub = MODE == kHalo2 ? 2 : 1;                 // 128-pixel sub-tiles (TMEM accumulators) per CTA tile
    constexpr int kHTaps = kUp ? 4 : (kRes2 ? 10 : 9);           // weight blocks per (channel block[, phase]) in the halo modes
    constexpr int kBG = kUp ? kUpBTaps : (kRes2 ? kResBTaps : (kHalo ? halo_btaps(NT, kSub) : 1));   // taps per streamed weight stage
    constexpr int kAccCols = (kRes2 ? 4 : 2 * kSub) * NT;        // TMEM columns: double-buffered accumulators
    constexpr int kBTap = NT * 128;                              // one [NT x 64] weight block (one tap of one channel block)
    constexpr int kBStage = kBG * kBTap;
    constexpr int kAStage = MODE == kHalo2 ? kHaloStage2 : (kHalo ? kHaloStage : 128 * 128);
    constexpr int kACopy = MODE == kHalo2 ? kHaloCopy2 : kHaloCopy;
    const int a_stages = a.a_stages, b_stages = a.b_stages;
    const uint32_t ringA = smem_u32(smem);
    const uint32_t ringB = ringA + a_stages * kAStage;
    SmemTail* tail = reinterpret_cast<SmemTail*>(smem + a_stages * kAStage + a.b_region_bytes + a.stage_bytes);
    const uint32_t stageS = ringB + a.b_region_bytes;          // epilogue staging block (a.stage_bytes > 0), 1024-B aligned
    const uint32_t bar_fullA = smem_u32(&tail->fullA[0]), bar_emptyA = smem_u32(&tail->emptyA[0]);
    const uint32_t bar_fullB = smem_u32(&tail->fullB[0]), bar_emptyB = smem_u32(&tail->emptyB[0]);
    const uint32_t bar_tfull = smem_u32(&tail->tmem_full[0]), bar_tempty = smem_u32(&tail->tmem_empty[0]);
    const uint32_t bar_xfA = smem_u32(&tail->xfA[0]);
    const uint32_t bar_readyA = XF ? bar_xfA : bar_fullA;      // what the MMA warp waits on before reading an A stage

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int CB = a.cb0 + a.cb1;
    const int TAPS = kUp ? 16 : (kHalo ? kHTaps : a.taps_y * a.taps_x);    // weight blocks per channel block
    const int nt = blockIdx.x % a.n_tiles;                       // this CTA's N tile for its whole life
    // contiguous range of M tiles per CTA: consecutive tiles share halo rows in L2 and (almost always) the sample index,
    // which lets the epilogue keep GroupNorm partial sums in registers across tiles
    const int m_total = a.total_tiles / a.n_tiles;
    const int cta_m = blockIdx.x / a.n_tiles, n_cta_m = gridDim.x / a.n_tiles;
    const int m_begin = static_cast<int>(static_cast<long long>(m_total) * cta_m / n_cta_m);
    const int m_end = static_cast<int>(static_cast<long long>(m_total) * (cta_m + 1) / n_cta_m);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a.tmA0);
        if (a.cb1 > 0) tma_prefetch_desc(&a.tmA1);
        tma_prefetch_desc(&a.tmB);
        if (a.stage_bytes) tma_prefetch_desc(&a.tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < a_stages; ++i) { mbar_init(&tail->fullA[i], 1); mbar_init(&tail->emptyA[i], 1); mbar_init(&tail->xfA[i], kXfWarps); }
        for (int i = 0; i < b_stages; ++i) { mbar_init(&tail->fullB[i], 1); mbar_init(&tail->emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tail->tmem_full[i], 1); mbar_init(&tail->tmem_empty[i], kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kAccCols>(&tail->tmem_base);
    if (threadIdx.x < NT) {
        tail->bias[threadIdx.x] = a.bias ? __ldg(a.bias + nt * NT + threadIdx.x) : 0.f;
        if (kRes2) tail->bias2[threadIdx.x] = a.bias2 ? __ldg(a.bias2 + nt * NT + threadIdx.x) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;
    pdl_trigger();

    if (warp == 0) {
        // ================================ TMA producer (whole warp walks the loop, one elected lane issues) ========
        int sa = 0, pa = 0, sb = 0, pb = 0;
        if constexpr (RES) {
            if (elect_one()) {
                const int nkb = CB * TAPS;
                mbar_expect_tx(bar_fullB, nkb * kBTap);
                for (int kb = 0; kb < nkb; ++kb) tma_load_2d(ringB + kb * kBTap, &a.tmB, bar_fullB, kb * 64, nt * NT);
            }
            __syncwarp();
        }
        // weights are constants; activations are the previous kernel's output.  Everything downstream of the first
        // activation load (MMA, epilogue reads and writes) is ordered behind this wait by the mbarrier chain.
        pdl_wait();
        for (int mt = m_begin; mt < m_end; ++mt) {
            int m = mt;
            const int phase = kUp ? (m & 3) : 0;
            if (kUp) m >>= 2;
            const int tx = m % a.tiles_x; m /= a.tiles_x;
            const int ty = m % a.tiles_y;
            const int b = m / a.tiles_y;
            const int x0 = tx * a.TW, y0 = ty * a.TH;
            int kcol = kUp ? phase * 4 * 64 : 0;
            for (int cb = 0; cb < CB; ++cb) {
                const CUtensorMap* tm = cb < a.cb0 ? &a.tmA0 : &a.tmA1;
                const int c0 = (cb < a.cb0 ? cb : cb - a.cb0) * 64;
                if constexpr (kHalo) {
                    mbar_wait(bar_emptyA + sa * 8, pa ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_fullA + sa * 8, kACopy);
                        tma_load_4d(ringA + sa * kAStage, tm, bar_fullA + sa * 8, c0, x0 - 1, y0 - 1, b);
                    }
                    __syncwarp();
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                    if constexpr (!RES) {
                        // weights stream in groups of kBG taps (kBG x [NT x 64]) per stage
#pragma unroll 1
                        for (int g = 0; g < kHTaps / kBG; ++g) {
                            mbar_wait(bar_emptyB + sb * 8, pb ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(bar_fullB + sb * 8, kBStage);
#pragma unroll
                                for (int t = 0; t < kBG; ++t)
                                    tma_load_2d(ringB + sb * kBStage + t * kBTap, &a.tmB, bar_fullB + sb * 8, kcol + t * 64,
                                                nt * NT);
                            }
                            __syncwarp();
                            kcol += 64 * kBG;
                            if (++sb == b_stages) { sb = 0; pb ^= 1; }
                        }
                        if (kUp) kcol += 12 * 64;      // skip the other three phases of this channel block
                    }
                } else if (RES && a.toeplitz) {
                    // Toeplitz operand: the rows of ALL taps of a tile share one stage (TAPS x 2 176 B, kToepPitch apart), so a
                    // stage is a tile and a_stages tiles are in flight instead of a_stages / TAPS
                    mbar_wait(bar_emptyA + sa * 8, pa ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_fullA + sa * 8, TAPS * a.a_copy_bytes);
                        for (int tap = 0; tap < TAPS; ++tap)
                            tma_load_4d(ringA + sa * kAStage + tap * kToepPitch, tm, bar_fullA + sa * 8, 0, x0 >> 3, y0 + tap * a.tap_sy - a.pad_y, b);
                    }
                    __syncwarp();
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                } else {
                    int ky = 0, kx = 0;
                    for (int tap = 0; tap < TAPS; ++tap) {
                        mbar_wait(bar_emptyA + sa * 8, pa ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(bar_fullA + sa * 8, a.a_copy_bytes);
                            if constexpr (MODE == kS2D)
                                tma_load_5d(ringA + sa * kAStage, tm, bar_fullA + sa * 8, c0, kx, x0, ky, b * a.H + y0);
                            else
                                tma_load_4d(ringA + sa * kAStage, tm, bar_fullA + sa * 8, c0, x0 + kx - a.pad_x,
                                            y0 + ky * a.tap_sy - a.pad_y, b);
                        }
                        __syncwarp();
                        if (++sa == a_stages) { sa = 0; pa ^= 1; }
                        if (++kx == a.taps_x) { kx = 0; ++ky; }
                        if constexpr (!RES) {
                            mbar_wait(bar_emptyB + sb * 8, pb ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(bar_fullB + sb * 8, kBStage);
                                tma_load_2d(ringB + sb * kBStage, &a.tmB, bar_fullB + sb * 8, kcol, nt * NT);
                            }
                            __syncwarp();
                            kcol += 64;
                            if (++sb == b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================================================
        // The issuing warp is a serial instruction stream: every barrier wait, election and descriptor fix-up between two
        // tcgen05.mma's is dead time for the tensor pipe.  So one election covers a whole group of MMAs (36 for a resident
        // 64-channel block, 12 per streamed kernel row) whose descriptors differ from a base by compile-time immediates.
        constexpr uint32_t idesc = umma_idesc_bf16(128, NT);
        constexpr uint32_t hiB = umma_desc_hi(1024);
        constexpr uint32_t hiA = umma_desc_hi(kHalo ? kHaloPitch : 1024);
        int sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 0;
        if constexpr (RES) mbar_wait(bar_fullB, 0);
        for (int mt = m_begin; mt < m_end; ++mt) {
            const int py = kUp ? ((mt >> 1) & 1) : 0, px = kUp ? (mt & 1) : 0;      // phase = mt & 3 = py * 2 + px
            mbar_wait(bar_tempty + acc * 8, pacc ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kSub * NT;
            for (int cb = 0; cb < CB; ++cb) {
                if constexpr (kHalo) {
                    mbar_wait(bar_readyA + sa * 8, pa);
                    tc_fence_after();
                    const uint32_t a_base = umma_desc_lo(ringA + sa * kAStage);
                    if constexpr (RES && kUp) {
                        const uint32_t b_base = umma_desc_lo(ringB + (cb * 16 + (py * 2 + px) * 4) * kBTap);
                        if (elect_one()) {
                            uint32_t b_cur = b_base, accum = cb > 0 ? 1u : 0u;
#pragma unroll 1
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t a_cur = a_base + (((py + (t >> 1)) * kHaloPitch + (px + (t & 1)) * 128) >> 4);
                                umma_bf16_lohi_pred(d_tmem, a_cur, hiA, b_cur, hiB, idesc, accum);
                                umma_bf16_lohi<true>(d_tmem, a_cur + 2, hiA, b_cur + 2, hiB, idesc);
                                umma_bf16_lohi<true>(d_tmem, a_cur + 4, hiA, b_cur + 4, hiB, idesc);
                                umma_bf16_lohi<true>(d_tmem, a_cur + 6, hiA, b_cur + 6, hiB, idesc);
                                accum = 1u;
                                b_cur += kBTap >> 4;
                            }
                            umma_commit(bar_emptyA + sa * 8);
                        }
                        __syncwarp();
                    } else if constexpr (RES) {
                        const uint32_t b_base = umma_desc_lo(ringB + cb * kHTaps * kBTap);
                        if (elect_one()) {
                            // rolled tap loops on purpose: unrolled, ptxas hoists all 72 descriptor updates ahead of the
                            // first MMA (spilling uniform registers) and the tensor pipe idles meanwhile
                            uint32_t a_row = a_base, b_cur = b_base, accum = cb > 0 ? 1u : 0u;
#pragma unroll 1
                            for (int ky = 0; ky < 3; ++ky) {
                                uint32_t a_cur = a_row;
#pragma unroll 1
                                for (int kx = 0; kx < 3; ++kx) {
                                    if constexpr (WS) {
                                        umma_bf16_ws_pair(d_tmem, d_tmem + NT, a_cur, a_cur + kSubStep, hiA, b_cur, hiB, idesc, accum);
                                    } else {
#pragma unroll
                                    for (int sub = 0; sub < kSub; ++sub) {
                                        const uint32_t d = d_tmem + sub * NT, as = a_cur + sub * kSubStep;
                                        umma_bf16_lohi_pred(d, as, hiA, b_cur, hiB, idesc, accum);
                                        umma_bf16_lohi<true>(d, as + 2, hiA, b_cur + 2, hiB, idesc);
                                        umma_bf16_lohi<true>(d, as + 4, hiA, b_cur + 4, hiB, idesc);
                                        umma_bf16_lohi<true>(d, as + 6, hiA, b_cur + 6, hiB, idesc);
                                    }
                                    }
                                    accum = 1u;
                                    a_cur += 128 >> 4;
                                    b_cur += kBTap >> 4;
                                }
                                a_row += kHaloPitch >> 4;
                            }
                            if constexpr (kRes2) {      // 1x1 residual conv: centre tap of the same halo box, weight block 9
                                const uint32_t d2 = tmem_base + (2 + acc) * NT, ac = a_base + ((kHaloPitch + 128) >> 4);
                                umma_bf16_lohi_pred(d2, ac, hiA, b_cur, hiB, idesc, cb > 0 ? 1u : 0u);
                                umma_bf16_lohi<true>(d2, ac + 2, hiA, b_cur + 2, hiB, idesc);
                                umma_bf16_lohi<true>(d2, ac + 4, hiA, b_cur + 4, hiB, idesc);
                                umma_bf16_lohi<true>(d2, ac + 6, hiA, b_cur + 6, hiB, idesc);
                            }
                            umma_commit(bar_emptyA + sa * 8);
                        }
                        __syncwarp();
                    } else {
#pragma unroll 1
                        for (int g = 0; g < kHTaps / kBG; ++g) {
                            mbar_wait(bar_fullB + sb * 8, pb);
                            tc_fence_after();
                            const uint32_t b_base = umma_desc_lo(ringB + sb * kBStage);
                            if (elect_one()) {
                                uint32_t b_cur = b_base, accum = (cb > 0 || g > 0) ? 1u : 0u;
#pragma unroll 1
                                for (int t = 0; t < kBG; ++t) {
                                    const int tap = g * kBG + t;
                                    const bool rtap = kRes2 && tap == 9;      // fused 1x1 residual conv: centre tap -> 2nd accumulator
                                    const int ky = kUp ? py + (tap >> 1) : (rtap ? 1 : (tap * 11) >> 5);
                                    const int kx = kUp ? px + (tap & 1) : (rtap ? 1 : tap - 3 * ky);
                                    const uint32_t a_cur = a_base + ((ky * kHaloPitch + kx * 128) >> 4);
                                    if (rtap) accum = cb > 0 ? 1u : 0u;
                                    if constexpr (WS) {
                                        umma_bf16_ws_pair(d_tmem, d_tmem + NT, a_cur, a_cur + kSubStep, hiA, b_cur, hiB, idesc, accum);
                                    } else {
#pragma unroll
                                    for (int sub = 0; sub < kSub; ++sub) {
                                        const uint32_t d = rtap ? tmem_base + (2 + acc) * NT : d_tmem + sub * NT, as = a_cur + sub * kSubStep;
                                        umma_bf16_lohi_pred(d, as, hiA, b_cur, hiB, idesc, accum);
                                        umma_bf16_lohi<true>(d, as + 2, hiA, b_cur + 2, hiB, idesc);
                                        umma_bf16_lohi<true>(d, as + 4, hiA, b_cur + 4, hiB, idesc);
                                        umma_bf16_lohi<true>(d, as + 6, hiA, b_cur + 6, hiB, idesc);
                                    }
                                    }
                                    accum = 1u;
                                    b_cur += kBTap >> 4;
                                }
                                umma_commit(bar_emptyB + sb * 8);
                                if (g == kHTaps / kBG - 1) umma_commit(bar_emptyA + sa * 8);
                            }
                            __syncwarp();
                            if (++sb == b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                } else if (RES && a.toeplitz) {
                    // Toeplitz operand (init_conv): no swizzle, SBO = 128 B (the LBO field of umma_desc_lo is already 16 B); a
                    // K = 16 step is two 16-byte chunks = the same +2 on the start address as in the swizzled layout
                    constexpr uint32_t hiT = ((128u >> 4) & 0x3FFFu) | (1u << 14);
                    mbar_wait(bar_fullA + sa * 8, pa);
                    tc_fence_after();
                    const uint32_t a_base = umma_desc_lo(ringA + sa * kAStage);
                    const uint32_t b_base = umma_desc_lo(ringB + cb * TAPS * kBTap);
                    if (elect_one()) {
                        uint32_t a_lo = a_base, b_lo = b_base;
#pragma unroll 1
                        for (int tap = 0; tap < TAPS; ++tap) {
                            umma_bf16_lohi_pred(d_tmem, a_lo, hiT, b_lo, hiB, idesc, (cb > 0 || tap > 0) ? 1u : 0u);
                            umma_bf16_lohi<true>(d_tmem, a_lo + 2, hiT, b_lo + 2, hiB, idesc);
                            umma_bf16_lohi<true>(d_tmem, a_lo + 4, hiT, b_lo + 4, hiB, idesc);
                            umma_bf16_lohi<true>(d_tmem, a_lo + 6, hiT, b_lo + 6, hiB, idesc);
                            a_lo += kToepPitch >> 4;
                            b_lo += kBTap >> 4;
                        }
                        umma_commit(bar_emptyA + sa * 8);
                    }
                    __syncwarp();
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                } else {
                    for (int tap = 0; tap < TAPS; ++tap) {
                        mbar_wait(bar_fullA + sa * 8, pa);
                        if constexpr (!RES) mbar_wait(bar_fullB + sb * 8, pb);
                        tc_fence_after();
                        const uint32_t a_lo = umma_desc_lo(ringA + sa * kAStage);
                        const uint32_t b_lo = umma_desc_lo(ringB + (RES ? (cb * TAPS + tap) * kBTap : sb * kBStage));
                        if (elect_one()) {
                            umma_bf16_lohi_pred(d_tmem, a_lo, hiA, b_lo, hiB, idesc, (cb > 0 || tap > 0) ? 1u : 0u);
                            umma_bf16_lohi<true>(d_tmem, a_lo + 2, hiA, b_lo + 2, hiB, idesc);
                            umma_bf16_lohi<true>(d_tmem, a_lo + 4, hiA, b_lo + 4, hiB, idesc);
                            umma_bf16_lohi<true>(d_tmem, a_lo + 6, hiA, b_lo + 6, hiB, idesc);
                            if constexpr (!RES) umma_commit(bar_emptyB + sb * 8);
                            umma_commit(bar_emptyA + sa * 8);
                        }
                        __syncwarp();
                        if constexpr (!RES) { if (++sb == b_stages) { sb = 0; pb ^= 1; } }
                        if (++sa == a_stages) { sa = 0; pa ^= 1; }
                    }
                }
            }
            if (elect_one()) umma_commit(bar_tfull + acc * 8);
            __syncwarp();
            if (++acc == 2) { acc = 0; pacc ^= 1; }
        }
    } else if (XF && (warp == 2 || warp == 3 || warp >= kXfWarp0)) {
        // ================================ in-place GroupNorm-apply + SiLU on the landed activation stages ===============
        if constexpr (XF && kHalo && !kUp) {
            const int t = (warp >= kXfWarp0 ? warp - kXfWarp0 + 2 : warp - 2) * 32 + lane;           // 0..191
            const int o = t & 7;                                 // this thread's channel octet inside a 64-channel block
            constexpr int kBoxW = kHaloTW + 2;
            constexpr int kPix = (kSub * kHaloTH + 2) * kBoxW;   // pixels of one halo box
            const int Cin = CB * 64;
            int sa = 0, pa = 0, cur_b = -1;
            pdl_wait();                                          // statistics / scale-shift come from the previous kernels
            for (int mt = m_begin; mt < m_end; ++mt) {
                int m = mt;
                const int tx = m % a.tiles_x; m /= a.tiles_x;
                const int ty = m % a.tiles_y;
                const int b = m / a.tiles_y;
                if (b != cur_b) {
                    cur_b = b;
                    named_bar_sync(1, kXfWarps * 32);            // nobody still reads the previous sample's table
                    for (int c = t; c < Cin; c += kXfWarps * 32) {
                        const int g = c >> a.xf_lgs;
                        const double inv_n = 1.0 / (static_cast<double>(a.H) * a.W * (1 << a.xf_lgs) * a.xf_real_frac);
                        const double s_ = static_cast<double>(static_cast<long long>(a.xf_stats[(b * a.xf_G + g) * 2])) * (1.0 / 16777216.0);
                        const double q_ = static_cast<double>(static_cast<long long>(a.xf_stats[(b * a.xf_G + g) * 2 + 1])) * (1.0 / 16777216.0);
                        const double meand = s_ * inv_n;
                        const float mean = static_cast<float>(meand);
                        const float var = fmaxf(static_cast<float>(q_ * inv_n - meand * meand), 0.f);
                        const float rstd = rsqrtf(var + a.xf_eps);
                        float Aj = rstd * __ldg(a.xf_gamma + c);
                        float Bj = __ldg(a.xf_beta + c) - mean * Aj;
                        if (a.xf_ss) {
                            const float sc = a.xf_ss[static_cast<size_t>(b) * a.xf_ss_ld + c] + 1.0f;
                            const float sh = a.xf_ss[static_cast<size_t>(b) * a.xf_ss_ld + Cin + c];
                            Aj *= sc;
                            Bj = Bj * sc + sh;
                        }
                        // stored halved: SiLU(y) = h + h tanh(h) with h = y / 2 = fma(x, A/2, B/2) — exact, powers of two
                        reinterpret_cast<float*>(tail->coef)[c] = 0.5f * Aj;            // A / 2 for channels 0..511,
                        reinterpret_cast<float*>(tail->coef)[512 + c] = 0.5f * Bj;      // then B / 2
                    }
                    named_bar_sync(1, kXfWarps * 32);
                }
                const int gx0 = tx * a.TW - 1, gy0 = ty * a.TH - 1;
                for (int cb = 0; cb < CB; ++cb) {
                    // explicit ld.shared (a generic load of this table showed 16 wavefronts per instruction in ncu); the table is
                    // stored as two float arrays so that the eight channel octets of a warp read 8 x 32 contiguous bytes each
                    float A[8], Bc[8];
                    {
                        const uint32_t ca = smem_u32(&tail->coef[0]) + (cb * 64 + o * 8) * 4;
                        const uint4 a0 = lds128(ca), a1 = lds128(ca + 16), b0 = lds128(ca + 2048), b1 = lds128(ca + 2048 + 16);
                        A[0] = __uint_as_float(a0.x); A[1] = __uint_as_float(a0.y); A[2] = __uint_as_float(a0.z); A[3] = __uint_as_float(a0.w);
                        A[4] = __uint_as_float(a1.x); A[5] = __uint_as_float(a1.y); A[6] = __uint_as_float(a1.z); A[7] = __uint_as_float(a1.w);
                        Bc[0] = __uint_as_float(b0.x); Bc[1] = __uint_as_float(b0.y); Bc[2] = __uint_as_float(b0.z); Bc[3] = __uint_as_float(b0.w);
                        Bc[4] = __uint_as_float(b1.x); Bc[5] = __uint_as_float(b1.y); Bc[6] = __uint_as_float(b1.z); Bc[7] = __uint_as_float(b1.w);
                    }
                    mbar_wait(bar_fullA + sa * 8, pa);
                    const uint32_t base = ringA + sa * kAStage;
                    // batches of kXfBatch chunks per thread: all 16-byte loads of a batch are issued before any arithmetic, and
                    // the loop body is branch-free (loads are always inside the stage; only the store is predicated).
                    // Interior boxes (no pixel outside the image: 70 % of the tiles at 256 x 256) skip the per-pixel image test and
                    // walk the box with compile-time offsets — a pass advances 24 pixels, a multiple of 8, so the swizzle term of a
                    // thread's address never changes.
                    constexpr int kXfBatch = kSub == 2 ? 5 : 4;      // 15 = 3 x 5 (34 x 10 box) / 8 = 2 x 4 (18 x 10 box) passes of 24 pixels
                    constexpr int kIters = (kPix + kXfWarps * 4 - 1) / (kXfWarps * 4);
                    constexpr int kPassPix = kXfWarps * 4;
                    const bool interior = gy0 >= 0 && gx0 >= 0 && gy0 + (kSub * kHaloTH + 2) <= a.H && gx0 + kBoxW <= a.W;
                    const uint32_t tbase = base + (t >> 3) * 128 + ((o ^ ((t >> 3) & 7)) << 4);
                    auto act8 = [&](const uint4& u) {
                        const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
                        auto act = [](float x, float ah, float bh) {
                            const float h = fmaf(x, ah, bh);
                            return fmaf(h, tanh_approx(h), h);
                        };
                        uint4 w;
                        w.x = pack_bf16(act(f0.x, A[0], Bc[0]), act(f0.y, A[1], Bc[1]));
                        w.y = pack_bf16(act(f1.x, A[2], Bc[2]), act(f1.y, A[3], Bc[3]));
                        w.z = pack_bf16(act(f2.x, A[4], Bc[4]), act(f2.y, A[5], Bc[5]));
                        w.w = pack_bf16(act(f3.x, A[6], Bc[6]), act(f3.y, A[7], Bc[7]));
                        return w;
                    };
                    if (interior) {
                        // every address is tbase + a compile-time offset (the last pass reads up to 2.4 KB past the box — still this
                        // CTA's shared memory, the next stage or the weight ring — and stores only its live pixels), and the batches
                        // are software-pipelined: the next batch's loads are in flight while this one is transformed (the kernel's
                        // register budget is set by the epilogue warps, so the second buffer is free).  Measured in-step on B200:
                        // XF / plain time of the 64 -> 64 layers 1.39 -> 1.33 with the interior path, no further change from the
                        // pipelining — what the XF form costs is its shared-memory traffic (the box is landed, read and rewritten:
                        // 491 KB per 256-pixel tile through the 128 B/clk pipe against 404 KB for the plain form), not instructions
                        // (profiles/ncu_r2_conv64_roles.json)
                        static_assert(kIters % kXfBatch == 0, "whole batches");
                        constexpr int kBatches = kIters / kXfBatch;
                        uint4 u[2][kXfBatch];
#pragma unroll
                        for (int k = 0; k < kXfBatch; ++k) u[0][k] = lds128(tbase + k * (kPassPix * 128));
#pragma unroll
                        for (int bi = 0; bi < kBatches; ++bi) {
                            if (bi + 1 < kBatches) {
#pragma unroll
                                for (int k = 0; k < kXfBatch; ++k) u[(bi + 1) & 1][k] = lds128(tbase + ((bi + 1) * kXfBatch + k) * (kPassPix * 128));
                            }
#pragma unroll
                            for (int k = 0; k < kXfBatch; ++k) {
                                const int pass = bi * kXfBatch + k;
                                const uint4 w = act8(u[bi & 1][k]);
                                if (pass * kPassPix + kPassPix <= kPix || (t >> 3) + pass * kPassPix < kPix)      // first term: compile time
                                    sts128(tbase + pass * (kPassPix * 128), w);
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int i0 = 0; i0 < kIters; i0 += kXfBatch) {
                            uint4 u[kXfBatch];
                            uint32_t addr[kXfBatch];
                            bool ok[kXfBatch];
#pragma unroll
                            for (int k = 0; k < kXfBatch; ++k) {
                                const int pr = (t >> 3) + (i0 + k) * kPassPix;
                                const int p = pr < kPix ? pr : kPix - 1;
                                const int hy = (p * 205) >> 11;          // p / 10 for p < 1024
                                const int hx = p - hy * kBoxW;
                                ok[k] = pr < kPix && static_cast<unsigned>(gy0 + hy) < static_cast<unsigned>(a.H) &&
                                        static_cast<unsigned>(gx0 + hx) < static_cast<unsigned>(a.W);
                                addr[k] = base + p * 128 + ((o ^ (p & 7)) << 4);
                                u[k] = lds128(addr[k]);
                            }
#pragma unroll
                            for (int k = 0; k < kXfBatch; ++k) {
                                const uint4 w = act8(u[k]);
                                if (ok[k]) sts128(addr[k], w);
                            }
                        }
                    }
                    fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_xfA + sa * 8);
                    if (++sa == a_stages) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
        // ================================ epilogue =============================================================
        // Warp w may read TMEM lanes 32*(w%4)..+31 (= 32 pixels of the tile); the two warps of a lane quarter split the
        // tile's 32-column chunks.  GroupNorm partial sums (per 8 columns) stay in registers across tiles and are flushed
        // as fixed-point atomics only when the sample index changes — no per-tile shuffles, barriers or atomics.
        constexpr int kSlots = NT / 64;         // chunks per warp per tile
        const int q = warp & 3;
        const int cset = (warp - kEpiWarp0) >> 2;
        const int n0 = nt * NT;
        float sacc[kSlots][4], qacc[kSlots][4];
#pragma unroll
        for (int k = 0; k < kSlots; ++k)
#pragma unroll
            for (int g = 0; g < 4; ++g) { sacc[k][g] = 0.f; qacc[k][g] = 0.f; }
        int stat_b = -1;
        auto flush_stats = [&](int bb) {
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float s_ = warp_sum(sacc[k][g]), q_ = warp_sum(qacc[k][g]);
                    if (lane == 0) {
                        const int gl = (n0 + (cset + 2 * k) * 32 + g * 8) >> a.lgs;
                        unsigned long long* dst = a.stats + (static_cast<size_t>(bb) * a.G + gl) * 2;
                        atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn(s_ * kStatScale)));
                        atomicAdd(dst + 1, static_cast<unsigned long long>(__float2ll_rn(q_ * kStatScale)));
                    }
                    sacc[k][g] = 0.f; qacc[k][g] = 0.f;
                }
            }
        };
        int acc = 0, pacc = 0;
        for (int mt = m_begin; mt < m_end; ++mt) {
            int m = mt;
            const int phase = kUp ? (m & 3) : 0;
            if (kUp) m >>= 2;
            const int tx = m % a.tiles_x; m /= a.tiles_x;
            const int ty = m % a.tiles_y;
            const int b = m / a.tiles_y;
            const int r = q * 32 + lane;
            if (a.stats && b != stat_b) {
                if (stat_b >= 0) flush_stats(stat_b);
                stat_b = b;
            }

            constexpr int kPasses = kRes2 ? 2 : kSub;      // kHalo1R: pass 1 drains the residual-conv accumulator
            if (a.res) {
                // the residual rows this lane adds in the epilogue are known before the accumulators are: pull them into L1 while
                // the MMAs of the tile are still running (the loads below then hit instead of exposing a trip to L2 / HBM per slot)
#pragma unroll
                for (int sub = 0; sub < (kRes2 ? 1 : kSub); ++sub) {
                    const int y = ty * a.TH + sub * kHaloTH + (r >> a.lgTW), x = tx * a.TW + (r & (a.TW - 1));
                    if (y < a.H && x < a.W) {
                        const size_t pix = kUp ? (static_cast<size_t>(b) * 2 * a.H + 2 * y + (phase >> 1)) * (2 * a.W) + 2 * x + (phase & 1)
                                               : (static_cast<size_t>(b) * a.H + y) * a.W + x;
#pragma unroll
                        for (int k = 0; k < kSlots; ++k)
                            prefetch_l1(a.res + pix * a.res_ld + n0 + (cset + 2 * k) * 32);
                    }
                }
            }
            mbar_wait(bar_tfull + acc * 8, pacc);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < kPasses; ++sub) {
                const bool pass2 = kRes2 && sub == 1;
                const int y = ty * a.TH + (kRes2 ? 0 : sub) * kHaloTH + (r >> a.lgTW), x = tx * a.TW + (r & (a.TW - 1));
                const bool valid = (y < a.H) && (x < a.W);
                // kHaloUp writes output pixel (2y + py, 2x + px) of the [B, 2H, 2W] grid
                const size_t pix = kUp ? (static_cast<size_t>(b) * 2 * a.H + 2 * y + (phase >> 1)) * (2 * a.W) + 2 * x + (phase & 1)
                                       : (static_cast<size_t>(b) * a.H + y) * a.W + x;
                uint32_t raw[kSlots][32];
#pragma unroll
                for (int k = 0; k < kSlots; ++k)
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (pass2 ? (2 + acc) : (acc * kSub + sub)) * NT +
                                  (cset + 2 * k) * 32,
                              raw[k]);
                tmem_ld_wait();
                if (sub == kPasses - 1) {
                    // everything this warp needs from the accumulators is in registers: hand the TMEM stage back right away
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + acc * 8);
                }
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    const int ch = cset + 2 * k;
                    const int nbase = n0 + ch * 32;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(pass2 ? &tail->bias2[ch * 32 + j] : &tail->bias[ch * 32 + j]);
                        v[j] = __uint_as_float(raw[k][j]) + bv.x; v[j + 1] = __uint_as_float(raw[k][j + 1]) + bv.y;
                        v[j + 2] = __uint_as_float(raw[k][j + 2]) + bv.z; v[j + 3] = __uint_as_float(raw[k][j + 3]) + bv.w;
                    }
                    if (a.act == kActGelu && !pass2) {
                        // packed-fp16 tanh form: one MUFU per PAIR of elements (the ex2 + rcp form needs four); its error
                        // (~5e-4 relative) is below the bf16 rounding of the store
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const uint32_t hh = gelu_f16x2(v[j], v[j + 1]);
                            const float2 ff = __half22float2(*reinterpret_cast<const __half2*>(&hh));
                            v[j] = ff.x; v[j + 1] = ff.y;
                        }
                    }
                    if (a.vec && !pass2) {
                        const float* vp = a.vec + static_cast<size_t>(b) * a.vec_ld + nbase;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(vp + j));
                            v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                        }
                    }
                    if (a.res && valid && !pass2) {
                        const __nv_bfloat16* rp = a.res + pix * a.res_ld + nbase;
                        const uint8x r0 = ldg256(rp), r1 = ldg256(rp + 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = j == 0 ? r0.lo : (j == 1 ? r0.hi : (j == 2 ? r1.lo : r1.hi));
                            float2 f;
                            f = unpack_bf16(u.x); v[j * 8 + 0] += f.x; v[j * 8 + 1] += f.y;
                            f = unpack_bf16(u.y); v[j * 8 + 2] += f.x; v[j * 8 + 3] += f.y;
                            f = unpack_bf16(u.z); v[j * 8 + 4] += f.x; v[j * 8 + 5] += f.y;
                            f = unpack_bf16(u.w); v[j * 8 + 6] += f.x; v[j * 8 + 7] += f.y;
                        }
                    }
                    if (a.stats && valid && !pass2) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float t0 = 0.f, t1 = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) { const float w = v[g * 8 + j]; t0 += w; t1 = fmaf(w, w, t1); }
                            sacc[k][g] += t0; qacc[k][g] += t1;
                        }
                    }
                    uint4 u[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        u[j].x = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]);
                        u[j].y = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
                        u[j].z = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]);
                        u[j].w = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
                    }
                    if (NT == 64 && a.stage_bytes) {
                        // Staged store: lane r owns row r of the sub-tile's [128 pixels][64 channels] block (SWIZZLE_128B, the order
                        // of the tensor-map box: y, x, channel) and writes its 32 channels as four conflict-free 16-byte chunks;
                        // one elected lane then moves the whole block out with a TMA store (which also clips pixels outside the
                        // image).  A lane-owned row would otherwise leave as two 32-byte sectors per lane, every lane of a store
                        // instruction in a different 128-byte line.
                        const int ew = warp - kEpiWarp0;
                        if (ew == 0) {      // the previous block's store must have read the staging block before it is rewritten
                            if (elect_one()) tma_store_wait_read();
                            __syncwarp();
                        }
                        named_bar_sync(2, kEpiWarps * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) sts128(stageS + r * 128 + (((ch * 4 + j) ^ (r & 7)) << 4), u[j]);
                        fence_proxy_async();
                        named_bar_sync(3, kEpiWarps * 32);
                        if (ew == 0) {
                            if (elect_one()) {
                                tma_store_4d(pass2 ? &a.tmOut2 : &a.tmOut, stageS, n0, tx * a.TW,
                                             ty * a.TH + (kRes2 ? 0 : sub) * kHaloTH, b);
                                tma_store_commit();
                            }
                            __syncwarp();
                        }
                    } else if (valid) {
                        __nv_bfloat16* op = pass2 ? a.out2 + pix * a.out2_ld + nbase : a.out + pix * a.out_ld + nbase;
                        stg256(op, u[0], u[1]);          // two full 32-byte sectors per lane
                        stg256(op + 16, u[2], u[3]);
                    }
                }
            }
            if (++acc == 2) { acc = 0; pacc ^= 1; }
        }
        if (a.stats && stat_b >= 0) flush_stats(stat_b);
        if (a.stage_bytes && warp == kEpiWarp0) {      // the last block must have left shared memory before the CTA exits
            if (elect_one()) tma_store_wait_all();
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kAccCols>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Weight-stationary MMA form of the N = 64 kHalo2 kernels: ON by default since round 2 (measured on B200, in-step: the plain
// 64 -> 64 layers at 256^2 300 -> 282 us, the XF ones 395 -> 385 us, step 23.03 -> 22.74 ms; parity test
// tests/test_gpu_ops.py::test_weight_stationary_conv_variant).  NDIFF_NO_WS=1 switches back to the plain form (A/B measurements).
bool ws_enabled() {
    static const bool on = [] { const char* v = getenv("NDIFF_NO_WS"); return !(v && v[0] == '1'); }();
    return on;
}

// Staged TMA store in the epilogue of the N = 64 kernels (default on; NDIFF_NO_STAGED_STORE=1 for A/B measurements)
bool staged_store_enabled() {
    static const bool on = [] { const char* v = getenv("NDIFF_NO_STAGED_STORE"); return !(v && v[0] == '1'); }();
    return on;
}

int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace

int encode_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    NDIFF_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the CUDA driver");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::string s = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (rank " +
                        std::to_string(rank) + ", dims";
        for (int i = 0; i < rank; ++i) s += " " + std::to_string(dims[i]);
        s += ", strides";
        for (int i = 0; i + 1 < rank; ++i) s += " " + std::to_string(strides_bytes[i]);
        s += ", box";
        for (int i = 0; i < rank; ++i) s += " " + std::to_string(box[i]);
        s += ")";
        set_error(s);
        return 1;
    }
    return 0;
}

int conv_gemm_plan(const ConvGemmDesc& d, int num_sms, ConvGemmPlan* plan) {
    {   // opt in to > 48 KB dynamic shared memory once per process (plans are never built under stream capture)
        static std::once_flag once;
        static int init_rc = 0;
        std::call_once(once, [] { init_rc = conv_gemm_init(); });
        if (init_rc) return 1;
    }
    ConvGemmArgs& a = plan->args;
    memset(&a, 0, sizeof(a));
    NDIFF_REQUIRE(d.C0 > 0 && d.C0 % 64 == 0 && d.C1 % 64 == 0, "channel counts must be multiples of 64");
    NDIFF_REQUIRE(d.Cout % 64 == 0, "C_out must be a multiple of 64");
    int NT = d.force_nt ? d.force_nt : (d.Cout % 128 == 0 ? 128 : 64);
    NDIFF_REQUIRE((NT == 64 || NT == 128) && d.Cout % NT == 0, "unsupported N tile");
    plan->NT = NT;
    a.mode = d.mode;
    a.B = d.B; a.H = d.H; a.W = d.W;
    int TW = d.TW;
    if (d.toeplitz) {
        NDIFF_REQUIRE(d.mode == kDirect && d.custom_src0 && d.C0 == 64 && d.C1 == 0 && d.W % 128 == 0 && d.taps_x == 1,
                      "Toeplitz operand: kDirect, one 64-value window per tap, W a multiple of 128");
        TW = 128;
    }
    if (TW == 0) {
        if (d.mode == kHalo1 || d.mode == kHalo2 || d.mode == kHaloUp || d.mode == kHalo1R) TW = kHaloTW;
        else TW = d.W >= 64 ? 64 : (d.W >= 32 ? 32 : (d.W >= 16 ? 16 : 8));
    }
    NDIFF_REQUIRE(TW >= 8 && TW <= 128 && (TW & (TW - 1)) == 0, "tile width must be a power of two in [8,128]");
    const int sub = d.mode == kHalo2 ? 2 : 1;
    a.TW = TW; a.TH = sub * (128 / TW); a.lgTW = ilog2(TW);
    a.tiles_x = (d.W + a.TW - 1) / a.TW;
    a.tiles_y = (d.H + a.TH - 1) / a.TH;
    a.cb0 = d.C0 / 64; a.cb1 = d.C1 / 64;
    const bool halo1 = d.mode == kHalo1 || d.mode == kHalo2 || d.mode == kHaloUp || d.mode == kHalo1R;
    const bool up = d.mode == kHaloUp;
    const bool res2 = d.mode == kHalo1R;
    NDIFF_REQUIRE(d.mode == kDirect || d.mode == kS2D || halo1, "unknown convolution mode");
    NDIFF_REQUIRE(!halo1 || TW == kHaloTW, "halo mode needs TW == 8 (one 8-row UMMA group per output row)");
    if (up) { a.taps_y = 4; a.taps_x = 4; a.pad_y = 1; a.pad_x = 1; }      // 16 weight blocks per channel block (4 phases x 4 taps)
    else if (halo1) { a.taps_y = 3; a.taps_x = 3; a.pad_y = 1; a.pad_x = 1; }
    else if (d.mode == kS2D) { a.taps_y = 2; a.taps_x = 2; a.pad_y = 0; a.pad_x = 0; }
    else { a.taps_y = d.taps_y; a.taps_x = d.taps_x; a.pad_y = d.pad_y; a.pad_x = d.pad_x; }
    a.tap_sy = d.tap_sy > 0 ? d.tap_sy : 1;
    a.toeplitz = d.toeplitz ? 1 : 0;

    a.n_tiles = d.Cout / NT;
    a.total_tiles = d.B * a.tiles_y * a.tiles_x * a.n_tiles * (up ? 4 : 1);
    const int b_tap = NT * 128;                          // one [NT x 64] weight block
    const int b_stage = (up ? kUpBTaps : (res2 ? kResBTaps : (halo1 ? halo_btaps(NT, sub) : 1))) * b_tap;   // streamed weights: bytes per ring stage
    if (d.mode == kHalo2) {
        a.a_copy_bytes = kHaloCopy2;
        a.a_stage_bytes = kHaloStage2;
        a.a_stages = NT == 128 ? 2 : 3;
        a.b_stages = NT == 128 ? 8 : 3;
    } else if (up) {
        a.a_copy_bytes = kHaloCopy;
        a.a_stage_bytes = kHaloStage;
        a.a_stages = 3;
        a.b_stages = NT == 128 ? 2 : 4;
    } else if (halo1) {
        a.a_copy_bytes = kHaloCopy;
        a.a_stage_bytes = kHaloStage;
        a.a_stages = 3;
        a.b_stages = res2 ? (NT == 128 ? 4 : 6) : (NT == 128 ? 3 : 5);
    } else {
        a.a_copy_bytes = d.toeplitz ? (128 + 8) * 16 : 128 * 128;
        a.a_stage_bytes = 128 * 128;
        a.a_stages = NT == 128 ? 6 : 8;
        a.b_stages = a.a_stages;
    }
    NDIFF_REQUIRE(a.a_stage_bytes % 1024 == 0, "operand stages must stay 1024-B aligned");
    // grid: persistent CTAs, a multiple of n_tiles so that every CTA keeps one N tile for its whole life
    int grid = a.total_tiles < num_sms ? a.total_tiles : num_sms;
    grid = grid / a.n_tiles * a.n_tiles;
    if (grid == 0) grid = a.n_tiles;
    plan->grid = grid;
    // weights resident in shared memory when this CTA's slice fits next to >= 2 activation stages and is reused
    const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(SmemTail));
    const int n_kb = (a.cb0 + a.cb1) * (a.taps_y * a.taps_x + (res2 ? 1 : 0));      // kHalo1R: + the 1x1 residual-conv block
    const int slice = n_kb * b_tap;
    const int m_per_cta = (a.total_tiles / a.n_tiles + grid / a.n_tiles - 1) / (grid / a.n_tiles);
    a.b_resident = (slice + 2 * a.a_stage_bytes <= budget && m_per_cta >= 2 && slice < (1 << 20)) ? 1 : 0;
    if (a.b_resident) {
        int st = (budget - slice) / a.a_stage_bytes;
        a.a_stages = st > 6 ? 6 : st;
        a.b_stages = 1;
        a.b_region_bytes = slice;
    } else {
        a.b_region_bytes = a.b_stages * b_stage;
    }
    NDIFF_REQUIRE(!d.toeplitz || (a.b_resident && a.taps_y * kToepPitch <= a.a_stage_bytes),
                  "Toeplitz operand: needs the resident-weight form (a stage holds the rows of all taps of a tile)");
    plan->smem_bytes = 1024 + a.a_stages * a.a_stage_bytes + a.b_region_bytes + static_cast<int>(sizeof(SmemTail));
    // staged epilogue (see ConvGemmArgs::stage_bytes): the N = 64 XF kernels with 16 KB of shared memory to spare.  Measured on one
    // box (profiles/staged_store_ab_r2.json): the XF layers, whose LSU pipe carries the in-place transform AND the stores, 408 ->
    // 367 us; the plain forms are bound by the tensor core's operand reads and do not move (302 -> 305 us), so they keep the
    // per-lane stores.  NDIFF_NO_STAGED_STORE=1 switches it off (A/B measurements).
    a.stage_bytes = 0;
    if (NT == 64 && !up && d.xf_stats != nullptr && staged_store_enabled() && plan->smem_bytes + 128 * 128 <= 227 * 1024 &&
        (a.a_stages * a.a_stage_bytes + a.b_region_bytes) % 1024 == 0) {
        a.stage_bytes = 128 * 128;
        plan->smem_bytes += a.stage_bytes;
        for (int o = 0; o < 2; ++o) {
            __nv_bfloat16* dst = o == 0 ? d.out : d.out2;
            const int ld = o == 0 ? d.out_ld : d.out2_ld;
            if (!dst) continue;
            uint64_t dims[4] = {static_cast<uint64_t>(d.Cout), static_cast<uint64_t>(d.W) * (up ? 2 : 1), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
            uint64_t str[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(d.W) * ld * 2, static_cast<uint64_t>(d.H) * d.W * ld * 2};
            uint32_t box[4] = {64, static_cast<uint32_t>(a.TW), static_cast<uint32_t>(128 / a.TW), 1};
            if (encode_tensor_map(o == 0 ? &a.tmOut : &a.tmOut2, dst, 4, dims, str, box, true)) return 1;
        }
    }
    NDIFF_REQUIRE(plan->smem_bytes <= 227 * 1024, "shared-memory budget exceeded");

    // ---- tensor maps -------------------------------------------------------------------------------------
    const int Hin = d.mode == kS2D ? 2 * d.H : d.H, Win = d.mode == kS2D ? 2 * d.W : d.W;
    for (int s = 0; s < 2; ++s) {
        const __nv_bfloat16* src = s == 0 ? d.src0 : d.src1;
        const int C = s == 0 ? d.C0 : d.C1;
        if (C == 0) continue;
        NDIFF_REQUIRE(src != nullptr, "null activation source");
        CUtensorMap* tm = s == 0 ? &a.tmA0 : &a.tmA1;
        if (s == 0 && d.toeplitz) {
            // one row of 136 sixteen-byte pixels, landed densely (no swizzle) as 17 box rows of 128 B (TMA moves 16-byte box rows
            // one request at a time; the tensor is contiguous along the row, so eight pixels make one 64-element inner row)
            NDIFF_REQUIRE(d.cdim[0] == 8 && d.cdim[1] % 8 == 0, "Toeplitz operand: rows of 16-byte pixels, a multiple of 8 per row");
            uint64_t dims[4] = {64, d.cdim[1] / 8, d.cdim[2], d.cdim[3]};
            uint64_t str[3] = {128, d.cstride[1], d.cstride[2]};
            uint32_t box[4] = {64, (128 + 8) / 8, 1, 1};
            if (encode_tensor_map(tm, src, 4, dims, str, box, false)) return 1;
        } else if (s == 0 && d.custom_src0) {
            uint32_t box[4] = {64, static_cast<uint32_t>(a.TW), static_cast<uint32_t>(a.TH), 1};
            if (encode_tensor_map(tm, src, 4, d.cdim, d.cstride, box, true)) return 1;
        } else if (d.mode == kS2D) {
            // input [B, 2H, 2W, C] viewed as (C, p2, W, p1, B*H)
            uint64_t dims[5] = {static_cast<uint64_t>(C), 2, static_cast<uint64_t>(d.W), 2,
                                static_cast<uint64_t>(d.B) * d.H};
            uint64_t str[4] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * 4,
                               static_cast<uint64_t>(Win) * C * 2, static_cast<uint64_t>(Win) * C * 4};
            uint32_t box[5] = {64, 1, static_cast<uint32_t>(a.TW), 1, static_cast<uint32_t>(a.TH)};
            if (encode_tensor_map(tm, src, 5, dims, str, box, true)) return 1;
        } else {
            uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                                static_cast<uint64_t>(d.B)};
            uint64_t str[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(Win) * C * 2,
                               static_cast<uint64_t>(Hin) * Win * C * 2};
            uint32_t box[4] = {64, static_cast<uint32_t>(halo1 ? a.TW + 2 : a.TW),
                               static_cast<uint32_t>(halo1 ? a.TH + 2 : a.TH), 1};
            if (encode_tensor_map(tm, src, 4, dims, str, box, true)) return 1;
        }
    }
    {
        const uint64_t Ktot = static_cast<uint64_t>(a.cb0 + a.cb1) * (a.taps_y * a.taps_x + (res2 ? 1 : 0)) * 64;
        uint64_t dims[2] = {Ktot, static_cast<uint64_t>(d.Cout)};
        uint64_t str[1] = {Ktot * 2};
        uint32_t box[2] = {64, static_cast<uint32_t>(NT)};
        NDIFF_REQUIRE(d.weight != nullptr, "null weight");
        if (encode_tensor_map(&a.tmB, d.weight, 2, dims, str, box, true)) return 1;
    }
    a.bias = d.bias; a.vec = d.vec; a.vec_ld = d.vec_ld; a.res = d.res; a.res_ld = d.res_ld;
    a.out = d.out; a.out_ld = d.out_ld; a.act = d.act;
    a.bias2 = d.bias2; a.out2 = d.out2; a.out2_ld = d.out2_ld;
    plan->xf = d.xf_stats != nullptr;
    plan->ws = d.mode == kHalo2 && NT == 64 && ws_enabled();      // see the comment above conv_gemm_kernel
    if (plan->xf) {
        NDIFF_REQUIRE((d.mode == kHalo1 || d.mode == kHalo2) && d.C1 == 0 && d.C0 <= 512, "fused GroupNorm input: single-source 3x3 conv with C_in <= 512");
        NDIFF_REQUIRE(d.xf_gamma && d.xf_beta && d.xf_groups > 0 && d.C0 % d.xf_groups == 0, "fused GroupNorm input: bad arguments");
        const int gs = d.C0 / d.xf_groups;
        NDIFF_REQUIRE(gs >= 8 && (gs & (gs - 1)) == 0, "fused GroupNorm input: group size must be a power of two >= 8");
        a.xf_stats = d.xf_stats; a.xf_gamma = d.xf_gamma; a.xf_beta = d.xf_beta; a.xf_ss = d.xf_ss; a.xf_ss_ld = d.xf_ss_ld;
        a.xf_G = d.xf_groups; a.xf_lgs = ilog2(gs); a.xf_eps = d.xf_eps;
        NDIFF_REQUIRE(d.xf_real_frac > 0.f && d.xf_real_frac <= 1.f, "fused GroupNorm input: live channel fraction must be in (0, 1]");
        a.xf_real_frac = d.xf_real_frac;
    }
    NDIFF_REQUIRE(!res2 || (d.out2 != nullptr && d.out2_ld % 8 == 0), "kHalo1R needs the residual-conv output");
    a.stats = d.stats; a.G = d.groups;
    if (d.stats) {
        const int gs = d.Cout / d.groups;
        NDIFF_REQUIRE(gs >= 8 && (gs & (gs - 1)) == 0 && gs <= NT, "GroupNorm group size must be a power of two in [8, NT]");
        a.lgs = ilog2(gs);
    }
    NDIFF_REQUIRE(d.out != nullptr && d.out_ld % 8 == 0, "output must be 16-byte aligned per pixel");
    return 0;
}

namespace {
template <int NT, int MODE, bool RES>
int launch_one(const ConvGemmPlan& plan, cudaStream_t stream) {
    if constexpr (MODE == kHalo2 && NT == 64) {
        if (plan.ws) {
            if (plan.xf)
                NDIFF_CUDA_OK(launch_pdl(conv_gemm_kernel<NT, MODE, RES, true, true>, dim3(plan.grid), dim3(kThreadsXf), plan.smem_bytes,
                                         stream, plan.args));
            else
                NDIFF_CUDA_OK(launch_pdl(conv_gemm_kernel<NT, MODE, RES, false, true>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes,
                                         stream, plan.args));
            return 0;
        }
    }
    if constexpr (MODE == kHalo1 || MODE == kHalo2) {
        if (plan.xf) {
            NDIFF_CUDA_OK(launch_pdl(conv_gemm_kernel<NT, MODE, RES, true>, dim3(plan.grid), dim3(kThreadsXf), plan.smem_bytes, stream,
                                     plan.args));
            return 0;
        }
    }
    NDIFF_CUDA_OK(launch_pdl(conv_gemm_kernel<NT, MODE, RES>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args));
    return 0;
}
template <int NT, int MODE>
int launch_res(const ConvGemmPlan& plan, cudaStream_t stream) {
    return plan.args.b_resident ? launch_one<NT, MODE, true>(plan, stream) : launch_one<NT, MODE, false>(plan, stream);
}
template <int NT, int MODE>
cudaError_t opt_in() {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if constexpr (MODE == kHalo1 || MODE == kHalo2) {
        e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    if constexpr (MODE == kHalo2 && NT == 64) {
      {
        e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
      }
    }
    return cudaFuncSetAttribute(conv_gemm_kernel<NT, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
}  // namespace

int conv_gemm_init() {
    NDIFF_CUDA_OK((opt_in<64, kDirect>()));
    NDIFF_CUDA_OK((opt_in<128, kDirect>()));
    NDIFF_CUDA_OK((opt_in<64, kS2D>()));
    NDIFF_CUDA_OK((opt_in<128, kS2D>()));
    NDIFF_CUDA_OK((opt_in<64, kHalo1>()));
    NDIFF_CUDA_OK((opt_in<128, kHalo1>()));
    NDIFF_CUDA_OK((opt_in<64, kHalo2>()));
    NDIFF_CUDA_OK((opt_in<128, kHalo2>()));
    NDIFF_CUDA_OK((opt_in<64, kHaloUp>()));
    NDIFF_CUDA_OK((opt_in<128, kHaloUp>()));
    NDIFF_CUDA_OK((opt_in<64, kHalo1R>()));
    NDIFF_CUDA_OK((opt_in<128, kHalo1R>()));
    return 0;
}

int conv_gemm_launch(const ConvGemmPlan& plan, cudaStream_t stream) {
    const int mode = plan.args.mode;
    if (plan.NT == 64) {
        if (mode == kHaloUp) return launch_res<64, kHaloUp>(plan, stream);
        if (mode == kHalo2) return launch_res<64, kHalo2>(plan, stream);
        if (mode == kHalo1R) return launch_res<64, kHalo1R>(plan, stream);
        if (mode == kHalo1) return launch_res<64, kHalo1>(plan, stream);
        if (mode == kS2D) return launch_res<64, kS2D>(plan, stream);
        return launch_res<64, kDirect>(plan, stream);
    }
    if (mode == kHaloUp) return launch_res<128, kHaloUp>(plan, stream);
    if (mode == kHalo2) return launch_res<128, kHalo2>(plan, stream);
    if (mode == kHalo1R) return launch_res<128, kHalo1R>(plan, stream);
    if (mode == kHalo1) return launch_res<128, kHalo1>(plan, stream);
    if (mode == kS2D) return launch_res<128, kS2D>(plan, stream);
    return launch_res<128, kDirect>(plan, stream);
}

}  // namespace ndiff
