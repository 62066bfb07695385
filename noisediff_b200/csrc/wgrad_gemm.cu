// noisediff_b200 — convolution weight gradients on tcgen05 with MN-major operands (sm_100a).  See wgrad_gemm.cuh.
//
// One CTA owns a (128 output channels) x (64 input channels) x (tap group) block of dW and a contiguous range of 16 x 8 pixel
// tiles (split-K over the pixels; partial results meet in fp32 atomics).  Per tile the TMA producer lands the dY tile
// (two 64-channel blocks: 128 pixel rows x 128 B each) and ONE box of X (18 x 10 halo for 3x3); the MMA warp walks the tile
// in K = 16 pixel steps (= two image rows of the 8-wide tile) and issues, per step and tap,
//     D[128 co x 64 ci] += dY[16 px x 128 co]^T * X_tap[16 px x 64 ci]      (tcgen05.mma, a_major = b_major = MN)
// where X_tap is the halo box shifted by (ky, kx) pixel rows — an address offset, exactly like the forward conv's taps.
// Accumulators (one 64-column block per tap of the group) stay in TMEM for the CTA's whole pixel range; four epilogue warps
// drain them once at the end.
#include "wgrad_gemm.cuh"
#include "conv_gemm.cuh"

#include <cstring>
#include <mutex>

namespace ndiff {

namespace {

constexpr int kWgThreads = 256;           // w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4..7 epilogue
constexpr int kDyBlk = 128 * 128;         // one [128 pixels x 64 channels] bf16 block
constexpr int kHaloPitchW = 10 * 128;     // bytes between halo rows (8 + 2 pixels)
constexpr int kHaloCopyW = 18 * kHaloPitchW;                          // 23040 B landed by the 3x3 X box
constexpr int kHaloStageW = (kHaloCopyW + 1023) / 1024 * 1024;
constexpr int kWgTmemCols = 256;          // >= 3 taps x 64 columns

struct WgTail {
    uint64_t full[4], empty[4], tmem_full;
    uint32_t tmem_base;
    uint32_t pad_;
};

// descriptor halves for an MN-major SWIZZLE_128B operand: lo = start >> 4 | (LBO >> 4) << 16; hi = SBO >> 4 | version | layout.
// LBO = byte distance between consecutive 64-element chunks along M / N, SBO = between consecutive 8-row groups along K.
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}

template <int MODE>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_gemm_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __builtin_assume(__isShared(smem));      // the manual alignment hides the address space: without the hint table reads are generic LD.E
    constexpr bool kHalo = MODE == kWg3x3;
    constexpr int kXBytes = kHalo ? kHaloStageW : kDyBlk;
    constexpr int kXCopy = kHalo ? kHaloCopyW : kDyBlk;
    constexpr int kStage = 2 * kDyBlk + kXBytes;
    const int stages = a.stages;
    WgTail* tail = reinterpret_cast<WgTail*>(smem + stages * kStage);
    const uint32_t ring = smem_u32(smem);
    const uint32_t bar_full = smem_u32(&tail->full[0]), bar_empty = smem_u32(&tail->empty[0]), bar_tfull = smem_u32(&tail->tmem_full);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int units = a.co_blocks * a.ci_blocks * a.tap_groups;
    const int u = blockIdx.x % units, chunk = blockIdx.x / units;
    const int tg = u % a.tap_groups, cib = (u / a.tap_groups) % a.ci_blocks, cob = u / (a.tap_groups * a.ci_blocks);
    const int t0 = static_cast<int>(static_cast<long long>(a.total_tiles) * chunk / a.chunks);
    const int t1 = static_cast<int>(static_cast<long long>(a.total_tiles) * (chunk + 1) / a.chunks);
    const bool two_blocks = a.Cout - cob * 128 > 64;        // C_out = 64: rows 64..127 of the accumulator are never read

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a.tmDY);
        tma_prefetch_desc(&a.tmX0);
        if (a.cb0 < a.ci_blocks) tma_prefetch_desc(&a.tmX1);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&tail->full[i], 1); mbar_init(&tail->empty[i], 1); }
        mbar_init(&tail->tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kWgTmemCols>(&tail->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp == 0) {
        // ================================ TMA producer ===========================================================
        const CUtensorMap* tmx = cib < a.cb0 ? &a.tmX0 : &a.tmX1;
        const int c0 = (cib < a.cb0 ? cib : cib - a.cb0) * 64;
        int s = 0, ph = 0;
        for (int tile = t0; tile < t1; ++tile) {
            int m = tile;
            const int tx = m % a.tiles_x; m /= a.tiles_x;
            const int ty = m % a.tiles_y;
            const int b = m / a.tiles_y;
            const int x0 = tx * 8, y0 = ty * 16;
            mbar_wait(bar_empty + s * 8, ph ^ 1);
            if (elect_one()) {
                const uint32_t st = ring + s * kStage, fb = bar_full + s * 8;
                mbar_expect_tx(fb, (two_blocks ? 2 : 1) * kDyBlk + kXCopy);
                tma_load_4d(st, &a.tmDY, fb, cob * 128, x0, y0, b);
                if (two_blocks) tma_load_4d(st + kDyBlk, &a.tmDY, fb, cob * 128 + 64, x0, y0, b);
                if constexpr (MODE == kWg3x3) tma_load_4d(st + 2 * kDyBlk, tmx, fb, c0, x0 - 1, y0 - 1, b);
                else if constexpr (MODE == kWgS2D) tma_load_5d(st + 2 * kDyBlk, tmx, fb, c0, tg & 1, x0, tg >> 1, b * a.H + y0);
                else tma_load_4d(st + 2 * kDyBlk, tmx, fb, c0, x0, y0, b);
            }
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ===============================================================
        // 3x3: ONE MMA per K step covers the three taps of the kernel row: N = 192 = three 64-channel chunks whose "leading
        // byte offset" is 128 B — chunk j is the same halo window shifted by j pixels (kx = j).  The issuing thread is a serial
        // instruction stream (a first version with one N = 64 MMA per tap spent ~80 cycles of address arithmetic per 32-cycle
        // MMA); now a tile is 8 MMAs of 96 tensor-pipe cycles whose descriptors differ by compile-time immediates.
        constexpr int kN = kHalo ? 192 : 64;
        constexpr uint32_t idesc = umma_idesc_bf16(128, kN) | (1u << 15) | (1u << 16);      // A and B MN-major
        constexpr uint32_t hiA = umma_desc_hi(1024);                                         // 8-pixel groups 1024 B apart
        constexpr uint32_t hiB = umma_desc_hi(kHalo ? kHaloPitchW : 1024);                   // ... one halo row apart for X
        constexpr uint32_t kBStep = (kHalo ? 2 * kHaloPitchW : 2048) >> 4;                   // two image rows per K step
        int s = 0, ph = 0;
        uint32_t accum = 0u;
        for (int tile = t0; tile < t1; ++tile) {
            mbar_wait(bar_full + s * 8, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = ring + s * kStage;
                const uint32_t a_lo = desc_lo_mn(st, kDyBlk);
                // kernel row ky = tg: the window starts tg halo rows down; chunk stride (LBO) 128 B = one pixel = one kx step
                const uint32_t b_lo = kHalo ? desc_lo_mn(st + 2 * kDyBlk + tg * kHaloPitchW, 128) : desc_lo_mn(st + 2 * kDyBlk, 1024);
                umma_bf16_lohi_pred(tmem_base, a_lo, hiA, b_lo, hiB, idesc, accum);
#pragma unroll
                for (int ks = 1; ks < 8; ++ks)              // 16 pixels = image rows 2 ks, 2 ks + 1 of the 16 x 8 tile
                    umma_bf16_lohi<true>(tmem_base, a_lo + ks * 128, hiA, b_lo + ks * kBStep, hiB, idesc);
                accum = 1u;
                umma_commit(bar_empty + s * 8);
                if (tile == t1 - 1) umma_commit(bar_tfull);
            }
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (warp >= 4) {
        // ================================ epilogue: TMEM -> fp32 atomics into dW[co][ci][tap] ===========================
        if (t1 > t0) {
            const int q = warp & 3;
            const int co = cob * 128 + q * 32 + lane;
            mbar_wait(bar_tfull, 0);
            tc_fence_after();
            for (int t = 0; t < a.taps_per_group; ++t) {
                const int tap = MODE == kWg3x3 ? tg * 3 + t : (MODE == kWgS2D ? tg : 0);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t raw[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 64 + half * 32, raw);
                    tmem_ld_wait();
                    if (co < a.Cout) {
                        float* dst = a.dw + (static_cast<size_t>(co) * a.Cin + cib * 64 + half * 32) * a.taps_total + tap;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + static_cast<size_t>(j) * a.taps_total, __uint_as_float(raw[j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kWgTmemCols>(tmem_base);
    }
}

int smem_for(int mode, int stages) {
    const int stage = 2 * kDyBlk + (mode == kWg3x3 ? kHaloStageW : kDyBlk);
    return 1024 + stages * stage + static_cast<int>(sizeof(WgTail));
}

int wgrad_init() {
    NDIFF_CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<kWg1x1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    NDIFF_CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<kWg3x3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    NDIFF_CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<kWgS2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return 0;
}

}  // namespace

int wgrad_gemm_plan(const WgradDesc& d, int num_sms, WgradPlan* plan) {
    {
        static std::once_flag once;
        static int init_rc = 0;
        std::call_once(once, [] { init_rc = wgrad_init(); });
        if (init_rc) return 1;
    }
    WgradArgs& a = plan->args;
    memset(&a, 0, sizeof(a));
    NDIFF_REQUIRE(d.mode == kWg1x1 || d.mode == kWg3x3 || d.mode == kWgS2D, "weight gradient: unknown mode");
    NDIFF_REQUIRE(d.dy && d.src0 && d.dw && d.B > 0 && d.H > 0 && d.W > 0, "weight gradient: null / empty argument");
    NDIFF_REQUIRE(d.Cout % 64 == 0 && d.C0 > 0 && d.C0 % 64 == 0 && d.C1 % 64 == 0, "weight gradient: channel counts must be multiples of 64");
    NDIFF_REQUIRE(d.mode != kWgS2D || d.C1 == 0, "weight gradient: the space-to-depth form has one source");
    a.mode = d.mode; a.B = d.B; a.H = d.H; a.W = d.W;
    a.tiles_x = (d.W + 7) / 8; a.tiles_y = (d.H + 15) / 16;
    a.total_tiles = d.B * a.tiles_y * a.tiles_x;
    a.Cout = d.Cout; a.Cin = d.C0 + d.C1; a.cb0 = d.C0 / 64;
    a.co_blocks = (d.Cout + 127) / 128; a.ci_blocks = a.Cin / 64;
    a.tap_groups = d.mode == kWg3x3 ? 3 : (d.mode == kWgS2D ? 4 : 1);
    a.taps_per_group = d.mode == kWg3x3 ? 3 : 1;
    a.taps_total = d.mode == kWg3x3 ? 9 : (d.mode == kWgS2D ? 4 : 1);
    const int units = a.co_blocks * a.ci_blocks * a.tap_groups;
    int chunks = num_sms / units;
    if (chunks < 1) chunks = 1;
    if (chunks > a.total_tiles) chunks = a.total_tiles;
    a.chunks = chunks;
    a.stages = 3;
    a.dw = d.dw;
    plan->grid = units * chunks;
    plan->smem_bytes = smem_for(d.mode, a.stages);
    NDIFF_REQUIRE(plan->smem_bytes <= 227 * 1024, "weight gradient: shared-memory budget exceeded");
    {
        const uint64_t C = static_cast<uint64_t>(d.Cout);
        uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
        uint64_t str[3] = {C * 2, static_cast<uint64_t>(d.W) * C * 2, static_cast<uint64_t>(d.H) * d.W * C * 2};
        uint32_t box[4] = {64, 8, 16, 1};
        if (encode_tensor_map(&a.tmDY, d.dy, 4, dims, str, box, true)) return 1;
    }
    for (int sidx = 0; sidx < 2; ++sidx) {
        const __nv_bfloat16* src = sidx == 0 ? d.src0 : d.src1;
        const uint64_t C = static_cast<uint64_t>(sidx == 0 ? d.C0 : d.C1);
        if (C == 0) continue;
        NDIFF_REQUIRE(src != nullptr, "weight gradient: null second source");
        CUtensorMap* tm = sidx == 0 ? &a.tmX0 : &a.tmX1;
        if (d.mode == kWgS2D) {
            // X [B, 2H, 2W, C] viewed as (C, p2, W, p1, B*H), as the forward kS2D kernel reads it
            const uint64_t Win = 2ull * d.W;
            uint64_t dims[5] = {C, 2, static_cast<uint64_t>(d.W), 2, static_cast<uint64_t>(d.B) * d.H};
            uint64_t str[4] = {C * 2, C * 4, Win * C * 2, Win * C * 4};
            uint32_t box[5] = {64, 1, 8, 1, 16};
            if (encode_tensor_map(tm, src, 5, dims, str, box, true)) return 1;
        } else {
            uint64_t dims[4] = {C, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
            uint64_t str[3] = {C * 2, static_cast<uint64_t>(d.W) * C * 2, static_cast<uint64_t>(d.H) * d.W * C * 2};
            uint32_t box[4] = {64, static_cast<uint32_t>(d.mode == kWg3x3 ? 10 : 8), static_cast<uint32_t>(d.mode == kWg3x3 ? 18 : 16), 1};
            if (encode_tensor_map(tm, src, 4, dims, str, box, true)) return 1;
        }
    }
    return 0;
}

int wgrad_gemm_launch(const WgradPlan& plan, cudaStream_t stream) {
    const dim3 grid(plan.grid), block(kWgThreads);
    if (plan.args.mode == kWg3x3) wgrad_gemm_kernel<kWg3x3><<<grid, block, plan.smem_bytes, stream>>>(plan.args);
    else if (plan.args.mode == kWgS2D) wgrad_gemm_kernel<kWgS2D><<<grid, block, plan.smem_bytes, stream>>>(plan.args);
    else wgrad_gemm_kernel<kWg1x1><<<grid, block, plan.smem_bytes, stream>>>(plan.args);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ndiff
