// noisediff_b200 — fused per-pixel MLP chains on tensor cores (sm_100a).  See pixel_chain.cuh.
//
// One CTA = three independent warpgroups of 128 threads.  Each warpgroup walks its own contiguous range of 128-pixel tiles;
// thread r owns pixel r of the tile (= TMEM lane r), so LayerNorm over channels and all epilogue math are thread-local.  A GEMM
// stage is: the 128 threads write the 16-bit A operand [128 x K] — into tensor memory as packed pairs (tcgen05.st; default) or
// into shared memory in the SWIZZLE_128B K-major layout — fence + named barrier, ONE elected lane issues the tcgen05.mma's
// against the layer's weights (resident in shared memory for the whole kernel, loaded once by TMA) and commits to an mbarrier,
// everyone waits and drains the fp32 accumulator from TMEM with tcgen05.ld.  Stages of one tile are strictly sequential; the
// warpgroups interleave, so one group's tensor-core work and TMEM/shared traffic overlaps the other groups' epilogue arithmetic.
// Input tiles arrive by TMA (attn) or as coalesced float4 loads (shot); outputs leave by TMA store from a staging block.
#include "pixel_chain.cuh"
#include "conv_gemm.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ndiff {

namespace {

constexpr int kTile = 128;                 // pixels per tile = TMEM lanes
constexpr int kBlk = kTile * 128;          // bytes of one [128 x 64] bf16 operand block
constexpr int kNWG = 3;                    // warpgroups per CTA (each owns 128 TMEM columns and two operand blocks)
constexpr int kWgBytes = 3 * kBlk;         // per warpgroup: X (input tile by TMA / s1 staging) | A0 | A1 (A0, A1 = the two K
                                           // blocks of the hidden layer; A1 doubles as the staging block of the output store)
constexpr int kThreads = kNWG * 128;
constexpr int kTmemCols = kNWG <= 2 ? 256 : 512;    // power of two >= kNWG * 160 (128 accumulator + 32 operand columns each)

struct ChainTail {
    uint64_t bar_w, bar_x[kNWG], bar_mma[kNWG];
    uint32_t tmem_base;
    uint32_t pad_;
    float fvec[kChainShotFloats];
    // per warpgroup, per sample slot (a 128-pixel tile touches at most two samples): [0] the collapsed attention vector c,
    // [1] b2 + c (bias of ff.net.2 plus the residual's per-sample part)
    alignas(16) float ctab[kNWG][2][2][64];
};

__device__ __forceinline__ uint32_t swz(uint32_t blk, int r, int j) {   // 16-byte chunk j of row r in a SWIZZLE_128B block
    return blk + r * 128 + ((j ^ (r & 7)) << 4);
}
__device__ __forceinline__ void store_half(uint32_t blk, int r, int h, const float (&v)[32]) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        uint4 u;
        u.x = pack_bf16(v[jj * 8 + 0], v[jj * 8 + 1]); u.y = pack_bf16(v[jj * 8 + 2], v[jj * 8 + 3]);
        u.z = pack_bf16(v[jj * 8 + 4], v[jj * 8 + 5]); u.w = pack_bf16(v[jj * 8 + 6], v[jj * 8 + 7]);
        sts128(swz(blk, r, h * 4 + jj), u);
    }
}
// 32 floats of a bias / per-sample table (shared memory, 16-byte aligned).  Called BETWEEN tcgen05.ld and tcgen05.wait::ld so
// that the table's LDS latency hides behind the accumulator drain instead of stalling the first add after the wait.
__device__ __forceinline__ void load32(const float* p, float (&b)[32]) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + j);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
    }
}
// GELU(acc + bias) of 32 accumulator columns -> fp16 operand half (the consuming GEMM runs with fp16 A and B)
__device__ __forceinline__ void store_half_gelu_f16(uint32_t blk, int r, int h, const uint32_t (&raw)[32], const float (&bias)[32]) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        uint4 u;
        u.x = gelu_f16x2(__uint_as_float(raw[jj * 8 + 0]) + bias[jj * 8 + 0], __uint_as_float(raw[jj * 8 + 1]) + bias[jj * 8 + 1]);
        u.y = gelu_f16x2(__uint_as_float(raw[jj * 8 + 2]) + bias[jj * 8 + 2], __uint_as_float(raw[jj * 8 + 3]) + bias[jj * 8 + 3]);
        u.z = gelu_f16x2(__uint_as_float(raw[jj * 8 + 4]) + bias[jj * 8 + 4], __uint_as_float(raw[jj * 8 + 5]) + bias[jj * 8 + 5]);
        u.w = gelu_f16x2(__uint_as_float(raw[jj * 8 + 6]) + bias[jj * 8 + 6], __uint_as_float(raw[jj * 8 + 7]) + bias[jj * 8 + 7]);
        sts128(swz(blk, r, h * 4 + jj), u);
    }
}
// GELU(acc + bias) of 32 accumulator columns -> 16 TMEM columns of packed fp16 pairs (A operand of a tensor-memory GEMM)
__device__ __forceinline__ void gelu_to_tmem(uint32_t taddr, const uint32_t (&raw)[32], const float (&bias)[32]) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
        pk[j] = gelu_f16x2(__uint_as_float(raw[2 * j]) + bias[2 * j], __uint_as_float(raw[2 * j + 1]) + bias[2 * j + 1]);
    tmem_st16(taddr, pk);
}
// D[tmem] = A[tmem: 128 lanes x (ksteps * 16) fp16, 8 columns per step] * W[N x K]^T (fp16 rows in shared memory, 64-wide K blocks)
template <int N, bool F16 = true>
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_blk, int ksteps) {
    constexpr uint32_t idesc = F16 ? umma_idesc_f16(128, N) : umma_idesc_bf16(128, N);
    constexpr uint32_t hi = umma_desc_hi(1024);
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t b_lo = umma_desc_lo(w_blk + (k >> 2) * N * 128) + 2 * (k & 3);
        if (k == 0) umma_ts_lohi<false>(d_tmem, a_tmem, b_lo, hi, idesc);
        else umma_ts_lohi<true>(d_tmem, a_tmem + 8 * k, b_lo, hi, idesc);
    }
}
// D[tmem] = A[128 x (kblocks*64)] * W[N x (kblocks*64)]^T ; k16 = MMAs per K block (4, or 1 when only K = 16 is live)
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

// kNWG warpgroups keep that many independent tiles in flight per SM (the chain of one tile is strictly sequential: operand
// write -> barrier -> MMA -> TMEM drain, five times for the shot program), so the tensor-core round trips and the epilogue
// arithmetic of different tiles overlap.  384 threads cap the kernel at 168 registers: the pixel's input row stays packed
// (xr, 32 registers), accumulators are drained 32 columns at a time, and LayerNorm re-derives y = x + c per pass instead of
// holding 64 floats.
// TS: the GELU outputs that feed the next GEMM (shot_mlp1.fc1 -> fc2, ff.net.0 -> the folded stage, shot_mlp2.fc1 -> fc2) are
// written to TENSOR MEMORY as packed fp16 — over accumulator columns the same warp has already drained — and the next GEMM reads
// its A operand from there (umma_ts_lohi); its accumulator goes to the upper 64 columns.  Takes the thread-written operands and
// their read-back off the shared-memory pipe, which bounds these kernels together with instruction issue.
template <int PROG, bool TS>
__global__ void __launch_bounds__(kThreads, 1) pixel_chain_kernel(const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __builtin_assume(__isShared(smem));      // the manual alignment hides the address space: without the hint every table read is a generic LD.E
    constexpr bool kShot = PROG == kProgShot;
    constexpr int kWRows = kShot ? kChainShotRows : kChainAttnRows;
    constexpr int kNF = kShot ? kChainShotFloats : kChainAttnFloats;
    constexpr int kWBytes = kWRows * 128;
    constexpr uint32_t kDup = TS ? 64 : 0;      // TS: accumulators of the GEMMs fed from tensor memory live in the upper 64 columns
    ChainTail* tail = reinterpret_cast<ChainTail*>(smem + kWBytes + kNWG * kWgBytes);

    // the warp index is broadcast from lane 0 so that the compiler knows it (and everything derived from it: the warpgroup's
    // operand blocks, barriers and TMEM columns) is warp-uniform
    const int tid = threadIdx.x, warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0), wg = warp_u >> 2, r = tid & 127, q = warp_u & 3;
    const uint32_t sW = smem_u32(smem);
    const uint32_t sX = sW + kWBytes + wg * kWgBytes, sA0 = sX + kBlk, sA1 = sA0 + kBlk;
    // weight blocks (row offsets of pixel_chain.cuh's blob layout, 128 B per row)
    const uint32_t sWattn = sW + (kShot ? 128 * 128 : 0);
    const uint32_t sW1 = sWattn, sW2 = sW1 + 128 * 128, sWp = sW2 + 128 * 128, sWm2 = sWp + 64 * 128;
    const float* fA = tail->fvec + (kShot ? 128 : 0);
    // (the first 128 floats of the attention block are reserved: LayerNorm's affine is folded into W1 / b1 by the packer)
    const float* f_b1 = fA + 128, *f_bm2 = fA + 448;
    (void)f_bm2;
    const uint32_t bar_w = smem_u32(&tail->bar_w), bar_x = smem_u32(&tail->bar_x[wg]), bar_mma = smem_u32(&tail->bar_mma[wg]);

    if (tid == 0) {
        tma_prefetch_desc(&a.tmW);
        tma_prefetch_desc(&a.tmOut);
        if (kShot) tma_prefetch_desc(&a.tmOut2); else tma_prefetch_desc(&a.tmX);
        mbar_init(&tail->bar_w, 1);
        for (int i = 0; i < kNWG; ++i) { mbar_init(&tail->bar_x[i], 1); mbar_init(&tail->bar_mma[i], 1); }
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc<kTmemCols>(&tail->tmem_base);
    for (int i = tid; i < kNF; i += kThreads) tail->fvec[i] = __ldg(a.fvec + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // this warpgroup's tensor-memory columns: 128 of accumulators (+ TS: 32 for LayerNorm's output, the A operand of ff.net.0)
    const uint32_t tmem_d = tail->tmem_base + wg * (TS ? 160 : 128);
    const uint32_t tmem_rd = tmem_d + (static_cast<uint32_t>(q * 32) << 16);     // + this warp's lane quarter

    if (tid == 0) {   // the whole weight blob, once
        mbar_expect_tx(bar_w, kWBytes);
        for (int i = 0; i < kWRows / 64; ++i) tma_load_2d(sW + i * 64 * 128, &a.tmW, bar_w, 0, i * 64);
    }
    pdl_trigger();
    pdl_wait();          // weights / parameters above are constants; everything below touches the previous kernel's output
    // every warpgroup walks a CONTIGUOUS range of tiles: the per-sample tables change once or twice per range (a strided walk of
    // 148 x 3 warpgroups jumps 0.87 samples per step at 256 x 256, i.e. reloaded them from global memory on almost every tile)
    const int n_wg = gridDim.x * kNWG, wg_id = blockIdx.x * kNWG + wg;
    const int t_begin = static_cast<int>(static_cast<long long>(a.n_tiles) * wg_id / n_wg);
    const int t_end = static_cast<int>(static_cast<long long>(a.n_tiles) * (wg_id + 1) / n_wg);
    // Single-thread work (TMA, tcgen05.mma, commits) is done by the warpgroup's first warp with ALL lanes walking the code and
    // one elected lane issuing: under a plain `if (r == 0)` the compiler cannot know that one lane is active and wraps every
    // uniform-register operand of UTCHMMA / UTMALDG in a lane-serialising loop (15 instructions per MMA on the stage's critical
    // path).  elect.sync picks the same lane every time, so the bulk-group waits pair with the stores of that lane.
    if (!kShot && q == 0 && t_begin < t_end) {
        if (elect_one()) {
            mbar_expect_tx(bar_x, kBlk);
            tma_load_2d(sX, &a.tmX, bar_x, 0, t_begin * kTile);
        }
        __syncwarp();
    }
    uint32_t xph = 0, mph = 0;
    bool w_ready = false;
    int tab_b0 = -1, tab_b1 = -1;            // samples currently held by this warpgroup's ctab slots
    float4 c4n = make_float4(0.f, 0.f, 0.f, 0.f), x4n = c4n;      // shot: this pixel's inputs of the NEXT tile (software prefetch)
    if (kShot && t_begin < t_end && t_begin * kTile + r < a.npix) { c4n = __ldg(a.clean + t_begin * kTile + r); x4n = a.x[t_begin * kTile + r]; }
    (void)c4n; (void)x4n;

    // one GEMM stage: publish my operand writes, one thread issues, everybody waits for the accumulator
#define NDIFF_STAGE(ISSUE)                                                   \
    do {                                                                     \
        fence_proxy_async();                                                 \
        tc_fence_before();                                                   \
        named_bar_sync(1 + wg, 128);                                         \
        if (q == 0) {                                                        \
            if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }           \
            tc_fence_after();                                                \
            if (elect_one()) {                                               \
                ISSUE;                                                       \
                umma_commit(bar_mma);                                        \
            }                                                                \
            __syncwarp();                                                    \
        }                                                                    \
        mbar_wait(bar_mma, mph);                                             \
        mph ^= 1;                                                            \
        tc_fence_after();                                                    \
    } while (0)

    for (int tile = t_begin; tile < t_end; ++tile) {
        const int p = tile * kTile + r;
        const bool live = p < a.npix;
        const int pc = live ? p : a.npix - 1;
        uint32_t xr[32];                      // this pixel's 64-channel attention-block input, packed bf16
        // per-sample vectors of the tile's (at most two) samples -> shared memory; refreshed only when the samples change.
        // Nobody still reads the old table: every thread's last read precedes the previous tile's proj_out stage barrier.
        const int b_first = (tile * kTile) / a.HW;
        {
            const int last = tile * kTile + kTile - 1;
            const int b_last = (last < a.npix ? last : a.npix - 1) / a.HW;
            if (b_first != tab_b0 || b_last != tab_b1) {
                tab_b0 = b_first; tab_b1 = b_last;
                const int slot = r >> 6, j = r & 63;
                const size_t co = static_cast<size_t>(slot ? b_last : b_first) * a.cvec_ld + j;
                const float cj = __ldg(a.cvec + co);
                tail->ctab[wg][slot][0][j] = cj;
                // the per-sample vector of the folded last linear stage (attn: Wp (b2 + c) + bp; shot: Wm1 Wp (b2 + c) + Wm1 bp
                // + bm1), computed once per condition (engine.cu attn_vec2_kernel)
                tail->ctab[wg][slot][1][j] = __ldg(a.cvec2 + co);
                named_bar_sync(1 + wg, 128);
            }
        }
        const float* ct = &tail->ctab[wg][(pc / a.HW) != b_first ? 1 : 0][0][0];      // c at ct[j], folded-stage vector at ct[64 + j]

        // X (shot) and A1 are sources of the previous tile's TMA stores: they must have been read before anyone rewrites them
        // (every thread's first write to either comes after the next named barrier, which the issuing warp joins after this wait)
        if (q == 0) {
            if (elect_one()) tma_store_wait_read();
            __syncwarp();
        }

        if constexpr (kShot) {
            // ---- shot_mlp1.fc1 on cat[clean, x_t] (ref Diffusion_arch.py:598; clean first) --------------------------
            // (this tile's inputs were requested one tile ago; the next tile's are requested now, a whole chain ahead of their use)
            uint4 u;
            u.x = pack_bf16(c4n.x, c4n.y); u.y = pack_bf16(c4n.z, c4n.w); u.z = pack_bf16(x4n.x, x4n.y); u.w = pack_bf16(x4n.z, x4n.w);
            {
                const int pn = p + kTile;
                c4n = x4n = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tile + 1 < t_end && pn < a.npix) { c4n = __ldg(a.clean + pn); x4n = a.x[pn]; }
            }
            sts128(swz(sA0, r, 0), u);
            sts128(swz(sA0, r, 1), make_uint4(0u, 0u, 0u, 0u));
            NDIFF_STAGE(issue_gemm<64>(tmem_d, sA0, sW, 1, 1));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[32];
                float bv[32];
                tmem_ld32(tmem_rd + h * 32, raw);
                load32(tail->fvec + h * 32, bv);
                tmem_ld_wait();
                if constexpr (TS) gelu_to_tmem(tmem_rd + h * 16, raw, bv);
                else store_half_gelu_f16(sA0, r, h, raw, bv);
            }
            // ---- shot_mlp1.fc2 -> s1 (stored: it is the branch's residual r_s, ref :599) ------------------------------
            if constexpr (TS) { tmem_st_wait(); NDIFF_STAGE(issue_gemm_ts<64>(tmem_d + kDup, tmem_d, sW + 64 * 128, 4)); }
            else NDIFF_STAGE((issue_gemm<64, true>(tmem_d, sA0, sW + 64 * 128, 1, 4)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[32];
                float bv[32];
                tmem_ld32(tmem_rd + kDup + h * 32, raw);
                load32(tail->fvec + 64 + h * 32, bv);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float v0 = __uint_as_float(raw[j]) + bv[j];
                    const float v1 = __uint_as_float(raw[j + 1]) + bv[j + 1];
                    xr[h * 16 + j / 2] = pack_bf16(v0, v1);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)      // s1 staged in the X slot; stored by TMA at the next barrier
                sts128(swz(sX, r, j), make_uint4(xr[j * 4], xr[j * 4 + 1], xr[j * 4 + 2], xr[j * 4 + 3]));
        } else {
            mbar_wait(bar_x, xph);
            xph ^= 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 u = lds128(swz(sX, r, j));
                xr[j * 4] = u.x; xr[j * 4 + 1] = u.y; xr[j * 4 + 2] = u.z; xr[j * 4 + 3] = u.w;
            }
        }

        // ---- y = x + c ; A0 = (y - mean) * rstd   (AttnBlock.norm2 on the collapsed attention, ref :438-439; the affine
        //      g, b of the LayerNorm lives in W1 / b1).  One pass for the moments, one to normalise; y is re-derived from the
        //      packed row instead of being held in 64 registers.
        {
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(ct + j);
                const float2 f0 = unpack_bf16(xr[j / 2]), f1 = unpack_bf16(xr[j / 2 + 1]);
                const float y0 = f0.x + c4.x, y1 = f0.y + c4.y, y2 = f1.x + c4.z, y3 = f1.y + c4.w;
                sum += (y0 + y1) + (y2 + y3);
                sq = fmaf(y0, y0, sq); sq = fmaf(y1, y1, sq); sq = fmaf(y2, y2, sq); sq = fmaf(y3, y3, sq);
            }
            // (zero-padded layouts: y is exactly 0 on the padding, so both moments are sums over the live channels)
            const float mean = sum * a.inv_c;
            const float rstd = rsqrtf(fmaxf(sq * a.inv_c - mean * mean, 0.f) + 1e-5f);
            const float nb = -mean * rstd;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 c4 = *reinterpret_cast<const float4*>(ct + h * 32 + j);
                    const float2 f0 = unpack_bf16(xr[(h * 32 + j) / 2]), f1 = unpack_bf16(xr[(h * 32 + j) / 2 + 1]);
                    v[j] = fmaf(f0.x + c4.x, rstd, nb); v[j + 1] = fmaf(f0.y + c4.y, rstd, nb);
                    v[j + 2] = fmaf(f1.x + c4.z, rstd, nb); v[j + 3] = fmaf(f1.y + c4.w, rstd, nb);
                }
                if constexpr (TS) {      // bf16 pairs -> columns 128 .. 159 of this group's tensor memory (A operand of ff.net.0)
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                    tmem_st16(tmem_rd + 128 + h * 16, pk);
                } else {
                    store_half(sA0, r, h, v);
                }
            }
            if constexpr (TS) tmem_st_wait();
        }
        // ---- FeedForward.net.0: Linear(C, 2C) + GELU   (ref :405-422); hidden K block hh -> operand block A0 / A1 ------------
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1 + wg, 128);
        if (q == 0) {
            if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
            tc_fence_after();
            if (elect_one()) {
                if (kShot) {                   // s1 staged in the X slot by every thread before this barrier
                    tma_store_2d(&a.tmOut2, sX, 0, tile * kTile);
                    tma_store_commit();
                }  // (attn: the X tile is an operand of the last GEMM stage, so the next tile is prefetched after that stage)
                if constexpr (TS) issue_gemm_ts<128, false>(tmem_d, tmem_d + 128, sW1, 4);
                else issue_gemm<128>(tmem_d, sA0, sW1, 1, 4);
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, mph);
        mph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[32];
                float bv[32];
                tmem_ld32(tmem_rd + hh * 64 + h * 32, raw);
                load32(f_b1 + hh * 64 + h * 32, bv);
                tmem_ld_wait();
                if constexpr (TS) gelu_to_tmem(tmem_rd + hh * 32 + h * 16, raw, bv);
                else store_half_gelu_f16(sA0 + hh * kBlk, r, h, raw, bv);
            }
        }
        if constexpr (!kShot) {
            // ---- FeedForward.net.2 and proj_out meet without a nonlinearity (ref :439-443): ONE stage
            //        out = (Wp W2) h + Wp x + [Wp (b2 + c) + bp] + x
            //      = fp16 GEMM over the hidden layer (K = 128, folded weight in W2's slot) accumulated with a bf16 GEMM over the
            //      input tile itself, which already sits in shared memory as TMA landed it (K = 64, Wp) -- z never exists.
            if constexpr (TS) {
                tmem_st_wait();
                NDIFF_STAGE((issue_gemm_ts<64>(tmem_d + kDup, tmem_d, sW2, 8), issue_gemm<64>(tmem_d + kDup, sX, sWp, 1, 4, true)));
            } else {
                NDIFF_STAGE((issue_gemm<64, true>(tmem_d, sA0, sW2, 2, 4), issue_gemm<64>(tmem_d, sX, sWp, 1, 4, true)));
            }
            if (q == 0 && tile + 1 < t_end) {      // the tensor core is done with X (the wait above): prefetch the next tile
                if (elect_one()) {
                    mbar_expect_tx(bar_x, kBlk);
                    tma_load_2d(sX, &a.tmX, bar_x, 0, (tile + 1) * kTile);
                }
                __syncwarp();
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[32];
                float v[32];
                tmem_ld32(tmem_rd + kDup + h * 32, raw);
                load32(ct + 64 + h * 32, v);                      // Wp (b2 + c) + bp
#pragma unroll
                for (int j = 0; j < 32; j += 2) {                 // + the residual x, while the accumulator drains
                    const float2 f = unpack_bf16(xr[(h * 32 + j) / 2]);
                    v[j] += f.x; v[j + 1] += f.y;
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(raw[j]);
                store_half(sA1, r, h, v);               // staging for the TMA store
            }
        }
        if constexpr (kShot) {
            // ---- ff.net.2, proj_out (+ both residuals) and shot_mlp2.fc1 are all linear (ref :439-443, :601): ONE stage
            //        fc1(Wp (W2 h + b2 + c + s1) + bp + s1) = (Wm1 Wp W2) h + (Wm1 Wp + Wm1) s1 + [per-sample vector]
            //      = fp16 GEMM over the hidden layer (K = 128) accumulated with a bf16 GEMM over s1, which still sits in the X slot
            //      where it was staged for its TMA store.  Neither z nor the attention block's output ever exists.  Then GELU, fc2.
            if constexpr (TS) {
                tmem_st_wait();
                NDIFF_STAGE((issue_gemm_ts<64>(tmem_d + kDup, tmem_d, sW2, 8), issue_gemm<64>(tmem_d + kDup, sX, sWp, 1, 4, true)));
            } else {
                NDIFF_STAGE((issue_gemm<64, true>(tmem_d, sA0, sW2, 2, 4), issue_gemm<64>(tmem_d, sX, sWp, 1, 4, true)));
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[32];
                float bv[32];
                tmem_ld32(tmem_rd + kDup + h * 32, raw);
                load32(ct + 64 + h * 32, bv);
                tmem_ld_wait();
                if constexpr (TS) gelu_to_tmem(tmem_rd + h * 16, raw, bv);     // (columns 0..31: the A operand there has been consumed)
                else store_half_gelu_f16(sA0, r, h, raw, bv);
            }
            if constexpr (TS) { tmem_st_wait(); NDIFF_STAGE(issue_gemm_ts<64>(tmem_d + kDup, tmem_d, sWm2, 4)); }
            else NDIFF_STAGE((issue_gemm<64, true>(tmem_d, sA0, sWm2, 1, 4)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t raw[32];
                float v[32];
                tmem_ld32(tmem_rd + kDup + h * 32, raw);
                load32(f_bm2 + h * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(raw[j]);
                store_half(sA1, r, h, v);
            }
        }
        // ---- output tile: staged in A1, stored by TMA (rows beyond npix are clipped by the tensor map) ---------------------
        fence_proxy_async();
        tc_fence_before();        // the accumulator drains above must be ordered before the next tile's MMA overwrites TMEM
        named_bar_sync(1 + wg, 128);
        if (q == 0) {
            if (elect_one()) {
                tma_store_2d(&a.tmOut, sA1, 0, tile * kTile);
                tma_store_commit();
            }
            __syncwarp();
        }
    }
#undef NDIFF_STAGE
    if (q == 0) {
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tail->tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Shot-branch tail: GroupNorm-apply + residuals -> fc1 -> GELU -> fc2 (see pixel_chain.cuh).  Same skeleton as the chain
// kernel: three warpgroups, thread r = pixel r of its warpgroup's tile; the three input tiles arrive by TMA into three
// blocks, the operand block A0 feeds both GEMMs.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTailWgBytes = 4 * kBlk;     // H | R1 | R2 | A0

struct TailSmem {
    uint64_t bar_w, bar_x[kNWG], bar_mma[kNWG];
    uint32_t tmem_base;
    uint32_t pad_;
    float fvec[kTailFloats];
    alignas(16) float ctab[kNWG][2][2][64];     // per warpgroup / sample slot: [0] A/2, [1] B/2 of the folded GroupNorm affine
};

// TS (see pixel_chain_kernel): both A operands go through tensor memory — the GroupNorm-apply output as bf16 pairs in columns
// 64 .. 95, fc1's GELU output as fp16 pairs over the drained accumulator columns 0 .. 31; fc2 accumulates in columns 96 .. 111.
template <bool TS>
__global__ void __launch_bounds__(kThreads, 1) tail_chain_kernel(const __grid_constant__ TailArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __builtin_assume(__isShared(smem));      // the manual alignment hides the address space: without the hint every table read is a generic LD.E
    constexpr int kWBytes = kTailRows * 128;
    TailSmem* tail = reinterpret_cast<TailSmem*>(smem + kWBytes + kNWG * kTailWgBytes);
    const int tid = threadIdx.x, warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0), wg = warp_u >> 2, r = tid & 127, q = warp_u & 3;
    const uint32_t sW = smem_u32(smem);
    const uint32_t sH = sW + kWBytes + wg * kTailWgBytes, sR1 = sH + kBlk, sR2 = sR1 + kBlk, sA0 = sR2 + kBlk;
    const uint32_t bar_w = smem_u32(&tail->bar_w), bar_x = smem_u32(&tail->bar_x[wg]), bar_mma = smem_u32(&tail->bar_mma[wg]);
    if (tid == 0) {
        tma_prefetch_desc(&a.tmW); tma_prefetch_desc(&a.tmH); tma_prefetch_desc(&a.tmR1); tma_prefetch_desc(&a.tmR2);
        mbar_init(&tail->bar_w, 1);
        for (int i = 0; i < kNWG; ++i) { mbar_init(&tail->bar_x[i], 1); mbar_init(&tail->bar_mma[i], 1); }
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc<kTmemCols>(&tail->tmem_base);
    for (int i = tid; i < kTailFloats; i += kThreads) tail->fvec[i] = __ldg(a.fvec + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tail->tmem_base + wg * 128;
    const uint32_t tmem_rd = tmem_d + (static_cast<uint32_t>(q * 32) << 16);
    if (tid == 0) {
        mbar_expect_tx(bar_w, kWBytes);
        for (int i = 0; i < kTailRows / 64; ++i) tma_load_2d(sW + i * 64 * 128, &a.tmW, bar_w, 0, i * 64);
    }
    pdl_trigger();
    pdl_wait();
    const int n_wg = gridDim.x * kNWG, wg_id = blockIdx.x * kNWG + wg;      // contiguous tile range per warpgroup (see pixel_chain_kernel)
    const int t_begin = static_cast<int>(static_cast<long long>(a.n_tiles) * wg_id / n_wg);
    const int t_end = static_cast<int>(static_cast<long long>(a.n_tiles) * (wg_id + 1) / n_wg);
    auto load_tile = [&](int tile) {
        mbar_expect_tx(bar_x, 3 * kBlk);
        tma_load_2d(sH, &a.tmH, bar_x, 0, tile * kTile);
        tma_load_2d(sR1, &a.tmR1, bar_x, 0, tile * kTile);
        tma_load_2d(sR2, &a.tmR2, bar_x, 0, tile * kTile);
    };
    if (q == 0 && t_begin < t_end) {        // one elected lane of the warpgroup's first warp issues (see pixel_chain_kernel)
        if (elect_one()) load_tile(t_begin);
        __syncwarp();
    }
    uint32_t xph = 0, mph = 0;
    bool w_ready = false;
    int tab_b0 = -1, tab_b1 = -1;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int p = tile * kTile + r;
        const bool live = p < a.npix;
        const int pc = live ? p : a.npix - 1;
        const int b_first = (tile * kTile) / a.HW;
        {
            const int last = tile * kTile + kTile - 1;
            const int b_last = (last < a.npix ? last : a.npix - 1) / a.HW;
            if (b_first != tab_b0 || b_last != tab_b1) {     // (the previous tile's readers are past its first stage barrier)
                tab_b0 = b_first; tab_b1 = b_last;
                const int slot = r >> 6, c = r & 63, b = slot ? b_last : b_first;
                const int g = c >> a.lgs;
                const double inv_n = 1.0 / (static_cast<double>(a.HW) * (1 << a.lgs) * a.real_frac);
                const double s_ = static_cast<double>(static_cast<long long>(a.stats[(b * a.G + g) * 2])) * (1.0 / 16777216.0);
                const double q_ = static_cast<double>(static_cast<long long>(a.stats[(b * a.G + g) * 2 + 1])) * (1.0 / 16777216.0);
                const double meand = s_ * inv_n;
                const float mean = static_cast<float>(meand);
                const float var = fmaxf(static_cast<float>(q_ * inv_n - meand * meand), 0.f);
                const float rstd = rsqrtf(var + a.eps);
                const float Aj = rstd * __ldg(a.gamma + c);
                const float Bj = __ldg(a.beta + c) - mean * Aj;
                tail->ctab[wg][slot][0][c] = 0.5f * Aj;      // SiLU(y) = h + h tanh(h), h = y / 2 (exact halving)
                tail->ctab[wg][slot][1][c] = 0.5f * Bj;
                named_bar_sync(1 + wg, 128);
            }
        }
        const float* ct = &tail->ctab[wg][(pc / a.HW) != b_first ? 1 : 0][0][0];
        // ---- y = SiLU(GN(h2)) + s4 + s1 -> A0 ---------------------------------------------------------------------------
        mbar_wait(bar_x, xph);
        xph ^= 1;
        uint32_t pk[16];
        (void)pk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 h = lds128(swz(sH, r, j)), r1 = lds128(swz(sR1, r, j)), r2 = lds128(swz(sR2, r, j));
            const float4 a0 = *reinterpret_cast<const float4*>(ct + j * 8), a1 = *reinterpret_cast<const float4*>(ct + j * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(ct + 64 + j * 8), b1 = *reinterpret_cast<const float4*>(ct + 64 + j * 8 + 4);
            auto act = [](float x, float ah, float bh) {
                const float hh = fmaf(x, ah, bh);
                return fmaf(hh, tanh_approx(hh), hh);
            };
            const float2 h0 = unpack_bf16(h.x), h1 = unpack_bf16(h.y), h2 = unpack_bf16(h.z), h3 = unpack_bf16(h.w);
            const float2 p0 = unpack_bf16(r1.x), p1 = unpack_bf16(r1.y), p2 = unpack_bf16(r1.z), p3 = unpack_bf16(r1.w);
            const float2 q0 = unpack_bf16(r2.x), q1 = unpack_bf16(r2.y), q2 = unpack_bf16(r2.z), q3 = unpack_bf16(r2.w);
            uint4 u;
            // same association as gn_apply_kernel: (SiLU + res1) + res2
            u.x = pack_bf16((act(h0.x, a0.x, b0.x) + p0.x) + q0.x, (act(h0.y, a0.y, b0.y) + p0.y) + q0.y);
            u.y = pack_bf16((act(h1.x, a0.z, b0.z) + p1.x) + q1.x, (act(h1.y, a0.w, b0.w) + p1.y) + q1.y);
            u.z = pack_bf16((act(h2.x, a1.x, b1.x) + p2.x) + q2.x, (act(h2.y, a1.y, b1.y) + p2.y) + q2.y);
            u.w = pack_bf16((act(h3.x, a1.z, b1.z) + p3.x) + q3.x, (act(h3.y, a1.w, b1.w) + p3.y) + q3.y);
            if constexpr (TS) {
                pk[(j & 3) * 4] = u.x; pk[(j & 3) * 4 + 1] = u.y; pk[(j & 3) * 4 + 2] = u.z; pk[(j & 3) * 4 + 3] = u.w;
                if ((j & 3) == 3) tmem_st16(tmem_rd + 64 + (j >> 2) * 16, pk);
            } else {
                sts128(swz(sA0, r, j), u);
            }
        }
        if constexpr (TS) tmem_st_wait();
        // ---- shot_mlp3.fc1 + GELU -----------------------------------------------------------------------------------------
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1 + wg, 128);
        if (q == 0) {
            if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
            tc_fence_after();
            if (elect_one()) {
                if (tile + 1 < t_end) load_tile(tile + 1);     // everybody has consumed the three input blocks
                if constexpr (TS) issue_gemm_ts<64, false>(tmem_d, tmem_d + 64, sW, 4);
                else issue_gemm<64>(tmem_d, sA0, sW, 1, 4);
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, mph);
        mph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t raw[32];
            float bv[32];
            tmem_ld32(tmem_rd + h * 32, raw);
            load32(tail->fvec + h * 32, bv);
            tmem_ld_wait();
            if constexpr (TS) gelu_to_tmem(tmem_rd + h * 16, raw, bv);
            else store_half_gelu_f16(sA0, r, h, raw, bv);
        }
        if constexpr (TS) tmem_st_wait();
        // ---- shot_mlp3.fc2 (64 -> 4, N padded to 16) ------------------------------------------------------------------------
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1 + wg, 128);
        if (q == 0) {
            tc_fence_after();
            if (elect_one()) {
                if constexpr (TS) issue_gemm_ts<16>(tmem_d + 96, tmem_d, sW + 64 * 128, 4);
                else issue_gemm<16, true>(tmem_d, sA0, sW + 64 * 128, 1, 4);
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, mph);
        mph ^= 1;
        tc_fence_after();
        {
            uint32_t raw[4];
            tmem_ld4(tmem_rd + (TS ? 96 : 0), raw);
            tmem_ld_wait();
            if (live)
                a.out[p] = make_float4(__uint_as_float(raw[0]) + tail->fvec[64], __uint_as_float(raw[1]) + tail->fvec[65],
                                       __uint_as_float(raw[2]) + tail->fvec[66], __uint_as_float(raw[3]) + tail->fvec[67]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tail->tmem_base);
    }
}

__global__ void pack_chain_weight_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int N, int K, int KB,
                                         int f16) {
    const int total = KB * N * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kk = i & 63, n = (i >> 6) % N, kb = (i >> 6) / N;
        const int k = kb * 64 + kk;
        const float v = k < K ? src[static_cast<size_t>(n) * K + k] : 0.f;
        dst[i] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
}


__global__ void fold_layernorm_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ g,
                                      const float* __restrict__ beta, float* __restrict__ w_out, float* __restrict__ b_out,
                                      int K) {
    __shared__ float red[32];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float wv = w[static_cast<size_t>(n) * K + k];
        w_out[static_cast<size_t>(n) * K + k] = wv * g[k];
        acc = fmaf(wv, beta[k], acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = b[n];
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
        b_out[n] = t;
    }
}

}  // namespace

namespace {
__global__ void fold_linear_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ b1,
                                   const float* __restrict__ b2, float* __restrict__ w_out, float* __restrict__ b_out, int N) {
    const int n = blockIdx.x, k = threadIdx.x;          // one output row per block, one column per thread
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(a[n * N + j], w[j * N + k], acc);
    w_out[n * N + k] = acc;
    if (k == 0) {
        float t = b2[n];
        for (int j = 0; j < N; ++j) t = fmaf(a[n * N + j], b1[j], t);
        b_out[n] = t;
    }
}
}  // namespace

int fold_linear_launch(const float* a, const float* w, const float* b1, const float* b2, float* w_out, float* b_out, int N,
                       cudaStream_t s) {
    fold_linear_kernel<<<N, N, 0, s>>>(a, w, b1, b2, w_out, b_out, N);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int fold_layernorm_launch(const float* w, const float* b, const float* g, const float* beta, float* w_out, float* b_out, int N,
                          int K, cudaStream_t s) {
    fold_layernorm_kernel<<<N, 64, 0, s>>>(w, b, g, beta, w_out, b_out, K);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pack_chain_weight_launch(const float* src, __nv_bfloat16* dst, int N, int K, bool f16, cudaStream_t s) {
    const int KB = (K + 63) / 64;
    pack_chain_weight_kernel<<<(KB * N * 64 + 255) / 256, 256, 0, s>>>(src, reinterpret_cast<uint16_t*>(dst), N, K, KB,
                                                                       f16 ? 1 : 0);
    NDIFF_CUDA_OK(cudaGetLastError());
    return 0;
}

// A operands of the chained GEMMs through tensor memory (see pixel_chain_kernel's TS); NDIFF_CHAIN_TS=0 keeps them in shared memory
static bool chain_operands_in_tmem() {
    static const bool ts = [] {
        const char* v = std::getenv("NDIFF_CHAIN_TS");
        return !(v && v[0] == '0');          // default on (measured on B200: -7 % on every chain); 0 = shared-memory operands
    }();
    return ts;
}

static int chain_smem_bytes(int prog) {
    const int rows = prog == kProgShot ? kChainShotRows : kChainAttnRows;
    return 1024 + rows * 128 + kNWG * kWgBytes + static_cast<int>(sizeof(ChainTail));
}

int pixel_chain_init() {
    NDIFF_CUDA_OK(cudaFuncSetAttribute(pixel_chain_kernel<kProgAttn, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       chain_smem_bytes(kProgAttn)));
    NDIFF_CUDA_OK(cudaFuncSetAttribute(pixel_chain_kernel<kProgShot, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       chain_smem_bytes(kProgShot)));
    NDIFF_CUDA_OK(cudaFuncSetAttribute(pixel_chain_kernel<kProgAttn, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       chain_smem_bytes(kProgAttn)));
    NDIFF_CUDA_OK(cudaFuncSetAttribute(pixel_chain_kernel<kProgShot, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       chain_smem_bytes(kProgShot)));
    return 0;
}

int pixel_chain_plan(const ChainDesc& d, int num_sms, ChainPlan* plan) {
    {
        static std::once_flag once;
        static int init_rc = 0;
        std::call_once(once, [] { init_rc = pixel_chain_init(); });
        if (init_rc) return 1;
    }
    ChainArgs& a = plan->args;
    memset(&a, 0, sizeof(a));
    NDIFF_REQUIRE(d.prog == kProgAttn || d.prog == kProgShot, "unknown pixel-chain program");
    NDIFF_REQUIRE(d.npix > 0 && d.HW > 0 && d.weights && d.fvec && d.cvec && d.out, "pixel chain: null argument");
    NDIFF_REQUIRE(d.cvec_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d.cvec) & 15) == 0,
                  "pixel chain: the per-sample attention vector must be 16-byte aligned");
    plan->prog = d.prog;
    a.npix = d.npix; a.HW = d.HW; a.n_tiles = (d.npix + kTile - 1) / kTile;
    a.fvec = d.fvec; a.cvec = d.cvec; a.cvec_ld = d.cvec_ld; a.cvec2 = d.cvec2;
    NDIFF_REQUIRE(d.cvec2 && (reinterpret_cast<uintptr_t>(d.cvec2) & 15) == 0,
                  "pixel chain: the folded stage needs its per-sample vector (attn: Wp (b2 + c) + bp)");
    NDIFF_REQUIRE(d.real_frac > 0.f && d.real_frac <= 1.f, "pixel chain: live channel fraction must be in (0, 1]");
    a.inv_c = 1.0f / (64.0f * d.real_frac);
    const uint64_t adims[2] = {64, static_cast<uint64_t>(d.npix)};
    const uint64_t astr[1] = {128};
    const uint32_t abox[2] = {64, kTile};
    if (d.prog == kProgAttn) {
        NDIFF_REQUIRE(d.x != nullptr, "pixel chain: null input");
        if (encode_tensor_map(&a.tmX, d.x, 2, adims, astr, abox, true)) return 1;
    } else {
        NDIFF_REQUIRE(d.clean && d.xt && d.out2, "pixel chain (shot): null input");
        a.clean = reinterpret_cast<const float4*>(d.clean);
        a.x = reinterpret_cast<const float4*>(d.xt);
        if (encode_tensor_map(&a.tmOut2, d.out2, 2, adims, astr, abox, true)) return 1;
    }
    if (encode_tensor_map(&a.tmOut, d.out, 2, adims, astr, abox, true)) return 1;
    const int rows = d.prog == kProgShot ? kChainShotRows : kChainAttnRows;
    const uint64_t wdims[2] = {64, static_cast<uint64_t>(rows)};
    const uint32_t wbox[2] = {64, 64};
    if (encode_tensor_map(&a.tmW, d.weights, 2, wdims, astr, wbox, true)) return 1;
    const int want = (a.n_tiles + kNWG - 1) / kNWG;
    plan->grid = want < num_sms ? want : num_sms;
    plan->smem_bytes = chain_smem_bytes(d.prog);
    NDIFF_REQUIRE(plan->smem_bytes <= 227 * 1024, "pixel chain: shared-memory budget exceeded");
    plan->ts = chain_operands_in_tmem();
    return 0;
}

static int tail_smem_bytes() { return 1024 + kTailRows * 128 + kNWG * kTailWgBytes + static_cast<int>(sizeof(TailSmem)); }

int tail_chain_plan(const TailDesc& d, int num_sms, TailPlan* plan) {
    {
        static std::once_flag once;
        static int init_rc = 0;
        std::call_once(once, [] {
            init_rc = (cudaFuncSetAttribute(tail_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem_bytes()) == cudaSuccess &&
                       cudaFuncSetAttribute(tail_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem_bytes()) == cudaSuccess) ? 0 : 1;
        });
        NDIFF_REQUIRE(init_rc == 0, "tail chain: cannot opt in to the shared-memory size");
    }
    TailArgs& a = plan->args;
    memset(&a, 0, sizeof(a));
    NDIFF_REQUIRE(d.npix > 0 && d.HW > 0 && d.h2 && d.r1 && d.r2 && d.weights && d.fvec && d.stats && d.gamma && d.beta && d.out,
                  "tail chain: null argument");
    NDIFF_REQUIRE(d.groups > 0 && 64 % d.groups == 0, "tail chain: bad group count");
    const int gs = 64 / d.groups;
    NDIFF_REQUIRE(gs >= 8 && (gs & (gs - 1)) == 0, "tail chain: group size must be a power of two >= 8");
    a.npix = d.npix; a.HW = d.HW; a.n_tiles = (d.npix + kTile - 1) / kTile;
    a.fvec = d.fvec; a.stats = d.stats; a.gamma = d.gamma; a.beta = d.beta; a.G = d.groups; a.eps = 1e-5f;
    NDIFF_REQUIRE(d.real_frac > 0.f && d.real_frac <= 1.f, "tail chain: live channel fraction must be in (0, 1]");
    a.real_frac = d.real_frac;
    a.lgs = 0;
    while ((1 << a.lgs) < gs) ++a.lgs;
    a.out = reinterpret_cast<float4*>(d.out);
    const uint64_t adims[2] = {64, static_cast<uint64_t>(d.npix)};
    const uint64_t astr[1] = {128};
    const uint32_t abox[2] = {64, kTile};
    if (encode_tensor_map(&a.tmH, d.h2, 2, adims, astr, abox, true)) return 1;
    if (encode_tensor_map(&a.tmR1, d.r1, 2, adims, astr, abox, true)) return 1;
    if (encode_tensor_map(&a.tmR2, d.r2, 2, adims, astr, abox, true)) return 1;
    const uint64_t wdims[2] = {64, static_cast<uint64_t>(kTailRows)};
    const uint32_t wbox[2] = {64, 64};
    if (encode_tensor_map(&a.tmW, d.weights, 2, wdims, astr, wbox, true)) return 1;
    const int want = (a.n_tiles + kNWG - 1) / kNWG;
    plan->grid = want < num_sms ? want : num_sms;
    plan->smem_bytes = tail_smem_bytes();
    NDIFF_REQUIRE(plan->smem_bytes <= 227 * 1024, "tail chain: shared-memory budget exceeded");
    plan->ts = chain_operands_in_tmem();
    return 0;
}

int tail_chain_launch(const TailPlan& plan, cudaStream_t stream) {
    NDIFF_CUDA_OK(plan.ts ? launch_pdl(tail_chain_kernel<true>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args)
                          : launch_pdl(tail_chain_kernel<false>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args));
    return 0;
}

int pixel_chain_launch(const ChainPlan& plan, cudaStream_t stream) {
    if (plan.prog == kProgShot)
        NDIFF_CUDA_OK(plan.ts ? launch_pdl(pixel_chain_kernel<kProgShot, true>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args)
                              : launch_pdl(pixel_chain_kernel<kProgShot, false>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args));
    else
        NDIFF_CUDA_OK(plan.ts ? launch_pdl(pixel_chain_kernel<kProgAttn, true>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args)
                              : launch_pdl(pixel_chain_kernel<kProgAttn, false>, dim3(plan.grid), dim3(kThreads), plan.smem_bytes, stream, plan.args));
    return 0;
}

}  // namespace ndiff
