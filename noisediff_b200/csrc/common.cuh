// noisediff_b200 — shared device/host helpers for the sm_100a kernels.
// Thin inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM), bf16 packing, and host-side error plumbing.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <utility>

namespace ndiff {

// ------------------------------------------------------------------------------------------------------------
// host-side error plumbing (the C ABI never throws; it records a message and returns a status)
// ------------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
const char* get_error();

#define NDIFF_CUDA_OK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::ndiff::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " @" + \
                               __FILE__ + ":" + std::to_string(__LINE__));                        \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define NDIFF_REQUIRE(cond, msg)                                                     \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::ndiff::set_error(std::string("requirement failed: ") + #cond + " — " + \
                               (msg) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

// Programmatic dependent launch: kernels of one reverse step are chained with programmatic edges, so a kernel's CTAs
// start (barrier init, TMEM allocation, weight prefetch) as soon as SM resources free up, and block in pdl_wait() until the
// previous kernel has completed and flushed.  g_use_pdl is process-wide (NDIFF_FLAG_PDL sets it; measured on B200 it
// buys nothing inside a CUDA graph, so it is off by default).
extern bool g_use_pdl;

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- programmatic dependent launch -------------------------------------------------------------------------
// wait: every memory operation of the prerequisite grids is complete and visible (no-op without a programmatic edge);
// trigger: this thread no longer holds back the launch of the dependent grid.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug becomes a trap (an error the host sees) instead of a hung GPU.  The common case (the
// phase already completed, or completes within the hardware's own suspend window) is one inline instruction; the
// retry loop lives out of line so the single-thread producer / MMA loops stay small.
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    printf("ndiff: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, addr,
           parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done) mbar_wait_slow(addr, parity);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait(smem_u32(bar), parity); }
// 32-bit-address flavours for hot single-thread loops (barrier addresses are computed once)
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- TMA ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ---- TMA stores, proxy fence, raw shared-memory vector access ------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 256-bit global accesses (sm_100): one instruction moves a full 32-byte sector per lane — half the LSU / L1 tag work of two
// 128-bit accesses for the row-per-thread epilogue pattern.  `p` must be 32-byte aligned.
struct alignas(32) uint8x { uint4 lo, hi; };
__device__ __forceinline__ uint8x ldg256(const void* p) {
    uint8x v;
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg256(void* p, const uint4& lo, const uint4& hi) {
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
                 "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, descriptors passed as (lo, hi) halves so the hot loop only touches the 32-bit address word; `kAccum` is a
// compile-time predicate (the first MMA of a tile overwrites the accumulator).
template <bool kAccum>
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc) {
    if constexpr (kAccum) {
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "setp.eq.u32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "setp.ne.u32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    }
}
// A operand from TENSOR MEMORY (M = 128 rows = TMEM lanes, 16-bit elements packed two per 32-bit column, K-major; 8 columns per
// K = 16 step), B from shared memory.  An epilogue that produces the next GEMM's A operand writes it with tcgen05.st instead of
// shared memory: no swizzled stores, and the tensor core does not read it back through the shared-memory pipe.
template <bool kAccum>
__device__ __forceinline__ void umma_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if constexpr (kAccum) {
        asm volatile(
            "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.eq.u32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
            ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.u32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
            ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    }
}
// Same with a run-time accumulate flag (first MMA of a tile clears the accumulator).
__device__ __forceinline__ void umma_bf16_lohi_pred(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                    uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.u32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Weight-stationary form (conv_gemm.cu WS, default for the N = 64 kHalo2 kernels): the B operand goes through collector buffer kBuf (b0..b3).  kReuse ==
// false reads B from shared memory and keeps it in the collector (SASS: UTCHMMA.WS ... B_KEEP); kReuse == true multiplies by the
// kept copy without touching shared memory again (B_REUSE).  Two M = 128 sub-tiles that share one weight block then read it once.
template <int kBuf, bool kReuse>
__device__ __forceinline__ void umma_bf16_ws_pred(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
    static_assert(kBuf >= 0 && kBuf < 4, "four collector buffers");
#define NDIFF_WS_MMA(BUF, OP)                                                                                              \
    asm volatile(                                                                                                          \
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"                                                                    \
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"                                                              \
        "setp.ne.u32 p, %6, 0;\n\t"                                                                                        \
        "tcgen05.mma.ws.cta_group::1.kind::f16.collector::" BUF "::" OP " [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),       \
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)                                            \
        : "memory")
    if constexpr (kBuf == 0 && !kReuse) NDIFF_WS_MMA("b0", "fill");
    if constexpr (kBuf == 1 && !kReuse) NDIFF_WS_MMA("b1", "fill");
    if constexpr (kBuf == 2 && !kReuse) NDIFF_WS_MMA("b2", "fill");
    if constexpr (kBuf == 3 && !kReuse) NDIFF_WS_MMA("b3", "fill");
    if constexpr (kBuf == 0 && kReuse) NDIFF_WS_MMA("b0", "lastuse");
    if constexpr (kBuf == 1 && kReuse) NDIFF_WS_MMA("b1", "lastuse");
    if constexpr (kBuf == 2 && kReuse) NDIFF_WS_MMA("b2", "lastuse");
    if constexpr (kBuf == 3 && kReuse) NDIFF_WS_MMA("b3", "lastuse");
#undef NDIFF_WS_MMA
}
// One weight block (K = 64 = four K16 steps) against the two stacked 128-pixel sub-tiles of a kHalo2 tile: sub-tile 0 fills the
// four collector buffers, sub-tile 1 reuses them.  Same issue order as the plain form (sub-tile outer, K step inner).
__device__ __forceinline__ void umma_bf16_ws_pair(uint32_t d0, uint32_t d1, uint32_t a0, uint32_t a1, uint32_t a_hi, uint32_t b_lo,
                                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    umma_bf16_ws_pred<0, false>(d0, a0, a_hi, b_lo, b_hi, idesc, accumulate);
    umma_bf16_ws_pred<1, false>(d0, a0 + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
    umma_bf16_ws_pred<2, false>(d0, a0 + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
    umma_bf16_ws_pred<3, false>(d0, a0 + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
    umma_bf16_ws_pred<0, true>(d1, a1, a_hi, b_lo, b_hi, idesc, accumulate);
    umma_bf16_ws_pred<1, true>(d1, a1 + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
    umma_bf16_ws_pred<2, true>(d1, a1 + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
    umma_bf16_ws_pred<3, true>(d1, a1 + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
}
// Running-descriptor form: bumps both descriptor address words by compile-time increments and issues an accumulating
// MMA.  The in-place "+r" operands chain consecutive calls, so ptxas emits add / add / UTCHMMA per MMA instead of
// materialising (and spilling) a long list of independent descriptors.
template <int kAInc, int kBInc>
__device__ __forceinline__ void umma_bf16_step(uint32_t d_tmem, uint32_t& a_lo, uint32_t a_hi, uint32_t& b_lo, uint32_t b_hi,
                                               uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "add.u32 %0, %0, %6;\n\tadd.u32 %1, %1, %7;\n\t"
        "mov.b64 da, {%0, %2};\n\tmov.b64 db, {%1, %3};\n\t"
        "setp.eq.u32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%4], da, db, %5, p;\n\t}"
        : "+r"(a_lo), "+r"(b_lo)
        : "r"(a_hi), "r"(b_hi), "r"(d_tmem), "r"(idesc), "n"(kAInc), "n"(kBInc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Arrives on `bar` once every previously issued MMA of this thread has completed (implies fence::before).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i of the warp owns lane i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 16 registers -> 16 columns of this warp's 32 TMEM lanes (thread = lane)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart.
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64) | [46,48) version=1
//   | [49,52) base offset | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t base_offset = 0,
                                                    uint32_t sbo_bytes = 1024) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// The two 32-bit halves of the same descriptor: lo = start>>4 | LBO(=1)<<16, hi = SBO>>4 | version | layout.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFF) | (1u << 16); }
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) {
    return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
}
// Instruction descriptor, kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), both K-major, N>>3 at 17, M>>4 at 24.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Same with A = B = fp16 (format code 0 in bits 7 and 10).
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- misc --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// GELU (the reference's exact-erf nn.GELU()) as x * Phi(x) with Phi(x) = sigmoid(x * (a0 + a1 x^2 + a2 x^4)), x^2 clamped
// where the odd polynomial peaks.  Minimax fit against 0.5 x (1 + erf(x / sqrt 2)): max |error| 2.5e-5 over all x
// (tests/test_host.py pins the coefficients against scipy/torch erf) — an order of magnitude below the bf16 rounding of
// every consumer.  9 FP32 instructions + 2 MUFU per element instead of ~30 for a polynomial erf.
__device__ __forceinline__ float gelu_erf(float x) {
    constexpr float kL2e = 1.4426950408889634f;
    const float x2 = fminf(x * x, 52.6f);
    float p = fmaf(7.03035067e-04f * kL2e, x2, -7.40113019e-02f * kL2e);
    p = fmaf(p, x2, -1.59501576f * kL2e);                       // -(a0 + a1 x2 + a2 x2^2) * log2(e)
    return x * rcp_approx(1.0f + ex2_approx(x * p));
}
// Two GELUs per instruction stream in packed fp16: 0.5 x (1 + tanh(x (c0 + c1 x^2))), the tanh form of the same sigmoid
// fit truncated to two coefficients (max |error| vs exact-erf GELU 2.7e-4 in exact arithmetic; fp16 evaluation adds
// ~5e-4 relative).  Used where the result is the fp16 A operand of the next tensor-core GEMM inside a fused chain.
// 6 HFMA2-class instructions + 1 MUFU per PAIR of elements.
__device__ __forceinline__ uint32_t gelu_f16x2(float a, float b) {
    const __half2 x = __floats2half2_rn(a, b);
    const __half2 c0 = __float2half2_rn(0.80015698f), c1 = __float2half2_rn(0.034700935f), half = __float2half2_rn(0.5f);
    const __half2 x2 = __hmul2(x, x);
    const __half2 u = __hmul2(x, __hfma2(c1, x2, c0));
    uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
    const __half2 th = *reinterpret_cast<const __half2*>(&ti);
    const __half2 hx = __hmul2(x, half);
    const __half2 o = __hfma2(hx, th, hx);
    return *reinterpret_cast<const uint32_t*>(&o);
}
// libdevice erf for the run-once kernels (time MLP, positional path) whose outputs stay in fp32
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// SiLU = x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE MUFU op (tanh.approx.f32, max relative error 2^-11) and two FMA-class
// instructions per element.  The GroupNorm-apply passes were MUFU-bound with the exp + reciprocal form (2 MUFU per element at
// 16 lanes/clk/SM); |error| <= 2.5e-4 |x|, below the bf16 rounding of the stored result except in the far negative tail.
__device__ __forceinline__ float tanh_approx(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float silu(float x) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__

}  // namespace ndiff
