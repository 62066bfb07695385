// noisediff_b200 — tcgen05 implicit-GEMM convolution over NHWC bf16 activations (declarations).
#pragma once
#include "common.cuh"

namespace ndiff {

enum ConvMode : int {
    kDirect = 0,  // every (channel-block, tap) k-block loads its own shifted 128-pixel box (1x1, 7x7-row trick, debug 3x3)
    kS2D = 2,     // 2x2 stride-2 (space-to-depth + 1x1) through a 5-D view of the input
    kHaloUp = 5,  // nearest x2 upsample + 3x3 pad 1 (Upsample, Diffusion_arch.py:72-76) WITHOUT materialising the upsampled
                  // tensor: output phase (py, px) of pixel (2y+py, 2x+px) is a 2x2 convolution of the low-resolution input
                  // with pre-summed weights (rows {y-1: w0, y: w1+w2} for py = 0, {y: w0+w1, y+1: w2} for py = 1; same in x).
                  // Tiles walk (low-res 16x8 tile, phase); the halo box is the kHalo1 one.  4/9 of the direct-form MACs.
                  // Weights: [Cout][cblk][phase(4)][tap(4)][64].  H, W in the descriptor are the LOW-resolution size.
    kHalo2 = 4,   // kHalo1 with TWO vertically stacked 16x8 sub-tiles per CTA tile (32x8 pixels, one 34x10 halo box, two TMEM
                  // accumulators): every weight block read from shared memory / streamed from L2 feeds 256 pixels
    kHalo1R = 6,  // kHalo1 plus the ResnetBlock's 1x1 res_conv of the SAME (concatenated) input (Diffusion_arch.py:157,169): the
                  // centre tap of the halo box is multiplied by a tenth weight block into a second TMEM accumulator, so the
                  // residual path costs 1/9 more MMAs instead of a second pass over the activations.
                  // Weights: [Cout][cblk][10][64] (taps 0..8 = 3x3, 9 = res_conv); second output / bias: out2 / bias2.
    kHalo1 = 3,   // 3x3 pad 1: ONE (TH+2)x(TW+2) = 18x10-pixel halo box per 64-channel block (TH x TW = 16 x 8); the nine
                  // taps are 128-byte row shifts of the UMMA start address with SBO = (TW+2)*128.  Works because the
                  // 128-B swizzle is a function of the absolute shared-memory address for both TMA and UMMA
                  // (verified on B200: tests/test_gpu_ops.py::test_conv3x3).
};
enum ConvAct : int { kActNone = 0, kActGelu = 1 };
constexpr float kStatScale = 16777216.0f;   // 2^24

// Kernel arguments (one struct, passed as a __grid_constant__ so the three tensor maps stay in param space).
struct ConvGemmArgs {
    CUtensorMap tmA0, tmA1, tmB;
    CUtensorMap tmOut, tmOut2;   // staged epilogue (stage_bytes > 0): output tensors as [B][H][W][C] maps, box = one 128-pixel sub-tile x 64 ch
    int mode;
    int B, H, W;           // OUTPUT spatial size (input is the same except kS2D: 2H x 2W)
    int TH, TW, lgTW;      // pixel tile (TH*TW == 128, TW power of two >= 8)
    int tiles_x, tiles_y;  // tiles per image
    int cb0, cb1;          // 64-channel blocks taken from source 0 / source 1 (concat order: 0 then 1)
    int taps_y, taps_x, pad_y, pad_x;  // tap grid of kDirect (kHalo3: 3,3,1,1; kS2D: 2,2,0,0)
    int tap_sy;            // kDirect: input-row step between consecutive ky taps (1; 2 for the row-paired 7x7 init_conv)
    int n_tiles;           // Cout / NT
    int total_tiles;
    int a_stages, b_stages;
    int b_region_bytes;    // bytes of shared memory holding B (ring, or the whole resident slice)
    int b_resident;        // 1: the CTA's whole [NT x Ktot] weight slice is loaded once and stays in shared memory
    int a_stage_bytes, a_copy_bytes;
    int stage_bytes;       // > 0 (the N = 64 XF kernels): the epilogue writes each 128-pixel x 64-channel sub-tile into a swizzled
                           // staging block and ONE TMA store moves it out, instead of two 32-byte-sector stores per lane — takes
                           // ~180k store wavefronts per SM off the LSU pipe that the in-place transform also uses
    int toeplitz;          // kDirect: the A operand of a tap is ONE row of TW + 8 sixteen-byte pixels read as overlapping 128-byte
                           // windows by a non-swizzled descriptor (LBO = 16 B, SBO = 128 B) — see ConvGemmDesc::toeplitz
    // epilogue
    const float* bias;           // [Cout] or null
    const float* vec;            // per-sample vector [B][vec_ld] added to every pixel, or null
    int vec_ld;
    const __nv_bfloat16* res;    // residual [B,H,W,res_ld] (same channel offset as the output), or null
    int res_ld;
    __nv_bfloat16* out;          // [B,H,W,out_ld]
    int out_ld;
    int act;
    const float* bias2;          // kHalo1R: bias and output of the fused 1x1 residual convolution
    __nv_bfloat16* out2;
    int out2_ld;
    // XF kernels: GroupNorm-apply + scale/shift + SiLU of the INPUT, evaluated in shared memory (see conv_gemm.cu)
    const unsigned long long* xf_stats;   // the producing conv's sums [B][xf_G][2]
    const float* xf_gamma; const float* xf_beta;
    const float* xf_ss; int xf_ss_ld;     // per-sample [scale C | shift C] at xf_ss[b * xf_ss_ld], or null
    int xf_G, xf_lgs; float xf_eps;
    float xf_real_frac;                   // live fraction of every input group's channels (statistics count)
    unsigned long long* stats;   // GroupNorm sums [B][G][2] (sum, sum of squares) in 2^-24 fixed point, or null
                                 // (integer atomics are associative: the result does not depend on tile order)
    int lgs;                     // log2(channels per group)
    int G;
};

// Host-side description of one convolution / GEMM launch, built once per layer at plan time.
struct ConvGemmPlan {
    ConvGemmArgs args;
    int NT;          // 64 or 128
    bool xf;         // input transform (fused GroupNorm-apply) variant
    bool ws;         // weight-stationary MMA form for the N = 64 kHalo2 kernels (default; NDIFF_NO_WS=1 selects the plain form)
    int grid;
    int smem_bytes;
};

struct ConvGemmDesc {
    int mode = kDirect;
    int B = 0, H = 0, W = 0;            // output spatial size
    const __nv_bfloat16* src0 = nullptr; int C0 = 0;   // NHWC bf16, C0 % 64 == 0
    const __nv_bfloat16* src1 = nullptr; int C1 = 0;   // optional second concat source
    int taps_y = 1, taps_x = 1, pad_y = 0, pad_x = 0;
    int tap_sy = 1;                      // kDirect: input-row step per ky tap
    // kDirect special: source 0 described by an explicit (possibly overlapping-window) 4-D map
    bool custom_src0 = false;
    uint64_t cdim[4] = {0, 0, 0, 0};     // dims innermost first
    uint64_t cstride[3] = {0, 0, 0};     // byte strides of dims 1..3
    // kDirect special (init_conv): source 0 is [B][rows][W + 8][8] bf16 (16-byte pixels) and output pixel x of a tap multiplies the
    // 64 values of pixels x .. x + 7.  Instead of fetching that 128-byte window per pixel through an overlapping tensor map (8x the
    // bytes from L2), the tile is ONE output row of 128 pixels, TMA lands its 136 input pixels once (2 176 B), and the MMA reads
    // the Toeplitz operand in place: K-major, no swizzle, core-matrix rows 16 B apart, LBO (next 16-byte K chunk) = 16 B,
    // SBO (next 8 rows) = 128 B.  Needs W % 128 == 0; cdim / cstride then describe the 16-byte-pixel tensor.
    bool toeplitz = false;
    const __nv_bfloat16* weight = nullptr;  // [Cout][Ktot] bf16, K order [cblk][tap][64]
    int Cout = 0;
    const float* bias = nullptr;
    const float* vec = nullptr; int vec_ld = 0;
    const __nv_bfloat16* res = nullptr; int res_ld = 0;
    __nv_bfloat16* out = nullptr; int out_ld = 0;
    int act = kActNone;
    const float* bias2 = nullptr; __nv_bfloat16* out2 = nullptr; int out2_ld = 0;   // kHalo1R
    unsigned long long* stats = nullptr; int groups = 0;
    // fused GroupNorm-apply of the input (kHalo1 / kHalo2, single source): statistics of the producing conv etc.
    const unsigned long long* xf_stats = nullptr; const float* xf_gamma = nullptr; const float* xf_beta = nullptr;
    const float* xf_ss = nullptr; int xf_ss_ld = 0; int xf_groups = 0; float xf_eps = 1e-5f;
    float xf_real_frac = 1.0f;           // zero-padded channel layouts: live fraction of every group (element count of the statistics)
    int force_nt = 0;                    // 0 = auto
    int TW = 0;                          // 0 = auto
};

// Fills `plan` (tensor maps, tiling, smem).  Returns 0 on success, 1 with set_error() otherwise.
int conv_gemm_plan(const ConvGemmDesc& d, int num_sms, ConvGemmPlan* plan);
int conv_gemm_launch(const ConvGemmPlan& plan, cudaStream_t stream);
int conv_gemm_init();   // one-time kernel attribute setup for the current device (call outside stream capture)

// Driver entry point for tensor-map encoding, resolved at run time (no link-time libcuda dependency).
int encode_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool swizzle128);

}  // namespace ndiff
