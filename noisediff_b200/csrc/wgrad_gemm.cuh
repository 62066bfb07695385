// noisediff_b200 — convolution WEIGHT gradients on tcgen05 (declarations).  Training row:
// loss.backward() through every nn.Conv2d / nn.Linear of NoiseDiffNet (models/archs/Diffusion_arch.py:128-443).
//
//   dW[co][ci][tap] += sum over pixels p of dY[p][co] * X[p + tap][ci]
//
// GEMM view: M = C_out, N = C_in (per tap), K = pixels.  Both operands are NHWC activations, i.e. the contraction index (the
// pixel) is the SLOW index of each shared-memory row: the tiles that TMA lands (one 128-byte SWIZZLE_128B row per pixel,
// 64 channels) are exactly tcgen05's MN-major canonical layout, so the MMAs run with a_major = b_major = MN and no
// transposing pass exists anywhere.  Taps are 128-byte row shifts of ONE halo box of X, as in the forward kernel.
#pragma once
#include "common.cuh"

namespace ndiff {

enum WgradMode : int {
    kWg1x1 = 0,   // 1x1 conv / token Linear: one tap, X tile = the dY tile's pixels
    kWg3x3 = 1,   // 3x3 pad 1: nine taps out of an 18 x 10 halo box of X (three per CTA: one kernel row)
    kWgS2D = 2,   // 2x2 stride-2 (space-to-depth + 1x1, Diffusion_arch.py:78-82): four taps through the 5-D view of X [B, 2H, 2W, C]
};

struct WgradArgs {
    CUtensorMap tmDY, tmX0, tmX1;
    int mode;
    int B, H, W;              // spatial size of dY (X is 2H x 2W for kWgS2D)
    int tiles_x, tiles_y;     // 16 x 8 pixel tiles per image
    int total_tiles;          // B * tiles_y * tiles_x
    int Cout, Cin;            // Cin = C0 + C1
    int cb0;                  // 64-channel blocks of source 0 (then source 1)
    int co_blocks, ci_blocks, tap_groups, taps_per_group, taps_total;
    int chunks;               // pixel-range splits (split-K): grid = units * chunks
    int stages;
    float* dw;                // fp32 [Cout][Cin][taps_total], accumulated with atomics
};
struct WgradPlan { WgradArgs args; int grid; int smem_bytes; };

struct WgradDesc {
    int mode = kWg1x1;
    int B = 0, H = 0, W = 0;
    const __nv_bfloat16* dy = nullptr; int Cout = 0;     // [B, H, W, Cout]
    const __nv_bfloat16* src0 = nullptr; int C0 = 0;     // [B, H(or 2H), W(or 2W), C0]
    const __nv_bfloat16* src1 = nullptr; int C1 = 0;     // optional second concat source
    float* dw = nullptr;
};
int wgrad_gemm_plan(const WgradDesc& d, int num_sms, WgradPlan* plan);
int wgrad_gemm_launch(const WgradPlan& plan, cudaStream_t stream);

}  // namespace ndiff
