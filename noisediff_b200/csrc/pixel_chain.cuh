// noisediff_b200 — fused per-pixel MLP chains on tensor cores (declarations).
//
// NoiseDiffNet has long runs of 1x1 layers where every pixel is processed independently: AttnBlock with its 1-token cross
// attention collapsed (LayerNorm -> Linear -> GELU -> Linear -> +res -> 1x1 conv -> +res, Diffusion_arch.py:425-443), and
// the head of the shot-noise branch (shot_mlp1 -> shot_attn -> shot_mlp2, :598-601).  Run layer by layer these are
// HBM-bound round trips of a 64-channel activation per layer; here a CTA keeps a 128-pixel tile on chip for the whole
// chain: weights and TMA-landed tiles in shared memory (SWIZZLE_128B rows), accumulators in tensor memory, and every epilogue
// writes the next layer's A operand straight back into tensor memory (tcgen05.st; the next GEMM reads A from TMEM).
#pragma once
#include "common.cuh"

namespace ndiff {

enum ChainProg : int {
    kProgAttn = 0,   // out = AttnBlock(x)                                   (C = 64)
    kProgShot = 1,   // s1 = shot_mlp1(cat[clean, x_t]); out = shot_mlp2(shot_attn(s1))   (dim = 64)
};

// rows of the bf16 weight blob ([rows][64], one 128-byte row per output channel per 64-wide K block) and floats of the
// parameter block, in program order:
//   attn : W1 (ff.net.0.0, 128 rows) | Wp W2 kblock 0, 1 (proj_out o ff.net.2 folded, 64 + 64, fp16) | Wp (proj_out, 64) = 320 rows
//          [reserved 128][b1 128][unused 128]                                                               = 384 floats
//          (AttnBlock.norm2's affine is folded by the packer: W1 <- W1 diag(ln_g), b1 <- b1 + W1 ln_b; ff.net.2 and proj_out
//          are one stage: out = (Wp W2) h + Wp x + [Wp (b2 + c) + bp] + x, the bracket arriving per sample as cvec2)
//   shot : W0 (shot_mlp1.fc1, K = 8 zero-padded, 64) | Wfc2 (shot_mlp1.fc2, 64, fp16) | W1 128 |
//          Wm1 Wp W2 kblock 0, 1 (64 + 64, fp16, applied to the hidden layer) | Wm1 Wp + Wm1 (64, applied to s1) | Wm2 (fp16) = 512 rows
//          (shot_attn.ff.net.2, shot_attn.proj_out and shot_mlp2.fc1 meet without a nonlinearity and are ONE stage:
//             fc1(Wp (W2 h + b2 + c + s1) + bp + s1) = (Wm1 Wp W2) h + (Wm1 Wp + Wm1) s1 + [Wm1 Wp (b2 + c) + Wm1 bp + bm1],
//           the bracket arriving per sample as cvec2)
//          [b0 64][bfc2 64] + attn 384 (only b1 is read) + [unused 64][bm2 64]                                = 640 floats
constexpr int kChainAttnRows = 320, kChainAttnFloats = 384;
constexpr int kChainShotRows = 512, kChainShotFloats = 640;

struct ChainArgs {
    CUtensorMap tmX;      // attn: input activation viewed as [npix][64] bf16, box {64, 128}
    CUtensorMap tmW;      // weight blob [rows][64] bf16, box {64, 64}
    CUtensorMap tmOut;    // output activation [npix][64]
    CUtensorMap tmOut2;   // shot: shot_mlp1 output (kept as the branch's residual r_s)
    int npix, HW, n_tiles;
    const float* fvec;    // parameter block (see above)
    const float* cvec;    // collapsed cross-attention vector of this block, per sample: cvec[b * cvec_ld + c]
    const float* cvec2;   // per-sample vector of the folded last linear stage (attn: Wp (b2 + c) + bp; shot: see above), same ld
    int cvec_ld;
    float inv_c;          // 1 / (live channels of the 64): LayerNorm's element count under zero-padded channel layouts
    const float4* clean;  // shot: fp32 NHWC4 clean image and chain state
    const float4* x;
};

struct ChainPlan {
    ChainArgs args;
    int prog;
    int grid;
    int smem_bytes;
    bool ts;              // GELU outputs reach the next GEMM as its A operand through tensor memory
};

struct ChainDesc {
    int prog = kProgAttn;
    int npix = 0, HW = 0;
    const __nv_bfloat16* x = nullptr;       // attn input [npix][64]
    const __nv_bfloat16* weights = nullptr; // blob
    const float* fvec = nullptr;
    const float* cvec = nullptr; int cvec_ld = 0;
    const float* cvec2 = nullptr;           // folded-stage vector (see ChainArgs)
    float real_frac = 1.0f;                 // live fraction of the 64 channels (engine.cu "physical channels")
    const float* clean = nullptr; const float* xt = nullptr;   // shot inputs (fp32 NHWC4)
    __nv_bfloat16* out = nullptr;
    __nv_bfloat16* out2 = nullptr;
};

// ---- tail of the shot-noise branch (Diffusion_arch.py:602-604) as ONE per-pixel kernel ---------------------------------------
//   y = SiLU(GroupNorm(h2)) + s4 + s1        shot_time.block2.norm + the ResnetBlock residual + r_s   (:135-170, :603)
//   shot_noise = fc2(GELU(fc1(y)))           shot_mlp3                                                  (:604)
// h2 is the RAW output of shot_time.block2.proj (its GroupNorm sums come from that conv's epilogue); the result is the
// 4-channel fp32 shot-noise image the heads kernel adds to final_conv's output.
// weight blob: [128][64] 16-bit rows = fc1 (64 rows, bf16) | fc2 (16 rows, fp16, rows 4..15 zero) | zero padding;
// fvec: [b_fc1 64][b_fc2 4][pad 60].
constexpr int kTailRows = 128, kTailFloats = 128;
struct TailArgs {
    CUtensorMap tmH, tmR1, tmR2, tmW;
    int npix, HW, n_tiles;
    const float* fvec;
    const unsigned long long* stats;    // [B][G][2] fixed-point sums of h2
    const float* gamma; const float* beta;
    int G, lgs; float eps;
    float real_frac;                    // live fraction of every group's channels (statistics count)
    float4* out;                        // [npix] fp32 x 4
};
struct TailPlan { TailArgs args; int grid; int smem_bytes; bool ts; };
struct TailDesc {
    int npix = 0, HW = 0;
    const __nv_bfloat16* h2 = nullptr; const __nv_bfloat16* r1 = nullptr; const __nv_bfloat16* r2 = nullptr;
    const __nv_bfloat16* weights = nullptr; const float* fvec = nullptr;
    const unsigned long long* stats = nullptr; const float* gamma = nullptr; const float* beta = nullptr; int groups = 0;
    float real_frac = 1.0f;
    float* out = nullptr;
};
int tail_chain_plan(const TailDesc& d, int num_sms, TailPlan* plan);
int tail_chain_launch(const TailPlan& plan, cudaStream_t stream);

int pixel_chain_plan(const ChainDesc& d, int num_sms, ChainPlan* plan);
int pixel_chain_launch(const ChainPlan& plan, cudaStream_t stream);
int pixel_chain_init();

// dst[(kb * N + n) * 64 + kk] = src[n * K + kb * 64 + kk] (zero beyond K): fp32 [N][K] -> 16-bit K-blocked rows.
// f16 = true for the layers whose A operand is a GELU output (kept in fp16 on chip: ff.net.2, shot_mlp1.fc2, shot_mlp2.fc2).
int pack_chain_weight_launch(const float* src, __nv_bfloat16* dst, int N, int K, bool f16, cudaStream_t s);
// w_out = a @ w (fp32 [N][N] row-major each);  b_out[n] = b2[n] + sum_k a[n][k] b1[k]   (two stacked Linears folded into one)
int fold_linear_launch(const float* a, const float* w, const float* b1, const float* b2, float* w_out, float* b_out, int N,
                       cudaStream_t s);
// w_out[n][k] = w[n][k] * g[k];  b_out[n] = b[n] + sum_k w[n][k] * beta[k]   (LayerNorm affine folded into the next Linear)
int fold_layernorm_launch(const float* w, const float* b, const float* g, const float* beta, float* w_out, float* b_out, int N,
                          int K, cudaStream_t s);

}  // namespace ndiff
