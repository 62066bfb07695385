// noisediff_b200 — HBM-bound kernels around the tensor-core convolutions (declarations).
#pragma once
#include "common.cuh"

namespace ndiff {

typedef __nv_bfloat16 bf16;

// One reverse-diffusion step's scalars (host builds the table from GaussianDiffusion's fp32 buffers).
//   x0   = clamp(p*x + q*out)                                (objective: pred_v / pred_noise / pred_x0)
//   eps  = (r1*x - x0) / r2
//   x'   = ((a*x0 + b*x) + c*eps) + sigma*z                  (DDPM: a,b = posterior coefs, c = 0; DDIM: b = 0)
struct StepParams {
    int t;
    float p, q, a, b, c, r1, r2, sigma;
    int clip;
    int pad_[2];
};
static_assert(sizeof(StepParams) == 48, "StepParams layout is part of the C ABI");

struct ChainState {      // device-resident, advanced by the last kernel of every step
    int step;            // index into the step table
    int n_steps;
    unsigned long long seed;
    StepParams cur;
    // I/O block rewritten by every ndiff_chain_run call (keep the order: engine.cu copies it as one blob)
    int base_step;       // step index at which the current run's noise / teacher / snapshot arrays start
    int pad_;
    const float* noise;  // injected N(0,1) [n][B,4,H,W] fp32 NCHW, or null -> in-kernel Philox
    const float* teacher;// teacher-forced network inputs [n][B,4,H,W] fp32 NCHW, or null
    float* snap;         // receives x after every step, [n][B,4,H,W] fp32 NCHW, or null
};

// GroupNorm(+affine) -> x*(scale+1)+shift -> SiLU (-> + residuals).  `x` is the conv output (bf16 NHWC) whose
// per-(sample, group) sums were accumulated by the conv epilogue.  ref Block.forward, Diffusion_arch.py:135-144.
struct GnApplyArgs {
    const bf16* x; bf16* out;
    const unsigned long long* stats;               // 2^-24 fixed-point sums from the conv epilogue
    const float* gamma; const float* beta;
    const float* ss; int ss_ld; int ss_off;       // per-sample [scale C | shift C] at ss[b*ss_ld + ss_off], or null
    const bf16* maps;                              // per-pixel [scale C | shift C] bf16 [B,HW,2C], or null
    const bf16* res1; const bf16* res2;            // optional residuals added after SiLU
    int B, HW, C, G;
    float eps;
    float real_frac;                               // live fraction of every group's channels (zero-padded layouts, engine.cu
                                                   // "physical channels"); 0 is read as 1.  Sums over the padding are 0, so
                                                   // only the element COUNT of the statistics changes.
};
int gn_apply_launch(const GnApplyArgs& a, cudaStream_t s);

// u = LayerNorm_C(x + vec[b]) * g + beta   (AttnBlock.norm2 on the collapsed attention; Diffusion_arch.py:438-439)
// real_frac: live fraction of the C channels (the rest is zero padding that must not enter mean / variance); 1 = all.
int layernorm_launch(const bf16* x, const float* vec, int vec_ld, const float* g, const float* beta, bf16* out, int B,
                     int HW, int C, cudaStream_t s, float real_frac = 1.0f);

// shot_mlp1.fc1 on cat[clean, x_t] (8 -> C) + GELU  (Diffusion_arch.py:598, Mlp :340-356)
int shot_in_launch(const float* clean, const float* x, const float* w, const float* bias, bf16* out, int npix, int C,
                   cudaStream_t s);

// init_conv 7x7 pad 3, 4 -> C (fp32 math on CUDA cores; C_in = 4 is too thin for a tensor-core tile)
int init_conv7_launch(const float* x, const float* w_tap_ci_co, const float* bias, bf16* out, int B, int H, int W, int C,
                      cudaStream_t s);

// nearest-neighbour x2 (Upsample, Diffusion_arch.py:72-76) — materialised, the 3x3 conv follows
int upsample2x_launch(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s);

// final_conv (C->4) + shot_mlp3.fc2 (C->4) + sum = network output; then either store it or apply the posterior update
struct FinalArgs {
    const bf16* xf; const bf16* sf;          // final_res_block output, shot_mlp3.fc1 (post-GELU) output  [npix, C]
    const float* wf; const float* bfin;      // final_conv  weight [4][C], bias [4]
    const float* ws; const float* bs;        // shot_mlp3.fc2 weight [4][C], bias [4]
    int npix; int C; int HW;
    float* v_out;                            // fp32 NHWC4 network output (forward API) or null
    // posterior update (null chain => skipped); noise / snapshot pointers live in the ChainState
    ChainState* chain;
    float* x;                                // fp32 NHWC4 state (= this step's network input), updated in place
    // optional (dim = 64): xf is a raw conv output and the heads read SiLU(GroupNorm(xf)) + gn_res instead — the last
    // GroupNorm-apply pass of final_res_block folded into this kernel.  gn_stats: that conv's fixed-point sums [B][G][2].
    const unsigned long long* gn_stats; const float* gn_gamma; const float* gn_beta; const bf16* gn_res;
    int gn_G; float gn_eps;
    float gn_real_frac;                      // as GnApplyArgs::real_frac
    // optional (dim = 64): the shot-noise head was already evaluated by the shot-branch tail kernel — sn[pix] = shot_mlp3 output
    // INCLUDING its bias; sf / ws / bs are then unused.
    const float4* sn;
};
int final_launch(const FinalArgs& a, cudaStream_t s);

// time path: sinusoidal(dim) -> Linear -> GELU -> Linear -> SiLU  (Diffusion_arch.py:94-107,502-507,163)
int time_mlp_launch(const int* t, int t_stride, int n, int dim, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* st_out, cudaStream_t s);
// out[n][rows] = W[rows][K] @ st[n][K] + bias  (all ResnetBlock.mlp Linears stacked)
int rows_gemv_launch(const float* W, const float* bias, const float* in, float* out, int n, int rows, int K,
                     cudaStream_t s);

// collapsed cross attention: c = to_out(to_v(iso_embed[idx]))  (softmax over ONE key == 1)
int iso_vec_launch(const float* emb_table, const long long* idx, const float* wv, const float* wo, const float* bo,
                   float* out, int out_ld, int out_off, int B, int C, cudaStream_t s);

// positional path (step-invariant): pos_enc -> pos_mlp -> the two ResnetBlock2 scale/shift maps
struct PosArgs {
    const float* position;                    // fp32 NCHW (B,2,H,W)
    const float* we; const float* be;         // pos_enc.weights [8][2], [8]
    const float* w1; const float* b1;         // pos_mlp.fc1 [16][24], [16]
    const float* w2; const float* b2;         // pos_mlp.fc2 [8][16], [8]
    const float* wm1; const float* bm1;       // pos_block1.mlp.1 [2C][8], [2C]
    const float* wm2; const float* bm2;       // pos_block2.mlp.1
    bf16* map1; bf16* map2;                   // [B,HW,2C]
    float* pos_emb;                           // optional fp32 [B,HW,8] (tests)
    int B, HW, C;
};
int pos_maps_launch(const PosArgs& a, cudaStream_t s);

// layout conversion at the API boundary
int nchw_to_nhwc4_launch(const float* in, float* out, int B, int HW, cudaStream_t s);
int nhwc4_to_nchw_launch(const float* in, float* out, int B, int HW, cudaStream_t s);

// consumer contract (dataloader/dataset_denoising.py:140-144): noisy = clip(clip(noise, -1, 1) + clean, 0, 1); clean_out = clip(clean, 0, 1)
// (fp32, any layout: purely elementwise; n4 = number of float4 elements; clean_out may be null)
int compose_noisy_launch(const float* noise, const float* clean, float* noisy_out, float* clean_out, size_t n4, cudaStream_t s);

int pointwise_init();   // one-time kernel attribute setup (call outside stream capture)
int philox_normal_launch(float* out, size_t n4, unsigned long long seed, unsigned long long stream_id, cudaStream_t s);

}  // namespace ndiff
