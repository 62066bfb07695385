// noisediff_b200 — backward / optimizer kernels of the diffusion TRAINING step (declarations).
//
// Reference: GaussianDiffusion.p_losses + loss.backward() + Adam.step() + EMA.update()
// (models/denoising_diffusion_pytorch.py:481-531, models/trainer_diffusion.py:176-191, :63-69).  Activations and activation
// gradients are NHWC bf16 (as in the sampling path), parameter gradients fp32 in the state_dict's own layouts, reductions fp32.
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace ndiff {

// out[b][c] (+)= sum over the pixels of sample b of x[b, p, c]  (per-sample column sums: bias / per-sample-vector gradients).
// per_sample == false: one row, out[c] += sum over all pixels of all samples.  `out` must be zeroed by the caller (atomics).
int colsum_launch(const bf16* x, int x_ld, int x_off, float* out, int out_ld, bool per_sample, int B, int HW, int C,
                  cudaStream_t s);

// GroupNorm + (scale+1)/shift + SiLU backward (Block.forward, Diffusion_arch.py:135-144).
//   h      raw conv output (the GroupNorm input), stats its fixed-point sums from the conv epilogue
//   dout   gradient of SiLU(...) (residual branches are routed by the caller)
// pass 1 accumulates acc[b][c][2] = (sum_p dy1, sum_p dy1 * xhat) with dy1 = dout * SiLU'(y2) * (scale + 1 | 1); with `maps`
// (ResnetBlock2) it also writes the per-pixel map gradients dmaps[b, p, 2C] = (dy2 * y1 | dy2).
// pass 2 writes dh; the parameter kernel adds dgamma / dbeta and, for per-sample scale/shift, writes dss[b][off .. off+2C).
struct GnBwdArgs {
    const bf16* h; const bf16* dout; bf16* dh;
    const unsigned long long* stats;
    const float* gamma; const float* beta;
    const float* ss; int ss_ld; int ss_off;       // per-sample [scale C | shift C] or null
    const bf16* maps; bf16* dmaps;                // per-pixel [scale C | shift C] and its gradient, or null
    float* acc;                                   // [B][C][2] fp32 scratch (zeroed by the launcher)
    float* dgamma; float* dbeta;                  // [C] fp32, accumulated
    float* dbias;                                 // optional [C]: bias gradient of the conv that produced h (= column sums of dh),
                                                  // accumulated by the second pass instead of a separate read of dh
    float* dss; int dss_ld;                       // [B][dss_ld] fp32 gradient of the scale/shift table (written at ss_off), or null
    int B, HW, C, G;
    float eps, real_frac;
};
int gn_backward_launch(const GnBwdArgs& a, cudaStream_t s);

// LayerNorm_C(x + vec[b]) * g + beta backward (AttnBlock.norm2, Diffusion_arch.py:438-439).  dx (+)= ...; dg / dbeta accumulated.
// (the per-sample vector's gradient is the per-sample column sum of dx's LayerNorm part: the caller gets it from `dy_out`.)
int layernorm_backward_launch(const bf16* x, const float* vec, int vec_ld, const float* g, const bf16* du, bf16* dy_out,
                              float* dg, float* dbeta, int B, int HW, int C, float real_frac, cudaStream_t s);

// dpre = dy * GELU'(pre)  (exact-erf GELU, nn.GELU())
int gelu_backward_launch(const bf16* pre, const bf16* dy, bf16* dpre, size_t n, cudaStream_t s);
// y = GELU(pre) (training forward keeps the pre-activation)
int gelu_forward_launch(const bf16* pre, bf16* y, size_t n, cudaStream_t s);

// dst[p][dst_off + c] (+)= src[p][src_off + c], c < C   (gradient routing: residual adds, concat splits)
int add_slice_launch(bf16* dst, int dst_ld, int dst_off, const bf16* src, int src_ld, int src_off, int C, size_t npix,
                     bool accumulate, cudaStream_t s);

// nearest x2 upsample backward: dx[b,h,w,c] (+)= sum of the 2x2 block of dy[b,2h..,2w..,c]
int upsample2x_backward_launch(const bf16* dy, bf16* dx, int B, int H, int W, int C, bool accumulate, cudaStream_t s);
// space-to-depth backward: dx[b, 2h+p1, 2w+p2, c] (+)= t[b, h, w, (p1*2+p2)*C + c]
int depth_to_space_launch(const bf16* t, bf16* dx, int B, int H, int W, int C, bool accumulate, cudaStream_t s);

// Heads + loss (Diffusion_arch.py:643-644, denoising_diffusion_pytorch.py:518-531, pred_v / pred_noise objectives):
//   v = final_conv(xf) + shot_mlp3.fc2(sf);  loss = mean_b [ w_b * mean_{c,h,w} (v - target)^2 ]
// writes dxf, dsf (bf16 [npix][C]), accumulates the four head parameter gradients and the loss value (double).
struct HeadsBwdArgs {
    const bf16* xf; const bf16* sf;               // [npix][C]
    const float* wf; const float* ws;             // [4][C]
    const float4* v; const float4* target;        // fp32 NHWC4 network output / regression target
    const float* w_b;                             // [B] loss weights (loss_weight[t_b])
    bf16* dxf; bf16* dsf;
    float* dwf; float* dbf; float* dws; float* dbs;
    double* loss;                                 // scalar, accumulated
    int B, HW, C;
};
int heads_backward_launch(const HeadsBwdArgs& a, cudaStream_t s);

// shot_mlp1.fc1 on cat[clean, x_t] (8 -> C) + GELU: parameter gradients from ds0 (gradient of the GELU output)
int shot_in_backward_launch(const float* clean, const float* x, const float* w, const float* bias, const bf16* ds0, float* dw,
                            float* db, size_t npix, int C, cudaStream_t s);

// init_conv 7x7 (4 -> C) parameter gradients: dw[co][ci][7][7] += sum_p dy[p][co] x[p + tap][ci]; db via colsum
int init_conv_wgrad_launch(const float* x_nhwc4, const bf16* dy, float* dw, int B, int H, int W, int C, cudaStream_t s);

// time path with saved intermediates and its backward (Diffusion_arch.py:94-107,502-507 + every ResnetBlock.mlp head)
//   saved[n] = [emb dim | a1 td | a2 td]  (pre-activations), st_out[n] = SiLU(a2)
int time_mlp_train_launch(const int* t, int n, int dim, const float* w1, const float* b1, const float* w2, const float* b2,
                          float* st_out, float* saved, cudaStream_t s);
int time_mlp_backward_launch(const float* dst, const float* saved, int n, int dim, const float* w1, const float* w2, float* dw1,
                             float* db1, float* dw2, float* db2, cudaStream_t s);
// C[M][N] (+)= op(A) op(B): tiny fp32 GEMMs of the time / iso paths (one thread per output element)
//   tA == false: A is [M][K] (lda), true: A is [K][M];  tB == false: B is [K][N] (ldb), true: B is [N][K]
int small_gemm_launch(bool tA, bool tB, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                      bool accumulate, cudaStream_t s);
// out[c] += sum_n in[n][c]
int rowsum_f32_launch(const float* in, int ld, int n, int C, float* out, cudaStream_t s);

// collapsed cross attention c = to_out(to_v(iso_embed[idx])) backward for one AttnBlock: dc [B][ld] at offset off
int iso_vec_backward_launch(const float* emb_table, const long long* idx, const float* wv, const float* wo, const float* dc,
                            int dc_ld, int dc_off, float* demb, float* dwv, float* dwo, float* dbo, int B, int C, cudaStream_t s);

// positional path backward (pos_enc -> pos_mlp -> SiLU -> the two ResnetBlock2 heads), from the two map gradients [B,HW,2C]
struct PosBwdArgs {
    PosArgs fwd;                                  // forward weights / position (map pointers unused)
    const bf16* dmap1; const bf16* dmap2;
    float* dwe; float* dbe; float* dw1; float* db1; float* dw2; float* db2; float* dwm1; float* dbm1; float* dwm2; float* dbm2;
};
int pos_backward_launch(const PosBwdArgs& a, cudaStream_t s);

// Adam (torch.optim.Adam semantics, models/trainer_diffusion.py:92) over a flat fp32 parameter buffer, and the EMA lerp
int adam_launch(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, float grad_scale, cudaStream_t s);
int ema_lerp_launch(float* ema, const float* p, size_t n, float weight, cudaStream_t s);   // ema += weight * (p - ema)

}  // namespace ndiff
