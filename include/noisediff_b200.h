/*
 * noisediff_b200 — C ABI of the B200-native NoiseDiff sampling hot path.
 *
 * The reference (IVRL/NoiseDiff) is pure Python/PyTorch and has no FFI; this header is the boundary a maintainer
 * binds from Python (ctypes — see INTEGRATION.md).  Each entry point names the reference interface it replaces
 * (paths relative to the reference tree).  Conventions:
 *   - every function returns 0 on success, non-zero on failure; ndiff_last_error() then describes the failure.
 *     Nothing throws across the boundary.
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors); "host" pointers are CPU
 *     memory.  The library owns only its packed weights, workspace and CUDA graphs.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are asynchronous on that
 *     stream unless stated otherwise.
 *   - an engine belongs to one device and one caller thread at a time (the reference is single-threaded too).
 *   - tensors use the reference's layouts: images fp32 NCHW (B,4,H,W); position fp32 (B,2,H,W); indices int64.
 */
#ifndef NOISEDIFF_B200_H_
#define NOISEDIFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDIFF_ABI_VERSION 1

#if defined(__GNUC__)
#define NDIFF_API __attribute__((visibility("default")))
#else
#define NDIFF_API
#endif

typedef struct ndiff_engine ndiff_engine;

/* Geometry of one engine: the network of NoiseDiffNet(args) (models/archs/Diffusion_arch.py:447-573) for a fixed
 * micro-batch and crop.  dim = args.dim (64), height/width = crop (multiples of 8, ref :578). */
typedef struct ndiff_config {
    int32_t dim;
    int32_t batch;
    int32_t height;
    int32_t width;
    int32_t device;      /* CUDA ordinal */
    int32_t flags;       /* NDIFF_FLAG_* */
} ndiff_config;

#define NDIFF_FLAG_CONV_DIRECT 1   /* debug: 3x3 convs load every tap from L2 instead of the halo layout */
#define NDIFF_FLAG_NO_GRAPH    2   /* debug: launch kernels eagerly instead of replaying a CUDA graph */
#define NDIFF_FLAG_INIT_SIMT   8   /* debug: 7x7 init_conv on CUDA cores instead of the tensor-core window trick */
#define NDIFF_FLAG_KEEP_ACTS   4   /* debug: never recycle activation buffers (ndiff_debug_tensor sees every layer) */
#define NDIFF_FLAG_PDL         32  /* chain the kernels of a step with programmatic dependent launch (measured: no gain inside a
                                      CUDA graph on B200, so it is off by default) */
#define NDIFF_FLAG_HALO1       64  /* debug: 3x3 convs always use 128-pixel CTA tiles (one accumulator) */
#define NDIFF_FLAG_NO_XF       128 /* debug: keep block1's GroupNorm-apply as its own pass instead of evaluating it inside block2's conv */
#define NDIFF_FLAG_INIT_WINDOWS 256 /* debug: init_conv fetches a 128-byte window per pixel and tap (round-1 form) instead of one landed row per tap */
#define NDIFF_FLAG_UNFUSED     16  /* debug: run the per-pixel 1x1 chains layer by layer instead of as fused tensor-core chains */

/* One reverse step's scalars; the caller derives them from GaussianDiffusion's fp32 buffers
 * (models/denoising_diffusion_pytorch.py:240-266) so the arithmetic constants are the reference's own:
 *   x0  = clamp(p*x + q*net_out, -1, 1) if clip          predict_start_from_v/_noise :298-320, clamp :361
 *   eps = (r1*x - x0) / r2                                predict_noise_from_start :304-308
 *   x'  = ((a*x0 + b*x) + c*eps) + sigma*z                q_posterior :322-329 + p_sample :366-373 (c = 0)
 *                                                         ddim_sample :427-437 (b = 0)                       */
typedef struct ndiff_step {
    int32_t t;
    float p, q, a, b, c, r1, r2, sigma;
    int32_t clip;
    int32_t reserved[2];
} ndiff_step;

NDIFF_API int32_t     ndiff_abi_version(void);
NDIFF_API const char* ndiff_last_error(void);

/* Lifetime.  Replaces NoiseDiffNet.__init__ + .to(device) (Diffusion_arch.py:447-573, models/modules.py:73-83). */
NDIFF_API int32_t ndiff_engine_create(const ndiff_config* cfg, ndiff_engine** out);
NDIFF_API void    ndiff_engine_destroy(ndiff_engine* e);

/* Weights.  Replaces load_state_dict (models/trainer_diffusion.py:333-349): call once per state_dict entry with the
 * reference's key (e.g. "downs.0.0.block1.proj.weight") and fp32 data (host or device pointer), then finalize —
 * which fails if any of the live keys is missing or mis-shaped.  Dead keys (attn.to_q/to_k, norm1) are accepted
 * and ignored. */
NDIFF_API int32_t ndiff_load_param(ndiff_engine* e, const char* name, const float* data, int32_t ndim, const int64_t* shape);
/* Same, ordered on `stream` (cudaMemcpyAsync): behind the work that produced `data` on that stream and ahead of the repack that
 * ndiff_finalize_params enqueues on it — the form to use when reloading weights while other streams are busy (the stream-less
 * form above fences the whole device instead).  Loading any parameter invalidates the condition: call ndiff_set_condition
 * again before the next forward / chain. */
NDIFF_API int32_t ndiff_load_param_async(ndiff_engine* e, const char* name, const float* data, int32_t ndim,
                                         const int64_t* shape, void* stream);
NDIFF_API int32_t ndiff_finalize_params(ndiff_engine* e, void* stream);

/* Condition.  Replaces the step-invariant head of NoiseDiffNet.forward (Diffusion_arch.py:580-591): position
 * encoding + pos MLP + ResnetBlock2 scale/shift maps, ISO embedding -> collapsed cross-attention vectors. */
NDIFF_API int32_t ndiff_set_condition(ndiff_engine* e, const float* clean_dev, const float* position_dev,
                            const int64_t* iso_idx_dev, void* stream);

/* One network evaluation.  Replaces NoiseDiffNet.forward(x, time, condition) (Diffusion_arch.py:577-646). */
NDIFF_API int32_t ndiff_forward(ndiff_engine* e, const float* x_dev, const int64_t* time_dev, float* out_dev, void* stream);

/* Reverse chain.  Replaces GaussianDiffusion.p_sample_loop / ddim_sample
 * (models/denoising_diffusion_pytorch.py:375-444).
 *   begin : uploads the step table (host), sets x_T from x_init_dev (fp32 NCHW) or, if NULL, from Philox(seed).
 *   run   : executes the next n steps (one CUDA-graph replay each).  noise_dev: injected N(0,1) draws
 *           [n][B,4,H,W] fp32 NCHW for those steps, or NULL for in-kernel Philox.  teacher_dev (optional):
 *           network inputs [n][B,4,H,W] replacing the running state (teacher-forced parity checks).
 *           snapshots_dev (optional): receives x after each of the n steps, [n][B,4,H,W].
 *   read  : copies the current state to out_dev (fp32 NCHW). */
NDIFF_API int32_t ndiff_chain_begin(ndiff_engine* e, const ndiff_step* steps_host, int32_t n_steps, const float* x_init_dev,
                          uint64_t seed, void* stream);
NDIFF_API int32_t ndiff_chain_run(ndiff_engine* e, int32_t n, const float* noise_dev, const float* teacher_dev,
                        float* snapshots_dev, void* stream);
NDIFF_API int32_t ndiff_chain_read(ndiff_engine* e, float* out_dev, void* stream);
/*   seek  : re-positions the chain at `step` with state x_dev (fp32 NCHW) — lets ONE engine advance several
 *           micro-batches of a larger batch in lock-step (their states are swapped in and out between chunks);
 *           `seed` is that micro-batch's Philox seed. */
NDIFF_API int32_t ndiff_chain_seek(ndiff_engine* e, int32_t step, const float* x_dev, uint64_t seed, void* stream);

/* End-to-end convenience with HOST buffers (what Trainer.test() does per batch, models/trainer_diffusion.py:256-317:
 * host->device copies, full chain, device->host copy).  Synchronous. */
NDIFF_API int32_t ndiff_sample_host(ndiff_engine* e, const float* clean_host, const float* position_host,
                          const int64_t* iso_idx_host, const ndiff_step* steps_host, int32_t n_steps, uint64_t seed,
                          float* out_host);

/* Consumer contract.  Replaces the composition SyntheticNoisDiffDenoisingDataset does with a generated crop
 * (dataloader/dataset_denoising.py:140-144): noisy = clip(clip(noise, -1, 1) + clean, 0, 1), clean_out = clip(clean, 0, 1)
 * (clean_out may be NULL).  fp32 device buffers of n elements (n % 4 == 0, 16-byte aligned), any layout. */
NDIFF_API int32_t ndiff_compose_noisy(const float* noise_dev, const float* clean_dev, float* noisy_out_dev, float* clean_out_dev,
                                      int64_t n, void* stream);

/* Introspection used by tests / bench. */
NDIFF_API int32_t ndiff_debug_tensor(ndiff_engine* e, const char* name, float* out_dev_nchw, int64_t* shape4, void* stream);
NDIFF_API int64_t ndiff_launches_per_step(const ndiff_engine* e);
NDIFF_API double  ndiff_conv_flops_per_step(const ndiff_engine* e);    /* executed tensor-core FLOPs per network evaluation */
/* Per-op durations measured INSIDE the step: `iters` whole steps are enqueued back to back with a CUDA event between consecutive
 * ops; ms_out[i] is the median over the iterations.  Row 0 is the step prologue, the last row the fused heads + posterior update.
 * names_out: one line "name;flops;bytes" per row (executed FLOPs and algorithmic HBM bytes of that op).  After ndiff_chain_begin
 * the iterations are real chain steps (iters + 1 steps must be left), otherwise forward evaluations.  ms_out == NULL: only *n_out. */
NDIFF_API int32_t ndiff_time_layers(ndiff_engine* e, int32_t iters, float* ms_out, char* names_out, int32_t names_cap,
                          int32_t* n_out, void* stream);

/* Single-operator entry points (parity tests of the individual kernels; pointers are device pointers, activations
 * bf16 NHWC).  GroupNorm statistics are uint64 [B][groups][2] = (sum, sum of squares) in 2^-24 fixed point. */
NDIFF_API int32_t ndiff_op_conv(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                      int32_t C1, int32_t taps_y, int32_t taps_x, int32_t pad_y, int32_t pad_x, const void* weight_packed,
                      int32_t Cout, const float* bias, const float* vec, int32_t vec_ld, const void* res, int32_t act,
                      void* stats, int32_t groups, void* out, int32_t force_nt, int32_t tile_w, void* stream);
/* The fused forms of the ResnetBlock 3x3 convolutions (pad 1), as single operators:
 *   mode 6 (kHalo1R): weight_packed = [Cout][Cin/64][10][64] — taps 0..8 the 3x3 kernel, tap 9 the block's 1x1 res_conv
 *     (Diffusion_arch.py:157,169); ex->out2 / ex->bias2 receive res_conv(x).
 *   ex->xf_stats != NULL (modes 3 / 4, one source): the input is the RAW output of the previous conv and
 *     SiLU(GroupNorm(x) * (scale + 1) + shift) (Block.forward, :135-144) is applied to it inside the kernel;
 *     xf_stats = that conv's fixed-point sums [B][xf_groups][2], xf_ss = per-sample [scale C | shift C] rows or NULL. */
typedef struct ndiff_conv_ex {
    void* out2; const float* bias2;
    const void* xf_stats; const float* xf_gamma; const float* xf_beta; const float* xf_ss; int32_t xf_ss_ld; int32_t xf_groups;
} ndiff_conv_ex;
NDIFF_API int32_t ndiff_op_conv_ex(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                         int32_t C1, const void* weight_packed, int32_t Cout, const float* bias, void* stats, int32_t groups,
                         void* out, const ndiff_conv_ex* ex, void* stream);
/* Tail of the shot-noise branch (Diffusion_arch.py:602-604): y = SiLU(GroupNorm(h2)) + r1 + r2; out = fc2(GELU(fc1(y))).
 * h2 / r1 / r2: bf16 [npix][64]; weights_blob / fvec as documented in noisediff_b200/csrc/pixel_chain.cuh; out: fp32 [npix][4]. */
NDIFF_API int32_t ndiff_op_tail_chain(int32_t npix, int32_t HW, const void* h2, const void* r1, const void* r2,
                            const void* weights_blob, const float* fvec, const void* stats, const float* gamma, const float* beta,
                            int32_t groups, float* out_npix4, void* stream);
/* Same operator, launched `iters` times back to back after 3 warm-up launches and timed with CUDA events on `stream`
 * (kernel-variant selection experiments and the per-shape roofline table; blocks until done). */
NDIFF_API int32_t ndiff_op_conv_time(int32_t mode, int32_t B, int32_t H, int32_t W, const void* src0, int32_t C0, const void* src1,
                           int32_t C1, int32_t taps_y, int32_t taps_x, int32_t pad_y, int32_t pad_x,
                           const void* weight_packed, int32_t Cout, const float* bias, void* stats, int32_t groups, void* out,
                           int32_t force_nt, int32_t tile_w, int32_t iters, float* ms_per_launch, void* stream);
NDIFF_API int32_t ndiff_op_gn_apply(const void* x, void* out, const void* stats, const float* gamma, const float* beta,
                          const float* ss, int32_t ss_ld, int32_t ss_off, const void* maps, const void* res1,
                          const void* res2, int32_t B, int32_t HW, int32_t C, int32_t G, void* stream);
NDIFF_API int32_t ndiff_op_layernorm(const void* x, const float* vec, int32_t vec_ld, const float* g, const float* beta, void* out,
                           int32_t B, int32_t HW, int32_t C, void* stream);
NDIFF_API int32_t ndiff_op_philox_normal(float* out, int64_t n4, uint64_t seed, uint64_t stream_id, void* stream);

/* Fused per-pixel chain (prog 0: AttnBlock with the 1-token cross attention collapsed, Diffusion_arch.py:425-443;
 * prog 1: shot_mlp1 -> shot_attn -> shot_mlp2, :598-601).  x/out/out2: bf16 [npix][64]; clean/xt: fp32 [npix][4];
 * weights_blob: bf16 [rows][64] K-blocked rows and fvec: fp32 parameter block in the order documented in
 * noisediff_b200/csrc/pixel_chain.cuh; cvec: per-sample collapsed attention vector [npix/HW][cvec_ld]; cvec2: the per-sample
 * vector of the folded last linear stage, same leading dimension (prog 0: Wp (b2 + c) + bp of ff.net.2 + proj_out; prog 1:
 * Wm1 Wp (b2 + c) + Wm1 bp + bm1 of ff.net.2 + proj_out + shot_mlp2.fc1). */
NDIFF_API int32_t ndiff_op_pixel_chain(int32_t prog, int32_t npix, int32_t HW, const void* x, const float* clean_nhwc4,
                             const float* xt_nhwc4, const void* weights_blob, const float* fvec, const float* cvec,
                             int32_t cvec_ld, const float* cvec2, void* out, void* out2, void* stream);

/* ------------------------------------------------------------------------------------------------------------------------
 * Training step.  Replaces, for one batch: GaussianDiffusion.p_losses -> NoiseDiffNet.forward with per-sample t
 * (models/denoising_diffusion_pytorch.py:481-531), loss.backward(), torch.optim.Adam.step() and the EMA lerp
 * (models/trainer_diffusion.py:63-69,92,176-191).  The host computes q_sample / the regression target / the loss weights (three
 * elementwise lines on its own tensors) and owns the schedule (cosine LR, EMA warm-up); gradients are exposed as ONE flat fp32
 * buffer so that data-parallel ranks all-reduce it with a single NCCL call before the Adam step.
 *   create   : batch / crop / device as ndiff_engine_create (dim = 64).  ndiff_trainer_engine() is the embedded engine: load the
 *              state_dict into it with ndiff_load_param[_async], then ndiff_trainer_finalize (flattens the parameters, packs
 *              the weights, builds forward + backward plans), and call ndiff_set_condition on it before every step.
 *   forward_backward : x_t, target fp32 NCHW [B,4,H,W]; time int64 [B]; loss_weight fp32 [B] (loss_weight[t_b]); all device
 *              pointers.  Gradients of all 416 parameters land in the flat gradient buffer; *loss_host (optional) receives
 *              mean_b( w_b * mean_chw (v - target)^2 ) and synchronises the stream.
 *   flat / slot : the flat buffers (0 parameters, 1 gradients, 2 Adam m, 3 Adam v, 4 EMA) and a parameter's (offset, numel).
 *   adam_step : torch.optim.Adam semantics over the live parameters, gradients pre-multiplied by grad_scale (1 / world size
 *              after an all-reduce SUM); re-packs every derived weight form.  The condition must be set again afterwards.
 *   ema_update : ema += weight * (param - ema)  (weight = 1 - decay; weight >= 1 copies).                                      */
typedef struct ndiff_trainer ndiff_trainer;
NDIFF_API int32_t ndiff_trainer_create(const ndiff_config* cfg, ndiff_trainer** out);
NDIFF_API void    ndiff_trainer_destroy(ndiff_trainer* t);
NDIFF_API ndiff_engine* ndiff_trainer_engine(ndiff_trainer* t);
NDIFF_API int32_t ndiff_trainer_finalize(ndiff_trainer* t, void* stream);
NDIFF_API int32_t ndiff_trainer_forward_backward(ndiff_trainer* t, const float* x_t_dev, const int64_t* time_dev,
                                                 const float* target_dev, const float* loss_weight_dev, double* loss_host,
                                                 void* stream);
NDIFF_API int32_t ndiff_trainer_flat(ndiff_trainer* t, int32_t which, float** ptr, int64_t* n_live, int64_t* n_total);
NDIFF_API int32_t ndiff_trainer_slot(ndiff_trainer* t, const char* name, int64_t* offset, int64_t* numel);
NDIFF_API int32_t ndiff_trainer_adam_step(ndiff_trainer* t, float lr, float beta1, float beta2, float eps, float weight_decay,
                                          float grad_scale, void* stream);
NDIFF_API int32_t ndiff_trainer_ema_update(ndiff_trainer* t, float weight, void* stream);
/* per-kernel-family durations of the last step's forward / backward launch lists ("fwd|bwd;family;launches;ms" lines) */
NDIFF_API int32_t ndiff_trainer_time(ndiff_trainer* t, char* out, int32_t cap, void* stream);
NDIFF_API int64_t ndiff_trainer_activation_bytes(const ndiff_trainer* t);
NDIFF_API int64_t ndiff_trainer_launches(const ndiff_trainer* t, int32_t backward);
/* Backward kernels one at a time (parity tests against torch autograd).  wgrad modes: 0 = 1x1, 1 = 3x3 pad 1, 2 = 2x2 stride 2
 * (space-to-depth); dw is fp32 [Cout][C0+C1][taps], ACCUMULATED (zero it first). */
NDIFF_API int32_t ndiff_op_wgrad(int32_t mode, int32_t B, int32_t H, int32_t W, const void* dy, int32_t Cout, const void* src0,
                                 int32_t C0, const void* src1, int32_t C1, float* dw, void* stream);
NDIFF_API int32_t ndiff_op_gn_backward(const void* h, const void* dout, void* dh, const void* stats, const float* gamma,
                                       const float* beta, const float* ss, int32_t ss_ld, int32_t ss_off, const void* maps,
                                       void* dmaps, float* dgamma, float* dbeta, float* dss, int32_t B, int32_t HW, int32_t C,
                                       int32_t G, void* stream);
NDIFF_API int32_t ndiff_op_layernorm_backward(const void* x, const float* vec, int32_t vec_ld, const float* g, const void* du,
                                              void* dy, float* dg, float* dbeta, int32_t B, int32_t HW, int32_t C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NOISEDIFF_B200_H_ */
