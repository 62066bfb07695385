// noisediff_b200 — engine internals shared by engine.cu (sampling path + C ABI) and train_engine.cu (training row).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/noisediff_b200.h"
#include "conv_gemm.cuh"
#include "pointwise.cuh"
#include "pixel_chain.cuh"

namespace ndiff {

struct Param {
    std::vector<int64_t> shape;
    float* dev = nullptr;
    size_t n = 0;
};

struct Act {
    bf16* p = nullptr;
    int C = 0, H = 0, W = 0;
};

struct Op {
    std::string name;
    std::function<int(cudaStream_t)> fn;
    double flops = 0.0;
    double bytes = 0.0;   // algorithmic HBM bytes: every input read once + every output written once (bf16 activations)
    int launches = 1;
    bool chain = false;   // fused per-pixel chain (its FLOPs are not convolution FLOPs)
};

struct RbSpec { std::string name; int cin, cout, groups; bool pos; int c0, c1; };   // cin = c0 + c1 (x, then the concatenated skip)
struct AttnSpec { std::string name; int C; };
std::vector<RbSpec> resblocks(int dim);      // every ResnetBlock / ResnetBlock2 of NoiseDiffNet, in state_dict naming
std::vector<AttnSpec> attnblocks(int dim);   // every AttnBlock

inline cudaStream_t as_stream(void* p) { return static_cast<cudaStream_t>(p); }

// Entry points run on the engine's device and put the caller's current device back on return: a model on cuda:1 must not
// silently switch the process (PyTorch's notion of the current device included) away from cuda:0.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace ndiff

using ndiff::Param; using ndiff::Act; using ndiff::Op; using ndiff::bf16; using ndiff::ChainState; using ndiff::StepParams;

struct ndiff_engine {

    ndiff_config cfg{};
    int num_sms = 148;
    int B = 0, H = 0, W = 0;
    int dim = 0;        // PHYSICAL base width the kernels run at (64); every channel count below is a multiple of it
    int dim_real = 0;   // the model's `dim` (args.dim): 64, or smaller (the shipped checkpoint: 48) embedded with zero padding
    float real_frac = 1.0f;   // dim_real / dim: live fraction of every GroupNorm group / LayerNorm row
    bool finalized = false, cond_set = false, plan_built = false;
    bool skip_plan = false;   // training engines pack weights and set conditions but never build the sampling plan
    std::map<std::string, Param> params;          // as loaded: the reference's state_dict, logical shapes
    std::map<std::string, Param> phys;            // dim_real != dim: zero-padded, channel-permuted copies the kernels read
    int* emb_maps = nullptr; size_t emb_maps_n = 0;   // device index tables of embed_params()
    std::vector<void*> owned;                       // every cudaMalloc'd pointer
    std::map<size_t, std::vector<void*>> pool_free;  // size -> free buffers
    std::map<void*, size_t> pool_size;
    bool keep_all = false;

    std::map<std::string, bf16*> packed;
    std::map<std::string, bf16*> chain_w;      // fused per-pixel chains: weight blob / parameter block per chain
    std::map<std::string, float*> chain_f;
    float* init_w = nullptr;
    float* fold_tmp = nullptr;       // scratch for folding LayerNorm affines into chain weights
    float* shot_fold = nullptr;      // fp32 shot_mlp2.fc1 x shot_attn.proj_out (64 x 64): per-sample vector of the shot chain's folded stage
    bf16* init_w_tc = nullptr; bf16* xpad = nullptr;
    // time path
    float* ss_w = nullptr; float* ss_b = nullptr; int ss_total = 0;
    std::map<std::string, int> ss_off;
    float* st_buf = nullptr;       // [max(B, n_steps)][4 dim]
    float* ss_cur = nullptr;       // [B][ss_total]
    float* ss_table = nullptr; int ss_table_rows = 0;
    int* t_buf = nullptr; int t_buf_n = 0;
    // iso path
    float* cvec = nullptr; int cv_total = 0;
    float* cvec2 = nullptr;        // C >= 128 AttnBlocks: Wp (b2 + c) + bp, the per-sample vector of the folded ff.net.2 + proj_out GEMM
    std::map<std::string, int> cv_off;
    // condition / state
    float* clean = nullptr;        // fp32 NHWC4
    bf16* map1 = nullptr; bf16* map2 = nullptr;
    float* pos_emb = nullptr;
    float* position = nullptr;     // fp32 NCHW (B,2,H,W) copy of the last condition's position maps (training backward re-derives pos_emb)
    long long* iso_idx = nullptr;  // [B] copy of the last condition's iso_ratio_idx (training backward of iso_embed)
    float* x = nullptr;            // fp32 NHWC4 chain state / network input
    float* v_out = nullptr;        // fp32 NHWC4 network output
    unsigned long long* stats = nullptr; int n_stats = 0; size_t stats_bytes = 0;
    ChainState* chain = nullptr;
    StepParams* step_table = nullptr; int n_steps = 0; int steps_done = 0;
    // plan
    std::vector<Op> net_ops;
    std::map<std::string, Act> named;
    Act xf{}, sf{};
    // final_res_block.block2.norm folded into the heads kernel (FinalArgs::gn_*): xf is then the raw block2 conv output
    const unsigned long long* xf_stats = nullptr; Act xf_res{}; int xf_groups = 0; std::string xf_norm;
    float* sn = nullptr;           // fp32 [npix][4] shot-noise image written by the fused shot-branch tail (pixel_chain.cuh), or unused
    bool tail_fused = false;
    cudaGraphExec_t step_exec = nullptr, fwd_exec = nullptr;
    cudaStream_t cap_stream = nullptr;
    double conv_flops = 0.0;

    ~ndiff_engine() {
        if (step_exec) cudaGraphExecDestroy(step_exec);
        if (fwd_exec) cudaGraphExecDestroy(fwd_exec);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        for (void* p : owned) cudaFree(p);
    }

    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        NDIFF_CUDA_OK(cudaMalloc(&p, count * sizeof(T) ? count * sizeof(T) : 16));
        owned.push_back(p);
        *out = static_cast<T*>(p);
        return 0;
    }
    // frees a buffer obtained from alloc() (regrown tables / resized parameters must not pile up until engine destroy)
    void release(void* p) {
        if (!p) return;
        auto it = std::find(owned.begin(), owned.end(), p);
        if (it != owned.end()) { owned.erase(it); cudaFree(p); }
    }
    bf16* pool_get(size_t elems) {
        const size_t bytes = elems * sizeof(bf16);
        auto& fl = pool_free[bytes];
        if (!fl.empty() && !keep_all) {
            void* p = fl.back();
            fl.pop_back();
            return static_cast<bf16*>(p);
        }
        bf16* p = nullptr;
        if (alloc(&p, elems)) return nullptr;
        pool_size[p] = bytes;
        return p;
    }
    void pool_put(const void* p) {
        auto it = pool_size.find(const_cast<void*>(p));
        if (it != pool_size.end()) pool_free[it->second].push_back(it->first);
    }
    const Param* param(const std::string& name) const {
        auto ip = phys.find(name);
        if (ip != phys.end()) return &ip->second;
        auto it = params.find(name);
        return it == params.end() ? nullptr : &it->second;
    }
    const float* pf(const std::string& name) const { return param(name)->dev; }
};

namespace ndiff {
// engine.cu: (re)pack every derived weight form from the fp32 parameters (and build the sampling plan on the first call)
int engine_finalize(ndiff_engine* e, cudaStream_t s);
}  // namespace ndiff
