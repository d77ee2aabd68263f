// Handle layout shared by ard_api.cu (forward schedule, C ABI) and ard_train.cu (saved activations + backward schedule).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "ard_internal.h"

namespace ard {

// ------------------------------------------------------------------------------------------------ device buffers
// bumped whenever a device buffer is (re)allocated: captured CUDA graphs hold raw pointers and are stale afterwards
inline unsigned long long& alloc_epoch() {
    static unsigned long long e = 0;
    return e;
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int ensure(size_t n) {
        if (n <= bytes) return 0;
        ++alloc_epoch();
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) { p = nullptr; return set_error(ARD_ERR_CUDA, "cudaMalloc(%zu): %s", n, cudaGetErrorString(e)); }
        bytes = n;
        return 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// padded per-head q/k/v weights of the window-resident attention block kernel (attn_block.cu), 96-channel stage
struct AttnBlockW {
    DevBuf wqk, bq, wv, bv, table, wp_plain, wp_fold, bp_plain, bp_fold;   // [256,96] bf16, [128] f32, [128,96] bf16, [96] f32 (un-padded v bias),
                                                                           // [4,15,24] f32, 2 x [96,128] bf16, 2 x [96] f32 (bias + Wp' bv)
    bool ready = false;
};

struct BlockW {
    DevBuf ln1_g, ln1_b, ln2_g, ln2_b;
    DevBuf qkv_w, qkv_b, proj_w, proj_b, proj_w_f32, fc1_w, fc1_b, fc1_b_half, fc2_w, fc2_b, rpb;
    // ResiDual (src/residual.py:14-42) injected after this block's attention
    bool has_res = false, lambda_set = false;
    int K = 0;
    std::vector<float> h_mean, h_basis;
    DevBuf res_basis, res_dmean, res_M, proj_w_fold, proj_w_fold_f32, proj_b_fold, lam_ones;
    AttnBlockW ab;
    // training (ard_train.cu): current lambda (padded to Kp), transposed / folded weights for the dgrad GEMMs, built lazily
    int Kp = 0;
    bool bwd_ready = false;
    DevBuf lam, fc1_wT, fc2_wT, qkv_wT, proj_wT, res_wc, res_wcT, res_basis_bf16, res_c0;
    // activations saved by a save_for_backward forward (views into ard_handle::tape)
    float *t_s = nullptr, *t_x1 = nullptr, *t_x3 = nullptr, *t_out = nullptr;
    __nv_bfloat16 *t_qkv = nullptr, *t_ao = nullptr;
};
struct LayerW {
    std::vector<BlockW> blocks;
    DevBuf mg_g, mg_b, mg_w, mg_wT;
};

}  // namespace ard

struct ard_handle {
    using DevBuf = ard::DevBuf;
    using LayerW = ard::LayerW;
    ard_config cfg;
    int nlayers = 4;
    int num_sms = 148;
    bool finalized = false;
    std::map<std::string, std::vector<float>> host;   // raw state_dict tensors (fp32)
    // front end
    DevBuf window, twiddle, melw, mstart, mlen, bn_scale, bn_shift;
    int band_max = 0;
    DevBuf f_window, f_melw, f_mstart, f_mlen;   // fusion featuriser (get_mel, data.py:363-399): htk filters, periodic hann
    int f_band_max = 0;
    DevBuf pe_w, pe_b, pe_g, pe_beta;
    std::vector<LayerW> layers;
    DevBuf norm_g, norm_b, tscam_w, tscam_b, p0_w, p0_b, p2_w, p2_b;
    // workspace
    DevBuf ws_logmel, ws_x, ws_y, ws_xn, ws_ao, ws_qkv, ws_h, ws_normed, ws_emb, ws_hid, ws_proj, ws_tscam_a, ws_tscam_y, ws_wave;
    int last_launches = 0;
    bool use_fused_ffn = true;   // ARD_FUSED_FFN=0 disables the fused 96-channel FFN kernel (A/B measurements)
    bool use_ln_qkv = true;      // ARD_LN_QKV=0: LayerNorm kernel + qkv GEMM instead of ln_qkv_96 (A/B measurements)
    bool use_attn_block = true;  // ARD_ATTN_BLOCK=0: ln_qkv + window_attention + proj GEMM instead of attn_block_96 (A/B measurements)
    int use_dual_gemm = 1;       // ARD_DUAL_GEMM: backward with gemm_dual (1, default) or the separate GEMMs + lambda_grad kernel (0; A/B measurements)
    int wide_skip = 0;           // ARD_FFN_WIDE_SKIP=<C>: that width takes the unfused FFN chain (development)
    int use_fused_ffn_wide = 1;    // ARD_FUSED_FFN_WIDE: 0 never, 1 C = 192 (default), 2 also C = 384, 3 also the HTSAT-base widths 128 / 256 (opt-in, see ard_api.cu)
    // training state
    DevBuf tape, p0_wT, p2_wT, t_emb, t_hid, t_proj;
    DevBuf bw_g, bw_gs, bw_t, bw_hpre, bw_dh, bw_gqkv, bw_gb, bw_coef, bw_gcoef, bw_gsc, bw_small;
    int tape_B = 0;          // batch of the forward whose activations the tape holds (0: none)
    long long tape_gen = 0;  // serial number of that forward (ard_tape_generation): a backward built on an older one is refused
    // CUDA-graph replay of the inference forward (ard_api.cu): the ~110 launches of a forward cost ~1.1 ms of fixed launch /
    // ramp latency whatever the batch; replaying a captured graph removes the host side of that and most of the device side.
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;   // null: this (input, batch) was seen once and ran eagerly; captured on its next use
        unsigned long long epoch = 0, aepoch = 0;
        int launches = 0;
        unsigned long long last_use = 0;
    };
    struct GraphKey {
        const void* src;
        int B, quantize, want_ae;
        bool operator<(const GraphKey& o) const {
            if (src != o.src) return src < o.src;
            if (B != o.B) return B < o.B;
            if (quantize != o.quantize) return quantize < o.quantize;
            return want_ae < o.want_ae;
        }
    };
    std::map<GraphKey, GraphEntry> graphs;
    unsigned long long graph_epoch = 0;   // bumped when weights / injected ResiDual modules change
    unsigned long long graph_clock = 0;
    int use_graphs = 1;                   // ARD_GRAPHS=0: always launch kernel by kernel
    DevBuf g_emb, g_ae;                   // outputs of the captured forward, copied to the caller's tensors after each replay
    cudaStream_t cap_stream = nullptr;    // capture happens on a private stream (the caller's may be the legacy default stream)
    ~ard_handle() {
        for (auto& kv : graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (cap_stream) cudaStreamDestroy(cap_stream);
    }
};


namespace ard {
inline int C_of(const ard_handle* h, int l) { return h->cfg.embed_dim << l; }
inline int R_of(int l) { return 64 >> l; }   // tokens per side
int upload(DevBuf& b, const void* src, size_t bytes);
int upload_f32(DevBuf& b, const std::vector<float>& v);
int upload_f16(DevBuf& b, const std::vector<float>& v);
int upload_bf16(DevBuf& b, const std::vector<float>& v);
int get(const ard_handle* h, const std::string& key, size_t numel, const std::vector<float>** out);
int ensure_workspace(ard_handle* h, int B);
int ensure_fold(ard_handle* h, int l, int b, cudaStream_t s);
// training (ard_train.cu)
int ensure_tape(ard_handle* h, int B);
int run_block_train(ard_handle* h, int l, int b, int B, float* attn_out, float attn_scale, int attn_acc, float* res_out,
                    long long res_bstride, cudaStream_t s);
int encoder_backward(ard_handle* h, const ard_backward_args* a, cudaStream_t s);
// window-resident attention block (attn_block.cu)
int attn_block_pack(AttnBlockW& w, const std::vector<float>& qkv_w, const std::vector<float>& qkv_b, const std::vector<float>& rpb, int C, int nH);
int attn_block_pad_proj(const __nv_bfloat16* w, const float* bp, const float* bv, __nv_bfloat16* out, float* bp_eff, int C, int nH, cudaStream_t s);
int attn_block_96(const float* x, float* out, const AttnBlockW& w, const __nv_bfloat16* wp_pad, const float* bp, const float* gamma,
                  const float* beta, int B, int R, int shift, int num_sms, cudaStream_t stream);
// fp32-grade mode (fp32_mode.cu)
int encoder_stages_fp32(ard_handle* h, const ard_forward_args* a, float* X, float* Y, float** x_final, cudaStream_t s);
int tscam_gemm_fp32(ard_handle* h, const float* normed, float* y, int ldy, int B, cudaStream_t s);
void fp32_release(const ard_handle* h);
}  // namespace ard
