// Internal op launchers (host side). All tensors are device pointers; NHWC / row-major.
#pragma once
#include "common.cuh"

namespace etai {

// ---------------- GEMM-shaped ops -------------------------------------------------------------
// C[M,N] = A * W^T (+ epilogue).  A is either a dense [M,K] matrix (lda) or the implicit im2col view of an
// NHWC image for a 3x3/pad-1 convolution (K = 9*Cin, k = (ky*3+kx)*Cin + c).
struct GemmArgs {
    const void* A = nullptr;   // dense: [M,lda]; conv: NHWC image [B,H,W,Cin]
    const void* W = nullptr;   // [N,K] row-major (K contiguous)
    void* C = nullptr;         // [M,ldc]  (geglu: [M, N/2] with ldc)
    const void* bias = nullptr;      // [N] or null
    const void* rowbias = nullptr;   // [M/rows_per_group, ldrb] fp32, added per (group,n); null = none
    const void* residual = nullptr;  // [M,ldr] or null
    long M = 0;
    int N = 0, K = 0;
    long lda = 0, ldc = 0, ldr = 0, ldrb = 0;
    long rows_per_group = 1;
    int geglu = 0;
    // conv geometry (conv != 0)
    int conv = 0, B = 0, H = 0, Wd = 0, Cin = 0, stride = 1, Ho = 0, Wo = 0;
    int pad = 1;  // leading zero padding (1 = symmetric pad-1 conv; 0 with stride 2 = the VAE's (0,1,0,1) asymmetric pad)
    int dtype = ETAI_F32;
};

void gemm_simt(const GemmArgs& a, cudaStream_t s);
// tcgen05/TMA path (f16/bf16 only).  `ws` is scratch for stride-2 im2col (may be null when not needed).
void gemm_tc(const GemmArgs& a, void* ws, size_t ws_bytes, cudaStream_t s);
bool gemm_tc_supported(const GemmArgs& a);


// out[M,N] = act(x[M,K] (fp32) * W[N,K]^T (T) + bias) for tiny M (<=8); fp32 output. act: 0 none, 1 SiLU
void skinny_linear(const float* x, const void* W, const void* bias, float* out, int M, int N, int K, int act,
                   int wdtype, cudaStream_t s);

// ---------------- normalisation ---------------------------------------------------------------
size_t groupnorm_workspace_bytes(int B, long HW, int C, int groups);
size_t groupnorm_ticket_offset(int B, int groups);  // byte offset of the B int tickets (must be zero before first use)
int groupnorm_launches(long HW, int C, int groups, int dtype);  // 1 (single cluster launch) or 2 (stats + apply)
void groupnorm(const void* x, void* y, const void* gamma, const void* beta, int B, long HW, int C, int groups,
               float eps, bool silu, int dtype, void* ws, cudaStream_t s);
void layernorm(const void* x, void* y, const void* gamma, const void* beta, long M, int C, float eps, int dtype,
               cudaStream_t s);

// ---------------- attention -------------------------------------------------------------------
struct RowMap {  // passed by value to kernels
    int q[ETAI_MAX_ROWS], k[ETAI_MAX_ROWS], v[ETAI_MAX_ROWS];
};
struct SelfAttnArgs {
    const void *q, *k, *v;  // [B,N,ld*]; head h at column h*d
    void* out;              // [B,Nq,ldo]
    int B, Nq, Nk, heads, d;
    long ldq, ldk, ldv, ldo;
    float scale;
    RowMap map;
    int dtype;
};
void attention_simt(const SelfAttnArgs& a, cudaStream_t s);
void attention_tc(const SelfAttnArgs& a, cudaStream_t s);
bool attention_tc_supported(const SelfAttnArgs& a);

// cross attention over the 77-token text context with the prompt-to-prompt edit + store fused in
struct CrossGroup {  // one CTA-column of work: a plain row (tgt<0) or a (base,tgt) pair
    int base, tgt, pair;      // UNet batch rows; pair = index into the edit tables
    int store_base, store_tgt;  // slot in the store accumulators or -1
};
struct CrossAttnArgs {
    const void* q;   // [B,N,ldq]
    const void* kv;  // [B,L,ldkv]: K at column koff + h*d, V at column voff + h*d
    void* out;       // [B,N,ldo]
    int B, N, L, heads, d;
    long ldq, ldkv, ldo;
    int koff, voff;
    float scale;
    int n_groups;
    CrossGroup groups[ETAI_MAX_ROWS];
    const float *mapper, *blend_a, *equalizer, *alpha_step;  // per pair tables or null (no edit)
    float* store;  // [slots][N][L] fp32 accumulators or null
    // tcgen05 path only: 16-bit transposed/padded mapper prepared by cross_attention_tc_prep_mapper, and a
    // [heads][slots][N][L] fp32 workspace for per-head store partials
    const void* map16;
    float* store_part;
    int dtype;
};
void cross_attention(const CrossAttnArgs& a, cudaStream_t s);       // SIMT, fp32 math (parity path)
int cross_attention_tc(const CrossAttnArgs& a, cudaStream_t s);     // tcgen05 (16-bit engine); returns #launches
size_t cross_attention_tc_mapper_bytes(int pairs);
void cross_attention_tc_prep_mapper(const float* mapper, void* map16, int pairs, int L, int dtype, cudaStream_t s);
bool cross_attention_tc_supported(const CrossAttnArgs& a);

// ---------------- data movement / elementwise -------------------------------------------------
void nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int B, int C, int Cpad, long HW, cudaStream_t s);
void nhwc_to_nchw(const void* in, int in_dtype, void* out, int out_dtype, int B, int C, int Cpad, long HW, cudaStream_t s);
void convert(const void* in, int in_dtype, void* out, int out_dtype, long n, cudaStream_t s);
void concat_channels(const void* a, int Ca, const void* b, int Cb, void* out, long rows, int dtype, cudaStream_t s);
void upsample2x(const void* in, void* out, int B, int H, int W, int C, int dtype, cudaStream_t s);
void im2col3x3(const void* in, void* out, int B, int H, int W, int C, int stride, int pad, int Ho, int Wo, int dtype,
               cudaStream_t s);
void copy_rows(void* base, long row_elems, int src_row, int n_src, int dst_row, int n_dst, int dtype, cudaStream_t s);
void timestep_sincos(const float* t_dev, float* out, int dim, cudaStream_t s);
void add_f32(float* y, const float* x, long n, cudaStream_t s);
void add_inplace(void* y, const void* x, long n, int dtype, cudaStream_t s);
// weight repack: OIHW -> O,(ky,kx),I ; optional GEGLU row interleave
void pack_conv_weight(const float* oihw, void* out, int O, int I, int Opad, int Ipad, int dtype, cudaStream_t s);
void pack_geglu_weight(const float* w, const float* b, void* wout, void* bout, int N2, int K, int dtype, cudaStream_t s);

// ---------------- VAE / CLIP text tower only (textvae_ops.cu) -----------------------------------
// in place, softmax(scale*x) per row; lse (optional): fp32 [rows] log-sum-exp of the scaled row
void softmax_rows(void* x, long rows, int n, float scale, int dtype, cudaStream_t s, float* lse = nullptr);
void clip_embed(const int* ids, const void* tok, const void* pos, void* out, long rows, int L, int C, int vocab, int dtype,
                cudaStream_t s);
void quick_gelu(void* x, long n, int dtype, cudaStream_t s);                            // in place, x*sigmoid(1.702x)
void clip_attention(const void* qkv, void* out, int B, int L, int heads, int d, float scale, int dtype, cudaStream_t s);

// ---------------- backward (dgrad only; null-text inversion) -- backward.cu -----------------------
void transpose_2d(const void* in, void* out, int R, int Cc, int dtype, cudaStream_t s);          // [R,Cc] -> [Cc,R]
void conv_weight_flip(const void* in, void* out, int O, int I, int dtype, cudaStream_t s);       // [O,3,3,I] -> [I,3,3,O], taps mirrored
void zero_stuff2x(const void* in, void* out, int B, int Ho, int Wo, int C, int dtype, cudaStream_t s);
void upsample2x_bwd(const void* dout, void* din, int B, int H, int W, int C, int dtype, cudaStream_t s);
void slice_cols(const void* in, long ld, int off, int C, void* out, bool accumulate, long rows, int dtype, cudaStream_t s);
void scale_by_device_scalar(const void* in, void* out, const float* s_dev, bool invert, long n, int dtype, cudaStream_t s);
void absmax_scale(const float* x, long n, float target, float* s_dev, cudaStream_t s);
void geglu_fwd(const void* u, void* y, long M, int F, int dtype, cudaStream_t s);
void geglu_bwd(const void* u, const void* dy, void* du, long M, int F, int dtype, cudaStream_t s);
size_t groupnorm_bwd_workspace_bytes(int B, int groups);
void groupnorm_bwd(const void* x, const void* dy, const void* gamma, const void* beta, void* dx, int B, long HW, int C,
                   int groups, float eps, bool silu, int dtype, void* ws, cudaStream_t s);
void layernorm_bwd(const void* x, const void* dy, const void* gamma, void* dx, long M, int C, float eps, int dtype,
                   cudaStream_t s);
struct SelfAttnBwdArgs {
    const void *q, *k, *v, *o, *dout;
    void *dq, *dk, *dv;
    float *lse, *dsum;  // [B,heads,N] scratch
    int B, N, heads, d;
    long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
    float scale;
    int dtype;
};
void attention_bwd(const SelfAttnBwdArgs& a, cudaStream_t s);
struct CrossAttnBwdArgs {
    const void *q, *kv, *dout;
    void* dq;      // [B,N,lddq] or null
    float* part;   // scratch, cross_attention_bwd_partial_bytes()
    float* dkv;    // [B*L, ld_dkv] fp32: columns [kv_off, kv_off+C) = dK, [kv_off+C, kv_off+2C) = dV (overwritten)
    int B, N, L, heads, d;
    long ldq, ldkv, lddo, lddq, ld_dkv;
    int koff, voff, kv_off;
    float scale;
    int dtype;
};
// ---- pieces of the GEMM-based self-attention backward (16-bit engines, N >= 256): every N x N product runs on gemm_tc_k
// head slices [N, d] (row stride ld) -> zero-padded [heads][N][DP] and its transpose [heads][DP][N] (dstT may be null)
void attn_pack_heads(const void* src, long ld, int N, int heads, int d, int DP, void* dst, void* dstT, int dtype, cudaStream_t s);
void attn_unpack_heads(const void* src, int N, int heads, int d, int DP, void* dst, long ld, int dtype, cudaStream_t s);
void attn_rowdot(const void* a, long lda, const void* b, long ldb, int N, int heads, int d, float* out, int dtype, cudaStream_t s);
// mode 0: x = p * (x - dsum[row]) * scale          (dS   from dP,   p = P   [row = query])
// mode 1: x = exp(x * scale - lse[col])             (P^T  from S^T,  columns = queries)
// mode 2: x = p * (x - dsum[col]) * scale          (dS^T from dP^T, p = P^T)
void attn_bwd_elementwise(void* x, const void* p, const float* vec, int rows, int cols, float scale, int mode, int dtype,
                          cudaStream_t s);
size_t cross_attention_bwd_partial_bytes(int B, int N, int L, int C, int d);
void cross_attention_bwd(const CrossAttnBwdArgs& a, cudaStream_t s);

// ---------------- scheduler -------------------------------------------------------------------
void cfg_ddim_step(const float* eps, int n, int has_cfg, float guidance, const float* x, float* x_out,
                   float* eps_cfg_out, float a_from, float a_to, float eta, float variance, const float* eta_map,
                   const float* noise_cand, const float* losses, int K, const float* pin_src, long E, cudaStream_t s);
void eta_noise_losses(const float* eps, int n, int has_cfg, float guidance, const float* x, const float* x_prev_inv,
                      float a_from, float a_to, float eta, float variance, const float* noise_cand, int K, long E,
                      float* losses, int* best_idx, cudaStream_t s);
// out = eps_u + guidance * prox(eps_c - eps_u): soft threshold at a global quantile of |eps_c - eps_u| (rank_lo >= 0: torch
// 'linear' interpolation between order statistics rank_lo / rank_hi with `weight`) or at fixed_thr (rank_lo < 0)
void prox_guidance(const float* eps_u, const float* eps_c, float* out, long n, long rank_lo, long rank_hi, float weight,
                   float fixed_thr, int l1, float guidance, float* thr_out, cudaStream_t s);

}  // namespace etai
