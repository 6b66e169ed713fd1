/* etai.h -- C ABI of the B200-native eta-inversion engine (libetai.so).
 *
 * The reference (furiosa-ai/eta-inversion) is 100% Python and has NO native boundary; its
 * "operator API" for the hot path is three Python call shapes:
 *
 *   unet(sample, t, encoder_hidden_states=ctx)["sample"]      modules/inversion/diffusion_inversion.py:264-280
 *                                                             modules/inversion/eta_inversion.py:321
 *   scheduler_{fwd,bwd}.step(eps, t, x, eta=, variance_noise=) modules/inversion/diffusion_inversion.py:288-312
 *                                                             modules/inverse_schedulers/scheduling_ddim_inverse.py:115-143
 *   controller(attn, is_cross, place) / editor(q,k,v,sim,attn,...)  modules/utils/ptp_utils.py:250,
 *                                                             modules/utils/masactrl_utils.py:122-124
 *
 * Each entry point below replaces one of those call shapes (cited per function).  Conventions:
 *   - plain C, no torch types; every tensor is a raw device pointer + sizes; caller owns all I/O buffers
 *   - return 0 on success, negative ETAI_ERR_* otherwise; etai_last_error() has the message (thread local)
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*); no host sync inside
 *   - latents / eps at the boundary are NCHW like the reference; the engine is NHWC inside
 *   - one handle per device; a handle is not thread-safe; different handles are independent
 */
#ifndef ETAI_H_
#define ETAI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETAI_ABI_VERSION 1

#if defined(__GNUC__)
#define ETAI_EXPORT __attribute__((visibility("default")))
#else
#define ETAI_EXPORT
#endif

/* dtype enum (storage type of a buffer; math is always fp32-accumulated) */
enum { ETAI_F32 = 0, ETAI_F16 = 1, ETAI_BF16 = 2 };

/* error codes */
enum {
    ETAI_OK = 0,
    ETAI_ERR_ARG = -1,         /* bad argument (shape / dtype / null) */
    ETAI_ERR_CUDA = -2,        /* CUDA runtime / driver error */
    ETAI_ERR_STATE = -3,       /* call order (e.g. forward before set_context) */
    ETAI_ERR_UNSUPPORTED = -4, /* configuration not built */
    ETAI_ERR_NOMEM = -5
};

/* math back end */
enum {
    ETAI_MATH_AUTO = 0,   /* f32 -> SIMT fp32 kernels (parity mode); f16/bf16 -> tcgen05 kernels */
    ETAI_MATH_SIMT = 1    /* force the SIMT kernels (fp32 accumulate) for any storage dtype */
};

typedef struct etai_unet etai_unet; /* opaque */

/* One named weight tensor, diffusers key names and layouts (OIHW convs, [out,in] linears),
 * i.e. exactly `unet.state_dict()` of the model the reference loads (modules/models/__init__.py:135). */
typedef struct {
    const char* name;
    const void* data;   /* host or device pointer, fp32/f16/bf16 */
    int32_t dtype;
    int32_t ndim;
    int64_t shape[4];
    int32_t on_device;  /* 0: host pointer, 1: device pointer */
} etai_tensor;

typedef struct {
    int32_t dtype;                 /* engine storage dtype for weights + activations */
    int32_t math_mode;             /* ETAI_MATH_* */
    int32_t block_out_channels[4]; /* SD-1.x: 320,640,1280,1280 */
    int32_t heads;                 /* 8 (diffusers attention_head_dim=8 means 8 heads) */
    int32_t cross_dim;             /* 768 */
    int32_t ctx_len;               /* 77 */
    int32_t latent_hw;             /* 64 (512x512 images) */
    int32_t max_batch;             /* max rows per forward (2 inversion, 4 edit, 4k co-batched) */
} etai_unet_cfg;

/* Attention control for one UNet forward.  Replaces the monkey-patched Attention.forward of
 *   modules/utils/ptp_utils.py:196-302 (+ modules/utils/ptp.py:107-119,150-171,194-274),
 *   modules/utils/masactrl_utils.py:74-153 (+ modules/utils/masactrl.py:41-72),
 *   modules/utils/pnp_utils.py:67-133
 * by data: nothing is materialised, the edit happens inside the attention kernels.
 * All pointer members tagged [dev] are device pointers, [host] are host arrays read during the call. */
#define ETAI_CTRL_SELF_REMAP 1
#define ETAI_CTRL_CROSS_EDIT 2
#define ETAI_CTRL_CROSS_STORE 4
#define ETAI_MAX_ROWS 64
#define ETAI_MAX_PAIRS 16

typedef struct {
    int32_t flags;

    /* SELF_REMAP: out[r] = softmax(Q[q_row[r]] K[k_row[r]]^T * scale) V[v_row[r]] for self-attention layers whose
     * transformer index i (0..15, call order) has bit i set in self_layer_mask and whose token count <= self_max_tokens.
     *   PtP self-replace (ptp.py:194-200,212-216): target cond row -> (q,k) of source cond row, own v
     *   MasaCtrl (masactrl.py:56-72):              every row -> (k,v) of the source row of its CFG half
     *   PnP (pnp_utils.py:76-88):                  rows 1,2 -> (q,k) of row 0, own v */
    const int32_t* self_q_row; /* [host][B] */
    const int32_t* self_k_row; /* [host][B] */
    const int32_t* self_v_row; /* [host][B] */
    uint32_t self_layer_mask;
    int32_t self_max_tokens;

    /* CROSS_EDIT (ptp.py:205-211,234-274): for each pair p,
     *   F[n]   = eq[p][n] * ( a[p][n] * sum_w P_base[w] * mapper[p][w][n] + (1 - a[p][n]) * P_tgt[n] )
     *   P_tgt' = alpha[p][n] * F[n] + (1 - alpha[p][n]) * P_tgt[n]          (no renormalisation) */
    int32_t n_pairs;
    const int32_t* edit_base_row; /* [host][n_pairs] */
    const int32_t* edit_tgt_row;  /* [host][n_pairs] */
    const float* mapper;          /* [dev][n_pairs][77][77] */
    const float* blend_a;         /* [dev][n_pairs][77] */
    const float* equalizer;       /* [dev][n_pairs][77] */
    const float* alpha_step;      /* [dev][n_pairs][77]  = cross_replace_alpha[cur_step] */

    /* CROSS_STORE (ptp.py:150-171; consumers ptp.py:18-47, ptp_editor.py:43-85): for cross-attention layers with
     * store_res^2 query tokens, acc[place][i][pix][w] += sum_heads P'[store_row[i]][head][pix][w]  (post-edit P) */
    int32_t store_res;
    int32_t n_store_rows;
    const int32_t* store_row; /* [host][n_store_rows] */
    float* store_down;        /* [dev][n_store_rows][res*res][77] fp32, accumulated in place; may be NULL */
    float* store_mid;
    float* store_up;

    /* PnP resnet feature injection (pnp_utils.py:172-177): after conv2 of up_blocks[1].resnets[1],
     * rows [inject_n, 3*inject_n) := rows [0, inject_n) (tiled).  0 disables. */
    int32_t conv_inject_rows;
} etai_attn_ctrl;

ETAI_EXPORT int etai_abi_version(void);
ETAI_EXPORT const char* etai_last_error(void);

/* ---- UNet -------------------------------------------------------------------------------------
 * replaces `model.unet` of the reference pipeline object (diffusion_inversion.py:38). */
ETAI_EXPORT int etai_unet_create(etai_unet** out, const etai_unet_cfg* cfg, const etai_tensor* weights, int32_t n_weights,
                     int32_t device);
ETAI_EXPORT int etai_unet_destroy(etai_unet* h);

/* A second handle on the SAME packed device weights (read-only, reference counted: they are freed when the last handle
 * that uses them is destroyed) with its own activation arena, CUDA graphs, staging buffers and stream.  Two lock-step
 * groups in flight on one GPU (eval.py's per-GPU worker) then stream one copy of the 1.7 GB of
 * weights instead of two.  max_batch <= 0 keeps the source's. */
ETAI_EXPORT int etai_unet_clone(etai_unet** out, const etai_unet* src, int32_t max_batch);

/* Pre-projects the text context through all 16 cross-attention to_k/to_v (the context is constant over a
 * whole loop: diffusion_inversion.py:411-413,432-434).  ctx: [B,77,768] of io_dtype. */
ETAI_EXPORT int etai_unet_set_context(etai_unet* h, const void* ctx, int32_t io_dtype, int32_t B, void* stream);

/* eps = unet(latent, t, ctx)["sample"].  latent/eps_out: [B,4,hw,hw] NCHW of io_dtype.  `t` is one host
 * scalar shared by all rows (diffusers broadcasts it).  ctrl may be NULL. */
ETAI_EXPORT int etai_unet_forward(etai_unet* h, const void* latent, float t, int32_t io_dtype, int32_t B,
                      const etai_attn_ctrl* ctrl, void* eps_out, void* stream);

/* The same with one timestep per batch row -- diffusers' UNet accepts a [B] timestep tensor ("t scalar
 * or [B]"), although every loop of the reference passes a scalar.  t_rows: HOST array of B floats.  When all B values are
 * equal this IS etai_unet_forward (same schedule, same CUDA graphs).  Otherwise every row gets its own time embedding and the
 * 22 resnets add a per-image [B, Cout] time bias: the tcgen05 conv epilogue fuses one shared vector only, so those 22 convs
 * run on the fp32-accumulating SIMT conv and a mixed-timestep forward is slower than a uniform one. */
ETAI_EXPORT int etai_unet_forward_rows(etai_unet* h, const void* latent, const float* t_rows, int32_t io_dtype, int32_t B,
                           const etai_attn_ctrl* ctrl, void* eps_out, void* stream);

/* ---- null-text inversion: gradient of the UNet output w.r.t. the text context ------------------------------------
 * replaces `loss.backward()` through `unet(latent_cur, t, uncond_embeddings)` at
 * modules/inversion/null_text_inversion.py:75-80 (the only differentiated input is the [1,77,768] uncond embedding).
 *   etai_unet_enable_backward  once per handle: sizes the gradient arena for `max_batch` rows (1 for NTI); the dgrad copies of
 *                              the weights (W^T, mirrored conv filters) are made on first use (+1x weight memory).
 *   etai_unet_forward_train    like etai_unet_forward (same context protocol, no attention control) but eager, with GEGLU
 *                              un-fused and every activation kept in the arena for the backward pass.
 *   etai_unet_backward_ctx     d_ctx[B,77,768] = (d eps / d ctx)^T d_eps for the train-mode forward made immediately before
 *                              on this handle (any other forward in between invalidates it).  d_eps: fp32 [B,4,hw,hw] NCHW;
 *                              d_ctx_out: fp32.  Data gradients only (conv dgrad = conv3x3 with the mirrored filter on the same
 *                              tcgen05 kernel, linear dgrad = GEMM with W^T, flash-style attention backward, norm backward);
 *                              16-bit engines scale the seed to max|.| = 16 on the device and undo it at the end. */
ETAI_EXPORT int etai_unet_enable_backward(etai_unet* h, int32_t max_batch);
ETAI_EXPORT int etai_unet_forward_train(etai_unet* h, const void* latent, float t, int32_t io_dtype, int32_t B, void* eps_out,
                            void* stream);
ETAI_EXPORT int etai_unet_backward_ctx(etai_unet* h, const float* d_eps, int32_t B, float* d_ctx_out, void* stream);

/* Instrumentation.  etai_unet_launch_count: kernels launched by this handle since create (the bench's
 * `gpu_launches` claim).  etai_unet_profile: reads (and clears) the per-category device time accumulated since the
 * previous call -- CUDA events recorded around every op on the caller's stream -- then switches recording on/off.
 * ms_out / launches_out: [ETAI_PROF_NCAT] or NULL.  Recording adds event overhead; never time a bench with it on. */
enum { ETAI_PROF_CONV = 0, ETAI_PROF_GEMM = 1, ETAI_PROF_SELF_ATTN = 2, ETAI_PROF_CROSS_ATTN = 3,
       ETAI_PROF_GROUPNORM = 4, ETAI_PROF_LAYERNORM = 5, ETAI_PROF_OTHER = 6, ETAI_PROF_NCAT = 7 };
ETAI_EXPORT int64_t etai_unet_launch_count(const etai_unet* h);
ETAI_EXPORT int etai_unet_profile(etai_unet* h, int32_t enable, float* ms_out, int32_t* launches_out);

/* Bytes of device memory held by the handle (weights + workspace). */
ETAI_EXPORT int64_t etai_unet_device_bytes(const etai_unet* h);

/* ---- VAE ----------------------------------------------------------------------------------------
 * replaces `model.vae` of the reference pipeline object: `vae.encode(image)['latent_dist'].mean` and
 * `vae.decode(latent)['sample']` (modules/inversion/diffusion_inversion.py:183-208; the 0.18215 latent scaling stays with
 * the caller like in the reference).  Weights: `vae.state_dict()` of diffusers' AutoencoderKL (SD-1.x layout).
 * A handle serialises its own calls, so several threads / streams may share it. */
typedef struct etai_vae etai_vae; /* opaque */
typedef struct {
    int32_t dtype;                 /* storage dtype of weights + activations */
    int32_t math_mode;             /* ETAI_MATH_* */
    int32_t block_out_channels[4]; /* SD-1.x: 128,256,512,512 */
    int32_t image_hw;              /* 512 */
    int32_t max_batch;             /* images per call */
} etai_vae_cfg;
ETAI_EXPORT int etai_vae_create(etai_vae** out, const etai_vae_cfg* cfg, const etai_tensor* weights, int32_t n_weights,
                    int32_t device);
ETAI_EXPORT int etai_vae_destroy(etai_vae* h);
/* image: [B,3,hw,hw] NCHW in [-1,1]; mean_out: [B,4,hw/8,hw/8] NCHW (the posterior mean), both of io_dtype */
ETAI_EXPORT int etai_vae_encode(etai_vae* h, const void* image, int32_t io_dtype, int32_t B, void* mean_out, void* stream);
/* latent: [B,4,hw/8,hw/8] (already divided by 0.18215); image_out: [B,3,hw,hw], both of io_dtype */
ETAI_EXPORT int etai_vae_decode(etai_vae* h, const void* latent, int32_t io_dtype, int32_t B, void* image_out, void* stream);
ETAI_EXPORT int64_t etai_vae_launch_count(const etai_vae* h);
ETAI_EXPORT int64_t etai_vae_device_bytes(const etai_vae* h);

/* ---- CLIP text tower ---------------------------------------------------------------------------
 * replaces `model.text_encoder(input_ids)[0]` (modules/inversion/diffusion_inversion.py:210-247).  Weights:
 * `text_encoder.state_dict()` of transformers' CLIPTextModel (keys "text_model.*").  Causal mask only (the reference
 * passes no attention mask).  Any number of prompts <= max_batch per call. */
typedef struct etai_clip etai_clip; /* opaque */
typedef struct {
    int32_t dtype, math_mode;
    int32_t vocab;     /* 49408 */
    int32_t hidden;    /* 768 */
    int32_t layers;    /* 12 */
    int32_t heads;     /* 12 (head dim must be 64) */
    int32_t ffn;       /* 3072, quick-GELU */
    int32_t max_len;   /* 77 */
    int32_t max_batch; /* prompts per call */
} etai_clip_cfg;
ETAI_EXPORT int etai_clip_create(etai_clip** out, const etai_clip_cfg* cfg, const etai_tensor* weights, int32_t n_weights,
                     int32_t device);
ETAI_EXPORT int etai_clip_destroy(etai_clip* h);
/* input_ids: HOST int32 [B,max_len]; out: device [B,max_len,hidden] of io_dtype = last_hidden_state (final LayerNorm applied) */
ETAI_EXPORT int etai_clip_encode(etai_clip* h, const int32_t* input_ids, int32_t B, void* out, int32_t io_dtype, void* stream);
ETAI_EXPORT int64_t etai_clip_launch_count(const etai_clip* h);

/* ---- scheduler --------------------------------------------------------------------------------
 * One fused kernel for: CFG combine (diffusion_inversion.py:283-284, eta_inversion.py:328)
 *   + DDIM step  x0=(x-sqrt(1-a_from)e)/sqrt(a_from);  sigma = eta*sqrt(var);
 *                x' = sqrt(a_to) x0 + sqrt(1-a_to-sigma^2) e + sigma z
 *     which is BOTH the inverse step of scheduling_ddim_inverse.py:84-100 (eta=0, a_from=a[t-d], a_to=a[t])
 *     and diffusers DDIMScheduler.step(eta, variance_noise) used at eta_inversion.py:245 (a_from=a[t], a_to=a[t-d])
 *   + eta map (eta_inversion.py:236-243: eta = mask * etas[t], a [1,C,H,W] tensor shared by rows)
 *   + variance-noise pick: argmin over `losses` of etai_eta_noise_losses (eta_inversion.py:360-365), on device
 *   + source-row pin x'[0] := pin_src (eta_inversion.py:247-249, direct_inversion.py:43-45)
 * eps: [2n,E] rows [uncond..., cond...] if has_cfg else [n,E];  x, x_out: [n,E];  eps_cfg_out: [n,E] or NULL.
 * eta_map: [E] fp32 or NULL (=1).  noise_cand: [K,E] fp32 or NULL; losses: [K] fp32 (K>1) or NULL (use cand 0).
 * All latent-side buffers are fp32. */
ETAI_EXPORT int etai_cfg_ddim_step(const float* eps, int32_t n, int32_t has_cfg, float guidance, const float* x, float* x_out,
                       float* eps_cfg_out, float a_from, float a_to, float eta, float variance,
                       const float* eta_map, const float* noise_cand, const float* losses, int32_t K,
                       const float* pin_src, int64_t E, void* stream);

/* losses[k] = mean((z_k - z*)^2),  z* = (x_prev_inv - step(eps_cfg_row0, eta, z=0)) / (eta*sqrt(var))
 * (eta_inversion.py:296-317,356-360).  eps: as above (row 0 of each half is the source).  Also writes the
 * picked index to best_idx[0] (int32, device) for inspection; nothing reads it on the host in the loop. */
ETAI_EXPORT int etai_eta_noise_losses(const float* eps, int32_t n, int32_t has_cfg, float guidance, const float* x,
                          const float* x_prev_inv, float a_from, float a_to, float eta, float variance,
                          const float* noise_cand, int32_t K, int64_t E, float* losses, int32_t* best_idx,
                          void* stream);

/* Proximal CFG of proximal negative-prompt inversion (proximal_negative_prompt_inversion.py:61-128, prox = "l0" | "l1"):
 *   delta = eps_c - eps_u;  thr = quantile_q(|delta|) over all n elements;  delta -= clamp(delta, -thr, thr);
 *   l1: delta = where(delta > 0, delta - thr, delta); delta = where(delta < 0, delta + thr, delta);
 *   out = eps_u + guidance * delta.
 * The quantile is torch.quantile's 'linear' rule: with r = float32(q) * float32(n - 1), rank_lo = floor(r), rank_hi = ceil(r),
 * weight = r - rank_lo, thr = lerp(sorted[rank_lo], sorted[rank_hi], weight) -- found by radix select, no sort.
 * rank_lo < 0 selects the fixed threshold `fixed_thr` (the reference's negative `quantile`).  thr_out: device float[1] or
 * NULL.  One launch, no host sync; all buffers fp32; out must not alias the inputs. */
ETAI_EXPORT int etai_prox_guidance(const float* eps_u, const float* eps_c, float* out, int64_t n, int64_t rank_lo, int64_t rank_hi,
                       float weight, float fixed_thr, int32_t l1, float guidance, float* thr_out, void* stream);

/* ---- single ops, exported for unit parity and roofline runs ------------------------------------- */
/* y = SiLU?(GroupNorm(x)) over NHWC x:[B,HW,C] */
ETAI_EXPORT int etai_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int32_t B, int64_t HW, int32_t C,
                   int32_t groups, float eps, int32_t silu, int32_t dtype, void* workspace, int64_t workspace_bytes,
                   void* stream);
/* y = LayerNorm(x) over rows x:[M,C] */
ETAI_EXPORT int etai_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t M, int32_t C, float eps,
                   int32_t dtype, void* stream);
/* C[M,N] = A[M,K] W[N,K]^T (+bias[N]) (+residual[M,N]);  geglu!=0: W rows interleaved (value,gate), C is [M,N/2] */
ETAI_EXPORT int etai_gemm(const void* A, const void* W, const void* bias, const void* residual, void* C, int64_t M, int32_t N,
              int32_t K, int32_t geglu, int32_t dtype, int32_t math_mode, void* stream);
/* y[B,Ho,Wo,Co] = conv3x3(x[B,H,W,Ci], w[Co,3,3,Ci], pad 1, stride) (+bias) (+residual) */
ETAI_EXPORT int etai_conv3x3(const void* x, const void* w, const void* bias, const void* residual, void* y, int32_t B, int32_t H,
                 int32_t Wd, int32_t Ci, int32_t Co, int32_t stride, int32_t dtype, int32_t math_mode,
                 void* workspace, int64_t workspace_bytes, void* stream);
/* out[b,n,h*d..] = softmax(Q K^T scale) V;  q:[B,Nq,ldq] k,v:[B,Nk,ldk] head h at column h*d; row remap optional */
ETAI_EXPORT int etai_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t Nq, int32_t Nk,
                   int32_t heads, int32_t d, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale,
                   const int32_t* q_row, const int32_t* k_row, const int32_t* v_row, int32_t dtype,
                   int32_t math_mode, void* stream);

/* Cross-attention over the text context with the prompt-to-prompt control of ONE layer applied between softmax and PV
 * (the control-aware attention the UNet runs at every attn2; exported for unit parity against an explicit
 * softmax -> edit -> PV, modules/utils/ptp_utils.py:238-253 + modules/utils/ptp.py:205-274).
 *   q:[B,N,ldq]; kv:[B,L,ldkv] with K of head h at column koff+h*d and V at voff+h*d; out:[B,N,ldo].
 *   n_pairs > 0: rows edit_tgt_row[p] are edited from rows edit_base_row[p] with the pair's tables (etai_attn_ctrl).
 *   n_store_rows > 0: store_acc[i][pix][w] += sum_heads P'[store_row[i]][head][pix][w]  (fp32 [n_store_rows,N,L]).
 * workspace (16-bit path only): >= 327680 + heads*n_store_rows*N*L*4 bytes. */
ETAI_EXPORT int etai_cross_attention(const void* q, const void* kv, void* out, int32_t B, int32_t N, int32_t L, int32_t heads,
                         int32_t d, int32_t ldq, int32_t ldkv, int32_t ldo, int32_t koff, int32_t voff, float scale,
                         int32_t n_pairs, const int32_t* edit_base_row, const int32_t* edit_tgt_row, const float* mapper,
                         const float* blend_a, const float* equalizer, const float* alpha_step, int32_t n_store_rows,
                         const int32_t* store_row, float* store_acc, int32_t dtype, int32_t math_mode, void* workspace,
                         int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ETAI_H_ */
