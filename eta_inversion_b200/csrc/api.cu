// C ABI for the scheduler kernels and the individually exported ops (see include/etai.h).
#include <cstring>
#include "ops.cuh"

using namespace etai;

#define ETAI_API_BEGIN try {
#define ETAI_API_END                                        \
    }                                                       \
    catch (const etai::Error& e) {                          \
        etai::set_last_error(e.what());                     \
        return e.code;                                      \
    }                                                       \
    catch (const std::exception& e) {                       \
        etai::set_last_error(e.what());                     \
        return ETAI_ERR_STATE;                              \
    }                                                       \
    return ETAI_OK;

static void fill_map(RowMap& m, int B, const int32_t* q, const int32_t* k, const int32_t* v) {
    ETAI_CHECK(B >= 1 && B <= ETAI_MAX_ROWS, ETAI_ERR_ARG, "attention: B in [1,64]");
    for (int r = 0; r < B; ++r) {
        m.q[r] = q ? q[r] : r;
        m.k[r] = k ? k[r] : r;
        m.v[r] = v ? v[r] : r;
        ETAI_CHECK(m.q[r] >= 0 && m.q[r] < B && m.k[r] >= 0 && m.k[r] < B && m.v[r] >= 0 && m.v[r] < B, ETAI_ERR_ARG,
                   "attention: remap row out of range");
    }
}

extern "C" {

int etai_cfg_ddim_step(const float* eps, int32_t n, int32_t has_cfg, float guidance, const float* x, float* x_out,
                       float* eps_cfg_out, float a_from, float a_to, float eta, float variance, const float* eta_map,
                       const float* noise_cand, const float* losses, int32_t K, const float* pin_src, int64_t E,
                       void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(eps && x && x_out && n >= 1 && E > 0, ETAI_ERR_ARG, "cfg_ddim_step: null/empty argument");
    ETAI_CHECK(a_from > 0.f && a_from <= 1.f && a_to > 0.f && a_to <= 1.f, ETAI_ERR_ARG, "cfg_ddim_step: alphas in (0,1]");
    cfg_ddim_step(eps, n, has_cfg, guidance, x, x_out, eps_cfg_out, a_from, a_to, eta, variance, eta_map, noise_cand,
                  losses, K, pin_src, E, (cudaStream_t)stream);
    ETAI_API_END
}

int etai_eta_noise_losses(const float* eps, int32_t n, int32_t has_cfg, float guidance, const float* x,
                          const float* x_prev_inv, float a_from, float a_to, float eta, float variance,
                          const float* noise_cand, int32_t K, int64_t E, float* losses, int32_t* best_idx,
                          void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(eps && x && x_prev_inv && noise_cand && losses && K >= 1 && n >= 1 && E > 0, ETAI_ERR_ARG,
               "eta_noise_losses: null/empty argument");
    eta_noise_losses(eps, n, has_cfg, guidance, x, x_prev_inv, a_from, a_to, eta, variance, noise_cand, K, E, losses,
                     best_idx, (cudaStream_t)stream);
    ETAI_API_END
}

int etai_prox_guidance(const float* eps_u, const float* eps_c, float* out, int64_t n, int64_t rank_lo, int64_t rank_hi,
                       float weight, float fixed_thr, int32_t l1, float guidance, float* thr_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(eps_u && eps_c && out && n >= 1, ETAI_ERR_ARG, "prox_guidance: null/empty argument");
    ETAI_CHECK(rank_lo < 0 || (rank_lo < n && rank_hi >= rank_lo && rank_hi <= rank_lo + 1 && rank_hi < n), ETAI_ERR_ARG,
               "prox_guidance: need 0 <= rank_lo <= rank_hi <= rank_lo + 1 < n (or rank_lo < 0 for a fixed threshold)");
    prox_guidance(eps_u, eps_c, out, n, rank_lo, rank_hi, weight, fixed_thr, l1, guidance, thr_out, (cudaStream_t)stream);
    ETAI_API_END
}

int etai_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int32_t B, int64_t HW, int32_t C,
                   int32_t groups, float eps, int32_t silu, int32_t dtype, void* workspace, int64_t workspace_bytes,
                   void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(x && y && gamma && beta && B >= 1 && HW >= 1, ETAI_ERR_ARG, "groupnorm: null/empty argument");
    ETAI_CHECK((size_t)workspace_bytes >= groupnorm_workspace_bytes(B, HW, C, groups), ETAI_ERR_ARG,
               "groupnorm: workspace too small (need 33024 + B*256*groups*16 bytes)");
    CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(workspace) + groupnorm_ticket_offset(B, groups), 0, 64 * sizeof(int),
                               (cudaStream_t)stream));
    groupnorm(x, y, gamma, beta, B, HW, C, groups, eps, silu != 0, dtype, workspace, (cudaStream_t)stream);
    ETAI_API_END
}

int etai_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t M, int32_t C, float eps,
                   int32_t dtype, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(x && y && gamma && beta && M >= 1, ETAI_ERR_ARG, "layernorm: null/empty argument");
    layernorm(x, y, gamma, beta, M, C, eps, dtype, (cudaStream_t)stream);
    ETAI_API_END
}

int etai_gemm(const void* A, const void* W, const void* bias, const void* residual, void* C, int64_t M, int32_t N,
              int32_t K, int32_t geglu, int32_t dtype, int32_t math_mode, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(A && W && C && M >= 1 && N >= 1 && K >= 1, ETAI_ERR_ARG, "gemm: null/empty argument");
    GemmArgs a;
    int nout = geglu ? N / 2 : N;
    a.A = A; a.W = W; a.C = C; a.bias = bias; a.residual = residual;
    a.M = M; a.N = N; a.K = K; a.lda = K; a.ldc = nout; a.ldr = nout; a.geglu = geglu; a.dtype = dtype;
    if (math_mode == ETAI_MATH_AUTO && dtype != ETAI_F32) {
        ETAI_CHECK(gemm_tc_supported(a), ETAI_ERR_UNSUPPORTED, "gemm: shape not supported by the tcgen05 path");
        gemm_tc(a, nullptr, 0, (cudaStream_t)stream);
    } else {
        gemm_simt(a, (cudaStream_t)stream);
    }
    ETAI_API_END
}

int etai_conv3x3(const void* x, const void* w, const void* bias, const void* residual, void* y, int32_t B, int32_t H,
                 int32_t Wd, int32_t Ci, int32_t Co, int32_t stride, int32_t dtype, int32_t math_mode,
                 void* workspace, int64_t workspace_bytes, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(x && w && y && B >= 1 && H >= 1 && Wd >= 1 && (stride == 1 || stride == 2), ETAI_ERR_ARG,
               "conv3x3: null/empty argument or bad stride");
    GemmArgs a;
    a.conv = 1; a.B = B; a.H = H; a.Wd = Wd; a.Cin = Ci; a.stride = stride;
    a.Ho = (H - 1) / stride + 1; a.Wo = (Wd - 1) / stride + 1;
    a.A = x; a.W = w; a.C = y; a.bias = bias; a.residual = residual;
    a.M = (long)B * a.Ho * a.Wo; a.N = Co; a.K = 9 * Ci; a.ldc = Co; a.ldr = Co; a.dtype = dtype;
    if (math_mode == ETAI_MATH_AUTO && dtype != ETAI_F32) {
        ETAI_CHECK(gemm_tc_supported(a), ETAI_ERR_UNSUPPORTED, "conv3x3: shape not supported by the tcgen05 path");
        gemm_tc(a, workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
    } else {
        gemm_simt(a, (cudaStream_t)stream);
    }
    ETAI_API_END
}

int etai_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t Nq, int32_t Nk,
                   int32_t heads, int32_t d, int32_t ldq, int32_t ldk, int32_t ldv, int32_t ldo, float scale,
                   const int32_t* q_row, const int32_t* k_row, const int32_t* v_row, int32_t dtype,
                   int32_t math_mode, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(q && k && v && out && Nq >= 1 && Nk >= 1 && heads >= 1, ETAI_ERR_ARG, "attention: null/empty argument");
    SelfAttnArgs a;
    a.q = q; a.k = k; a.v = v; a.out = out;
    a.B = B; a.Nq = Nq; a.Nk = Nk; a.heads = heads; a.d = d;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.scale = scale; a.dtype = dtype;
    fill_map(a.map, B, q_row, k_row, v_row);
    if (math_mode == ETAI_MATH_AUTO && dtype != ETAI_F32) {
        ETAI_CHECK(attention_tc_supported(a), ETAI_ERR_UNSUPPORTED, "attention: shape not supported by the tcgen05 path");
        attention_tc(a, (cudaStream_t)stream);
    } else {
        attention_simt(a, (cudaStream_t)stream);
    }
    ETAI_API_END
}

int etai_cross_attention(const void* q, const void* kv, void* out, int32_t B, int32_t N, int32_t L, int32_t heads, int32_t d,
                         int32_t ldq, int32_t ldkv, int32_t ldo, int32_t koff, int32_t voff, float scale, int32_t n_pairs,
                         const int32_t* edit_base_row, const int32_t* edit_tgt_row, const float* mapper,
                         const float* blend_a, const float* equalizer, const float* alpha_step, int32_t n_store_rows,
                         const int32_t* store_row, float* store_acc, int32_t dtype, int32_t math_mode, void* workspace,
                         int64_t workspace_bytes, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(q && kv && out && B >= 1 && B <= ETAI_MAX_ROWS && N >= 1 && L >= 1 && heads >= 1, ETAI_ERR_ARG,
               "cross_attention: null/empty argument");
    ETAI_CHECK(n_pairs >= 0 && n_pairs <= ETAI_MAX_PAIRS && n_store_rows >= 0 && n_store_rows <= ETAI_MAX_ROWS, ETAI_ERR_ARG,
               "cross_attention: pair / store row count out of range");
    if (n_pairs > 0)
        ETAI_CHECK(edit_base_row && edit_tgt_row && mapper && blend_a && equalizer && alpha_step, ETAI_ERR_ARG,
                   "cross_attention: edit tables missing");
    if (n_store_rows > 0) ETAI_CHECK(store_row && store_acc, ETAI_ERR_ARG, "cross_attention: store arguments missing");
    cudaStream_t s = (cudaStream_t)stream;
    CrossAttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = q; a.kv = kv; a.out = out;
    a.B = B; a.N = N; a.L = L; a.heads = heads; a.d = d;
    a.ldq = ldq; a.ldkv = ldkv; a.ldo = ldo; a.koff = koff; a.voff = voff; a.scale = scale; a.dtype = dtype;
    int slot_of[ETAI_MAX_ROWS];
    bool used[ETAI_MAX_ROWS] = {false};
    for (int r = 0; r < B; ++r) slot_of[r] = -1;
    for (int i = 0; i < n_store_rows; ++i) {
        ETAI_CHECK(store_row[i] >= 0 && store_row[i] < B, ETAI_ERR_ARG, "cross_attention: store row out of range");
        slot_of[store_row[i]] = i;
    }
    if (n_store_rows > 0) a.store = store_acc;
    int ng = 0;
    if (n_pairs > 0) {
        a.mapper = mapper; a.blend_a = blend_a; a.equalizer = equalizer; a.alpha_step = alpha_step;
        for (int p = 0; p < n_pairs; ++p) {
            int br = edit_base_row[p], tr = edit_tgt_row[p];
            ETAI_CHECK(br >= 0 && br < B && tr >= 0 && tr < B && br != tr && !used[br] && !used[tr], ETAI_ERR_ARG,
                       "cross_attention: bad edit pair");
            used[br] = used[tr] = true;
            a.groups[ng++] = CrossGroup{br, tr, p, slot_of[br], slot_of[tr]};
        }
    }
    for (int r = 0; r < B; ++r)
        if (!used[r]) a.groups[ng++] = CrossGroup{r, -1, 0, slot_of[r], -1};
    a.n_groups = ng;
    if (math_mode == ETAI_MATH_AUTO && dtype != ETAI_F32) {
        // workspace: [16-bit mapper operand][per-head store partials]
        size_t mb = (cross_attention_tc_mapper_bytes(ETAI_MAX_PAIRS) + 255) & ~size_t(255);
        size_t sb = n_store_rows > 0 ? (size_t)heads * n_store_rows * N * L * sizeof(float) : 0;
        ETAI_CHECK(workspace && (size_t)workspace_bytes >= mb + sb, ETAI_ERR_ARG,
                   "cross_attention: workspace too small (need 327680 + heads*n_store_rows*N*L*4 bytes)");
        if (n_pairs > 0) {
            CUDA_CHECK(cudaMemsetAsync(workspace, 0, mb, s));
            cross_attention_tc_prep_mapper(mapper, workspace, n_pairs, L, dtype, s);
            a.map16 = workspace;
        }
        a.store_part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + mb);
        ETAI_CHECK(cross_attention_tc_supported(a), ETAI_ERR_UNSUPPORTED,
                   "cross_attention: shape not supported by the tcgen05 path");
        cross_attention_tc(a, s);
    } else {
        cross_attention(a, s);
    }
    ETAI_API_END
}

}  // extern "C"
