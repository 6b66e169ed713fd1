// C ABI for the scheduler kernels and the individually exported ops (see include/etai.h).
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

}  // extern "C"
