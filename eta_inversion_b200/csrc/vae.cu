// SD-1.x AutoencoderKL (diffusers 0.21.1 layout) as an explicit kernel schedule on the engine's own kernels.
//
// Replaces `model.vae.encode(image)['latent_dist'].mean` and `model.vae.decode(latent)['sample']` of the reference
// (modules/inversion/diffusion_inversion.py:183-208): 1 encode + 2 decodes sit inside every timed edit
// (edit_image.py:113-115).  Same building blocks as the UNet: NHWC activations, conv3x3 as implicit GEMM on tcgen05
// (image rows wider than one 128-pixel tile are walked tile by tile, the halo comes from TMA out-of-bounds zero fill),
// GroupNorm(+SiLU) kernels, dense GEMMs.  The mid-block attention is single-head with d = 512 over 64x64 tokens: its two
// products are plain GEMMs around an in-place row softmax of the materialised [4096,4096] logits of one image
// (34 GFLOP per image; V is produced already transposed by a GEMM with swapped operands, its bias is added after PV
// because the rows of P sum to one).
//
// Memory: activations ping-pong between two slots sized for the largest tensor (B x 512 x 512 x 256), temporaries of one
// block live in a bump arena that is rewound after the block.  A handle serialises its calls (mutex + event), so the
// lanes of a lock-step group and the two pipelined groups of one GPU can share it.
#include <mutex>
#include "engine_base.cuh"

using namespace etai;

namespace {
struct VRes { Norm n1, n2; Conv c1, c2; Lin sc; bool has_sc = false; int cin = 0, cout = 0; };
struct VAttn { Norm gn; Lin q, k, v, o; int C = 0; };
}  // namespace

struct etai_vae : etai::OpCtx {
    etai_vae_cfg cfg;
    // encoder
    Conv e_in, e_out, e_down[3];
    std::vector<VRes> e_res[4];
    VRes e_mid[2];
    VAttn e_attn;
    Norm e_norm;
    Lin quant;
    // decoder
    Lin post_quant;
    Conv d_in, d_out, d_up[3];
    VRes d_mid[2];
    VAttn d_attn;
    std::vector<VRes> d_res[4];
    Norm d_norm;

    char* slot[2] = {nullptr, nullptr};
    size_t slot_bytes = 0, slot_need = 0;
    size_t workspace_bytes = 0;
    std::mutex mu;
    cudaEvent_t ev_done = nullptr;
    bool has_done = false;

    bool planning() const { return arena.base == nullptr; }
    void* use_slot(int i, size_t bytes) {
        if (bytes > slot_need) slot_need = bytes;
        if (planning()) return reinterpret_cast<void*>(size_t(256));
        ETAI_CHECK(bytes <= slot_bytes, ETAI_ERR_NOMEM, "vae: activation slot too small");
        return slot[i];
    }

    VRes load_vres(const std::string& p, int cin, int cout) {
        VRes r;
        r.cin = cin; r.cout = cout;
        r.n1 = load_norm(p + ".norm1", cin);
        r.c1 = load_conv(p + ".conv1", cin, cout);
        r.n2 = load_norm(p + ".norm2", cout);
        r.c2 = load_conv(p + ".conv2", cout, cout);
        r.has_sc = cin != cout;
        if (r.has_sc) r.sc = load_lin(p + ".conv_shortcut", cout, cin, true, true);
        return r;
    }
    VAttn load_vattn(const std::string& p, int C) {
        VAttn a;
        a.C = C;
        a.gn = load_norm(p + ".group_norm", C);
        a.q = load_lin(p + ".to_q", C, C, true);
        a.k = load_lin(p + ".to_k", C, C, true);
        a.v = load_lin(p + ".to_v", C, C, true);
        a.o = load_lin(p + ".to_out.0", C, C, true);
        return a;
    }

    void build(const etai_tensor* weights, int n_weights) {
        for (int i = 0; i < n_weights; ++i) table[weights[i].name] = &weights[i];
        const int* c = cfg.block_out_channels;
        stage_elems = (size_t)c[3] * c[3] * 9;
        CUDA_CHECK(cudaMalloc(&stage, stage_elems * 4 + stage_elems * 2));
        const int io_pad = tc ? 64 : 0, out_pad = tc ? 32 : 0;
        // ---- encoder ----
        e_in = load_conv("encoder.conv_in", 3, c[0], io_pad, 0);
        for (int i = 0; i < 4; ++i) {
            const std::string p = "encoder.down_blocks." + std::to_string(i);
            for (int j = 0; j < 2; ++j)
                e_res[i].push_back(load_vres(p + ".resnets." + std::to_string(j), j == 0 ? c[i > 0 ? i - 1 : 0] : c[i], c[i]));
            if (i < 3) e_down[i] = load_conv(p + ".downsamplers.0.conv", c[i], c[i]);
        }
        e_mid[0] = load_vres("encoder.mid_block.resnets.0", c[3], c[3]);
        e_attn = load_vattn("encoder.mid_block.attentions.0", c[3]);
        e_mid[1] = load_vres("encoder.mid_block.resnets.1", c[3], c[3]);
        e_norm = load_norm("encoder.conv_norm_out", c[3]);
        e_out = load_conv("encoder.conv_out", c[3], 8, 0, out_pad);
        quant = load_lin("quant_conv", 8, 8, true, true);
        // ---- decoder ----
        post_quant = load_lin("post_quant_conv", 4, 4, true, true);
        d_in = load_conv("decoder.conv_in", 4, c[3], io_pad, 0);
        d_mid[0] = load_vres("decoder.mid_block.resnets.0", c[3], c[3]);
        d_attn = load_vattn("decoder.mid_block.attentions.0", c[3]);
        d_mid[1] = load_vres("decoder.mid_block.resnets.1", c[3], c[3]);
        for (int i = 0; i < 4; ++i) {
            const std::string p = "decoder.up_blocks." + std::to_string(i);
            const int cout = c[3 - i], cin = i == 0 ? c[3] : c[3 - i + 1];
            for (int j = 0; j < 3; ++j) d_res[i].push_back(load_vres(p + ".resnets." + std::to_string(j), j == 0 ? cin : cout, cout));
            if (i < 3) d_up[i] = load_conv(p + ".upsamplers.0.conv", cout, cout);
        }
        d_norm = load_norm("decoder.conv_norm_out", c[0]);
        d_out = load_conv("decoder.conv_out", c[0], 3, 0, out_pad);
        CUDA_CHECK(cudaFree(stage));
        stage = nullptr;
        table.clear();
    }

    // ---- blocks: input in slot[cur], result in slot[cur ^ 1]; temporaries in the arena (rewound by the caller) ----
    void* resnet(const void* x, int B, int H, int W, const VRes& r, int dst, cudaStream_t s) {
        const long HW = (long)H * W, M = B * HW;
        void* a1 = gnorm(x, B, HW, r.n1, 1e-6f, true, s);
        void* h = conv3x3(a1, B, H, W, r.c1, 1, nullptr, nullptr, s);
        void* a2 = gnorm(h, B, HW, r.n2, 1e-6f, true, s);
        const void* skip = x;
        if (r.has_sc) skip = linear(x, M, r.sc, nullptr, s);
        void* out = use_slot(dst, (size_t)M * r.cout * esz);
        return conv3x3(a2, B, H, W, r.c2, 1, nullptr, skip, s, out);
    }
    void* attention(const void* x, int B, int H, int W, const VAttn& t, int dst, cudaStream_t s) {
        const int C = t.C;
        const long N = (long)H * W, M = B * N;
        void* g = gnorm(x, B, N, t.gn, 1e-6f, false, s);
        void* q = linear(g, M, t.q, nullptr, s);
        void* k = linear(g, M, t.k, nullptr, s);
        void* o = arena.alloc((size_t)M * C * esz);
        void* vt = arena.alloc((size_t)C * N * esz);
        void* S = arena.alloc((size_t)N * N * esz);
        if (!planning()) {
            for (int b = 0; b < B; ++b) {
                const char* gb = (const char*)g + (size_t)b * N * C * esz;
                GemmArgs a;  // V^T[C,N] = Wv[C,C] * g_b[N,C]^T
                a.A = t.v.w; a.W = gb; a.C = vt; a.M = C; a.N = (int)N; a.K = C; a.lda = C; a.ldc = N;
                gemm(a, s);
                GemmArgs l;  // S[N,N] = q_b k_b^T
                l.A = (const char*)q + (size_t)b * N * C * esz; l.W = (const char*)k + (size_t)b * N * C * esz; l.C = S;
                l.M = N; l.N = (int)N; l.K = C; l.lda = C; l.ldc = N;
                gemm(l, s);
                cudaEvent_t e = prof_begin(s);
                softmax_rows(S, N, (int)N, 1.0f / sqrtf((float)C), dt, s);
                prof_end(ETAI_PROF_SELF_ATTN, e, 1, s);
                GemmArgs pv;  // o_b[N,C] = P[N,N] * (V^T)[C,N]^T + bv
                pv.A = S; pv.W = vt; pv.C = (char*)o + (size_t)b * N * C * esz; pv.bias = t.v.b;
                pv.M = N; pv.N = C; pv.K = (int)N; pv.lda = N; pv.ldc = C;
                gemm(pv, s);
            }
        }
        void* out = use_slot(dst, (size_t)M * C * esz);
        return linear(o, M, t.o, x, s, 0, out);
    }

    void encode_body(const void* image, int io_dtype, int B, void* mean_out, cudaStream_t s) {
        int H = cfg.image_hw, W = cfg.image_hw;
        arena.reset();
        int cur = 0;
        void* x0 = arena.alloc((size_t)B * H * W * e_in.cin * esz);
        if (!planning()) {
            nchw_to_nhwc(image, io_dtype, x0, dt, B, 3, e_in.cin, (long)H * W, s);
            launches += 1;
        }
        void* h = conv3x3(x0, B, H, W, e_in, 1, nullptr, nullptr, s, use_slot(cur, (size_t)B * H * W * e_in.cout * esz));
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 2; ++j) {
                arena.reset();
                h = resnet(h, B, H, W, e_res[i][j], cur ^ 1, s);
                cur ^= 1;
            }
            if (i < 3) {
                arena.reset();
                void* o = use_slot(cur ^ 1, (size_t)B * (H / 2) * (W / 2) * e_down[i].cout * esz);
                h = conv3x3(h, B, H, W, e_down[i], 2, nullptr, nullptr, s, o, /*pad=*/0);
                cur ^= 1;
                H /= 2; W /= 2;
            }
        }
        arena.reset(); h = resnet(h, B, H, W, e_mid[0], cur ^ 1, s); cur ^= 1;
        arena.reset(); h = attention(h, B, H, W, e_attn, cur ^ 1, s); cur ^= 1;
        arena.reset(); h = resnet(h, B, H, W, e_mid[1], cur ^ 1, s); cur ^= 1;
        arena.reset();
        const long M = (long)B * H * W;
        void* a = gnorm(h, B, (long)H * W, e_norm, 1e-6f, true, s);
        void* m = conv3x3(a, B, H, W, e_out, 1, nullptr, nullptr, s);     // [M, 8 (padded to e_out.cout)]
        void* qz = arena.alloc((size_t)M * 8 * esz);
        if (!planning()) {
            GemmArgs g;  // quant_conv (1x1, 8 -> 8) over the first 8 columns of the padded conv_out result
            g.A = m; g.W = quant.w; g.C = qz; g.bias = quant.b; g.M = M; g.N = 8; g.K = 8; g.lda = e_out.cout; g.ldc = 8;
            gemm(g, s);
            nhwc_to_nchw(qz, dt, mean_out, io_dtype, B, 4, 8, (long)H * W, s);  // latent_dist.mean = channels 0..3
            launches += 1;
        }
    }

    void decode_body(const void* latent, int io_dtype, int B, void* image_out, cudaStream_t s) {
        int H = cfg.image_hw / 8, W = cfg.image_hw / 8;
        arena.reset();
        int cur = 0;
        const long M0 = (long)B * H * W;
        void* z = arena.alloc((size_t)M0 * 8 * esz);
        void* zp = arena.alloc((size_t)M0 * d_in.cin * esz);
        if (!planning()) {
            nchw_to_nhwc(latent, io_dtype, z, dt, B, 4, 8, (long)H * W, s);
            CUDA_CHECK(cudaMemsetAsync(zp, 0, (size_t)M0 * d_in.cin * esz, s));  // channel padding of conv_in's input
            GemmArgs g;  // post_quant_conv (1x1, 4 -> 4)
            g.A = z; g.W = post_quant.w; g.C = zp; g.bias = post_quant.b; g.M = M0; g.N = 4; g.K = 4; g.lda = 8; g.ldc = d_in.cin;
            gemm(g, s);
            launches += 2;
        }
        void* h = conv3x3(zp, B, H, W, d_in, 1, nullptr, nullptr, s, use_slot(cur, (size_t)M0 * d_in.cout * esz));
        arena.reset(); h = resnet(h, B, H, W, d_mid[0], cur ^ 1, s); cur ^= 1;
        arena.reset(); h = attention(h, B, H, W, d_attn, cur ^ 1, s); cur ^= 1;
        arena.reset(); h = resnet(h, B, H, W, d_mid[1], cur ^ 1, s); cur ^= 1;
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 3; ++j) {
                arena.reset();
                h = resnet(h, B, H, W, d_res[i][j], cur ^ 1, s);
                cur ^= 1;
            }
            if (i < 3) {
                arena.reset();
                const int C = d_up[i].cin;
                void* up = arena.alloc((size_t)B * 4 * H * W * C * esz);
                if (!planning()) {
                    cudaEvent_t e = prof_begin(s);
                    upsample2x(h, up, B, H, W, C, dt, s);
                    prof_end(ETAI_PROF_OTHER, e, 1, s);
                }
                H *= 2; W *= 2;
                h = conv3x3(up, B, H, W, d_up[i], 1, nullptr, nullptr, s, use_slot(cur ^ 1, (size_t)B * H * W * d_up[i].cout * esz));
                cur ^= 1;
            }
        }
        arena.reset();
        void* a = gnorm(h, B, (long)H * W, d_norm, 1e-6f, true, s);
        void* o = conv3x3(a, B, H, W, d_out, 1, nullptr, nullptr, s);
        if (!planning()) {
            nhwc_to_nchw(o, dt, image_out, io_dtype, B, 3, d_out.cout, (long)H * W, s);
            launches += 1;
        }
    }

    void plan_workspace() {
        arena.base = nullptr; arena.cap = 0; arena.peak = 0;
        encode_body(nullptr, ETAI_F32, cfg.max_batch, nullptr, 0);
        decode_body(nullptr, ETAI_F32, cfg.max_batch, nullptr, 0);
        slot_bytes = (slot_need + 255) & ~size_t(255);
        CUDA_CHECK(cudaMalloc((void**)&slot[0], slot_bytes));
        CUDA_CHECK(cudaMalloc((void**)&slot[1], slot_bytes));
        size_t need = arena.peak + 4096;
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, need));
        arena.base = (char*)p; arena.cap = need; arena.reset();
        size_t gws = groupnorm_workspace_bytes(cfg.max_batch, 0, 0, 32);
        CUDA_CHECK(cudaMalloc(&gn_ws, gws));
        CUDA_CHECK(cudaMemset(gn_ws, 0, gws));
        tc_ws_bytes = 0;
        if (tc) {  // im2col scratch of the three stride-2 convs: largest is B x 256 x 256 x 9*c0
            const int* c = cfg.block_out_channels;
            size_t m = 0;
            long hw = (long)cfg.image_hw * cfg.image_hw / 4;
            for (int i = 0; i < 3; ++i) {
                size_t b = (size_t)cfg.max_batch * hw * 9 * c[i] * esz;
                if (b > m) m = b;
                hw /= 4;
            }
            tc_ws_bytes = (m + 255) & ~size_t(255);
            CUDA_CHECK(cudaMalloc(&tc_ws, tc_ws_bytes));
        }
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
        workspace_bytes = 2 * slot_bytes + need + gws + tc_ws_bytes;
    }

    template <typename F>
    void run(cudaStream_t s, F&& body) {
        std::lock_guard<std::mutex> lk(mu);  // the buffers are the handle's: calls are enqueued one after the other ...
        if (has_done) CUDA_CHECK(cudaStreamWaitEvent(s, ev_done, 0));  // ... and execute one after the other across streams
        body();
        CUDA_CHECK(cudaEventRecord(ev_done, s));
        has_done = true;
    }
};

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

extern "C" {

int etai_vae_destroy(etai_vae* h) {
    if (!h) return ETAI_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    h->wstore.reset();
    void* extra[] = {h->slot[0], h->slot[1], h->arena.base, h->gn_ws, h->tc_ws, h->stage};
    for (void* p : extra)
        if (p) cudaFree(p);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    delete h;
    return ETAI_OK;
}

int etai_vae_create(etai_vae** out, const etai_vae_cfg* cfg, const etai_tensor* weights, int32_t n_weights, int32_t device) {
    ETAI_API_BEGIN
    ETAI_CHECK(out && cfg && weights && n_weights > 0, ETAI_ERR_ARG, "vae_create: null argument");
    ETAI_CHECK(cfg->dtype == ETAI_F32 || cfg->dtype == ETAI_F16 || cfg->dtype == ETAI_BF16, ETAI_ERR_ARG, "vae_create: dtype");
    ETAI_CHECK(cfg->max_batch >= 1 && cfg->max_batch <= 16, ETAI_ERR_ARG, "vae_create: max_batch in [1,16]");
    ETAI_CHECK(cfg->image_hw >= 64 && cfg->image_hw % 64 == 0 && cfg->image_hw <= 1024, ETAI_ERR_ARG,
               "vae_create: image_hw must be a multiple of 64 in [64,1024]");
    for (int i = 0; i < 4; ++i)
        ETAI_CHECK(cfg->block_out_channels[i] % 64 == 0, ETAI_ERR_ARG, "vae_create: channels must be multiples of 64");
    int ndev = 0;
    CUDA_CHECK(cudaGetDeviceCount(&ndev));
    ETAI_CHECK(device >= 0 && device < ndev, ETAI_ERR_ARG, "vae_create: no such CUDA device");
    CUDA_CHECK(cudaSetDevice(device));
    etai_vae* h = new etai_vae();
    try {
        h->cfg = *cfg;
        h->device = device;
        h->wstore = std::make_shared<WeightStore>();
        h->wstore->device = device;
        h->dt = cfg->dtype;
        h->esz = dtype_size(cfg->dtype);
        h->tc = cfg->dtype != ETAI_F32 && cfg->math_mode == ETAI_MATH_AUTO;
        h->build(weights, n_weights);
        h->plan_workspace();
    } catch (...) {
        etai_vae_destroy(h);
        throw;
    }
    *out = h;
    ETAI_API_END
}

int etai_vae_encode(etai_vae* h, const void* image, int32_t io_dtype, int32_t B, void* mean_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && image && mean_out, ETAI_ERR_ARG, "vae_encode: null argument");
    ETAI_CHECK(B >= 1 && B <= h->cfg.max_batch, ETAI_ERR_ARG, "vae_encode: batch out of range");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->run((cudaStream_t)stream, [&] { h->encode_body(image, io_dtype, B, mean_out, (cudaStream_t)stream); });
    ETAI_API_END
}

int etai_vae_decode(etai_vae* h, const void* latent, int32_t io_dtype, int32_t B, void* image_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && latent && image_out, ETAI_ERR_ARG, "vae_decode: null argument");
    ETAI_CHECK(B >= 1 && B <= h->cfg.max_batch, ETAI_ERR_ARG, "vae_decode: batch out of range");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->run((cudaStream_t)stream, [&] { h->decode_body(latent, io_dtype, B, image_out, (cudaStream_t)stream); });
    ETAI_API_END
}

int64_t etai_vae_launch_count(const etai_vae* h) { return h ? h->launches : 0; }
int64_t etai_vae_device_bytes(const etai_vae* h) { return h ? (int64_t)(h->weight_bytes + h->workspace_bytes) : 0; }

}  // extern "C"
