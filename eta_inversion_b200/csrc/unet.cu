// SD-1.x UNet2DConditionModel as an explicit kernel schedule (NHWC inside, one stream, no host sync).
//
// Replaces `model.unet(sample, t, encoder_hidden_states=ctx)["sample"]` of the reference
// (modules/inversion/diffusion_inversion.py:264-280, eta_inversion.py:321); architecture per SURVEY.md Appendix A
// (diffusers 0.21.1).  Weight names/layouts are the diffusers state_dict; they are repacked once at create():
//   conv OIHW -> [O][ky][kx][I] (implicit-GEMM K order), to_q/to_k/to_v of attn1 stacked to one [3C,C] GEMM,
//   all 16 cross-attention to_k/to_v stacked to one [sum 2C, 768] GEMM that runs once per context,
//   all 22 time_emb_proj stacked to one skinny GEMM, GEGLU rows interleaved (value,gate) for the fused epilogue.
#include <unordered_set>
#include "engine_base.cuh"

namespace etai {

namespace {

struct Res { Norm n1, n2; Conv c1, c2; Lin sc; bool has_sc = false; int cin = 0, cout = 0, temb_off = 0; };
struct Tfm {
    Norm gn, ln1, ln2, ln3;
    Lin pin, qkv, o1, q2, o2, ff1, ff2, pout;
    int C = 0, index = 0, place = 0, kv_off = 0;  // place: 0 down, 1 mid, 2 up
};

// fp32 scratch layout of the time-embedding path
constexpr int TB_SIN = 0, TB_H1 = 4096, TB_ST = 8192, TB_PROJ = 16384;

}  // namespace

}  // namespace etai

using namespace etai;

struct etai_unet : etai::OpCtx {
    etai_unet_cfg cfg;

    Conv conv_in, conv_out;
    Norm norm_out;
    Lin time1, time2, temb_all;   // temb_all: stacked time_emb_proj of every resnet
    Lin kv_all;                   // stacked cross-attention to_k/to_v of every transformer
    std::vector<Res> down_res[4], up_res[4];
    std::vector<Tfm> down_tf[4], up_tf[4];
    Conv down_samp[3], up_samp[3];
    Res mid_res[2];
    Tfm mid_tf;
    int n_tf = 0, temb_total = 0, kv_total = 0;

    // per-forward workspace
    void* kv_cache = nullptr;  // [max_batch*ctx_len, kv_total]
    void* ctx_buf = nullptr;   // [max_batch*ctx_len, cross_dim] in storage dtype
    int ctx_rows = 0;          // batch rows of the cached context (0 = none)
    float* tbuf = nullptr;     // time embedding scratch (fp32)
    // ---- CUDA-graph replay: all per-call inputs are staged into engine-owned buffers (stable addresses), the kernel
    // schedule of one forward is captured once per (batch, control structure) key and replayed afterwards.
    cudaStream_t gs = nullptr;          // engine stream (capture on the legacy default stream is not allowed)
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    char* in_stage = nullptr;           // latent  [max_batch,4,hw,hw] (io dtype, <= 4 bytes/elem)
    char* out_stage = nullptr;          // eps out
    float* t_dev = nullptr;
    // one timestep per batch row (etai_unet_forward_rows): per-row copies of the time-embedding scratch, made on first use
    float* tbuf_rows = nullptr;
    const float* t_rows_host = nullptr;  // non-null only inside a mixed-timestep forward()
    size_t tb_stride() const { return (size_t)(TB_PROJ + temb_total + 64); }
    float *c_mapper = nullptr, *c_blend = nullptr, *c_eq = nullptr, *c_alpha = nullptr;  // staged PtP tables
    float* c_store[3] = {nullptr, nullptr, nullptr};  // per-forward attention-store sums per place
    bool use_graphs = true;
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int seen = 0; int64_t launches = 0; };
    std::unordered_map<std::string, GraphEntry> graphs;
    int64_t graph_replays = 0;
    void* map16_buf = nullptr;      // prompt-to-prompt mapper in the tcgen05 B-operand form (per forward)
    float* store_part_buf = nullptr;  // per-head attention-store partials
    bool map16_ready = false;
    size_t workspace_bytes = 0;

    Res load_res(const std::string& p, int cin, int cout, int temb) {
        Res r;
        r.cin = cin; r.cout = cout;
        r.n1 = load_norm(p + ".norm1", cin);
        r.c1 = load_conv(p + ".conv1", cin, cout);
        r.n2 = load_norm(p + ".norm2", cout);
        r.c2 = load_conv(p + ".conv2", cout, cout);
        r.has_sc = cin != cout;
        if (r.has_sc) r.sc = load_lin(p + ".conv_shortcut", cout, cin, true, true);
        // time_emb_proj goes into the stacked matrix
        r.temb_off = temb_total;
        put(p + ".time_emb_proj.weight", {cout, temb}, (char*)temb_all.w + (size_t)temb_total * temb * esz);
        put(p + ".time_emb_proj.bias", {cout}, (char*)temb_all.b + (size_t)temb_total * esz);
        temb_total += cout;
        return r;
    }
    Tfm load_tfm(const std::string& p, int C, int place) {
        Tfm t;
        t.C = C; t.place = place; t.index = n_tf++;
        const std::string b = p + ".transformer_blocks.0";
        int X = cfg.cross_dim;
        t.gn = load_norm(p + ".norm", C);
        t.pin = load_lin(p + ".proj_in", C, C, true, true);
        t.ln1 = load_norm(b + ".norm1", C);
        t.ln2 = load_norm(b + ".norm2", C);
        t.ln3 = load_norm(b + ".norm3", C);
        t.qkv.n = 3 * C; t.qkv.k = C;
        t.qkv.w = dmalloc((size_t)3 * C * C * esz);
        put(b + ".attn1.to_q.weight", {C, C}, t.qkv.w);
        put(b + ".attn1.to_k.weight", {C, C}, (char*)t.qkv.w + (size_t)C * C * esz);
        put(b + ".attn1.to_v.weight", {C, C}, (char*)t.qkv.w + (size_t)2 * C * C * esz);
        t.o1 = load_lin(b + ".attn1.to_out.0", C, C, true);
        t.q2 = load_lin(b + ".attn2.to_q", C, C, false);
        t.kv_off = kv_total;
        put(b + ".attn2.to_k.weight", {C, X}, (char*)kv_all.w + (size_t)kv_total * X * esz);
        put(b + ".attn2.to_v.weight", {C, X}, (char*)kv_all.w + (size_t)(kv_total + C) * X * esz);
        kv_total += 2 * C;
        t.o2 = load_lin(b + ".attn2.to_out.0", C, C, true);
        // GEGLU projection with (value,gate) row interleave
        t.ff1.n = 8 * C; t.ff1.k = C;
        t.ff1.w = dmalloc((size_t)8 * C * C * esz);
        t.ff1.b = dmalloc((size_t)8 * C * esz);
        {
            const float* w = staged(b + ".ff.net.0.proj.weight", {8 * C, C});
            // bias staged right behind the weight in the staging buffer
            const etai_tensor& bt = find(b + ".ff.net.0.proj.bias");
            ETAI_CHECK(bt.ndim == 1 && bt.shape[0] == 8 * C && bt.dtype == ETAI_F32, ETAI_ERR_ARG, "geglu bias must be fp32 [8C]");
            float* bs = stage + (size_t)8 * C * C;
            CUDA_CHECK(cudaMemcpy(bs, bt.data, (size_t)8 * C * 4, bt.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
            pack_geglu_weight(w, bs, t.ff1.w, t.ff1.b, 8 * C, C, dt, 0);
            CUDA_CHECK(cudaStreamSynchronize(0));
        }
        t.ff2 = load_lin(b + ".ff.net.2", C, 4 * C, true);
        t.pout = load_lin(p + ".proj_out", C, C, true, true);
        return t;
    }

    // ---- backward w.r.t. the text context (null-text inversion) ----
    bool bwd_enabled = false;
    int bwd_batch = 0;            // rows of the recorded train-mode forward (0 = none)
    int bwd_max_batch = 0;
    void* last_out = nullptr;     // conv_out result of the last forward (NHWC, padded channels)
    std::unordered_map<const void*, void*> wT;   // packed forward weight -> dgrad weight (W^T / flipped conv filter)
    std::vector<void*> bwd_owned;
    Arena garena;                 // gradients + temporaries of one backward pass
    float* dkv = nullptr;         // [bwd_max_batch*ctx_len, kv_total] fp32: d(K|V) of all 16 cross-attention layers
    float* cross_part = nullptr;  // per-tile dK/dV partials of one cross-attention layer
    float* lse_buf = nullptr;     // [2][bwd_max_batch*heads*HW] log-sum-exp and rowsum(dO*O) of one self-attention layer
    float* scale_dev = nullptr;   // loss scale of the 16-bit backward pass (device scalar)
    void* bwd_seed = nullptr;     // dL/d(conv_out result), NHWC
    void* bwd_dctx = nullptr;     // result of the last walk: dL/d(ctx) in storage dtype
    size_t cross_part_need = 0;
    void enable_backward(int max_batch);
    void* make_wT(const void* w, int n, int k, bool conv);
    void backward_walk(int B, cudaStream_t s);
    void self_attention_bwd_gemm(const TapeRec& r, const void* dy, char* dqkv, bool plan, cudaStream_t s);
    void backward_ctx(const float* d_eps, int B, float* d_ctx, cudaStream_t user);

    void build(const etai_tensor* weights, int n_weights);
    void plan_workspace();
    void set_context(const void* ctx, int io_dtype, int B, cudaStream_t s);
    void forward(const void* latent, float t, int io_dtype, int B, const etai_attn_ctrl* ctrl, void* eps_out,
                 cudaStream_t s, bool train = false);
    void run_body(int io_dtype, int B, const etai_attn_ctrl* ctrl, cudaStream_t s);

    void* resnet(const void* x, int B, int H, int W, const Res& r, const etai_attn_ctrl* ctrl, bool inject_here,
                 cudaStream_t s) {
        long HW = (long)H * W, M = B * HW;
        void* a1 = gnorm(x, B, HW, r.n1, 1e-5f, true, s);
        // time-embedding bias: one fp32 vector for all rows (fused by the tcgen05 epilogue) or, with one timestep per row,
        // a [B, cout] table applied per image (rows_per_group = HW; the SIMT conv serves that case)
        void* h = t_rows_host
                      ? conv3x3(a1, B, H, W, r.c1, 1, tbuf_rows + TB_PROJ + r.temb_off, nullptr, s, nullptr, 1, HW, (long)tb_stride())
                      : conv3x3(a1, B, H, W, r.c1, 1, tbuf + TB_PROJ + r.temb_off, nullptr, s);
        void* a2 = gnorm(h, B, HW, r.n2, 1e-5f, true, s);
        const void* skip = x;
        if (r.has_sc) skip = linear(x, M, r.sc, nullptr, s);
        int inj = (ctrl && inject_here) ? ctrl->conv_inject_rows : 0;
        if (inj > 0) {
            // pnp_utils.py:172-177: conv2 output of rows [inj,3inj) := rows [0,inj) BEFORE the skip add
            ETAI_CHECK(B == 3 * inj, ETAI_ERR_ARG, "pnp injection expects B == 3*inject_rows");
            void* o = conv3x3(a2, B, H, W, r.c2, 1, nullptr, nullptr, s);
            if (arena.base) {
                cudaEvent_t e = prof_begin(s);
                copy_rows(o, HW * r.cout, 0, inj, inj, 2 * inj, dt, s);
                add_inplace(o, skip, M * r.cout, dt, s);
                prof_end(ETAI_PROF_OTHER, e, 2, s);
            }
            return o;
        }
        return conv3x3(a2, B, H, W, r.c2, 1, nullptr, skip, s);
    }
    void* transformer(const void* x, int B, int H, int W, const Tfm& t, const etai_attn_ctrl* ctrl, cudaStream_t s);
};

// -------------------------------------------------------------------------------------------------
void etai_unet::build(const etai_tensor* weights, int n_weights) {
    for (int i = 0; i < n_weights; ++i) table[weights[i].name] = &weights[i];
    const int* c = cfg.block_out_channels;
    const int temb = c[0] * 4, X = cfg.cross_dim;
    ETAI_CHECK(temb <= 4096, ETAI_ERR_ARG, "time embedding too wide");
    // staging: largest tensor is 3x3 conv 2*c3 -> c3 (or GEGLU 8C x C) in fp32, plus room for a 16-bit source copy
    size_t big = (size_t)c[3] * (2 * c[3]) * 9;
    if ((size_t)8 * c[3] * c[3] + 8 * c[3] > big) big = (size_t)8 * c[3] * c[3] + 8 * c[3];
    stage_elems = big;
    CUDA_CHECK(cudaMalloc(&stage, stage_elems * 4 + stage_elems * 2));

    // totals for the stacked matrices
    int sum_res = 0, sum_c = 0;
    {
        int cout = c[0];
        for (int i = 0; i < 4; ++i) { cout = c[i]; sum_res += 2 * cout; if (i < 3) sum_c += 2 * cout; }
        sum_res += 2 * c[3]; sum_c += c[3];
        for (int i = 0; i < 4; ++i) { int co = c[3 - i]; sum_res += 3 * co; if (i > 0) sum_c += 3 * co; }
    }
    temb_all.n = sum_res; temb_all.k = temb;
    temb_all.w = dmalloc((size_t)sum_res * temb * esz);
    temb_all.b = dmalloc((size_t)sum_res * esz);
    kv_all.n = 2 * sum_c; kv_all.k = X;
    kv_all.w = dmalloc((size_t)2 * sum_c * X * esz);

    conv_in = load_conv("conv_in", 4, c[0], tc ? 64 : 0, 0);
    time1 = load_lin("time_embedding.linear_1", temb, c[0], true);
    time2 = load_lin("time_embedding.linear_2", temb, temb, true);
    int cout = c[0];
    for (int i = 0; i < 4; ++i) {
        int cin = cout;
        cout = c[i];
        for (int j = 0; j < 2; ++j) {
            std::string p = "down_blocks." + std::to_string(i);
            down_res[i].push_back(load_res(p + ".resnets." + std::to_string(j), j == 0 ? cin : cout, cout, temb));
            if (i < 3) down_tf[i].push_back(load_tfm(p + ".attentions." + std::to_string(j), cout, 0));
        }
        if (i < 3) down_samp[i] = load_conv("down_blocks." + std::to_string(i) + ".downsamplers.0.conv", cout, cout);
    }
    mid_res[0] = load_res("mid_block.resnets.0", c[3], c[3], temb);
    mid_tf = load_tfm("mid_block.attentions.0", c[3], 1);
    mid_res[1] = load_res("mid_block.resnets.1", c[3], c[3], temb);
    int rev[4] = {c[3], c[2], c[1], c[0]};
    cout = rev[0];
    for (int i = 0; i < 4; ++i) {
        int prev = cout;
        cout = rev[i];
        int cin = rev[i + 1 < 3 ? i + 1 : 3];
        for (int j = 0; j < 3; ++j) {
            int skip = j == 2 ? cin : cout, rin = j == 0 ? prev : cout;
            std::string p = "up_blocks." + std::to_string(i);
            up_res[i].push_back(load_res(p + ".resnets." + std::to_string(j), rin + skip, cout, temb));
            if (i > 0) up_tf[i].push_back(load_tfm(p + ".attentions." + std::to_string(j), cout, 2));
        }
        if (i < 3) up_samp[i] = load_conv("up_blocks." + std::to_string(i) + ".upsamplers.0.conv", cout, cout);
    }
    norm_out = load_norm("conv_norm_out", c[0]);
    conv_out = load_conv("conv_out", c[0], 4, 0, tc ? 32 : 0);
    ETAI_CHECK(temb_total == sum_res && kv_total == 2 * sum_c && n_tf == 16, ETAI_ERR_ARG, "architecture bookkeeping mismatch");
    CUDA_CHECK(cudaFree(stage));
    stage = nullptr;
    table.clear();
}

void* etai_unet::transformer(const void* x, int B, int H, int W, const Tfm& t, const etai_attn_ctrl* ctrl,
                             cudaStream_t s) {
    const int C = t.C, heads = cfg.heads, d = C / heads, L = cfg.ctx_len;
    const long HW = (long)H * W, M = B * HW;
    const float scale = 1.0f / sqrtf((float)d);
    void* g = gnorm(x, B, HW, t.gn, 1e-6f, false, s);
    void* h = linear(g, M, t.pin, nullptr, s);
    // ---- self attention ----
    void* n1 = lnorm(h, M, t.ln1, s);
    void* qkv = linear(n1, M, t.qkv, nullptr, s);
    void* ao = arena.alloc((size_t)M * C * esz);
    if (arena.base) {
        SelfAttnArgs a;
        a.q = qkv; a.k = (char*)qkv + (size_t)C * esz; a.v = (char*)qkv + (size_t)2 * C * esz;
        a.out = ao;
        a.B = B; a.Nq = (int)HW; a.Nk = (int)HW; a.heads = heads; a.d = d;
        a.ldq = a.ldk = a.ldv = 3 * C; a.ldo = C; a.scale = scale; a.dtype = dt;
        bool remap = ctrl && (ctrl->flags & ETAI_CTRL_SELF_REMAP) && ((ctrl->self_layer_mask >> t.index) & 1u) &&
                     HW <= ctrl->self_max_tokens;
        for (int r = 0; r < B; ++r) {
            a.map.q[r] = remap ? ctrl->self_q_row[r] : r;
            a.map.k[r] = remap ? ctrl->self_k_row[r] : r;
            a.map.v[r] = remap ? ctrl->self_v_row[r] : r;
            ETAI_CHECK(a.map.q[r] >= 0 && a.map.q[r] < B && a.map.k[r] >= 0 && a.map.k[r] < B && a.map.v[r] >= 0 &&
                           a.map.v[r] < B, ETAI_ERR_ARG, "self remap row out of range");
        }
        cudaEvent_t e = prof_begin(s);
        if (tc && attention_tc_supported(a)) attention_tc(a, s);
        else attention_simt(a, s);
        prof_end(ETAI_PROF_SELF_ATTN, e, 1, s);
    }
    if (tape_on) {
        ETAI_CHECK(!ctrl, ETAI_ERR_UNSUPPORTED, "train-mode forward does not take an attention control");
        TapeRec r{T_SELF_ATTN, GemmArgs(), qkv, nullptr, ao};
        r.B = B; r.HW = HW; r.heads = heads; r.d = d; r.C = C; r.scale = scale;
        tape.push_back(r);
    }
    h = linear(ao, M, t.o1, h, s);
    // ---- cross attention ----
    void* n2 = lnorm(h, M, t.ln2, s);
    void* q2 = linear(n2, M, t.q2, nullptr, s);
    void* co = arena.alloc((size_t)M * C * esz);
    if (arena.base) {
        CrossAttnArgs a;
        memset(&a, 0, sizeof(a));
        a.q = q2; a.kv = kv_cache; a.out = co;
        a.B = B; a.N = (int)HW; a.L = L; a.heads = heads; a.d = d;
        a.ldq = C; a.ldkv = kv_total; a.ldo = C; a.koff = t.kv_off; a.voff = t.kv_off + C; a.scale = scale;
        a.dtype = dt;
        int slot_of[ETAI_MAX_ROWS];
        for (int r = 0; r < B; ++r) slot_of[r] = -1;
        if (ctrl && (ctrl->flags & ETAI_CTRL_CROSS_STORE) && HW == (long)ctrl->store_res * ctrl->store_res) {
            float* acc = t.place == 0 ? ctrl->store_down : t.place == 1 ? ctrl->store_mid : ctrl->store_up;
            if (acc) {
                a.store = acc;
                for (int i = 0; i < ctrl->n_store_rows; ++i) {
                    ETAI_CHECK(ctrl->store_row[i] >= 0 && ctrl->store_row[i] < B, ETAI_ERR_ARG, "store row out of range");
                    slot_of[ctrl->store_row[i]] = i;
                }
            }
        }
        bool used[ETAI_MAX_ROWS] = {false};
        int ng = 0;
        if (ctrl && (ctrl->flags & ETAI_CTRL_CROSS_EDIT)) {
            a.mapper = ctrl->mapper; a.blend_a = ctrl->blend_a; a.equalizer = ctrl->equalizer; a.alpha_step = ctrl->alpha_step;
            for (int p = 0; p < ctrl->n_pairs; ++p) {
                int br = ctrl->edit_base_row[p], tr = ctrl->edit_tgt_row[p];
                ETAI_CHECK(br >= 0 && br < B && tr >= 0 && tr < B && br != tr && !used[br] && !used[tr], ETAI_ERR_ARG,
                           "bad cross-edit pair");
                used[br] = used[tr] = true;
                a.groups[ng++] = CrossGroup{br, tr, p, slot_of[br], slot_of[tr]};
            }
        }
        for (int r = 0; r < B; ++r)
            if (!used[r]) a.groups[ng++] = CrossGroup{r, -1, 0, slot_of[r], -1};
        a.n_groups = ng;
        cudaEvent_t e = prof_begin(s);
        int nl = 1;
        a.map16 = (a.mapper && map16_ready) ? map16_buf : nullptr;
        a.store_part = store_part_buf;
        if (tc && cross_attention_tc_supported(a)) nl = cross_attention_tc(a, s);
        else cross_attention(a, s);
        prof_end(ETAI_PROF_CROSS_ATTN, e, nl, s);
    }
    if (tape_on) {
        TapeRec r{T_CROSS_ATTN, GemmArgs(), q2, nullptr, co};
        r.B = B; r.HW = HW; r.heads = heads; r.d = d; r.C = C; r.scale = scale; r.kv_off = t.kv_off;
        tape.push_back(r);
    }
    h = linear(co, M, t.o2, h, s);
    // ---- feed forward (GEGLU) ----
    void* n3 = lnorm(h, M, t.ln3, s);
    void* f = linear(n3, M, t.ff1, nullptr, s, /*geglu=*/1);
    h = linear(f, M, t.ff2, h, s);
    return linear(h, M, t.pout, x, s);
}

void etai_unet::set_context(const void* ctx, int io_dtype, int B, cudaStream_t user) {
    ETAI_CHECK(B >= 1 && B <= cfg.max_batch, ETAI_ERR_ARG, "set_context: batch out of range");
    long M = (long)B * cfg.ctx_len;
    cudaStream_t s = gs;
    CUDA_CHECK(cudaEventRecord(ev_in, user));
    CUDA_CHECK(cudaStreamWaitEvent(gs, ev_in, 0));
    convert(ctx, io_dtype, ctx_buf, dt, M * cfg.cross_dim, s);
    launches += 1;
    GemmArgs a;
    a.A = ctx_buf; a.W = kv_all.w; a.C = kv_cache;
    a.M = M; a.N = kv_all.n; a.K = kv_all.k; a.lda = kv_all.k; a.ldc = kv_all.n;
    gemm(a, s);
    ctx_rows = B;
    CUDA_CHECK(cudaEventRecord(ev_out, gs));
    CUDA_CHECK(cudaStreamWaitEvent(user, ev_out, 0));
}

// The kernel schedule of one forward.  Inputs/outputs are the engine-owned staging buffers; `ctrl` already points at
// engine-owned tables, so the same schedule can be captured into a CUDA graph and replayed.
void etai_unet::run_body(int io_dtype, int B, const etai_attn_ctrl* ctrl, cudaStream_t s) {
    const bool planning = arena.base == nullptr;
    const void* latent = in_stage;
    void* eps_out = out_stage;
    const int* c = cfg.block_out_channels;
    const int temb = c[0] * 4;
    int H = cfg.latent_hw, W = cfg.latent_hw;
    arena.reset();

    // time embedding (t is shared by all rows): all fp32, M = 1 skinny GEMMs
    if (!planning && t_rows_host) {  // one timestep per row: B independent M = 1 chains into the per-row scratch
        cudaEvent_t e = prof_begin(s);
        for (int b = 0; b < B; ++b) {
            float* tb = tbuf_rows + (size_t)b * tb_stride();
            timestep_sincos(t_dev + b, tb + TB_SIN, c[0], s);
            skinny_linear(tb + TB_SIN, time1.w, time1.b, tb + TB_H1, 1, temb, c[0], 1, dt, s);
            skinny_linear(tb + TB_H1, time2.w, time2.b, tb + TB_ST, 1, temb, temb, 1, dt, s);
            skinny_linear(tb + TB_ST, temb_all.w, temb_all.b, tb + TB_PROJ, 1, temb_all.n, temb, 0, dt, s);
        }
        prof_end(ETAI_PROF_OTHER, e, 4 * B, s);
    } else if (!planning) {
        cudaEvent_t e = prof_begin(s);
        timestep_sincos(t_dev, tbuf + TB_SIN, c[0], s);
        skinny_linear(tbuf + TB_SIN, time1.w, time1.b, tbuf + TB_H1, 1, temb, c[0], 1, dt, s);        // silu(linear_1)
        skinny_linear(tbuf + TB_H1, time2.w, time2.b, tbuf + TB_ST, 1, temb, temb, 1, dt, s);          // silu(temb)
        skinny_linear(tbuf + TB_ST, temb_all.w, temb_all.b, tbuf + TB_PROJ, 1, temb_all.n, temb, 0, dt, s);
        prof_end(ETAI_PROF_OTHER, e, 4, s);
    }

    if (!planning && ctrl && (ctrl->flags & ETAI_CTRL_CROSS_STORE)) {
        size_t n = (size_t)ctrl->n_store_rows * ctrl->store_res * ctrl->store_res * cfg.ctx_len * sizeof(float);
        float* acc[3] = {ctrl->store_down, ctrl->store_mid, ctrl->store_up};
        for (int i = 0; i < 3; ++i)
            if (acc[i]) CUDA_CHECK(cudaMemsetAsync(acc[i], 0, n, s));
    }
    map16_ready = false;
    if (!planning && tc && ctrl && (ctrl->flags & ETAI_CTRL_CROSS_EDIT)) {
        cudaEvent_t e = prof_begin(s);
        cross_attention_tc_prep_mapper(ctrl->mapper, map16_buf, ctrl->n_pairs, cfg.ctx_len, dt, s);
        prof_end(ETAI_PROF_OTHER, e, 1, s);
        map16_ready = true;
    }
    void* x = arena.alloc((size_t)B * H * W * conv_in.cin * esz);
    if (!planning) {
        cudaEvent_t e = prof_begin(s);
        nchw_to_nhwc(latent, io_dtype, x, dt, B, 4, conv_in.cin, (long)H * W, s);
        prof_end(ETAI_PROF_OTHER, e, 1, s);
    }
    void* h = conv3x3(x, B, H, W, conv_in, 1, nullptr, nullptr, s);

    struct Skip { void* p; int C; };
    std::vector<Skip> skips;
    skips.push_back({h, c[0]});
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 2; ++j) {
            h = resnet(h, B, H, W, down_res[i][j], ctrl, false, s);
            if (i < 3) h = transformer(h, B, H, W, down_tf[i][j], ctrl, s);
            skips.push_back({h, c[i]});
        }
        if (i < 3) {
            h = conv3x3(h, B, H, W, down_samp[i], 2, nullptr, nullptr, s);
            H /= 2; W /= 2;
            skips.push_back({h, c[i]});
        }
    }
    h = resnet(h, B, H, W, mid_res[0], ctrl, false, s);
    h = transformer(h, B, H, W, mid_tf, ctrl, s);
    h = resnet(h, B, H, W, mid_res[1], ctrl, false, s);
    int hc = c[3];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 3; ++j) {
            Skip sk = skips.back();
            skips.pop_back();
            long rows = (long)B * H * W;
            void* cat = arena.alloc((size_t)rows * (hc + sk.C) * esz);
            if (!planning) {
                cudaEvent_t e = prof_begin(s);
                concat_channels(h, hc, sk.p, sk.C, cat, rows, dt, s);
                prof_end(ETAI_PROF_OTHER, e, 1, s);
            }
            if (tape_on) {
                TapeRec r{T_CONCAT, GemmArgs(), h, sk.p, cat};
                r.M = rows; r.C = hc; r.C2 = sk.C;
                tape.push_back(r);
            }
            const Res& r = up_res[i][j];
            ETAI_CHECK(r.cin == hc + sk.C, ETAI_ERR_STATE, "skip bookkeeping mismatch");
            h = resnet(cat, B, H, W, r, ctrl, i == 1 && j == 1, s);
            hc = r.cout;
            if (i > 0) h = transformer(h, B, H, W, up_tf[i][j], ctrl, s);
        }
        if (i < 3) {
            void* up = arena.alloc((size_t)B * H * W * 4 * hc * esz);
            if (!planning) {
                cudaEvent_t e = prof_begin(s);
                upsample2x(h, up, B, H, W, hc, dt, s);
                prof_end(ETAI_PROF_OTHER, e, 1, s);
            }
            if (tape_on) {
                TapeRec r{T_UPSAMPLE, GemmArgs(), h, nullptr, up};
                r.B = B; r.H = H; r.W = W; r.C = hc;
                tape.push_back(r);
            }
            H *= 2; W *= 2;
            h = conv3x3(up, B, H, W, up_samp[i], 1, nullptr, nullptr, s);
        }
    }
    void* a = gnorm(h, B, (long)H * W, norm_out, 1e-5f, true, s);
    void* o = conv3x3(a, B, H, W, conv_out, 1, nullptr, nullptr, s);
    last_out = o;
    if (!planning) {
        cudaEvent_t e = prof_begin(s);
        nhwc_to_nchw(o, dt, eps_out, io_dtype, B, 4, conv_out.cout, (long)H * W, s);
        prof_end(ETAI_PROF_OTHER, e, 1, s);
    }
}

static std::string graph_key(int io_dtype, int B, const etai_attn_ctrl* c) {
    std::string k;
    auto put = [&](const void* p, size_t n) { k.append(reinterpret_cast<const char*>(p), n); };
    put(&io_dtype, 4); put(&B, 4);
    int flags = c ? c->flags : 0;
    put(&flags, 4);
    if (!c) return k;
    if (flags & ETAI_CTRL_SELF_REMAP) {
        put(c->self_q_row, 4 * B); put(c->self_k_row, 4 * B); put(c->self_v_row, 4 * B);
        put(&c->self_layer_mask, 4); put(&c->self_max_tokens, 4);
    }
    if (flags & ETAI_CTRL_CROSS_EDIT) {
        put(&c->n_pairs, 4); put(c->edit_base_row, 4 * c->n_pairs); put(c->edit_tgt_row, 4 * c->n_pairs);
    }
    if (flags & ETAI_CTRL_CROSS_STORE) {
        put(&c->store_res, 4); put(&c->n_store_rows, 4); put(c->store_row, 4 * c->n_store_rows);
        int places = (c->store_down ? 1 : 0) | (c->store_mid ? 2 : 0) | (c->store_up ? 4 : 0);
        put(&places, 4);
    }
    put(&c->conv_inject_rows, 4);
    return k;
}

void etai_unet::forward(const void* latent, float t, int io_dtype, int B, const etai_attn_ctrl* ctrl, void* eps_out,
                        cudaStream_t user, bool train) {
    bwd_batch = 0;  // any forward overwrites the activations a pending backward pass would read
    ETAI_CHECK(B >= 1 && B <= cfg.max_batch, ETAI_ERR_ARG, "forward: batch out of range");
    ETAI_CHECK(ctx_rows == B, ETAI_ERR_STATE, "forward: etai_unet_set_context must be called with the same batch first");
    const int L = cfg.ctx_len;
    const size_t io_bytes = (size_t)B * 4 * cfg.latent_hw * cfg.latent_hw * dtype_size(io_dtype);
    // ---- hand over from the caller's stream to the engine stream ----
    CUDA_CHECK(cudaEventRecord(ev_in, user));
    CUDA_CHECK(cudaStreamWaitEvent(gs, ev_in, 0));
    // ---- prologue: stage every per-call input at a stable address ----
    CUDA_CHECK(cudaMemcpyAsync(in_stage, latent, io_bytes, cudaMemcpyDeviceToDevice, gs));
    // pageable copy of a few bytes: staged by the driver before the call returns
    if (t_rows_host) CUDA_CHECK(cudaMemcpyAsync(t_dev, t_rows_host, (size_t)B * sizeof(float), cudaMemcpyHostToDevice, gs));
    else CUDA_CHECK(cudaMemcpyAsync(t_dev, &t, sizeof(float), cudaMemcpyHostToDevice, gs));
    etai_attn_ctrl ic;
    const etai_attn_ctrl* body_ctrl = nullptr;
    if (ctrl) {
        ic = *ctrl;
        if (ctrl->flags & ETAI_CTRL_CROSS_EDIT) {
            const int P = ctrl->n_pairs;
            CUDA_CHECK(cudaMemcpyAsync(c_mapper, ctrl->mapper, (size_t)P * L * L * 4, cudaMemcpyDeviceToDevice, gs));
            CUDA_CHECK(cudaMemcpyAsync(c_blend, ctrl->blend_a, (size_t)P * L * 4, cudaMemcpyDeviceToDevice, gs));
            CUDA_CHECK(cudaMemcpyAsync(c_eq, ctrl->equalizer, (size_t)P * L * 4, cudaMemcpyDeviceToDevice, gs));
            CUDA_CHECK(cudaMemcpyAsync(c_alpha, ctrl->alpha_step, (size_t)P * L * 4, cudaMemcpyDeviceToDevice, gs));
            ic.mapper = c_mapper; ic.blend_a = c_blend; ic.equalizer = c_eq; ic.alpha_step = c_alpha;
        }
        if (ctrl->flags & ETAI_CTRL_CROSS_STORE) {
            ETAI_CHECK(ctrl->store_res <= 32, ETAI_ERR_ARG, "attention store resolution must be <= 32");
            ic.store_down = ctrl->store_down ? c_store[0] : nullptr;
            ic.store_mid = ctrl->store_mid ? c_store[1] : nullptr;
            ic.store_up = ctrl->store_up ? c_store[2] : nullptr;
        }
        body_ctrl = &ic;
    }
    // ---- the schedule: eager the first time a key is seen, captured the second time, replayed afterwards ----
    bool done = false;
    if (train) {  // eager, recorded on the tape, GEGLU un-fused: the activations stay in the arena for backward_ctx
        ETAI_CHECK(bwd_enabled && B <= bwd_max_batch && !ctrl, ETAI_ERR_STATE,
                   "forward_train: call etai_unet_enable_backward first (batch <= its max_batch, no attention control)");
        tape.clear();
        tape_on = true;
        try {
            run_body(io_dtype, B, nullptr, gs);
        } catch (...) {
            tape_on = false;
            throw;
        }
        tape_on = false;
        bwd_batch = B;
        done = true;
    }
    if (!done && use_graphs && !prof_on) {
        GraphEntry& ge = graphs[graph_key(io_dtype, B, ctrl) + (t_rows_host ? "R" : "")];
        if (ge.exec) {
            CUDA_CHECK(cudaGraphLaunch(ge.exec, gs));
            launches += ge.launches;
            ++graph_replays;
            done = true;
        } else if (ge.seen >= 1) {
            int64_t l0 = launches;
            cudaGraph_t graph = nullptr;
            CUDA_CHECK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed));
            try {
                run_body(io_dtype, B, body_ctrl, gs);
            } catch (...) {
                cudaStreamEndCapture(gs, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            CUDA_CHECK(cudaStreamEndCapture(gs, &graph));
            ge.launches = launches - l0;
            CUDA_CHECK(cudaGraphInstantiate(&ge.exec, graph, 0));
            CUDA_CHECK(cudaGraphDestroy(graph));
            CUDA_CHECK(cudaGraphLaunch(ge.exec, gs));
            done = true;
        } else {
            ge.seen = 1;
        }
    }
    if (!done) run_body(io_dtype, B, body_ctrl, gs);
    // ---- epilogue: results back to the caller's buffers ----
    CUDA_CHECK(cudaMemcpyAsync(eps_out, out_stage, io_bytes, cudaMemcpyDeviceToDevice, gs));
    if (ctrl && (ctrl->flags & ETAI_CTRL_CROSS_STORE)) {
        long n = (long)ctrl->n_store_rows * ctrl->store_res * ctrl->store_res * L;
        float* user_acc[3] = {ctrl->store_down, ctrl->store_mid, ctrl->store_up};
        for (int i = 0; i < 3; ++i)
            if (user_acc[i]) { add_f32(user_acc[i], c_store[i], n, gs); launches += 1; }
    }
    CUDA_CHECK(cudaEventRecord(ev_out, gs));
    CUDA_CHECK(cudaStreamWaitEvent(user, ev_out, 0));
}

void etai_unet::plan_workspace() {
    // dry run with a null arena to size it for max_batch
    arena.base = nullptr;
    arena.cap = 0;
    arena.peak = 0;
    run_body(ETAI_F32, cfg.max_batch, nullptr, 0);
    size_t need = arena.peak + 4096;
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, need));
    arena.base = (char*)p;
    arena.cap = need;
    arena.reset();
    workspace_bytes = need;
    long ctx_m = (long)cfg.max_batch * cfg.ctx_len;
    CUDA_CHECK(cudaMalloc(&kv_cache, (size_t)ctx_m * kv_total * esz));
    CUDA_CHECK(cudaMalloc(&ctx_buf, (size_t)ctx_m * cfg.cross_dim * esz));
    CUDA_CHECK(cudaMalloc((void**)&tbuf, (size_t)(TB_PROJ + temb_total + 64) * sizeof(float)));
    {
        size_t io = (size_t)cfg.max_batch * 4 * cfg.latent_hw * cfg.latent_hw * 4;
        CUDA_CHECK(cudaMalloc((void**)&in_stage, io));
        CUDA_CHECK(cudaMalloc((void**)&out_stage, io));
        CUDA_CHECK(cudaMalloc((void**)&t_dev, 256));
        const int L = cfg.ctx_len;
        CUDA_CHECK(cudaMalloc((void**)&c_mapper, (size_t)ETAI_MAX_PAIRS * L * L * 4));
        CUDA_CHECK(cudaMalloc((void**)&c_blend, (size_t)ETAI_MAX_PAIRS * L * 4));
        CUDA_CHECK(cudaMalloc((void**)&c_eq, (size_t)ETAI_MAX_PAIRS * L * 4));
        CUDA_CHECK(cudaMalloc((void**)&c_alpha, (size_t)ETAI_MAX_PAIRS * L * 4));
        for (int i = 0; i < 3; ++i) CUDA_CHECK(cudaMalloc((void**)&c_store[i], (size_t)cfg.max_batch * 1024 * L * 4));
        CUDA_CHECK(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
        workspace_bytes += 2 * io + 3 * (size_t)cfg.max_batch * 1024 * L * 4;
    }
    size_t gws = groupnorm_workspace_bytes(cfg.max_batch, 0, 0, 32);
    CUDA_CHECK(cudaMalloc(&gn_ws, gws));
    CUDA_CHECK(cudaMemset(gn_ws, 0, gws));  // tickets start at zero and reset themselves
    // im2col scratch for the stride-2 convs on the tcgen05 path: largest is B*32*32 x 9*c0
    tc_ws_bytes = 0;
    if (tc) {
        const int* c = cfg.block_out_channels;
        long hw = (long)cfg.latent_hw * cfg.latent_hw / 4;
        size_t m = 0;
        for (int i = 0; i < 3; ++i) {
            size_t b = (size_t)cfg.max_batch * hw * 9 * c[i] * esz;
            if (b > m) m = b;
            hw /= 4;
        }
        // + split-K partials: up to 16 fp32 copies of the largest low-resolution output (M <= max_batch*256, N <= c3)
        size_t splitk = (size_t)16 * cfg.max_batch * 256 * c[3] * sizeof(float);
        tc_ws_bytes = ((m + 255) & ~size_t(255)) + splitk;
        CUDA_CHECK(cudaMalloc(&tc_ws, tc_ws_bytes));
        CUDA_CHECK(cudaMalloc(&map16_buf, cross_attention_tc_mapper_bytes(ETAI_MAX_PAIRS)));
        CUDA_CHECK(cudaMemset(map16_buf, 0, cross_attention_tc_mapper_bytes(ETAI_MAX_PAIRS)));
        size_t sp = (size_t)cfg.heads * cfg.max_batch * 1024 * cfg.ctx_len * sizeof(float);  // store maps up to 32x32
        CUDA_CHECK(cudaMalloc((void**)&store_part_buf, sp));
        workspace_bytes += sp;
    }
    workspace_bytes += (size_t)ctx_m * (kv_total + cfg.cross_dim) * esz + gws + tc_ws_bytes;
}


// -------------------------------------------------------------------------------------------------
// backward w.r.t. the text context (null-text inversion: modules/inversion/null_text_inversion.py:62-80)
// -------------------------------------------------------------------------------------------------
void* etai_unet::make_wT(const void* w, int n, int k, bool conv) {
    auto it = wT.find(w);
    if (it != wT.end()) return it->second;
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, (conv ? (size_t)n * 9 * k : (size_t)n * k) * esz));
    bwd_owned.push_back(p);
    if (conv) conv_weight_flip(w, p, n, k, dt, gs);
    else transpose_2d(w, p, n, k, dt, gs);
    launches += 1;
    wT[w] = p;
    return p;
}

// Reverse walk over the tape of the last train-mode forward.  grad[] maps an activation to the buffer holding dL/d(it);
// only activations that depend on the context (everything downstream of the first cross-attention) get one.  The seed
// (dL/d(conv_out result)) is taken from `seed`; the result is left in dkv (per-layer dK | dV columns).
void etai_unet::backward_walk(int B, cudaStream_t s) {
    const bool plan = garena.base == nullptr;
    const int L = cfg.ctx_len;
    std::unordered_set<const void*> dep;
    for (const TapeRec& r : tape) {
        bool d = r.kind == T_CROSS_ATTN || dep.count(r.x) || (r.x2 && dep.count(r.x2)) ||
                 (r.kind == T_GEMM && r.g.residual && dep.count(r.g.residual));
        if (d) dep.insert(r.y);
    }
    std::unordered_map<const void*, void*> grad;
    auto add_grad = [&](const void* t, void* g, long n) {
        auto it = grad.find(t);
        if (it == grad.end()) { grad[t] = g; return; }
        if (!plan) { add_inplace(it->second, g, n, dt, s); launches += 1; }
    };
    ETAI_CHECK(last_out && dep.count(last_out), ETAI_ERR_STATE, "backward: the recorded forward does not reach the context");
    grad[last_out] = bwd_seed;
    for (size_t ii = tape.size(); ii-- > 0;) {
        const TapeRec& r = tape[ii];
        if (!dep.count(r.y)) continue;
        auto gi = grad.find(r.y);
        if (gi == grad.end()) continue;
        void* dy = gi->second;
        switch (r.kind) {
        case T_GEMM: {
            const GemmArgs& g = r.g;
            ETAI_CHECK(!g.geglu && g.ldc == g.N, ETAI_ERR_STATE, "backward: unexpected GEMM layout on the tape");
            if (g.residual && dep.count(g.residual)) add_grad(g.residual, dy, g.M * g.N);
            if (!dep.count(g.A)) break;
            if (!g.conv) {
                ETAI_CHECK(g.lda == g.K, ETAI_ERR_STATE, "backward: strided GEMM input on the tape");
                void* dx = garena.alloc((size_t)g.M * g.K * esz);
                if (!plan) {
                    GemmArgs b;  // dx[M,K] = dy[M,N] * W[N,K]
                    b.A = dy; b.W = make_wT(g.W, g.N, g.K, false); b.C = dx;
                    b.M = g.M; b.N = g.K; b.K = g.N; b.lda = g.N; b.ldc = g.K; b.ldr = g.K;
                    gemm(b, s);
                }
                add_grad(g.A, dx, g.M * g.K);
            } else {
                const int Cout = g.N, Cin = g.Cin, H = g.H, W = g.Wd;
                const void* img = dy;
                if (g.stride == 2) {  // adjoint of the stride-2 conv: zero-stuff dy to the input resolution, then a stride-1 conv
                    void* z = garena.alloc((size_t)g.B * H * W * Cout * esz);
                    if (!plan) { zero_stuff2x(dy, z, g.B, g.Ho, g.Wo, Cout, dt, s); launches += 1; }
                    img = z;
                }
                const long Mi = (long)g.B * H * W;
                void* dx = garena.alloc((size_t)Mi * Cin * esz);
                if (!plan) {
                    GemmArgs b;  // dx = conv3x3(dy, filter mirrored and transposed)
                    b.A = img; b.W = make_wT(g.W, Cout, Cin, true); b.C = dx;
                    b.M = Mi; b.N = Cin; b.K = 9 * Cout; b.ldc = Cin; b.ldr = Cin;
                    b.conv = 1; b.B = g.B; b.H = H; b.Wd = W; b.Cin = Cout; b.stride = 1; b.Ho = H; b.Wo = W; b.pad = 1;
                    gemm(b, s);
                }
                add_grad(g.A, dx, Mi * Cin);
            }
            break;
        }
        case T_GN: {
            const long n = (long)r.B * r.HW * r.n.c;
            void* dx = garena.alloc((size_t)n * esz);
            void* gws = garena.alloc(groupnorm_bwd_workspace_bytes(r.B, 32));
            if (!plan) { groupnorm_bwd(r.x, dy, r.n.g, r.n.b, dx, r.B, r.HW, r.n.c, 32, r.eps, r.silu, dt, gws, s); launches += 3; }
            add_grad(r.x, dx, n);
            break;
        }
        case T_LN: {
            const long n = r.M * r.n.c;
            void* dx = garena.alloc((size_t)n * esz);
            if (!plan) { layernorm_bwd(r.x, dy, r.n.g, dx, r.M, r.n.c, r.eps, dt, s); launches += 1; }
            add_grad(r.x, dx, n);
            break;
        }
        case T_SELF_ATTN: {
            const int C = r.C;
            const long M = (long)r.B * r.HW;
            char* dqkv = (char*)garena.alloc((size_t)M * 3 * C * esz);
            static const bool simt_only = [] { const char* e = getenv("ETAI_ATTN_BWD_SIMT"); return e && e[0] == '1'; }();
            if (tc && r.HW >= 256 && !simt_only) {
                const size_t mark = garena.off;
                self_attention_bwd_gemm(r, dy, dqkv, plan, s);
                garena.off = mark;  // the per-layer scratch is dead once the layer's kernels are enqueued (stream order)
            } else if (!plan) {
                SelfAttnBwdArgs a;
                const char* qkv = (const char*)r.x;
                a.q = qkv; a.k = qkv + (size_t)C * esz; a.v = qkv + (size_t)2 * C * esz; a.o = r.y; a.dout = dy;
                a.dq = dqkv; a.dk = dqkv + (size_t)C * esz; a.dv = dqkv + (size_t)2 * C * esz;
                a.lse = lse_buf; a.dsum = lse_buf + (size_t)bwd_max_batch * r.heads * cfg.latent_hw * cfg.latent_hw;
                a.B = r.B; a.N = (int)r.HW; a.heads = r.heads; a.d = r.d;
                a.ldq = a.ldk = a.ldv = 3 * C; a.ldo = C; a.lddo = C; a.lddq = a.lddk = a.lddv = 3 * C;
                a.scale = r.scale; a.dtype = dt;
                attention_bwd(a, s);
                launches += 2;
            }
            add_grad(r.x, dqkv, M * 3 * C);
            break;
        }
        case T_CROSS_ATTN: {
            const int C = r.C;
            const long M = (long)r.B * r.HW;
            const bool want_dq = dep.count(r.x) != 0;
            void* dq = want_dq ? garena.alloc((size_t)M * C * esz) : nullptr;
            size_t pb = cross_attention_bwd_partial_bytes(r.B, (int)r.HW, L, C, r.d);
            if (pb > cross_part_need) cross_part_need = pb;
            if (!plan) {
                CrossAttnBwdArgs a;
                a.q = r.x; a.kv = kv_cache; a.dout = dy; a.dq = dq; a.part = cross_part; a.dkv = dkv;
                a.B = r.B; a.N = (int)r.HW; a.L = L; a.heads = r.heads; a.d = r.d;
                a.ldq = C; a.ldkv = kv_total; a.lddo = C; a.lddq = C; a.ld_dkv = kv_total;
                a.koff = r.kv_off; a.voff = r.kv_off + C; a.kv_off = r.kv_off; a.scale = r.scale; a.dtype = dt;
                cross_attention_bwd(a, s);
                launches += 2;
            }
            if (want_dq) add_grad(r.x, dq, M * C);
            break;
        }
        case T_GEGLU: {
            const long n = r.M * 2 * r.C;
            void* du = garena.alloc((size_t)n * esz);
            if (!plan) { geglu_bwd(r.x, dy, du, r.M, r.C, dt, s); launches += 1; }
            add_grad(r.x, du, n);
            break;
        }
        case T_CONCAT: {
            const void* in[2] = {r.x, r.x2};
            const int cw[2] = {r.C, r.C2}, off[2] = {0, r.C};
            for (int k = 0; k < 2; ++k) {
                if (!dep.count(in[k])) continue;
                auto it = grad.find(in[k]);
                const bool acc = it != grad.end();
                void* dst = acc ? it->second : garena.alloc((size_t)r.M * cw[k] * esz);
                if (!plan) { slice_cols(dy, r.C + r.C2, off[k], cw[k], dst, acc, r.M, dt, s); launches += 1; }
                if (!acc) grad[in[k]] = dst;
            }
            break;
        }
        case T_UPSAMPLE: {
            const long n = (long)r.B * r.H * r.W * r.C;
            void* din = garena.alloc((size_t)n * esz);
            if (!plan) { upsample2x_bwd(dy, din, r.B, r.H, r.W, r.C, dt, s); launches += 1; }
            add_grad(r.x, din, n);
            break;
        }
        }
    }
    // d(ctx) = d(K|V of all layers)[B*L, kv_total] * W_kv[kv_total, cross_dim]
    const long Mc = (long)B * L;
    void* dkv_t = dkv;
    if (dt != ETAI_F32) {
        dkv_t = garena.alloc((size_t)Mc * kv_total * esz);
        if (!plan) { convert(dkv, ETAI_F32, dkv_t, dt, Mc * kv_total, s); launches += 1; }
    }
    bwd_dctx = garena.alloc((size_t)Mc * cfg.cross_dim * esz);
    if (!plan) {
        GemmArgs b;
        b.A = dkv_t; b.W = make_wT(kv_all.w, kv_all.n, kv_all.k, false); b.C = bwd_dctx;
        b.M = Mc; b.N = cfg.cross_dim; b.K = kv_total; b.lda = kv_total; b.ldc = cfg.cross_dim; b.ldr = cfg.cross_dim;
        gemm(b, s);
    }
}


// Self-attention backward of one layer with every N x N product on the tensor-core GEMM kernel (16-bit engines, N >= 256;
// the SIMT flash-style kernels of backward.cu stay the fp32 parity path and serve the short sequences).  Per head, with
// Qp / Kp / Vp / dOp the head's [N, DP] zero-padded slices and XpT their transposes:
//   P    = softmax(scale Qp Kp^T) (+ lse)     dP   = dOp Vp^T       dS   = P   o (dP   - D[row]) scale      dQ = dS   Kp
//   P^T  = exp(scale Kp Qp^T - lse[col])      dP^T = Vp dOp^T       dS^T = P^T o (dP^T - D[col]) scale      dK = dS^T Qp
//                                                                                                          dV = P^T  dOp
// Seven GEMMs and four elementwise passes over two [N, N] 16-bit buffers; nothing is accumulated across launches.
void etai_unet::self_attention_bwd_gemm(const TapeRec& r, const void* dy, char* dqkv, bool plan, cudaStream_t s) {
    const int C = r.C, heads = r.heads, d = r.d, N = (int)r.HW;
    const int DP = (d + 63) / 64 * 64;
    const size_t hp = (size_t)heads * N * DP * esz;
    char* qp = (char*)garena.alloc(hp);  char* qpT = (char*)garena.alloc(hp);
    char* kp = (char*)garena.alloc(hp);  char* kpT = (char*)garena.alloc(hp);
    char* vp = (char*)garena.alloc(hp);
    char* dop = (char*)garena.alloc(hp); char* dopT = (char*)garena.alloc(hp);
    char* dqp = (char*)garena.alloc(hp); char* dkp = (char*)garena.alloc(hp); char* dvp = (char*)garena.alloc(hp);
    char* b1 = (char*)garena.alloc((size_t)N * N * esz);
    char* b2 = (char*)garena.alloc((size_t)N * N * esz);
    float* lse = (float*)garena.alloc((size_t)N * sizeof(float));
    float* dsum = (float*)garena.alloc((size_t)heads * N * sizeof(float));
    if (plan) return;
    auto mm = [&](const void* A, long lda, const void* W, void* Cc, long M, int Nn, int K, long ldc) {
        GemmArgs g;
        g.A = A; g.W = W; g.C = Cc; g.M = M; g.N = Nn; g.K = K; g.lda = lda; g.ldc = ldc; g.ldr = ldc;
        gemm(g, s);
    };
    for (int b = 0; b < r.B; ++b) {
        const char* qkv = (const char*)r.x + (size_t)b * N * 3 * C * esz;
        const char* o = (const char*)r.y + (size_t)b * N * C * esz;
        const char* dO = (const char*)dy + (size_t)b * N * C * esz;
        attn_pack_heads(qkv, 3 * C, N, heads, d, DP, qp, qpT, dt, s);
        attn_pack_heads(qkv + (size_t)C * esz, 3 * C, N, heads, d, DP, kp, kpT, dt, s);
        attn_pack_heads(qkv + (size_t)2 * C * esz, 3 * C, N, heads, d, DP, vp, nullptr, dt, s);
        attn_pack_heads(dO, C, N, heads, d, DP, dop, dopT, dt, s);
        attn_rowdot(dO, C, o, C, N, heads, d, dsum, dt, s);
        launches += 5;
        for (int h = 0; h < heads; ++h) {
            const size_t ho = (size_t)h * N * DP * esz;  // same offset for [N,DP] and [DP,N] blocks
            const float* D = dsum + (size_t)h * N;
            mm(qp + ho, DP, kp + ho, b1, N, N, DP, N);                                   // S
            softmax_rows(b1, N, N, r.scale, dt, s, lse);                                 // P, lse
            mm(dop + ho, DP, vp + ho, b2, N, N, DP, N);                                  // dP
            attn_bwd_elementwise(b2, b1, D, N, N, r.scale, 0, dt, s);                    // dS
            mm(b2, N, kpT + ho, dqp + ho, N, DP, N, DP);                                 // dQ = dS K
            mm(kp + ho, DP, qp + ho, b1, N, N, DP, N);                                   // S^T
            attn_bwd_elementwise(b1, nullptr, lse, N, N, r.scale, 1, dt, s);             // P^T
            mm(b1, N, dopT + ho, dvp + ho, N, DP, N, DP);                                // dV = P^T dO
            mm(vp + ho, DP, dop + ho, b2, N, N, DP, N);                                  // dP^T
            attn_bwd_elementwise(b2, b1, D, N, N, r.scale, 2, dt, s);                    // dS^T
            mm(b2, N, qpT + ho, dkp + ho, N, DP, N, DP);                                 // dK = dS^T Q
            launches += 4;
        }
        char* dst = dqkv + (size_t)b * N * 3 * C * esz;
        attn_unpack_heads(dqp, N, heads, d, DP, dst, 3 * C, dt, s);
        attn_unpack_heads(dkp, N, heads, d, DP, dst + (size_t)C * esz, 3 * C, dt, s);
        attn_unpack_heads(dvp, N, heads, d, DP, dst + (size_t)2 * C * esz, 3 * C, dt, s);
        launches += 3;
    }
}

void etai_unet::enable_backward(int max_batch) {
    if (bwd_enabled && max_batch <= bwd_max_batch) return;
    ETAI_CHECK(!bwd_enabled, ETAI_ERR_STATE, "enable_backward: already enabled with a smaller max_batch");
    ETAI_CHECK(max_batch >= 1 && max_batch <= cfg.max_batch, ETAI_ERR_ARG, "enable_backward: max_batch out of range");
    CUDA_CHECK(cudaStreamSynchronize(gs));
    bwd_max_batch = max_batch;
    // dry run: a train-mode forward + its backward walk with null arenas give the gradient arena size
    char* base = arena.base;
    const size_t peak = arena.peak;
    arena.base = nullptr;
    arena.peak = 0;
    tape.clear();
    tape_on = true;
    run_body(ETAI_F32, max_batch, nullptr, 0);
    tape_on = false;
    const size_t train_peak = arena.peak;
    arena.base = base;
    arena.peak = peak > train_peak ? peak : train_peak;
    ETAI_CHECK(train_peak + 4096 <= arena.cap, ETAI_ERR_NOMEM,
               "enable_backward: the activation arena is too small for a train-mode forward of this batch (create the engine "
               "with a larger max_batch)");
    garena.base = nullptr; garena.cap = 0; garena.peak = 0; garena.reset();
    const long hw = (long)cfg.latent_hw * cfg.latent_hw;
    garena.alloc((size_t)max_batch * 4 * hw * sizeof(float));              // scaled copy of d_eps
    bwd_seed = garena.alloc((size_t)max_batch * hw * conv_out.cout * esz);
    cross_part_need = 0;
    backward_walk(max_batch, 0);
    tape.clear();
    size_t need = garena.peak + 4096;
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, need));
    garena.base = (char*)p; garena.cap = need; garena.reset();
    CUDA_CHECK(cudaMalloc((void**)&dkv, (size_t)max_batch * cfg.ctx_len * kv_total * sizeof(float)));
    CUDA_CHECK(cudaMalloc((void**)&cross_part, cross_part_need ? cross_part_need : 256));
    CUDA_CHECK(cudaMalloc((void**)&lse_buf, (size_t)2 * max_batch * cfg.heads * hw * sizeof(float)));
    CUDA_CHECK(cudaMalloc((void**)&scale_dev, 256));
    workspace_bytes += need + (size_t)max_batch * cfg.ctx_len * kv_total * 4 + cross_part_need + 2 * max_batch * cfg.heads * hw * 4;
    bwd_enabled = true;
}

void etai_unet::backward_ctx(const float* d_eps, int B, float* d_ctx, cudaStream_t user) {
    ETAI_CHECK(bwd_enabled && bwd_batch == B && B >= 1, ETAI_ERR_STATE,
               "backward_ctx: needs the train-mode forward of the same batch immediately before");
    const long hw = (long)cfg.latent_hw * cfg.latent_hw, ne = (long)B * 4 * hw;
    CUDA_CHECK(cudaEventRecord(ev_in, user));
    CUDA_CHECK(cudaStreamWaitEvent(gs, ev_in, 0));
    cudaStream_t s = gs;
    garena.reset();
    const bool scaled = dt != ETAI_F32;  // 16-bit gradients: normalise the seed to max|.| = 16, undo at the end (the pass is linear)
    float* de = (float*)garena.alloc((size_t)bwd_max_batch * 4 * hw * sizeof(float));
    bwd_seed = garena.alloc((size_t)bwd_max_batch * hw * conv_out.cout * esz);
    const float* src = d_eps;
    if (scaled) {
        absmax_scale(d_eps, ne, 16.0f, scale_dev, s);
        scale_by_device_scalar(d_eps, de, scale_dev, false, ne, ETAI_F32, s);
        launches += 2;
        src = de;
    }
    nchw_to_nhwc(src, ETAI_F32, bwd_seed, dt, B, 4, conv_out.cout, hw, s);
    launches += 1;
    backward_walk(B, s);
    const long nc = (long)B * cfg.ctx_len * cfg.cross_dim;
    convert(bwd_dctx, dt, d_ctx, ETAI_F32, nc, s);
    launches += 1;
    if (scaled) { scale_by_device_scalar(d_ctx, d_ctx, scale_dev, true, nc, ETAI_F32, s); launches += 1; }
    CUDA_CHECK(cudaEventRecord(ev_out, gs));
    CUDA_CHECK(cudaStreamWaitEvent(user, ev_out, 0));
}

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
namespace etai {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
}  // namespace etai

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

int etai_abi_version(void) { return ETAI_ABI_VERSION; }
const char* etai_last_error(void) { return etai::g_last_error.c_str(); }

int etai_unet_create(etai_unet** out, const etai_unet_cfg* cfg, const etai_tensor* weights, int32_t n_weights,
                     int32_t device) {
    ETAI_API_BEGIN
    ETAI_CHECK(out && cfg && weights && n_weights > 0, ETAI_ERR_ARG, "create: null argument");
    ETAI_CHECK(cfg->dtype == ETAI_F32 || cfg->dtype == ETAI_F16 || cfg->dtype == ETAI_BF16, ETAI_ERR_ARG, "create: dtype");
    ETAI_CHECK(cfg->max_batch >= 1 && cfg->max_batch <= ETAI_MAX_ROWS, ETAI_ERR_ARG, "create: max_batch in [1,64]");
    ETAI_CHECK(cfg->heads >= 1 && cfg->ctx_len >= 1 && cfg->ctx_len <= 80, ETAI_ERR_ARG, "create: heads/ctx_len");
    ETAI_CHECK(cfg->latent_hw >= 8 && cfg->latent_hw % 8 == 0, ETAI_ERR_ARG, "create: latent_hw must be a multiple of 8");
    for (int i = 0; i < 4; ++i)
        ETAI_CHECK(cfg->block_out_channels[i] % (8 * cfg->heads) == 0 && cfg->block_out_channels[i] % 32 == 0,
                   ETAI_ERR_ARG, "create: channels must be multiples of 32 and of 8*heads");
    int ndev = 0;
    CUDA_CHECK(cudaGetDeviceCount(&ndev));
    ETAI_CHECK(device >= 0 && device < ndev, ETAI_ERR_ARG, "create: no such CUDA device");
    CUDA_CHECK(cudaSetDevice(device));
    etai_unet* h = new etai_unet();
    try {
        h->cfg = *cfg;
        h->device = device;
        h->wstore = std::make_shared<WeightStore>();
        h->wstore->device = device;
        h->dt = cfg->dtype;
        h->esz = dtype_size(cfg->dtype);
        h->tc = cfg->dtype != ETAI_F32 && cfg->math_mode == ETAI_MATH_AUTO;
        if (const char* e = getenv("ETAI_NO_GRAPHS")) h->use_graphs = !(e[0] == '1');
        h->build(weights, n_weights);
        h->plan_workspace();
    } catch (...) {
        etai_unet_destroy(h);
        throw;
    }
    *out = h;
    ETAI_API_END
}

int etai_unet_destroy(etai_unet* h) {
    if (!h) return ETAI_OK;
    cudaSetDevice(h->device);
    if (h->gs) cudaStreamSynchronize(h->gs);
    h->wstore.reset();  // frees the packed weights when this was the last handle using them
    if (h->arena.base) cudaFree(h->arena.base);
    if (h->kv_cache) cudaFree(h->kv_cache);
    if (h->ctx_buf) cudaFree(h->ctx_buf);
    if (h->tbuf) cudaFree(h->tbuf);
    if (h->tbuf_rows) cudaFree(h->tbuf_rows);
    if (h->gn_ws) cudaFree(h->gn_ws);
    for (auto& kv : h->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (h->gs) cudaStreamSynchronize(h->gs);
    void* extra[] = {h->in_stage, h->out_stage, h->t_dev, h->c_mapper, h->c_blend, h->c_eq, h->c_alpha, h->c_store[0],
                     h->c_store[1], h->c_store[2]};
    for (void* p : extra)
        if (p) cudaFree(p);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->gs) cudaStreamDestroy(h->gs);
    if (h->tc_ws) cudaFree(h->tc_ws);
    if (h->map16_buf) cudaFree(h->map16_buf);
    if (h->store_part_buf) cudaFree(h->store_part_buf);
    if (h->stage) cudaFree(h->stage);
    for (void* p : h->bwd_owned) cudaFree(p);
    void* bw[] = {h->garena.base, h->dkv, h->cross_part, h->lse_buf, h->scale_dev};
    for (void* p : bw)
        if (p) cudaFree(p);
    delete h;
    return ETAI_OK;
}

int etai_unet_clone(etai_unet** out, const etai_unet* src, int32_t max_batch) {
    ETAI_API_BEGIN
    ETAI_CHECK(out && src, ETAI_ERR_ARG, "clone: null argument");
    ETAI_CHECK(max_batch <= ETAI_MAX_ROWS, ETAI_ERR_ARG, "clone: max_batch in [1,64]");
    CUDA_CHECK(cudaSetDevice(src->device));
    etai_unet* h = new etai_unet();
    try {
        h->cfg = src->cfg;
        if (max_batch > 0) h->cfg.max_batch = max_batch;
        h->device = src->device;
        h->dt = src->dt; h->tc = src->tc; h->esz = src->esz;
        h->use_graphs = src->use_graphs;
        h->wstore = src->wstore;  // shared, read-only
        h->weight_bytes = 0;      // accounted once, by the handle that loaded them
        h->conv_in = src->conv_in; h->conv_out = src->conv_out; h->norm_out = src->norm_out;
        h->time1 = src->time1; h->time2 = src->time2; h->temb_all = src->temb_all; h->kv_all = src->kv_all;
        for (int i = 0; i < 4; ++i) {
            h->down_res[i] = src->down_res[i]; h->up_res[i] = src->up_res[i];
            h->down_tf[i] = src->down_tf[i]; h->up_tf[i] = src->up_tf[i];
        }
        for (int i = 0; i < 3; ++i) { h->down_samp[i] = src->down_samp[i]; h->up_samp[i] = src->up_samp[i]; }
        h->mid_res[0] = src->mid_res[0]; h->mid_res[1] = src->mid_res[1];
        h->mid_tf = src->mid_tf;
        h->n_tf = src->n_tf; h->temb_total = src->temb_total; h->kv_total = src->kv_total;
        h->plan_workspace();
    } catch (...) {
        etai_unet_destroy(h);
        throw;
    }
    *out = h;
    ETAI_API_END
}

int64_t etai_unet_launch_count(const etai_unet* h) { return h ? h->launches : 0; }

int etai_unet_profile(etai_unet* h, int32_t enable, float* ms_out, int32_t* launches_out) {
    ETAI_API_BEGIN
    ETAI_CHECK(h, ETAI_ERR_ARG, "profile: null handle");
    CUDA_CHECK(cudaSetDevice(h->device));
    float ms[ETAI_PROF_NCAT] = {0};
    int32_t cnt[ETAI_PROF_NCAT] = {0};
    for (auto& r : h->prof_recs) {
        CUDA_CHECK(cudaEventSynchronize(r.b));
        float t = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cat] += t;
        cnt[r.cat] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->prof_recs.clear();
    if (ms_out) for (int i = 0; i < ETAI_PROF_NCAT; ++i) ms_out[i] = ms[i];
    if (launches_out) for (int i = 0; i < ETAI_PROF_NCAT; ++i) launches_out[i] = cnt[i];
    h->prof_on = enable != 0;
    ETAI_API_END
}

int64_t etai_unet_device_bytes(const etai_unet* h) { return h ? (int64_t)(h->weight_bytes + h->workspace_bytes) : 0; }

int etai_unet_set_context(etai_unet* h, const void* ctx, int32_t io_dtype, int32_t B, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && ctx, ETAI_ERR_ARG, "set_context: null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->set_context(ctx, io_dtype, B, (cudaStream_t)stream);
    ETAI_API_END
}

static void check_ctrl(const etai_attn_ctrl* ctrl) {
    if (!ctrl) return;
    if (ctrl->flags & ETAI_CTRL_SELF_REMAP)
        ETAI_CHECK(ctrl->self_q_row && ctrl->self_k_row && ctrl->self_v_row, ETAI_ERR_ARG, "ctrl: self remap arrays");
    if (ctrl->flags & ETAI_CTRL_CROSS_EDIT)
        ETAI_CHECK(ctrl->n_pairs >= 1 && ctrl->n_pairs <= ETAI_MAX_PAIRS && ctrl->edit_base_row && ctrl->edit_tgt_row &&
                       ctrl->mapper && ctrl->blend_a && ctrl->equalizer && ctrl->alpha_step,
                   ETAI_ERR_ARG, "ctrl: cross edit tables");
    if (ctrl->flags & ETAI_CTRL_CROSS_STORE)
        ETAI_CHECK(ctrl->store_res > 0 && ctrl->n_store_rows >= 1 && ctrl->n_store_rows <= ETAI_MAX_ROWS && ctrl->store_row,
                   ETAI_ERR_ARG, "ctrl: store rows");
}

int etai_unet_forward(etai_unet* h, const void* latent, float t, int32_t io_dtype, int32_t B,
                      const etai_attn_ctrl* ctrl, void* eps_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && latent && eps_out, ETAI_ERR_ARG, "forward: null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    check_ctrl(ctrl);
    h->forward(latent, t, io_dtype, B, ctrl, eps_out, (cudaStream_t)stream);
    ETAI_API_END
}

int etai_unet_forward_rows(etai_unet* h, const void* latent, const float* t_rows, int32_t io_dtype, int32_t B,
                           const etai_attn_ctrl* ctrl, void* eps_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && latent && eps_out && t_rows, ETAI_ERR_ARG, "forward_rows: null argument");
    ETAI_CHECK(B >= 1 && B <= h->cfg.max_batch && B <= 64, ETAI_ERR_ARG, "forward_rows: batch out of range");
    CUDA_CHECK(cudaSetDevice(h->device));
    check_ctrl(ctrl);
    bool same = true;
    for (int b = 1; b < B; ++b) same = same && t_rows[b] == t_rows[0];
    if (same) {  // the usual case (every loop of the reference): the shared-bias schedule and its CUDA graphs
        h->forward(latent, t_rows[0], io_dtype, B, ctrl, eps_out, (cudaStream_t)stream);
    } else {
        if (!h->tbuf_rows)
            CUDA_CHECK(cudaMalloc((void**)&h->tbuf_rows, (size_t)h->cfg.max_batch * h->tb_stride() * sizeof(float)));
        h->t_rows_host = t_rows;
        try {
            h->forward(latent, t_rows[0], io_dtype, B, ctrl, eps_out, (cudaStream_t)stream);
        } catch (...) {
            h->t_rows_host = nullptr;
            throw;
        }
        h->t_rows_host = nullptr;
    }
    ETAI_API_END
}

/* ---- null-text inversion support ---- */
int etai_unet_enable_backward(etai_unet* h, int32_t max_batch) {
    ETAI_API_BEGIN
    ETAI_CHECK(h, ETAI_ERR_ARG, "enable_backward: null handle");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->enable_backward(max_batch);
    ETAI_API_END
}

int etai_unet_forward_train(etai_unet* h, const void* latent, float t, int32_t io_dtype, int32_t B, void* eps_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && latent && eps_out, ETAI_ERR_ARG, "forward_train: null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->forward(latent, t, io_dtype, B, nullptr, eps_out, (cudaStream_t)stream, /*train=*/true);
    ETAI_API_END
}

int etai_unet_backward_ctx(etai_unet* h, const float* d_eps, int32_t B, float* d_ctx_out, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && d_eps && d_ctx_out, ETAI_ERR_ARG, "backward_ctx: null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->backward_ctx(d_eps, B, d_ctx_out, (cudaStream_t)stream);
    ETAI_API_END
}

}  // extern "C"
