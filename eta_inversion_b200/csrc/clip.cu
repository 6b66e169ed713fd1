// CLIP ViT-L/14 text tower (transformers `CLIPTextModel` layout) on the engine's own kernels.
//
// Replaces `model.text_encoder(input_ids)[0]` of the reference (modules/inversion/diffusion_inversion.py:210-247: four
// un-batched passes per edit -- "", source prompt, "", target prompt).  12 pre-LN layers over 77 tokens: LayerNorm ->
// fused QKV GEMM -> causal attention (12 heads of 64, one CTA per (row, head)) -> out projection (+residual) ->
// LayerNorm -> fc1 -> quick-GELU -> fc2 (+residual); final LayerNorm.  Any number of prompts up to max_batch go through
// in ONE pass (the lock-step groups batch their prompts), and the ~100 launches of a pass are replayed as one CUDA graph
// per batch size.
#include <mutex>
#include <unordered_map>
#include "engine_base.cuh"

using namespace etai;

namespace {
struct Layer { Norm ln1, ln2; Lin qkv, o, fc1, fc2; };
}

struct etai_clip : etai::OpCtx {
    etai_clip_cfg cfg;
    void* tok = nullptr;   // [vocab, C]
    void* pos = nullptr;   // [max_len, C]
    std::vector<Layer> layers;
    Norm final_ln;
    int* ids_dev = nullptr;
    void* out_stage = nullptr;  // [max_batch*L, C] storage dtype
    size_t workspace_bytes = 0;
    std::mutex mu;
    cudaStream_t gs = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    bool use_graphs = true;
    std::unordered_map<int, cudaGraphExec_t> graphs;
    std::unordered_map<int, int64_t> graph_launches;

    void build(const etai_tensor* weights, int n_weights) {
        for (int i = 0; i < n_weights; ++i) table[weights[i].name] = &weights[i];
        const int C = cfg.hidden, F = cfg.ffn, V = cfg.vocab, L = cfg.max_len;
        stage_elems = (size_t)V * C;
        if ((size_t)F * C > stage_elems) stage_elems = (size_t)F * C;
        CUDA_CHECK(cudaMalloc(&stage, stage_elems * 4 + stage_elems * 2));
        const std::string tm = "text_model.";
        tok = dmalloc((size_t)V * C * esz);
        put(tm + "embeddings.token_embedding.weight", {V, C}, tok);
        pos = dmalloc((size_t)L * C * esz);
        put(tm + "embeddings.position_embedding.weight", {L, C}, pos);
        for (int i = 0; i < cfg.layers; ++i) {
            const std::string p = tm + "encoder.layers." + std::to_string(i);
            Layer l;
            l.ln1 = load_norm(p + ".layer_norm1", C);
            l.ln2 = load_norm(p + ".layer_norm2", C);
            l.qkv.n = 3 * C; l.qkv.k = C;
            l.qkv.w = dmalloc((size_t)3 * C * C * esz);
            l.qkv.b = dmalloc((size_t)3 * C * esz);
            const char* names[3] = {".self_attn.q_proj", ".self_attn.k_proj", ".self_attn.v_proj"};
            for (int j = 0; j < 3; ++j) {
                put(p + names[j] + ".weight", {C, C}, (char*)l.qkv.w + (size_t)j * C * C * esz);
                put(p + names[j] + ".bias", {C}, (char*)l.qkv.b + (size_t)j * C * esz);
            }
            l.o = load_lin(p + ".self_attn.out_proj", C, C, true);
            l.fc1 = load_lin(p + ".mlp.fc1", F, C, true);
            l.fc2 = load_lin(p + ".mlp.fc2", C, F, true);
            layers.push_back(l);
        }
        final_ln = load_norm(tm + "final_layer_norm", C);
        CUDA_CHECK(cudaFree(stage));
        stage = nullptr;
        table.clear();
    }

    void run_body(int B, cudaStream_t s) {
        const int C = cfg.hidden, L = cfg.max_len, heads = cfg.heads, d = C / heads;
        const long M = (long)B * L;
        arena.reset();
        void* x = arena.alloc((size_t)M * C * esz);
        if (arena.base) {
            clip_embed(ids_dev, tok, pos, x, M, L, C, cfg.vocab, dt, s);
            launches += 1;
        }
        for (const Layer& l : layers) {
            void* n1 = lnorm(x, M, l.ln1, s);
            void* qkv = linear(n1, M, l.qkv, nullptr, s);
            void* ao = arena.alloc((size_t)M * C * esz);
            if (arena.base) {
                cudaEvent_t e = prof_begin(s);
                clip_attention(qkv, ao, B, L, heads, d, 1.0f / sqrtf((float)d), dt, s);
                prof_end(ETAI_PROF_SELF_ATTN, e, 1, s);
            }
            x = linear(ao, M, l.o, x, s);
            void* n2 = lnorm(x, M, l.ln2, s);
            void* f = linear(n2, M, l.fc1, nullptr, s);
            if (arena.base) {
                cudaEvent_t e = prof_begin(s);
                quick_gelu(f, M * cfg.ffn, dt, s);
                prof_end(ETAI_PROF_OTHER, e, 1, s);
            }
            x = linear(f, M, l.fc2, x, s);
        }
        lnorm(x, M, final_ln, s, out_stage);
    }

    void plan_workspace() {
        arena.base = nullptr; arena.cap = 0; arena.peak = 0;
        const long M = (long)cfg.max_batch * cfg.max_len;
        run_body(cfg.max_batch, 0);
        size_t need = arena.peak + 4096;
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, need));
        arena.base = (char*)p; arena.cap = need; arena.reset();
        CUDA_CHECK(cudaMalloc((void**)&ids_dev, (size_t)M * sizeof(int)));
        CUDA_CHECK(cudaMalloc(&out_stage, (size_t)M * cfg.hidden * esz));
        CUDA_CHECK(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
        workspace_bytes = need + (size_t)M * (sizeof(int) + cfg.hidden * esz);
    }

    // ids: HOST int32 [B, L]
    void encode(const int32_t* ids, int B, void* out, int io_dtype, cudaStream_t user) {
        std::lock_guard<std::mutex> lk(mu);
        const long M = (long)B * cfg.max_len;
        CUDA_CHECK(cudaEventRecord(ev_in, user));
        CUDA_CHECK(cudaStreamWaitEvent(gs, ev_in, 0));
        CUDA_CHECK(cudaMemcpyAsync(ids_dev, ids, (size_t)M * sizeof(int), cudaMemcpyHostToDevice, gs));  // 308 B per prompt
        bool done = false;
        if (use_graphs && !prof_on) {
            auto it = graphs.find(B);
            if (it != graphs.end()) {
                CUDA_CHECK(cudaGraphLaunch(it->second, gs));
                launches += graph_launches[B];
                done = true;
            } else {
                int64_t l0 = launches;
                cudaGraph_t graph = nullptr;
                CUDA_CHECK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed));
                try {
                    run_body(B, gs);
                } catch (...) {
                    cudaStreamEndCapture(gs, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                CUDA_CHECK(cudaStreamEndCapture(gs, &graph));
                graph_launches[B] = launches - l0;
                cudaGraphExec_t exec = nullptr;
                CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
                CUDA_CHECK(cudaGraphDestroy(graph));
                graphs[B] = exec;
                CUDA_CHECK(cudaGraphLaunch(exec, gs));
                done = true;
            }
        }
        if (!done) run_body(B, gs);
        convert(out_stage, dt, out, io_dtype, M * cfg.hidden, gs);
        launches += 1;
        CUDA_CHECK(cudaEventRecord(ev_out, gs));
        CUDA_CHECK(cudaStreamWaitEvent(user, ev_out, 0));
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

int etai_clip_destroy(etai_clip* h) {
    if (!h) return ETAI_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& kv : h->graphs)
        if (kv.second) cudaGraphExecDestroy(kv.second);
    h->wstore.reset();
    void* extra[] = {h->arena.base, h->ids_dev, h->out_stage, h->stage};
    for (void* p : extra)
        if (p) cudaFree(p);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->gs) cudaStreamDestroy(h->gs);
    delete h;
    return ETAI_OK;
}

int etai_clip_create(etai_clip** out, const etai_clip_cfg* cfg, const etai_tensor* weights, int32_t n_weights, int32_t device) {
    ETAI_API_BEGIN
    ETAI_CHECK(out && cfg && weights && n_weights > 0, ETAI_ERR_ARG, "clip_create: null argument");
    ETAI_CHECK(cfg->dtype == ETAI_F32 || cfg->dtype == ETAI_F16 || cfg->dtype == ETAI_BF16, ETAI_ERR_ARG, "clip_create: dtype");
    ETAI_CHECK(cfg->max_batch >= 1 && cfg->max_batch <= 256, ETAI_ERR_ARG, "clip_create: max_batch in [1,256]");
    ETAI_CHECK(cfg->hidden % 64 == 0 && cfg->heads >= 1 && cfg->hidden / cfg->heads == 64, ETAI_ERR_UNSUPPORTED,
               "clip_create: head dim must be 64");
    ETAI_CHECK(cfg->max_len >= 1 && cfg->max_len <= 96 && cfg->layers >= 1 && cfg->ffn % 64 == 0 && cfg->vocab >= 1, ETAI_ERR_ARG,
               "clip_create: max_len <= 96, ffn % 64 == 0");
    int ndev = 0;
    CUDA_CHECK(cudaGetDeviceCount(&ndev));
    ETAI_CHECK(device >= 0 && device < ndev, ETAI_ERR_ARG, "clip_create: no such CUDA device");
    CUDA_CHECK(cudaSetDevice(device));
    etai_clip* h = new etai_clip();
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
        etai_clip_destroy(h);
        throw;
    }
    *out = h;
    ETAI_API_END
}

int etai_clip_encode(etai_clip* h, const int32_t* input_ids, int32_t B, void* out, int32_t io_dtype, void* stream) {
    ETAI_API_BEGIN
    ETAI_CHECK(h && input_ids && out, ETAI_ERR_ARG, "clip_encode: null argument");
    ETAI_CHECK(B >= 1 && B <= h->cfg.max_batch, ETAI_ERR_ARG, "clip_encode: batch out of range");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->encode(input_ids, B, out, io_dtype, (cudaStream_t)stream);
    ETAI_API_END
}

int64_t etai_clip_launch_count(const etai_clip* h) { return h ? h->launches : 0; }

}  // extern "C"
