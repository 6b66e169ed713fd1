// Shared plumbing of the engine handles (etai_unet, etai_vae, etai_clip): packed-weight ownership, the bump arena for
// activations, weight loading / repacking helpers and the op wrappers that pick the tcgen05 or the SIMT kernel, count
// launches and (opt-in) time every op with CUDA events.
#pragma once
#include <memory>
#include <unordered_map>
#include <vector>
#include <string>
#include <cstring>
#include <cmath>
#include <cstdlib>

#include "ops.cuh"

namespace etai {

struct Conv { void* w = nullptr; void* b = nullptr; int cin = 0, cout = 0; };
struct Lin { void* w = nullptr; void* b = nullptr; int n = 0, k = 0; };
struct Norm { void* g = nullptr; void* b = nullptr; int c = 0; };

// Packed device weights, shared (read-only) between a handle and its clones.
struct WeightStore {
    std::vector<void*> owned;
    size_t bytes = 0;
    int device = 0;
    ~WeightStore() {
        cudaSetDevice(device);
        for (void* p : owned) cudaFree(p);
    }
};

struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    void* alloc(size_t bytes) {
        size_t a = (off + 255) & ~size_t(255);
        off = a + bytes;
        if (off > peak) peak = off;
        if (base == nullptr) return reinterpret_cast<void*>(size_t(256) + a);  // planning pass: unique, never dereferenced
        ETAI_CHECK(off <= cap, ETAI_ERR_NOMEM, "activation arena exhausted");
        return base + a;
    }
    void reset() { off = 0; }
};


// One recorded op of a "train mode" forward (null-text inversion differentiates the UNet w.r.t. the text context):
// enough to run the op's data-gradient afterwards.  Weight gradients do not exist.
enum TapeKind { T_GEMM, T_GN, T_LN, T_SELF_ATTN, T_CROSS_ATTN, T_GEGLU, T_CONCAT, T_UPSAMPLE };
struct TapeRec {
    TapeKind kind;
    GemmArgs g;                 // T_GEMM (dense / conv geometry, A, W, C, residual)
    const void* x = nullptr;    // primary input
    const void* x2 = nullptr;   // second input (concat)
    void* y = nullptr;          // output
    Norm n;                     // T_GN / T_LN
    int B = 0, H = 0, W = 0, C = 0, C2 = 0, heads = 0, d = 0, kv_off = 0;
    long M = 0, HW = 0;
    float eps = 0.f, scale = 0.f;
    bool silu = false;
};

struct OpCtx {
    // ---- tape (train mode) ----
    bool tape_on = false;
    std::vector<TapeRec> tape;

    int device = 0;
    int dt = ETAI_F32;   // storage dtype
    bool tc = false;     // tcgen05 path enabled
    size_t esz = 4;
    std::shared_ptr<WeightStore> wstore;  // device allocations of the packed weights (shared with clones)
    size_t weight_bytes = 0;
    Arena arena;
    void* gn_ws = nullptr;
    void* tc_ws = nullptr;
    size_t tc_ws_bytes = 0;

    // ---- instrumentation: launch counter (always) + per-category CUDA-event timing (opt-in) -----
    int64_t launches = 0;
    bool prof_on = false;
    struct ProfRec { int cat; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    cudaEvent_t prof_begin(cudaStream_t s) {
        if (!prof_on || !arena.base) return nullptr;
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreate(&e));
        CUDA_CHECK(cudaEventRecord(e, s));
        return e;
    }
    void prof_end(int cat, cudaEvent_t a, int n_launches, cudaStream_t s) {
        if (arena.base) launches += n_launches;
        if (!a) return;
        cudaEvent_t b;
        CUDA_CHECK(cudaEventCreate(&b));
        CUDA_CHECK(cudaEventRecord(b, s));
        prof_recs.push_back({cat, a, b});
    }

    // ---- weight loading helpers -------------------------------------------------------------
    std::unordered_map<std::string, const etai_tensor*> table;
    float* stage = nullptr;
    size_t stage_elems = 0;

    void* dmalloc(size_t bytes) {
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 256));
        wstore->owned.push_back(p);
        wstore->bytes += bytes;
        weight_bytes += bytes;
        return p;
    }
    const etai_tensor& find(const std::string& name) {
        auto it = table.find(name);
        ETAI_CHECK(it != table.end(), ETAI_ERR_ARG, ("missing weight: " + name).c_str());
        return *it->second;
    }
    static size_t numel(const etai_tensor& t) {
        size_t n = 1;
        for (int i = 0; i < t.ndim; ++i) n *= (size_t)t.shape[i];
        return n;
    }
    // fp32 copy of a named tensor in the staging buffer (device); valid until the next call
    const float* staged(const std::string& name, std::initializer_list<int64_t> shape) {
        const etai_tensor& t = find(name);
        ETAI_CHECK(t.ndim == (int)shape.size(), ETAI_ERR_ARG, ("bad rank for " + name).c_str());
        int i = 0;
        for (int64_t s : shape) {
            ETAI_CHECK(t.shape[i] == s, ETAI_ERR_ARG, ("bad shape for " + name).c_str());
            ++i;
        }
        size_t n = numel(t);
        ETAI_CHECK(n <= stage_elems, ETAI_ERR_ARG, "staging buffer too small");
        if (t.dtype == ETAI_F32) {
            CUDA_CHECK(cudaMemcpy(stage, t.data, n * 4, t.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
        } else {
            void* tmp = reinterpret_cast<char*>(stage) + stage_elems * 4;  // second half of the staging area
            CUDA_CHECK(cudaMemcpy(tmp, t.data, n * 2, t.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
            convert(tmp, t.dtype, stage, ETAI_F32, (long)n, 0);
            CUDA_CHECK(cudaStreamSynchronize(0));
        }
        return stage;
    }
    // plain tensor converted to storage dtype at dst (device)
    void put(const std::string& name, std::initializer_list<int64_t> shape, void* dst) {
        const etai_tensor& t = find(name);
        if (t.dtype == dt) {  // already in the storage dtype (the Python loader converts plain tensors on the host): one copy, no kernel
            ETAI_CHECK(t.ndim == (int)shape.size(), ETAI_ERR_ARG, ("bad rank for " + name).c_str());
            int i = 0;
            for (int64_t d : shape) {
                ETAI_CHECK(t.shape[i] == d, ETAI_ERR_ARG, ("bad shape for " + name).c_str());
                ++i;
            }
            CUDA_CHECK(cudaMemcpy(dst, t.data, numel(t) * esz, t.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
            return;
        }
        const float* s = staged(name, shape);
        size_t n = numel(t);
        convert(s, ETAI_F32, dst, dt, (long)n, 0);
        CUDA_CHECK(cudaStreamSynchronize(0));
    }
    Norm load_norm(const std::string& p, int c) {
        Norm n;
        n.c = c;
        n.g = dmalloc(c * esz);
        n.b = dmalloc(c * esz);
        put(p + ".weight", {c}, n.g);
        put(p + ".bias", {c}, n.b);
        return n;
    }
    Lin load_lin(const std::string& p, int n, int k, bool bias, bool conv1x1 = false) {
        Lin l;
        l.n = n; l.k = k;
        l.w = dmalloc((size_t)n * k * esz);
        if (conv1x1) put(p + ".weight", {n, k, 1, 1}, l.w);
        else put(p + ".weight", {n, k}, l.w);
        if (bias) {
            l.b = dmalloc(n * esz);
            put(p + ".bias", {n}, l.b);
        }
        return l;
    }
    // cin_pad / cout_pad > 0: zero-pad to a tcgen05-friendly shape (conv_in 4 -> 64 input channels = one K block,
    // conv_out 4 -> 32 output channels); the Conv then describes the padded problem.
    Conv load_conv(const std::string& p, int cin, int cout, int cin_pad = 0, int cout_pad = 0) {
        Conv c;
        c.cin = cin_pad > 0 ? cin_pad : cin;
        c.cout = cout_pad > 0 ? cout_pad : cout;
        c.w = dmalloc((size_t)c.cout * 9 * c.cin * esz);
        c.b = dmalloc(c.cout * esz);
        CUDA_CHECK(cudaMemset(c.b, 0, c.cout * esz));
        const float* s = staged(p + ".weight", {cout, cin, 3, 3});
        pack_conv_weight(s, c.w, cout, cin, c.cout, c.cin, dt, 0);
        CUDA_CHECK(cudaStreamSynchronize(0));
        put(p + ".bias", {cout}, c.b);
        return c;
    }

    // ---- op wrappers --------------------------------------------------------------------------
    void gemm(GemmArgs& a, cudaStream_t s) {
        a.dtype = dt;
        cudaEvent_t e = prof_begin(s);
        bool use_tc = tc && gemm_tc_supported(a);
        if (use_tc) gemm_tc(a, tc_ws, tc_ws_bytes, s);
        else gemm_simt(a, s);
        prof_end(a.conv ? ETAI_PROF_CONV : ETAI_PROF_GEMM, e, (use_tc && a.conv && a.stride == 2) ? 2 : 1, s);
    }
    // `out` (optional, every wrapper): write there instead of bump-allocating the result in the arena
    void* linear(const void* x, long M, const Lin& l, const void* residual, cudaStream_t s, int geglu = 0,
                 void* out = nullptr) {
        int nout = geglu ? l.n / 2 : l.n;
        if (geglu && tape_on) {  // train mode: keep the pre-activation (value, gate) pairs for the backward pass
            void* u = arena.alloc((size_t)M * l.n * esz);
            GemmArgs a;
            a.A = x; a.W = l.w; a.C = u; a.bias = l.b;
            a.M = M; a.N = l.n; a.K = l.k; a.lda = l.k; a.ldc = l.n; a.ldr = l.n;
            a.dtype = dt;
            if (arena.base) gemm(a, s);
            tape.push_back(TapeRec{T_GEMM, a, x, nullptr, u});
            void* y = out ? out : arena.alloc((size_t)M * nout * esz);
            if (arena.base) {
                cudaEvent_t e = prof_begin(s);
                geglu_fwd(u, y, M, nout, dt, s);
                prof_end(ETAI_PROF_OTHER, e, 1, s);
            }
            TapeRec r{T_GEGLU, GemmArgs(), u, nullptr, y};
            r.M = M; r.C = nout;
            tape.push_back(r);
            return y;
        }
        void* y = out ? out : arena.alloc((size_t)M * nout * esz);
        GemmArgs a;
        a.A = x; a.W = l.w; a.C = y; a.bias = l.b; a.residual = residual;
        a.M = M; a.N = l.n; a.K = l.k; a.lda = l.k; a.ldc = nout; a.ldr = nout; a.geglu = geglu;
        a.dtype = dt;
        if (arena.base) gemm(a, s);
        if (tape_on) tape.push_back(TapeRec{T_GEMM, a, x, nullptr, y});
        return y;
    }
    // pad = 1: the usual pad-1 conv; pad = 0 with stride 2: zero padding on the right / bottom only (AutoencoderKL's
    // downsampler, F.pad(x, (0,1,0,1)) + conv(stride 2, padding 0))
    // rowbias: fp32 bias added per (row group, n): one vector for all rows (rows_per_group = 0 -> M) or a
    // [M / rows_per_group, ldrb] table (one row of it per image when rows_per_group = Ho * Wo)
    void* conv3x3(const void* x, int B, int H, int W, const Conv& c, int stride, const float* rowbias,
                  const void* residual, cudaStream_t s, void* out = nullptr, int pad = 1, long rows_per_group = 0,
                  long ldrb = 0) {
        int Ho = (H + 1 + pad - 3) / stride + 1, Wo = (W + 1 + pad - 3) / stride + 1;
        long M = (long)B * Ho * Wo;
        void* y = out ? out : arena.alloc((size_t)M * c.cout * esz);
        GemmArgs a;
        a.A = x; a.W = c.w; a.C = y; a.bias = c.b; a.residual = residual; a.rowbias = rowbias;
        a.rows_per_group = rows_per_group > 0 ? rows_per_group : M; a.ldrb = rows_per_group > 0 ? ldrb : 0;
        a.M = M; a.N = c.cout; a.K = 9 * c.cin; a.ldc = c.cout; a.ldr = c.cout;
        a.conv = 1; a.B = B; a.H = H; a.Wd = W; a.Cin = c.cin; a.stride = stride; a.Ho = Ho; a.Wo = Wo; a.pad = pad;
        a.dtype = dt;
        if (arena.base) gemm(a, s);
        if (tape_on) tape.push_back(TapeRec{T_GEMM, a, x, nullptr, y});
        return y;
    }
    void* gnorm(const void* x, int B, long HW, const Norm& n, float eps, bool silu, cudaStream_t s, void* out = nullptr) {
        void* y = out ? out : arena.alloc((size_t)B * HW * n.c * esz);
        if (arena.base) {
            cudaEvent_t e = prof_begin(s);
            groupnorm(x, y, n.g, n.b, B, HW, n.c, 32, eps, silu, dt, gn_ws, s);
            prof_end(ETAI_PROF_GROUPNORM, e, groupnorm_launches(HW, n.c, 32, dt), s);
        }
        if (tape_on) {
            TapeRec r{T_GN, GemmArgs(), x, nullptr, y, n};
            r.B = B; r.HW = HW; r.eps = eps; r.silu = silu;
            tape.push_back(r);
        }
        return y;
    }
    void* lnorm(const void* x, long M, const Norm& n, cudaStream_t s, void* out = nullptr) {
        void* y = out ? out : arena.alloc((size_t)M * n.c * esz);
        if (arena.base) {
            cudaEvent_t e = prof_begin(s);
            layernorm(x, y, n.g, n.b, M, n.c, 1e-5f, dt, s);
            prof_end(ETAI_PROF_LAYERNORM, e, 1, s);
        }
        if (tape_on) {
            TapeRec r{T_LN, GemmArgs(), x, nullptr, y, n};
            r.M = M; r.eps = 1e-5f;
            tape.push_back(r);
        }
        return y;
    }
};

}  // namespace etai
