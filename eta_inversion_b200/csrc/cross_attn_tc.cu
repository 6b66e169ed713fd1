// tcgen05 cross-attention over the 77-token text context with prompt-to-prompt control fused in (16-bit engine).
//
// CTA = 128 query rows x one group (a plain UNet row, or a (base, target) PtP pair) x one head; grid (q tiles, groups,
// heads) so that even the 8x8 / 16x16 layers spread over the SMs.  Per-head store partials go to a workspace and are
// folded into the accumulators in head order by a second tiny launch (deterministic, no atomics).
// Per (head, row):   S = Q K^T            UMMA 128 x 80 x d      (K tile: 77 keys, TMA zero-fills 77..79 and the d tail)
//                    softmax per thread   (one query row per thread, 77 probabilities in registers, fp32)
//   target row only: R = P_base Mapper    UMMA 128 x 80 x 80     (P_base is still in smem from the base pass)
//                    P' = alpha*eq*(a*R + (1-a)*P) + (1-alpha)*P          ptp.py:205-211,234-274, no renormalisation
//                    store: acc[slot][pix][w] += P'                      ptp.py:150-171 (post-edit probabilities)
//                    O = P' V             UMMA 128 x d_pad x 80  (V tile as it landed: MN-major)
// warp 0 = TMA producer (Q, K, V of the next (head,row) prefetched when smem allows), warp 1 = MMA issuer,
// warps 2..5 = 128 softmax/edit threads.  Nothing of size [B*8, N, 77] is ever written to memory.
#include <cstring>
#include <type_traits>
#include "ops.cuh"
#include "tc_common.cuh"

namespace etai {

namespace {

using namespace tc;

constexpr int XT_THREADS = 192, QATOM = 128 * 128, KATOM = 80 * 128, LP = 80;

struct XParams {
    void* out;
    int N, L, heads, d;
    long ldo;
    float scale_log2e;
    CrossGroup g;  // filled per CTA from the table below
    int n_groups;
    CrossGroup groups[ETAI_MAX_ROWS];
    const float *mapper, *blend_a, *equalizer, *alpha_step;
    float* store;        // accumulators [slots][N][L] (used by the reduce launch)
    float* store_part;   // per-head partials [heads][slots][N][L]
    int n_slots;
    int fmt;
};

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int NCOL>
__device__ __forceinline__ void tmem_ld_row(uint32_t taddr, float* v) {  // NCOL in {48, 80, 160}
    constexpr int N32 = NCOL / 32;
#pragma unroll
    for (int b = 0; b < N32; ++b) {
        float t[32];
        tmem_ld32(taddr + b * 32, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[b * 32 + i] = t[i];
    }
    if constexpr (NCOL % 32 != 0) {
        uint32_t rr[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(taddr + N32 * 32)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) v[N32 * 32 + i] = __uint_as_float(rr[i]);
    }
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (std::is_same<T, __half>::value) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

template <typename T, int ATOMS, int STAGES>
__global__ void __launch_bounds__(XT_THREADS, ATOMS == 1 ? 2 : 1)
cross_attn_tc_k(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmM,
                const __grid_constant__ XParams p) {
    constexpr int STAGE_BYTES = ATOMS * (QATOM + 2 * KATOM);
    // TMEM columns: S (80) | R (80) | O (d_pad).  d = 40 fits 256 columns (two CTAs per SM), larger heads take 512
    constexpr int TMEM_COLS = ATOMS == 1 ? 256 : 512;
    constexpr int R_COL = ATOMS == 1 ? 96 : 128, O_COL = ATOMS == 1 ? 192 : 256;
    constexpr int PA_OFF = STAGES * STAGE_BYTES;       // P (A operand): 2 atoms of [128 x 128 B]
    constexpr int MAP_OFF = PA_OFF + 2 * QATOM;        // Mapper^T (B operand, K-major): 2 atoms of [80 x 128 B]
    constexpr int BAR_OFF = MAP_OFF + 2 * KATOM;
    constexpr int DPAD = ATOMS == 1 ? 48 : ATOMS == 2 ? 80 : 160;
    constexpr int D = ATOMS == 1 ? 40 : ATOMS == 2 ? 80 : 160;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;
    uint64_t* p_ready = s_full + 1;
    uint64_t* o_full = p_ready + 1;
    uint64_t* map_full = o_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(map_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CrossGroup g = p.groups[blockIdx.y];
    const int q0 = blockIdx.x * 128;
    const int nrows = g.tgt >= 0 ? 2 : 1;
    const bool edit = g.tgt >= 0 && p.mapper != nullptr;
    const int iters = nrows;
    const int head = blockIdx.z;
    const int ksteps = (D + 15) / 16;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
        mbar_init(map_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_r = tmem_base + R_COL, tmem_o = tmem_base + O_COL;

    if (warp == 0) {
        if (lane == 0) {
            if (edit) {  // Mapper^T (16-bit, [80 n][128 w] zero padded, prepared once per forward): B operand of R = P_base M
                mbar_expect_tx(map_full, 2 * KATOM);
                tma_load_3d(smem + MAP_OFF, &tmM, map_full, 0, 0, g.pair);
                tma_load_3d(smem + MAP_OFF + KATOM, &tmM, map_full, 64, 0, g.pair);
            }
            for (int it = 0; it < iters; ++it) {
                int row = it == 0 ? g.base : g.tgt;
                int s = it % STAGES;
                mbar_wait(&kv_empty[s], ((it / STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], STAGE_BYTES);
                unsigned char* sb = smem + s * STAGE_BYTES;
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) {
                    tma_load_4d(sb + a * QATOM, &tmQ, &kv_full[s], a * 64, head, q0, row);
                    tma_load_4d(sb + ATOMS * QATOM + a * KATOM, &tmK, &kv_full[s], a * 64, head, 0, row);
                    tma_load_4d(sb + ATOMS * (QATOM + KATOM) + a * KATOM, &tmV, &kv_full[s], a * 64, head, 0, row);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_f16(p.fmt, 128, LP);
            const uint32_t idesc_o = make_idesc_f16(p.fmt, 128, DPAD) | (1u << 16);
            const uint32_t pa = smem_u32(smem + PA_OFF), mp = smem_u32(smem + MAP_OFF);
            for (int it = 0; it < iters; ++it) {
                int s = it % STAGES;
                bool is_tgt = (it % nrows) == 1;
                mbar_wait(&kv_full[s], (it / STAGES) & 1);
                tc_fence_after();
                uint32_t qa = smem_u32(smem + s * STAGE_BYTES), ka = qa + ATOMS * QATOM, va = ka + ATOMS * KATOM;
                for (int k = 0; k < ksteps; ++k) {
                    uint32_t ko = (uint32_t)(k & 3) * 32;
                    umma_f16(tmem_base, make_smem_desc_sw128(qa + (k >> 2) * QATOM + ko),
                             make_smem_desc_sw128(ka + (k >> 2) * KATOM + ko), idesc_s, k != 0);
                }
                if (edit && is_tgt) {
                    mbar_wait(map_full, 0);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < LP / 16; ++k) {
                        uint32_t ko = (uint32_t)(k & 3) * 32;
                        umma_f16(tmem_r, make_smem_desc_sw128(pa + (k >> 2) * QATOM + ko),
                                 make_smem_desc_sw128(mp + (k >> 2) * KATOM + ko), idesc_s, k != 0);
                    }
                }
                umma_commit(s_full);
                mbar_wait(p_ready, it & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < LP / 16; ++k) {
                    uint32_t ko = (uint32_t)(k & 3) * 32;
                    umma_f16(tmem_o, make_smem_desc_sw128(pa + (k >> 2) * QATOM + ko),
                             desc_mn_sw128(va + (uint32_t)k * 16 * 128, KATOM), idesc_o, k != 0);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(o_full);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const int q = q0 + r;
        const bool q_ok = q < p.N;
        unsigned char* pa_row = smem + PA_OFF + r * 128;
        const float* al = edit ? p.alpha_step + g.pair * p.L : nullptr;
        const float* eq = edit ? p.equalizer + g.pair * p.L : nullptr;
        const float* ba = edit ? p.blend_a + g.pair * p.L : nullptr;
        for (int it = 0; it < iters; ++it) {
            const bool is_tgt = it == 1;
            const int row = is_tgt ? g.tgt : g.base;
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            float pr[LP];
            tmem_ld_row<LP>(tmem_base + lane_addr, pr);
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < LP; ++i)
                if (i < p.L) mx = fmaxf(mx, pr[i]);
            float sum = 0.f;
            const float mb = mx * p.scale_log2e;
#pragma unroll
            for (int i = 0; i < LP; ++i) {
                pr[i] = (i < p.L) ? exp2f(fmaf(pr[i], p.scale_log2e, -mb)) : 0.f;
                sum += pr[i];
            }
            const float inv = 1.f / sum;
#pragma unroll
            for (int i = 0; i < LP; ++i) pr[i] *= inv;
            if (edit && is_tgt) {
                float rr[LP];
                tmem_ld_row<LP>(tmem_r + lane_addr, rr);
#pragma unroll
                for (int i = 0; i < LP; ++i) {
                    if (i < p.L) {
                        float a = al[i];
                        if (a != 0.f) {
                            float b = ba[i];
                            float f = eq[i] * (b * rr[i] + (1.f - b) * pr[i]);
                            pr[i] = a * f + (1.f - a) * pr[i];
                        }
                    }
                }
            }
            const int slot = is_tgt ? g.store_tgt : g.store_base;
            if (p.store_part && slot >= 0 && q_ok) {  // per-head partial; folded in head order by store_reduce_k
                float* part = p.store_part + (((long)head * p.n_slots + slot) * p.N + q) * p.L;
#pragma unroll
                for (int i = 0; i < LP; ++i)
                    if (i < p.L) part[i] = pr[i];
            }
            // P' -> shared memory as the A operand (K-major, 128B swizzle): 80 values = atom 0 (64) + atom 1 (16)
#pragma unroll
            for (int c = 0; c < LP / 8; ++c) {
                uint4 v4 = make_uint4(pack2<T>(pr[8 * c], pr[8 * c + 1]), pack2<T>(pr[8 * c + 2], pr[8 * c + 3]),
                                      pack2<T>(pr[8 * c + 4], pr[8 * c + 5]), pack2<T>(pr[8 * c + 6], pr[8 * c + 7]));
                int chunk = (c & 7) ^ (r & 7);
                *reinterpret_cast<uint4*>(pa_row + (c >> 3) * QATOM + chunk * 16) = v4;
            }
            tc_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(p_ready);

            mbar_wait(o_full, it & 1);
            tc_fence_after();
            float o[DPAD];
            tmem_ld_row<DPAD>(tmem_o + lane_addr, o);
            if (q_ok) {
                T* dst = reinterpret_cast<T*>(p.out) + ((long)row * p.N + q) * p.ldo + head * D;
#pragma unroll
                for (int c = 0; c < D; c += 8) {
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o8[i] = o[c + i];
                    store8<T>(dst + c, o8);
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

__global__ void store_reduce_k(float* __restrict__ acc, const float* __restrict__ part, int heads, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = acc[i];
    for (int h = 0; h < heads; ++h) a += part[(long)h * n + i];
    acc[i] = a;
}

template <typename T, int ATOMS, int STAGES>
void launch_x(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const CUtensorMap& m, const XParams& p, dim3 grid,
              cudaStream_t s) {
    constexpr int SMEM = STAGES * ATOMS * (QATOM + 2 * KATOM) + 2 * QATOM + 2 * KATOM + 256 + 1024;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(cross_attn_tc_k<T, ATOMS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    cross_attn_tc_k<T, ATOMS, STAGES><<<grid, XT_THREADS, SMEM, s>>>(q, k, v, m, p);
    KERNEL_CHECK();
}

CUtensorMap tmap4(const void* base, int dtype, int d, int heads, int N, int B, long ld, int box_rows) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * N};
    uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    return make_tmap_16bit(base, dtype, 4, dims, str, box);
}

// mapper fp32 [P][L][L] (row w, column n)  ->  16-bit [P][80 n][128 w], zero padded: K-major B operand for R = P_base M
template <typename T>
__global__ void prep_mapper_k(const float* __restrict__ m, T* __restrict__ out, int P, int L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * LP * 128) return;
    int w = i % 128, n = (i / 128) % LP, p = i / (128 * LP);
    out[i] = (w < L && n < L) ? from_f<T>(m[((long)p * L + w) * L + n]) : from_f<T>(0.f);
}

}  // namespace

size_t cross_attention_tc_mapper_bytes(int pairs) { return (size_t)pairs * LP * 128 * 2; }

void cross_attention_tc_prep_mapper(const float* mapper, void* map16, int pairs, int L, int dtype, cudaStream_t s) {
    int n = pairs * LP * 128;
    if (dtype == ETAI_F16) prep_mapper_k<__half><<<cdiv(n, 256), 256, 0, s>>>(mapper, (__half*)map16, pairs, L);
    else prep_mapper_k<__nv_bfloat16><<<cdiv(n, 256), 256, 0, s>>>(mapper, (__nv_bfloat16*)map16, pairs, L);
    KERNEL_CHECK();
}

bool cross_attention_tc_supported(const CrossAttnArgs& a) {
    if (a.dtype != ETAI_F16 && a.dtype != ETAI_BF16) return false;
    if (a.d != 40 && a.d != 80 && a.d != 160) return false;
    if (a.L > 80 || a.L < 1 || a.ldq % 8 || a.ldkv % 8 || a.ldo % 8 || a.koff % 8 || a.voff % 8) return false;
    if (a.mapper && !a.map16) return false;
    if (a.store && !a.store_part) return false;
    return true;
}

int cross_attention_tc(const CrossAttnArgs& a, cudaStream_t s) {
    ETAI_CHECK(cross_attention_tc_supported(a), ETAI_ERR_UNSUPPORTED, "cross_attention_tc: unsupported problem");
    XParams p;
    memset(&p, 0, sizeof(p));
    p.out = a.out; p.N = a.N; p.L = a.L; p.heads = a.heads; p.d = a.d; p.ldo = a.ldo;
    p.scale_log2e = a.scale * 1.4426950408889634f;
    p.n_groups = a.n_groups;
    int n_slots = 0;
    for (int i = 0; i < a.n_groups; ++i) {
        p.groups[i] = a.groups[i];
        if (a.groups[i].store_base + 1 > n_slots) n_slots = a.groups[i].store_base + 1;
        if (a.groups[i].store_tgt + 1 > n_slots) n_slots = a.groups[i].store_tgt + 1;
    }
    p.mapper = a.mapper; p.blend_a = a.blend_a; p.equalizer = a.equalizer; p.alpha_step = a.alpha_step;
    p.store = a.store;
    p.store_part = (a.store && n_slots > 0) ? a.store_part : nullptr;
    p.n_slots = n_slots;
    p.fmt = a.dtype == ETAI_BF16 ? 1 : 0;
    const char* kv = reinterpret_cast<const char*>(a.kv);
    CUtensorMap tq = tmap4(a.q, a.dtype, a.d, a.heads, a.N, a.B, a.ldq, 128);
    CUtensorMap tk = tmap4(kv + (size_t)a.koff * 2, a.dtype, a.d, a.heads, a.L, a.B, a.ldkv, 80);
    CUtensorMap tv = tmap4(kv + (size_t)a.voff * 2, a.dtype, a.d, a.heads, a.L, a.B, a.ldkv, 80);
    CUtensorMap tm = tq;  // placeholder when no edit is active (never dereferenced)
    if (a.mapper) {
        uint64_t dims[3] = {128, (uint64_t)LP, (uint64_t)ETAI_MAX_PAIRS};
        uint64_t str[2] = {128 * 2, (uint64_t)128 * 2 * LP};
        uint32_t box[3] = {64, (uint32_t)LP, 1};
        tm = make_tmap_16bit(a.map16, a.dtype, 3, dims, str, box);
    }
    dim3 grid(cdiv(a.N, 128), a.n_groups, a.heads);
#define LAUNCH(T)                                                              \
    do {                                                                       \
        if (a.d == 40) launch_x<T, 1, 1>(tq, tk, tv, tm, p, grid, s);          \
        else if (a.d == 80) launch_x<T, 2, 1>(tq, tk, tv, tm, p, grid, s);     \
        else launch_x<T, 3, 1>(tq, tk, tv, tm, p, grid, s);                    \
    } while (0)
    if (a.dtype == ETAI_F16) LAUNCH(__half);
    else LAUNCH(__nv_bfloat16);
#undef LAUNCH
    if (p.store_part) {
        long n = (long)n_slots * a.N * a.L;
        store_reduce_k<<<cdiv(n, 256), 256, 0, s>>>(a.store, a.store_part, a.heads, n);
        KERNEL_CHECK();
        return 2;
    }
    return 1;
}

}  // namespace etai
