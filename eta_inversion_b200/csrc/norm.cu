// GroupNorm(+SiLU) over NHWC and LayerNorm over rows.  HBM-bound: 16-byte loads, warp-shuffle / fixed-order
// reductions only (no float atomics, so results are run-to-run identical: the reference's A/B/A determinism test,
// test/test_edit.py:259-289, must hold bit-exactly).
//
// GroupNorm = two launches:
//   1) gn_stats_k : grid (chunks, B). Each CTA streams `rows_per_chunk` pixel rows fully coalesced (all channels of a
//      pixel are contiguous in NHWC), keeps per-channel partial sums in registers, folds them per group in shared
//      memory in a fixed order and writes (sum, sumsq) as doubles to ws[b][chunk][g][2].
//   2) gn_apply_k : every CTA first reduces the chunk partials of its batch row (double, fixed order) to mean/rstd
//      in shared memory, then normalises + affine (+SiLU) its slice with 16-byte loads/stores.
// Algorithmic bytes: read x + write y (2*elements*sizeof); the second read of x in (2) hits L2 for UNet-sized
// tensors (<= 10.5 MB per row in fp16).
#include "ops.cuh"

namespace etai {

static constexpr int GN_MAX_CHUNKS = 256;

struct GnPlan {
    int cvecs;           // C / 8 (8 channels per thread)
    int R;               // pixel rows processed in parallel by one CTA
    int threads;         // cvecs * R
    int rows_per_chunk;  // pixel rows per CTA
    int chunks;
};
static GnPlan gn_plan(long HW, int C) {
    GnPlan p;
    p.cvecs = C / 8;
    p.R = 256 / p.cvecs;
    if (p.R < 1) p.R = 1;
    p.threads = (p.cvecs * p.R + 31) / 32 * 32;  // whole warps; the tail threads only help in the fold
    long rpc = 32;
    rpc = 8;
    while (cdiv(HW, rpc) > 128) rpc *= 2;  // <= 128 chunks per batch row; the last CTA folds them
    if (rpc > HW) rpc = HW;
    p.rows_per_chunk = (int)rpc;
    p.chunks = cdiv(HW, rpc);
    return p;
}

// workspace layout (offsets independent of the call's batch size, so calls with different B can share one workspace):
//   [tickets: 64 ints, self-resetting, zero before first use][stats: 64 x 64 float2][partials: B x 256 x G x 2 doubles]
static constexpr size_t GN_STATS_OFF = 256, GN_PART_OFF = 256 + 64 * 64 * sizeof(float2);
size_t groupnorm_workspace_bytes(int B, long HW, int C, int groups) {
    (void)HW; (void)C;
    return GN_PART_OFF + (size_t)B * GN_MAX_CHUNKS * groups * 2 * sizeof(double);
}
size_t groupnorm_ticket_offset(int B, int groups) { (void)B; (void)groups; return 0; }

template <typename T>
__global__ void gn_stats_k(const T* __restrict__ x, double* __restrict__ ws, float2* __restrict__ stats,
                           int* __restrict__ tickets, long HW, int C, int G, int cvecs, int R, int rows_per_chunk,
                           int chunks, float eps) {
    extern __shared__ float sm[];  // [R][C] sums, [R][C] sumsq
    __shared__ int s_last;
    int b = blockIdx.y, chunk = blockIdx.x;
    int cv = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
    long row0 = (long)chunk * rows_per_chunk;
    long row1 = row0 + rows_per_chunk < HW ? row0 + rows_per_chunk : HW;
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    const T* xb = x + (long)b * HW * C;
    if (r < R) {
        long row = row0 + r;
        for (; row + 3L * R < row1; row += 4L * R) {  // 4 independent 16-byte loads in flight per thread
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) load8<T>(xb + (row + (long)u * R) * C + cv * 8, v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) { s[j] += v[u][j]; ss[j] = fmaf(v[u][j], v[u][j], ss[j]); }
        }
        for (; row < row1; row += R) {
            float v[8];
            load8<T>(xb + row * C + cv * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
        }
        float* S = sm + (long)r * C + cv * 8;
        float* SS = sm + (long)R * C + (long)r * C + cv * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { S[j] = s[j]; SS[j] = ss[j]; }
    }
    __syncthreads();
    // one warp per group slice: fixed-order fold over R x cpg entries
    int cpg = C / G;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < G; g += nwarps) {
        double a = 0.0, aa = 0.0;
        for (int i = lane; i < R * cpg; i += 32) {
            int rr = i / cpg, c = g * cpg + i % cpg;
            a += (double)sm[(long)rr * C + c];
            aa += (double)sm[(long)R * C + (long)rr * C + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            aa += __shfl_xor_sync(0xffffffffu, aa, o);
        }
        if (lane == 0) {
            double* w = ws + (((long)b * chunks + chunk) * G + g) * 2;
            w[0] = a;
            w[1] = aa;
        }
    }
    // The last CTA of this batch row to finish folds all chunk partials (fixed order -> deterministic) into mean/rstd.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = atomicAdd(&tickets[b], 1);
        s_last = (t == chunks - 1);
        if (s_last) tickets[b] = 0;  // self-resetting for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int g = warp; g < G; g += nwarps) {
        double a = 0.0, aa = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            const double* w = ws + (((long)b * chunks + c) * G + g) * 2;
            a += __ldcg(w);
            aa += __ldcg(w + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            aa += __shfl_xor_sync(0xffffffffu, aa, o);
        }
        if (lane == 0) {
            double n = (double)HW * (C / G);
            double mean = a / n;
            double var = aa / n - mean * mean;
            if (var < 0) var = 0;
            stats[b * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
        }
    }
}

// apply: thread = fixed 8-channel vector (scale/shift folded once into registers), CTA = slab of pixel rows
template <typename T, bool SILU>
__global__ void gn_apply_k(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ gamma,
                           const T* __restrict__ beta, const float2* __restrict__ stats, long HW, int C, int G, int cvecs,
                           int R, int rows_per_cta) {
    const int b = blockIdx.y;
    const int cv = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
    if (r >= R) return;
    const int cpg = C / G, c0 = cv * 8;
    float sc[8], sh[8];
    {
        float ga[8], be[8];
        load8<T>(gamma + c0, ga);
        load8<T>(beta + c0, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float2 st = stats[b * G + (c0 + j) / cpg];  // (mean, rstd)
            sc[j] = st.y * ga[j];
            sh[j] = be[j] - st.x * sc[j];
        }
    }
    const T* xb = x + (long)b * HW * C + c0;
    T* yb = y + (long)b * HW * C + c0;
    const long row0 = (long)blockIdx.x * rows_per_cta;
    const long row1 = row0 + rows_per_cta < HW ? row0 + rows_per_cta : HW;
    long row = row0 + r;
    for (; row + R < row1; row += 2L * R) {  // two independent 16-byte loads in flight
        float v0[8], v1[8];
        load8<T>(xb + row * C, v0);
        load8<T>(xb + (row + R) * C, v1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o0 = fmaf(v0[j], sc[j], sh[j]), o1 = fmaf(v1[j], sc[j], sh[j]);
            v0[j] = SILU ? silu_f(o0) : o0;
            v1[j] = SILU ? silu_f(o1) : o1;
        }
        store8<T>(yb + row * C, v0);
        store8<T>(yb + (row + R) * C, v1);
    }
    for (; row < row1; row += R) {
        float v0[8];
        load8<T>(xb + row * C, v0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o0 = fmaf(v0[j], sc[j], sh[j]);
            v0[j] = SILU ? silu_f(o0) : o0;
        }
        store8<T>(yb + row * C, v0);
    }
}

void groupnorm(const void* x, void* y, const void* gamma, const void* beta, int B, long HW, int C, int groups,
               float eps, bool silu, int dtype, void* ws, cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0 && C % groups == 0 && groups <= 64, ETAI_ERR_ARG, "groupnorm: C%8, C%groups, groups<=64");
    ETAI_CHECK(C / 8 <= 512, ETAI_ERR_ARG, "groupnorm: C too large");
    ETAI_CHECK(ws != nullptr, ETAI_ERR_ARG, "groupnorm: workspace required");
    GnPlan p = gn_plan(HW, C);
    size_t smem = (size_t)2 * p.R * C * sizeof(float);
    // apply grid: about 6 CTAs per SM in total, each a slab of whole pixel rows
    long want = (HW * B + 148L * 6 - 1) / (148L * 6);
    int rows_per_cta = (int)((want + p.R - 1) / p.R) * p.R;
    if (rows_per_cta < 2 * p.R) rows_per_cta = 2 * p.R;
    int ablocks = cdiv(HW, rows_per_cta);
    ETAI_CHECK(B <= 64, ETAI_ERR_ARG, "groupnorm: B <= 64");
    int* tickets = reinterpret_cast<int*>(ws);
    float2* stats = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) + GN_STATS_OFF);
    double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + GN_PART_OFF);
    ETAI_DISPATCH_DTYPE(dtype, T, {
        gn_stats_k<T><<<dim3(p.chunks, B), p.threads, smem, s>>>((const T*)x, part, stats, tickets, HW, C, groups, p.cvecs,
                                                                 p.R, p.rows_per_chunk, p.chunks, eps);
        KERNEL_CHECK();
        if (silu)
            gn_apply_k<T, true><<<dim3(ablocks, B), p.threads, 0, s>>>((const T*)x, (T*)y, (const T*)gamma, (const T*)beta,
                                                                       stats, HW, C, groups, p.cvecs, p.R, rows_per_cta);
        else
            gn_apply_k<T, false><<<dim3(ablocks, B), p.threads, 0, s>>>((const T*)x, (T*)y, (const T*)gamma, (const T*)beta,
                                                                        stats, HW, C, groups, p.cvecs, p.R, rows_per_cta);
        KERNEL_CHECK();
    });
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C <= 1280 -> <= 5 vectors of 8 per lane),
// exact two-pass mean/variance like torch.
// ---------------------------------------------------------------------------------------------
template <typename T, int MAXV>
__global__ void layernorm_k(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ gamma,
                            const T* __restrict__ beta, long M, int C, float eps) {
    long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= M) return;
    int nv = C / 8;
    float v[MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int vi = lane + 32 * i;
        if (vi < nv) {
            load8<T>(x + row * C + vi * 8, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[i][j];
        }
    }
    float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int vi = lane + 32 * i;
        if (vi < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float d = v[i][j] - mean; sq = fmaf(d, d, sq); }
        }
    }
    float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int vi = lane + 32 * i;
        if (vi < nv) {
            float ga[8], be[8], o[8];
            load8<T>(gamma + vi * 8, ga);
            load8<T>(beta + vi * 8, be);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * ga[j] + be[j];
            store8<T>(y + row * C + vi * 8, o);
        }
    }
}

void layernorm(const void* x, void* y, const void* gamma, const void* beta, long M, int C, float eps, int dtype,
               cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0 && C <= 8 * 32 * 8, ETAI_ERR_ARG, "layernorm: C%8==0 and C<=2048");
    int warps = 8;
    ETAI_DISPATCH_DTYPE(dtype, T, {
        if (C <= 8 * 32 * 2)
            layernorm_k<T, 2><<<cdiv(M, warps), warps * 32, 0, s>>>((const T*)x, (T*)y, (const T*)gamma, (const T*)beta, M, C, eps);
        else if (C <= 8 * 32 * 5)
            layernorm_k<T, 5><<<cdiv(M, warps), warps * 32, 0, s>>>((const T*)x, (T*)y, (const T*)gamma, (const T*)beta, M, C, eps);
        else
            layernorm_k<T, 8><<<cdiv(M, warps), warps * 32, 0, s>>>((const T*)x, (T*)y, (const T*)gamma, (const T*)beta, M, C, eps);
    });
    KERNEL_CHECK();
}

}  // namespace etai
