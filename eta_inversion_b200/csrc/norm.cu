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
    int rows_per_chunk;  // pixel rows per CTA = 8 * R * batches: every thread has 8 independent 16-byte loads in flight
    int batches;
    int chunks;
};
// The partition depends on (HW, C) only -- never on the batch size -- so a row's statistics are bit-identical whatever
// it is batched with (lock-step groups, A/B/A runs).
static GnPlan gn_plan(long HW, int C) {
    GnPlan p;
    p.cvecs = C / 8;
    p.R = 256 / p.cvecs;
    if (p.R < 1) p.R = 1;
    p.threads = (p.cvecs * p.R + 31) / 32 * 32;  // whole warps; the tail threads only help in the fold
    long per_batch = 8L * p.R;
    p.batches = (int)cdiv(cdiv(HW, per_batch), (long)GN_MAX_CHUNKS);
    if (p.batches < 1) p.batches = 1;
    p.rows_per_chunk = (int)(per_batch * p.batches);
    p.chunks = (int)cdiv(HW, (long)p.rows_per_chunk);
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
__global__ void __launch_bounds__(512, 2) gn_stats_k(const T* __restrict__ x, double* __restrict__ ws, float2* __restrict__ stats,
                           int* __restrict__ tickets, long HW, int C, int G, int cvecs, int R, int rows_per_chunk,
                           int batches, int chunks, float eps) {
    extern __shared__ float sm[];  // [R][C] sums, [R][C] sumsq
    __shared__ int s_last;
    int b = blockIdx.y, chunk = blockIdx.x;
    int cv = threadIdx.x % cvecs, r = threadIdx.x / cvecs;
    long row0 = (long)chunk * rows_per_chunk;
    long row1 = row0 + rows_per_chunk < HW ? row0 + rows_per_chunk : HW;
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    const T* xb = x + (long)b * HW * C + cv * 8;
    if (r < R) {
        long row = row0 + r;
        for (int it = 0; it < batches; ++it, row += 8L * R) {  // 8 independent 16-byte loads in flight per thread
            // branch-free: rows past the chunk end re-read its last row and are masked out afterwards (a conditional
            // load puts every load in its own reconvergence region and serialises the eight DRAM round trips)
            float v[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                long rr = row + (long)u * R;
                load8<T>(xb + (rr < row1 ? rr : row1 - 1) * C, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool live = row + (long)u * R < row1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = live ? v[u][j] : 0.f;
                    s[j] += t;
                    ss[j] = fmaf(t, t, ss[j]);
                }
            }
        }
        float* S = sm + (long)r * C + cv * 8;
        float* SS = sm + (long)R * C + (long)r * C + cv * 8;
        *reinterpret_cast<float4*>(S) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4*>(S + 4) = make_float4(s[4], s[5], s[6], s[7]);
        *reinterpret_cast<float4*>(SS) = make_float4(ss[0], ss[1], ss[2], ss[3]);
        *reinterpret_cast<float4*>(SS + 4) = make_float4(ss[4], ss[5], ss[6], ss[7]);
    }
    __syncthreads();
    // fold the R row-slots per channel (fixed order), in place into slot 0
    if (R > 1) {
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
            float* base = sm + (c < C ? c : (long)R * C + (c - C));
            float a = base[0];
            for (int rr = 1; rr < R; ++rr) a += base[(long)rr * C];
            base[0] = a;
        }
        __syncthreads();
    }
    // one warp per group: lanes over the group's channels, fixed-order butterfly
    int cpg = C / G;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < G; g += nwarps) {
        double a = 0.0, aa = 0.0;
        for (int i = lane; i < cpg; i += 32) {
            a += (double)sm[g * cpg + i];
            aa += (double)sm[(long)R * C + g * cpg + i];
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
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // cumulative: covers the partials the other warps wrote before the barrier
        int t = atomicAdd(&tickets[b], 1);
        s_last = (t == chunks - 1);
        if (s_last) tickets[b] = 0;  // self-resetting for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // All threads take part: thread = (group, slice of the chunk list), loads unrolled so that they are all in flight
    // (a dependent-load loop here is the tail of the whole kernel).  Fixed slice -> fixed order -> deterministic.
    double* red = reinterpret_cast<double*>(sm);  // [nslices][G][2]; the float partials are dead by now
    const int nslices = blockDim.x / G;
    {
        const int g = threadIdx.x % G, slice = threadIdx.x / G;
        if (slice < nslices) {
            double a = 0.0, aa = 0.0;
            const double* w = ws + ((long)b * chunks * G + g) * 2;
#pragma unroll 8
            for (int c = slice; c < chunks; c += nslices) {
                a += __ldcg(w + (long)c * G * 2);
                aa += __ldcg(w + (long)c * G * 2 + 1);
            }
            red[(slice * G + g) * 2] = a;
            red[(slice * G + g) * 2 + 1] = aa;
        }
    }
    __syncthreads();
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        double a = 0.0, aa = 0.0;
        for (int sl = 0; sl < nslices; ++sl) {
            a += red[(sl * G + g) * 2];
            aa += red[(sl * G + g) * 2 + 1];
        }
        double n = (double)HW * (C / G);
        double mean = a / n;
        double var = aa / n - mean * mean;
        if (var < 0) var = 0;
        stats[b * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
}

// apply: thread = fixed 8-channel vector (scale/shift folded once into registers), CTA = slab of pixel rows
template <typename T, bool SILU>
__global__ void __launch_bounds__(512, 2) gn_apply_k(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ gamma,
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
    for (long row = row0 + r; row < row1; row += 8L * R) {  // eight independent 16-byte loads in flight
        Pack<T, 8> raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            long rr = row + (long)u * R;
            raw[u] = *reinterpret_cast<const Pack<T, 8>*>(xb + (rr < row1 ? rr : row1 - 1) * C);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = fmaf(to_f<T>(raw[u].v[j]), sc[j], sh[j]);
                o[j] = SILU ? silu_f(t) : t;
            }
            long rr = row + (long)u * R;
            if (rr < row1) store8<T>(yb + rr * C, o);
        }
    }
}

void groupnorm(const void* x, void* y, const void* gamma, const void* beta, int B, long HW, int C, int groups,
               float eps, bool silu, int dtype, void* ws, cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0 && C % groups == 0 && groups <= 64, ETAI_ERR_ARG, "groupnorm: C%8, C%groups, groups<=64");
    ETAI_CHECK(C / 8 <= 512, ETAI_ERR_ARG, "groupnorm: C too large");
    ETAI_CHECK(ws != nullptr, ETAI_ERR_ARG, "groupnorm: workspace required");
    GnPlan p = gn_plan(HW, C);
    size_t smem = (size_t)2 * p.R * C * sizeof(float);
    if (smem < (size_t)p.threads * 2 * sizeof(double)) smem = (size_t)p.threads * 2 * sizeof(double);  // final-fold scratch
    // apply grid: at most one wave (4 resident CTAs of <= 512 threads per SM), each CTA whole batches of 8R pixel rows
    long per_batch = 8L * p.R;
    long nb = cdiv(cdiv(HW * B, 148L * 4), per_batch);
    if (nb < 1) nb = 1;
    int rows_per_cta = (int)(nb * per_batch);
    int ablocks = cdiv(HW, rows_per_cta);
    ETAI_CHECK(B <= 64, ETAI_ERR_ARG, "groupnorm: B <= 64");
    int* tickets = reinterpret_cast<int*>(ws);
    float2* stats = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) + GN_STATS_OFF);
    double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + GN_PART_OFF);
    ETAI_DISPATCH_DTYPE(dtype, T, {
        gn_stats_k<T><<<dim3(p.chunks, B), p.threads, smem, s>>>((const T*)x, part, stats, tickets, HW, C, groups, p.cvecs,
                                                                 p.R, p.rows_per_chunk, p.batches, p.chunks, eps);
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
template <typename T, int MAXV, int ROWS>
__global__ void layernorm_k(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ gamma,
                            const T* __restrict__ beta, long M, int C, float eps) {
    const long row0 = (blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
    const int lane = threadIdx.x & 31;
    if (row0 >= M) return;
    const int nv = C / 8;
    // all ROWS * MAXV 16-byte loads are issued before the first use; out-of-range slots re-read a valid address and
    // are masked (conditional loads would serialise the DRAM round trips)
    float v[ROWS][MAXV][8];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const long row = row0 + rr < M ? row0 + rr : M - 1;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int vi = lane + 32 * i;
            load8<T>(x + row * C + (vi < nv ? vi : 0) * 8, v[rr][i]);
        }
    }
    float mean[ROWS], rstd[ROWS];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const bool ok = lane + 32 * i < nv;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += ok ? v[rr][i][j] : 0.f;
        }
        mean[rr] = warp_sum(sum) / (float)C;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const bool ok = lane + 32 * i < nv;
#pragma unroll
            for (int j = 0; j < 8; ++j) { float d = ok ? v[rr][i][j] - mean[rr] : 0.f; sq = fmaf(d, d, sq); }
        }
        rstd[rr] = rsqrtf(warp_sum(sq) / (float)C + eps);
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nv) {
            float ga[8], be[8];
            load8<T>(gamma + vi * 8, ga);
            load8<T>(beta + vi * 8, be);
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                if (row0 + rr < M) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = (v[rr][i][j] - mean[rr]) * rstd[rr] * ga[j] + be[j];
                    store8<T>(y + (row0 + rr) * C + vi * 8, o);
                }
            }
        }
    }
}

void layernorm(const void* x, void* y, const void* gamma, const void* beta, long M, int C, float eps, int dtype,
               cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0 && C <= 8 * 32 * 8, ETAI_ERR_ARG, "layernorm: C%8==0 and C<=2048");
    const int warps = 8;
#define LN_LAUNCH(MAXV, ROWS)                                                                                          \
    layernorm_k<T, MAXV, ROWS><<<(unsigned)cdiv(M, (long)warps * ROWS), warps * 32, 0, s>>>(                           \
        (const T*)x, (T*)y, (const T*)gamma, (const T*)beta, M, C, eps)
    ETAI_DISPATCH_DTYPE(dtype, T, {
        if (C <= 8 * 32 * 2) LN_LAUNCH(2, 4);
        else if (C <= 8 * 32 * 5) LN_LAUNCH(5, 2);
        else LN_LAUNCH(8, 1);
    });
#undef LN_LAUNCH
    KERNEL_CHECK();
}

}  // namespace etai
