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
#include <cstdlib>
#include <cooperative_groups.h>
#include "ops.cuh"

namespace cg = cooperative_groups;

namespace etai {

// ---------------------------------------------------------------------------------------------------------------
// Single-launch GroupNorm(+SiLU) for UNet-sized tensors: x is read ONCE and y written once.
//   work unit   = (batch row b, block of GB consecutive groups) = a [HW x cw] slab of the NHWC tensor (cw = GB * C/G
//                 channels, contiguous cw*sizeof(T) bytes per pixel row);
//   cluster     = CS thread blocks split the slab's pixel rows; every CTA copies its rows into shared memory while it
//                 accumulates per-channel sum / sum-of-squares, publishes its per-group partials (double) in its own
//                 shared memory, and after one cluster barrier every CTA folds the CS partials of its cluster through
//                 distributed shared memory in rank order (fixed order: run-to-run and batch invariant), then normalises
//                 its rows straight out of shared memory.
// The partition (GB, CS) is a function of (HW, C, dtype) only, never of the batch size.  Slabs that do not fit in the
// cluster's shared memory re-read x from global/L2 in the apply pass ("resident = 0"); tensors too large for eight CTAs
// per unit (VAE resolutions) keep the two-launch path below.  Roofline: HBM / L2 bandwidth, 2 * elements * sizeof bytes.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GNC_THREADS = 256, GNC_UNROLL = 8;

struct GnCPlan {
    bool ok = false;
    int GB = 0, cw = 0, nv = 0, CS = 1, rows_per_cta = 0, resident = 0, RP = 0;
    size_t smem = 0;
};

static GnCPlan gn_cluster_plan(long HW, int C, int G, size_t esz) {
    GnCPlan best;
    const int cpg = C / G;
    const size_t tiers[3] = {48 * 1024, 96 * 1024, 200 * 1024};  // per-CTA slab budget: many CTAs / 2 per SM / 1 per SM
    for (int strict = 1; strict >= 0 && !best.ok; --strict) {
        for (int tier = 0; tier < 3 && !best.ok; ++tier) {
            for (int GB = 1; GB <= G && !best.ok; GB *= 2) {
                if (G % GB) continue;
                const int cw = GB * cpg;
                if (cw % 8 || cw / 8 > GNC_THREADS) continue;
                if (cw * esz < 128 || (strict && (cw * esz) % 32)) continue;  // >= one line per row; whole 32-byte sectors
                for (int CS = 1; CS <= 8; CS *= 2) {
                    const long rows = (HW + CS - 1) / CS;
                    if ((size_t)rows * cw * esz > tiers[tier]) continue;
                    best.ok = true; best.GB = GB; best.cw = cw; best.nv = cw / 8; best.CS = CS;
                    best.rows_per_cta = (int)rows; best.resident = 1;
                    break;
                }
            }
        }
    }
    if (!best.ok) {  // not resident: stats pass + apply pass both stream from global (the second read hits L2)
        for (int GB = 1; GB <= G && !best.ok; GB *= 2) {
            const int cw = GB * cpg;
            if (G % GB || cw % 8 || cw / 8 > GNC_THREADS || (cw * esz) % 32 || cw * esz < 128) continue;
            const long rows = (HW + 7) / 8;
            if (rows > 1024) continue;  // too few CTAs per unit for a tensor this large: two-launch path
            best.ok = true; best.GB = GB; best.cw = cw; best.nv = cw / 8; best.CS = 8; best.rows_per_cta = (int)rows;
            best.resident = 0;
        }
    }
    if (best.ok) {
        best.RP = GNC_THREADS / best.nv;
        size_t slab = best.resident ? (size_t)best.rows_per_cta * best.cw * esz : 0;
        size_t red = (size_t)best.RP * best.cw * 2 * sizeof(float);
        best.smem = ((slab + 15) & ~size_t(15)) + red + (size_t)best.GB * 2 * sizeof(double) + (size_t)best.GB * sizeof(float2) + 64;
    }
    return best;
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(GNC_THREADS) gn_cluster_k(const T* __restrict__ x, T* __restrict__ y,
                                                            const T* __restrict__ gamma, const T* __restrict__ beta,
                                                            long HW, int C, int cpg, int GB, int cw, int nv, int RP,
                                                            int rows_per_cta, int resident, float eps) {
    extern __shared__ __align__(32) unsigned char gsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CS = cluster.num_blocks(), rank = cluster.block_rank();
    const int gblock = blockIdx.x / CS, b = blockIdx.y;
    const int tid = threadIdx.x, v = tid % nv, rl = tid / nv;
    const bool active = rl < RP;
    const size_t slab_bytes = resident ? (((size_t)rows_per_cta * cw * sizeof(T) + 15) & ~size_t(15)) : 0;
    T* slab = reinterpret_cast<T*>(gsm);
    float* red = reinterpret_cast<float*>(gsm + slab_bytes);                    // [RP][2][cw]
    double* part = reinterpret_cast<double*>(gsm + slab_bytes + (size_t)RP * cw * 2 * sizeof(float));  // [GB][2]
    float2* stat = reinterpret_cast<float2*>(part + 2 * GB);                    // [GB] (mean, rstd)

    const long r0 = (long)rank * rows_per_cta;
    const long r1 = r0 + rows_per_cta < HW ? r0 + rows_per_cta : HW;
    const long col = (long)gblock * cw + v * 8;
    const T* xb = x + (long)b * HW * C + col;
    T* yb = y + (long)b * HW * C + col;

    // ---- pass 1: bring the rows in and accumulate per-channel partial sums in registers ----
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    if (active && r1 > r0) {
        if (resident) {
            // Every 16-byte piece of this thread's rows is put in flight at once with cp.async (global -> shared, no
            // register staging): ~80 KB outstanding per CTA.  With register-staged loads (4 x 16 B per thread per round
            // trip) the kernel was latency-bound at 1.5 TB/s.  Each thread later reads back only what it copied itself, so
            // cp.async.wait_all is the only synchronisation the slab needs.
            for (long row = r0 + rl; row < r1; row += RP) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + (row - r0) * cw + v * 8);
                const T* src = xb + row * C;
#pragma unroll
                for (int q = 0; q < (int)(8 * sizeof(T) / 16); ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * q),
                                 "l"(reinterpret_cast<const char*>(src) + 16 * q) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
            for (long row = r0 + rl; row < r1; row += RP) {
                const Pack<T, 8> raw = *reinterpret_cast<const Pack<T, 8>*>(slab + (row - r0) * cw + v * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = to_f<T>(raw.v[j]);
                    s[j] += t;
                    ss[j] = fmaf(t, t, ss[j]);
                }
            }
        } else {
            for (long row = r0 + rl; row < r1; row += (long)RP * GNC_UNROLL) {
                Pack<T, 8> raw[GNC_UNROLL];
#pragma unroll
                for (int u = 0; u < GNC_UNROLL; ++u) {  // branch-free loads (rows past the end re-read the last row, masked below)
                    const long rr = row + (long)u * RP;
                    raw[u] = *reinterpret_cast<const Pack<T, 8>*>(xb + (rr < r1 ? rr : r1 - 1) * C);
                }
#pragma unroll
                for (int u = 0; u < GNC_UNROLL; ++u) {
                    if (row + (long)u * RP < r1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float t = to_f<T>(raw[u].v[j]);
                            s[j] += t;
                            ss[j] = fmaf(t, t, ss[j]);
                        }
                    }
                }
            }
        }
    }
    if (active) {
        float* S = red + (long)rl * 2 * cw + v * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { S[j] = s[j]; S[cw + j] = ss[j]; }
    }
    __syncthreads();
    // fold the RP row lanes per channel (fixed order), then the channels of each group (double)
    for (int c = tid; c < 2 * cw; c += GNC_THREADS) {
        float a = red[c];
        for (int r = 1; r < RP; ++r) a += red[(long)r * 2 * cw + c];
        red[c] = a;
    }
    __syncthreads();
    if (tid < GB) {
        double a = 0.0, aa = 0.0;
        for (int i = 0; i < cpg; ++i) {
            a += (double)red[tid * cpg + i];
            aa += (double)red[cw + tid * cpg + i];
        }
        part[2 * tid] = a;
        part[2 * tid + 1] = aa;
    }
    cluster.sync();  // every CTA's partials are visible cluster-wide
    if (tid < GB) {
        double a = 0.0, aa = 0.0;
        for (unsigned r = 0; r < CS; ++r) {  // rank order: deterministic
            const double* rp = cluster.map_shared_rank(part, r);
            a += rp[2 * tid];
            aa += rp[2 * tid + 1];
        }
        const double n = (double)HW * cpg;
        const double mean = a / n;
        double var = aa / n - mean * mean;
        if (var < 0) var = 0;
        stat[tid] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    cluster.sync();  // remote reads are done (a CTA may now run ahead and exit); also orders stat[] for this CTA
    if (!active || r1 <= r0) return;
    // ---- pass 2: normalise + affine (+SiLU) out of shared memory ----
    float sc[8], sh[8];
    {
        float ga[8], be[8];
        load8<T>(gamma + col, ga);
        load8<T>(beta + col, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 st = stat[(v * 8 + j) / cpg];
            sc[j] = st.y * ga[j];
            sh[j] = be[j] - st.x * sc[j];
        }
    }
    for (long row = r0 + rl; row < r1; row += (long)RP * GNC_UNROLL) {
        Pack<T, 8> raw[GNC_UNROLL];
#pragma unroll
        for (int u = 0; u < GNC_UNROLL; ++u) {
            const long rr = row + (long)u * RP;
            const long rc = rr < r1 ? rr : r1 - 1;
            raw[u] = resident ? *reinterpret_cast<const Pack<T, 8>*>(slab + (rc - r0) * cw + v * 8)
                              : *reinterpret_cast<const Pack<T, 8>*>(xb + rc * C);
        }
#pragma unroll
        for (int u = 0; u < GNC_UNROLL; ++u) {
            const long rr = row + (long)u * RP;
            if (rr < r1) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = fmaf(to_f<T>(raw[u].v[j]), sc[j], sh[j]);
                    o[j] = SILU ? silu_for<T>(t) : t;
                }
                store8<T>(yb + rr * C, o);
            }
        }
    }
}

template <typename T, bool SILU>
static void gn_cluster_launch(const GnCPlan& p, const void* x, void* y, const void* gamma, const void* beta, int B, long HW,
                              int C, int G, float eps, cudaStream_t s) {
    static size_t configured = 0;  // largest dynamic shared-memory size this instantiation was opted in for
    if (p.smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_k<T, SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        configured = p.smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((G / p.GB) * p.CS), (unsigned)B, 1);
    cfg.blockDim = dim3(GNC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, gn_cluster_k<T, SILU>, (const T*)x, (T*)y, (const T*)gamma, (const T*)beta, HW, C,
                                  C / G, p.GB, p.cw, p.nv, p.RP, p.rows_per_cta, p.resident, eps));
}

// returns false when the shape is left to the two-launch path
static bool groupnorm_cluster(const void* x, void* y, const void* gamma, const void* beta, int B, long HW, int C, int G,
                              float eps, bool silu, int dtype, cudaStream_t s) {
    static const bool disabled = [] { const char* e = getenv("ETAI_GN_TWO_PASS"); return e && e[0] == '1'; }();
    if (disabled || B > 65535) return false;
    GnCPlan p = gn_cluster_plan(HW, C, G, dtype_size(dtype));
    if (!p.ok) return false;
    ETAI_DISPATCH_DTYPE(dtype, T, {
        if (silu) gn_cluster_launch<T, true>(p, x, y, gamma, beta, B, HW, C, G, eps, s);
        else gn_cluster_launch<T, false>(p, x, y, gamma, beta, B, HW, C, G, eps, s);
    });
    return true;
}


// ---------------------------------------------------------------------------------------------
// Direct GroupNorm for the low-resolution layers (16-bit storage): one CTA per (batch row, group), two passes over the
// group's HW x cpg elements (<= 123 KB: the second pass hits L1 / L2).  The cluster kernel above costs 13-24 us on these
// shapes whatever their size (ncu, profiles/r02_ncu_shapes_summary.txt: 17.9 us for the 5 MB of 8x8x2560 at B = 16) --
// cluster launch, cp.async slab and two cluster barriers are fixed latency; 30 of the UNet's 61 GroupNorms are this small.
// Measured (ncu, B = 16): 8x8x2560 17.9 -> 7.1 us, 16x16x1280 13.4 -> 11.2 us; at 32x32 the cluster kernel is as fast or
// faster (23.5 vs 22.4 us at C = 640, 50.6 vs 60.8 us at C = 1920), hence the HW <= 256 default.
// Partition and reduction order depend on (HW, C) only: batch-invariant and deterministic like the other paths.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__device__ __forceinline__ void ldvec(const T* p, float (&v)[VEC]) {
    if constexpr (VEC == 8) load8<T>(p, v);
    else if constexpr (VEC == 4) load4<T>(p, v);
    else { v[0] = to_f<T>(p[0]); v[1] = to_f<T>(p[1]); }
}
template <typename T, int VEC>
__device__ __forceinline__ void stvec(T* p, const float (&v)[VEC]) {
    if constexpr (VEC == 8) store8<T>(p, v);
    else if constexpr (VEC == 4) store4<T>(p, v);
    else { Pack<T, 2> o; o.v[0] = from_f<T>(v[0]); o.v[1] = from_f<T>(v[1]); *reinterpret_cast<Pack<T, 2>*>(p) = o; }
}

template <typename T, int VEC, bool SILU>
__global__ void __launch_bounds__(256) gn_direct_k(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ gamma,
                                                   const T* __restrict__ beta, int HW, int C, int G, float eps) {
    __shared__ double red[2][8];
    __shared__ float s_mean, s_rstd;
    const int b = blockIdx.y, g = blockIdx.x, cpg = C / G, vpr = cpg / VEC, nvec = HW * vpr;
    const T* xb = x + (long)b * HW * C + g * cpg;
    T* yb = y + (long)b * HW * C + g * cpg;
    float s = 0.f, ss = 0.f;
    for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * 256) {  // four independent loads in flight per thread
        float v[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256, ic = i < nvec ? i : nvec - 1;
            ldvec<T, VEC>(xb + (long)(ic / vpr) * C + (ic % vpr) * VEC, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * 256 < nvec) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) { s += v[u][j]; ss = fmaf(v[u][j], v[u][j], ss); }
            }
        }
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = (double)s; red[1][warp] = (double)ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, aa = 0.0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; aa += red[1][w]; }
        const double n = (double)HW * cpg, mean = a / n;
        double var = aa / n - mean * mean;
        if (var < 0) var = 0;
        s_mean = (float)mean;
        s_rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const float mean = s_mean, rstd = s_rstd;
    for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * 256) {
        float v[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256, ic = i < nvec ? i : nvec - 1;
            ldvec<T, VEC>(xb + (long)(ic / vpr) * C + (ic % vpr) * VEC, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256;
            if (i < nvec) {
                const int c = (i % vpr) * VEC;
                float ga[VEC], be[VEC], o[VEC];
                ldvec<T, VEC>(gamma + g * cpg + c, ga);
                ldvec<T, VEC>(beta + g * cpg + c, be);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float sc = rstd * ga[j];
                    const float t = fmaf(v[u][j], sc, be[j] - mean * sc);
                    o[j] = SILU ? silu_for<T>(t) : t;
                }
                stvec<T, VEC>(yb + (long)(i / vpr) * C + c, o);
            }
        }
    }
}

static int gn_direct_max_hw() {
    static const int v = [] { const char* e = getenv("ETAI_GN_DIRECT_MAX_HW"); return e ? atoi(e) : 256; }();
    return v;
}
static bool gn_direct_ok(long HW, int C, int G, int dtype) {
    return dtype != ETAI_F32 && HW <= gn_direct_max_hw() && C % G == 0 && (C / G) % 2 == 0 && G <= 65535;
}
template <typename T, bool SILU>
static void gn_direct_launch(const void* x, void* y, const void* gamma, const void* beta, int B, long HW, int C, int G, float eps,
                             cudaStream_t s) {
    const int cpg = C / G;
    dim3 grid(G, B);
#define GND(V) gn_direct_k<T, V, SILU><<<grid, 256, 0, s>>>((const T*)x, (T*)y, (const T*)gamma, (const T*)beta, (int)HW, C, G, eps)
    if (cpg % 8 == 0) GND(8);
    else if (cpg % 4 == 0) GND(4);
    else GND(2);
#undef GND
    KERNEL_CHECK();
}

int groupnorm_launches(long HW, int C, int groups, int dtype) {
    static const bool disabled = [] { const char* e = getenv("ETAI_GN_TWO_PASS"); return e && e[0] == '1'; }();
    if (gn_direct_ok(HW, C, groups, dtype)) return 1;
    if (disabled || C % 8 || C % groups) return 2;
    return gn_cluster_plan(HW, C, groups, dtype_size(dtype)).ok ? 1 : 2;
}

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
                o[j] = SILU ? silu_for<T>(t) : t;
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
    if (gn_direct_ok(HW, C, groups, dtype)) {
        ETAI_CHECK(B <= 65535, ETAI_ERR_ARG, "groupnorm: B too large");
        if (dtype == ETAI_F16) {
            if (silu) gn_direct_launch<__half, true>(x, y, gamma, beta, B, HW, C, groups, eps, s);
            else gn_direct_launch<__half, false>(x, y, gamma, beta, B, HW, C, groups, eps, s);
        } else {
            if (silu) gn_direct_launch<__nv_bfloat16, true>(x, y, gamma, beta, B, HW, C, groups, eps, s);
            else gn_direct_launch<__nv_bfloat16, false>(x, y, gamma, beta, B, HW, C, groups, eps, s);
        }
        return;
    }
    if (groupnorm_cluster(x, y, gamma, beta, B, HW, C, groups, eps, silu, dtype, s)) return;
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
