// Data-gradient (dgrad) kernels of the UNet, used by null-text inversion only (reference:
// modules/inversion/null_text_inversion.py:75-80, `loss.backward()` through the UNet w.r.t. the [1,77,768] uncond
// embedding).  No weight gradients exist anywhere.  The contractions of the backward pass (conv dgrad = conv3x3 with the
// flipped/transposed filter, linear dgrad = GEMM with W^T) reuse the forward GEMM kernels on re-packed weights; this
// file holds what has no forward counterpart:
//   transposes of the packed weights, GroupNorm(+SiLU) / LayerNorm / GEGLU backward, flash-style self-attention backward
//   (dQ kernel + dK/dV kernel, probabilities recomputed tile by tile, nothing N x N materialised), cross-attention
//   backward over the 77 text keys with a fixed-order fold of the per-tile dK/dV partials, and the adjoints of concat,
//   nearest-2x upsample and the stride-2 downsample (zero stuffing).
// All math in fp32 on the SIMT pipes; B = 1 per step and at most 10 inner steps per DDIM step, so these kernels are sized
// for simplicity and determinism (no atomics), not for the tensor pipe.
#include "ops.cuh"

namespace etai {

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// weight re-packing
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_k(const T* __restrict__ in, T* __restrict__ out, int R, int Cc) {  // in [R][Cc] -> out [Cc][R]
    __shared__ T tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < Cc) tile[i][threadIdx.x] = in[(long)r * Cc + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Cc) out[(long)c * R + r] = tile[threadIdx.x][i];
    }
}

// packed conv filter [O][ky][kx][I] -> dgrad filter [I][2-ky][2-kx][O]
template <typename T>
__global__ void conv_flip_k(const T* __restrict__ in, T* __restrict__ out, int O, int I, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int o = (int)(i % O);
    long r = i / O;
    int tap = (int)(r % 9), ci = (int)(r / 9);
    out[i] = in[((long)o * 9 + (8 - tap)) * I + ci];
}

// ---------------------------------------------------------------------------------------------------------------------
// elementwise adjoints
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void zero_stuff2x_k(const T* __restrict__ in, T* __restrict__ out, int Ho, int Wo, int C, long total) {
    constexpr int V = 16 / sizeof(T);
    const int Cv = C / V;
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over the [B,2Ho,2Wo,C/V] output
    if (i >= total) return;
    int cv = (int)(i % Cv);
    long p = i / Cv;
    int x = (int)(p % (2 * Wo)), y = (int)((p / (2 * Wo)) % (2 * Ho));
    long b = p / (4L * Wo * Ho);
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((x & 1) == 0 && (y & 1) == 0) v = reinterpret_cast<const uint4*>(in + ((b * Ho + y / 2) * Wo + x / 2) * (long)C)[cv];
    reinterpret_cast<uint4*>(out)[i] = v;
}

template <typename T>
__global__ void upsample2x_bwd_k(const T* __restrict__ dout, T* __restrict__ din, int H, int W, int C, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over [B,H,W,C/8]
    if (i >= total) return;
    const int Cv = C / 8;
    int cv = (int)(i % Cv);
    long p = i / Cv;
    int x = (int)(p % W), y = (int)((p / W) % H);
    long b = p / ((long)W * H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, v[8];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            load8<T>(dout + ((b * 2 * H + 2 * y + dy) * (2L * W) + 2 * x + dx) * C + cv * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
    store8<T>(din + p * C + cv * 8, acc);
}

// out[r][0..C) (=|+=) in[r][off..off+C)
template <typename T>
__global__ void slice_cols_k(const T* __restrict__ in, long ld, int off, int C, T* __restrict__ out, int accumulate, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over rows * C/8
    if (i >= total) return;
    const int Cv = C / 8;
    int cv = (int)(i % Cv);
    long r = i / Cv;
    float v[8];
    load8<T>(in + r * ld + off + cv * 8, v);
    if (accumulate) {
        float o[8];
        load8<T>(out + r * C + cv * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += o[j];
    }
    store8<T>(out + r * C + cv * 8, v);
}

template <typename T>
__global__ void scale_k(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ s, int invert, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float f = invert ? 1.0f / s[0] : s[0];
    out[i] = from_f<T>(to_f<T>(in[i]) * f);
}

// s[0] = target / max|x| (1 if x == 0): the loss scale of the 16-bit backward pass, computed on the device
__global__ void __launch_bounds__(1024) absmax_scale_k(const float* __restrict__ x, long n, float target, float* __restrict__ s) {
    __shared__ float sm[32];
    float m = 0.f;
    for (long i = threadIdx.x; i < n; i += 1024) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = warp_max(sm[threadIdx.x]);
        if (threadIdx.x == 0) s[0] = m > 0.f ? target / m : 1.0f;
    }
}

// GEGLU on the un-fused projection u [M,2F] with (value, gate) interleaved columns: y[:,j] = u[:,2j] * gelu(u[:,2j+1])
template <typename T>
__global__ void geglu_fwd_k(const T* __restrict__ u, T* __restrict__ y, long total) {  // total = M*F/4
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    float v[8], o[4];
    load8<T>(u + i * 8, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = v[2 * j] * gelu_f(v[2 * j + 1]);
    store4<T>(y + i * 4, o);
}
template <typename T>
__global__ void geglu_bwd_k(const T* __restrict__ u, const T* __restrict__ dy, T* __restrict__ du, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    float v[8], d[4], o[8];
    load8<T>(u + i * 8, v);
    load4<T>(dy + i * 4, d);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], g = v[2 * j + 1];
        const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752f));
        const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
        o[2 * j] = d[j] * g * cdf;                  // d/da
        o[2 * j + 1] = d[j] * a * (cdf + g * pdf);  // d/dg
    }
    store8<T>(du + i * 8, o);
}

// ---------------------------------------------------------------------------------------------------------------------
// normalisation backward (statistics recomputed; fixed-order reductions)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_d(double v, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += sm[i];
    __syncthreads();
    return r;
}

// GroupNorm(+SiLU) backward in three launches of (chunks, groups, B) CTAs; each CTA owns a slab of pixel rows of one
// (batch row, group).  y = act(xhat * gamma + beta), act = SiLU or identity.
//   pass 0: per-chunk (sum x, sum x^2)                                   -> ws[0]
//   pass 1: mean / rstd from ws[0] (fixed order), per-chunk (sum dxhat, sum dxhat * xhat)  -> ws[1]
//   pass 2: dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
// (one CTA per (row, group) took 73 us per layer at B = 1: 32 CTAs on 148 SMs)
template <typename T, int PASS>
__global__ void __launch_bounds__(256) gn_bwd_k(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ gamma,
                                                const T* __restrict__ beta, T* __restrict__ dx, double* __restrict__ ws, long HW, int C,
                                                int G, float eps, int silu) {
    __shared__ double sm[16];
    const int chunk = blockIdx.x, chunks = gridDim.x, g = blockIdx.y, b = blockIdx.z, cpg = C / G;
    const long rows = (HW + chunks - 1) / chunks, r0 = chunk * rows, r1 = r0 + rows < HW ? r0 + rows : HW;
    const T* xb = x + (long)b * HW * C + g * cpg;
    const T* dyb = dy + (long)b * HW * C + g * cpg;
    T* dxb = dx + (long)b * HW * C + g * cpg;
    const long n_all = HW * cpg, i0 = r0 * cpg, i1 = r1 * cpg;
    double* w0 = ws + (((long)b * G + g) * chunks) * 2;                         // [chunks][2] of pass 0
    double* w1 = ws + ((long)gridDim.z * G * chunks + ((long)b * G + g) * chunks) * 2;  // [chunks][2] of pass 1
    float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
    if (PASS >= 1) {
        double s = 0.0, ss = 0.0;
        for (int c = 0; c < chunks; ++c) { s += w0[2 * c]; ss += w0[2 * c + 1]; }
        const double mean_d = s / n_all;
        double var = ss / n_all - mean_d * mean_d;
        if (var < 0.0) var = 0.0;
        mean = (float)mean_d;
        rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
    if (PASS == 2) {
        double a1 = 0.0, a2 = 0.0;
        for (int c = 0; c < chunks; ++c) { a1 += w1[2 * c]; a2 += w1[2 * c + 1]; }
        m1 = (float)(a1 / n_all);
        m2 = (float)(a2 / n_all);
    }
    double acc1 = 0.0, acc2 = 0.0;
    for (long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int c = (int)(i % cpg);
        const long off = (i / cpg) * C + c;
        const float xv = to_f<T>(xb[off]);
        if (PASS == 0) {
            acc1 += xv; acc2 += (double)xv * xv;
        } else {
            const float xh = (xv - mean) * rstd;
            const float gm = to_f<T>(gamma[g * cpg + c]);
            float d = to_f<T>(dyb[off]);
            if (silu) {
                const float z = fmaf(xh, gm, to_f<T>(beta[g * cpg + c]));
                const float sg = 1.0f / (1.0f + __expf(-z));
                d *= sg * (1.0f + z * (1.0f - sg));
            }
            const float dxh = d * gm;
            if (PASS == 1) { acc1 += dxh; acc2 += (double)dxh * xh; }
            else dxb[off] = from_f<T>(rstd * (dxh - m1 - xh * m2));
        }
    }
    if (PASS < 2) {
        acc1 = block_sum_d(acc1, sm);
        acc2 = block_sum_d(acc2, sm);
        if (threadIdx.x == 0) {
            double* w = PASS == 0 ? w0 : w1;
            w[2 * chunk] = acc1;
            w[2 * chunk + 1] = acc2;
        }
    }
}

// one warp per row
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_k(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ gamma,
                                                T* __restrict__ dx, long M, int C, float eps) {
    const long row = blockIdx.x * 8L + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const T* xr = x + row * C;
    const T* dr = dy + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += to_f<T>(xr[c]);
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = to_f<T>(xr[c]) - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    float a1 = 0.f, a2 = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float xh = (to_f<T>(xr[c]) - mean) * rstd, dxh = to_f<T>(dr[c]) * to_f<T>(gamma[c]);
        a1 += dxh; a2 += dxh * xh;
    }
    a1 = warp_sum(a1) / C; a2 = warp_sum(a2) / C;
    for (int c = lane; c < C; c += 32) {
        const float xh = (to_f<T>(xr[c]) - mean) * rstd, dxh = to_f<T>(dr[c]) * to_f<T>(gamma[c]);
        dx[row * C + c] = from_f<T>(rstd * (dxh - a1 - xh * a2));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention backward.  256 threads = 16 x 16; thread (ty,tx) owns rows ty+16i and columns tx+16j of every tile product.
// All shared-memory operand tiles are stored transposed ([k][row], row stride 65 / 97 words): reads by tx are
// conflict-free, reads by ty are broadcasts.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_tile_t(float* dst, int stride, const T* __restrict__ src, long ld, int rows, int D, int DP,
                                            int valid_rows) {
    // dst[k][r] = src[r][k] for r < rows, k < DP (zero beyond D / valid_rows)
    for (int i = threadIdx.x; i < rows * DP; i += 256) {
        const int r = i / DP, k = i % DP;
        dst[k * stride + r] = (k < D && r < valid_rows) ? to_f<T>(src[(long)r * ld + k]) : 0.f;
    }
}

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct AttnBwdParams {
    const void *q, *k, *v, *o, *dout;  // [B,N,ld*], head h at column h*d
    void *dq, *dk, *dv;                // same layouts as q, k, v
    float* lse;                        // [B,heads,N] (written by the dQ kernel, read by the dK/dV kernel)
    float* dsum;                       // [B,heads,N] rowsum(dO * O)
    int N, d;
    long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
    float scale;
};

// grid (N/64, heads, B).  pass 1: log-sum-exp of every query row; pass 2: dQ = sum_j dS_j K_j
template <typename T, int DP>
__global__ void __launch_bounds__(256) attn_bwd_dq_k(AttnBwdParams p) {
    extern __shared__ float smf[];
    constexpr int ST = 65, NC = DP / 16;
    float* Qt = smf;                 // [DP][65]
    float* dOt = Qt + DP * ST;
    float* Kt = dOt + DP * ST;
    float* Vt = Kt + DP * ST;
    float* Ss = Vt + DP * ST;        // [64 rows][65]: dS
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * 64, d = p.d, N = p.N;
    const T* qg = (const T*)p.q + ((long)b * N + q0) * p.ldq + h * d;
    const T* og = (const T*)p.o + ((long)b * N + q0) * p.ldo + h * d;
    const T* dog = (const T*)p.dout + ((long)b * N + q0) * p.lddo + h * d;
    load_tile_t<T>(Qt, ST, qg, p.ldq, 64, d, DP, 64);
    load_tile_t<T>(dOt, ST, dog, p.lddo, 64, d, DP, 64);
    // D[row] = sum_c dO * O
    float Dr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * i;
        float a = 0.f;
        for (int c = tx; c < d; c += 16) a += to_f<T>(dog[(long)r * p.lddo + c]) * to_f<T>(og[(long)r * p.ldo + c]);
        Dr[i] = half_warp_sum(a);
    }
    float mrow[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, lrow[4] = {0.f, 0.f, 0.f, 0.f};
    const int ntiles = N / 64;
    for (int pass = 0; pass < 2; ++pass) {
        float acc[4][NC];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;
        float lse[4];
        if (pass == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) lse[i] = mrow[i] + __logf(lrow[i]);
        }
        for (int j = 0; j < ntiles; ++j) {
            __syncthreads();
            const T* kg = (const T*)p.k + ((long)b * N + j * 64) * p.ldk + h * d;
            load_tile_t<T>(Kt, ST, kg, p.ldk, 64, d, DP, 64);
            if (pass == 1) {
                const T* vg = (const T*)p.v + ((long)b * N + j * 64) * p.ldv + h * d;
                load_tile_t<T>(Vt, ST, vg, p.ldv, 64, d, DP, 64);
            }
            __syncthreads();
            float s[4][4], dp[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) { s[i][jj] = 0.f; dp[i][jj] = 0.f; }
            for (int k = 0; k < DP; ++k) {
                float a[4], bb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Qt[k * ST + ty + 16 * i];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) bb[jj] = Kt[k * ST + tx + 16 * jj];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) s[i][jj] = fmaf(a[i], bb[jj], s[i][jj]);
                if (pass == 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = dOt[k * ST + ty + 16 * i];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) bb[jj] = Vt[k * ST + tx + 16 * jj];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) dp[i][jj] = fmaf(a[i], bb[jj], dp[i][jj]);
                }
            }
            if (pass == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3])) * p.scale;
                    mx = half_warp_max(mx);
                    const float mn = fmaxf(mrow[i], mx);
                    float e = 0.f;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) e += __expf(s[i][jj] * p.scale - mn);
                    e = half_warp_sum(e);
                    lrow[i] = lrow[i] * __expf(mrow[i] - mn) + e;
                    mrow[i] = mn;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float pr = __expf(s[i][jj] * p.scale - lse[i]);
                        Ss[(ty + 16 * i) * ST + tx + 16 * jj] = pr * (dp[i][jj] - Dr[i]) * p.scale;
                    }
                __syncthreads();
                for (int key = 0; key < 64; ++key) {
                    float a[4], bb[NC];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = Ss[(ty + 16 * i) * ST + key];
#pragma unroll
                    for (int c = 0; c < NC; ++c) bb[c] = Kt[(tx + 16 * c) * ST + key];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < NC; ++c) acc[i][c] = fmaf(a[i], bb[c], acc[i][c]);
                }
            }
        }
        if (pass == 1) {
            T* dqg = (T*)p.dq + ((long)b * N + q0) * p.lddq + h * d;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int col = tx + 16 * c;
                    if (col < d) dqg[(long)(ty + 16 * i) * p.lddq + col] = from_f<T>(acc[i][c]);
                }
            if (tx == 0) {
                const long base = ((long)b * gridDim.y + h) * N + q0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { p.lse[base + ty + 16 * i] = lse[i]; p.dsum[base + ty + 16 * i] = Dr[i]; }
            }
        }
    }
}

// grid (N/64 key tiles, heads, B): dK = sum_i dS_i^T Q_i, dV = sum_i P_i^T dO_i
template <typename T, int DP>
__global__ void __launch_bounds__(256) attn_bwd_dkv_k(AttnBwdParams p) {
    extern __shared__ float smf[];
    constexpr int ST = 65, NC = DP / 16;
    float* Qt = smf;
    float* dOt = Qt + DP * ST;
    float* Kt = dOt + DP * ST;
    float* Vt = Kt + DP * ST;
    float* Ps = Vt + DP * ST;        // [64 q rows][65]
    float* Ss = Ps + 64 * ST;        // dS
    float* lse_s = Ss + 64 * ST;     // [64]
    float* d_s = lse_s + 64;         // [64]
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int h = blockIdx.y, b = blockIdx.z, k0 = blockIdx.x * 64, d = p.d, N = p.N;
    load_tile_t<T>(Kt, ST, (const T*)p.k + ((long)b * N + k0) * p.ldk + h * d, p.ldk, 64, d, DP, 64);
    load_tile_t<T>(Vt, ST, (const T*)p.v + ((long)b * N + k0) * p.ldv + h * d, p.ldv, 64, d, DP, 64);
    float dk[4][NC], dv[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) { dk[i][c] = 0.f; dv[i][c] = 0.f; }
    const int ntiles = N / 64;
    for (int it = 0; it < ntiles; ++it) {
        __syncthreads();
        const long row0 = (long)b * N + it * 64;
        load_tile_t<T>(Qt, ST, (const T*)p.q + row0 * p.ldq + h * d, p.ldq, 64, d, DP, 64);
        load_tile_t<T>(dOt, ST, (const T*)p.dout + row0 * p.lddo + h * d, p.lddo, 64, d, DP, 64);
        if (threadIdx.x < 64) {
            const long base = ((long)b * gridDim.y + h) * N + it * 64;
            lse_s[threadIdx.x] = p.lse[base + threadIdx.x];
            d_s[threadIdx.x] = p.dsum[base + threadIdx.x];
        }
        __syncthreads();
        float s[4][4], dp[4][4];  // [query row ty+16i][key tx+16jj]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) { s[i][jj] = 0.f; dp[i][jj] = 0.f; }
        for (int k = 0; k < DP; ++k) {
            float a[4], bb[4], a2[4], b2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Qt[k * ST + ty + 16 * i]; a2[i] = dOt[k * ST + ty + 16 * i]; }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) { bb[jj] = Kt[k * ST + tx + 16 * jj]; b2[jj] = Vt[k * ST + tx + 16 * jj]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    s[i][jj] = fmaf(a[i], bb[jj], s[i][jj]);
                    dp[i][jj] = fmaf(a2[i], b2[jj], dp[i][jj]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int r = ty + 16 * i;
                const float pr = __expf(s[i][jj] * p.scale - lse_s[r]);
                Ps[r * ST + tx + 16 * jj] = pr;
                Ss[r * ST + tx + 16 * jj] = pr * (dp[i][jj] - d_s[r]) * p.scale;
            }
        __syncthreads();
        // this thread: keys ty+16i, columns tx+16c
        for (int r = 0; r < 64; ++r) {
            float ps[4], ds[4], qv[NC], dov[NC];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ps[i] = Ps[r * ST + ty + 16 * i]; ds[i] = Ss[r * ST + ty + 16 * i]; }
#pragma unroll
            for (int c = 0; c < NC; ++c) { qv[c] = Qt[(tx + 16 * c) * ST + r]; dov[c] = dOt[(tx + 16 * c) * ST + r]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    dk[i][c] = fmaf(ds[i], qv[c], dk[i][c]);
                    dv[i][c] = fmaf(ps[i], dov[c], dv[i][c]);
                }
        }
    }
    T* dkg = (T*)p.dk + ((long)b * N + k0) * p.lddk + h * d;
    T* dvg = (T*)p.dv + ((long)b * N + k0) * p.lddv + h * d;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int col = tx + 16 * c;
            if (col < d) {
                dkg[(long)(ty + 16 * i) * p.lddk + col] = from_f<T>(dk[i][c]);
                dvg[(long)(ty + 16 * i) * p.lddv + col] = from_f<T>(dv[i][c]);
            }
        }
}

// cross-attention backward over L <= 96 keys.  grid (N/TR, heads, B).  part: [B][tiles][L][2C] fp32 partial dK | dV
struct CrossBwdParams {
    const void *q, *kv, *dout;
    void* dq;
    float* part;
    int N, L, d, C;
    long ldq, ldkv, lddo, lddq;
    int koff, voff;
    float scale;
};
template <typename T, int DP, int TR>
__global__ void __launch_bounds__(256) cross_attn_bwd_k(CrossBwdParams p) {
    extern __shared__ float smf[];
    constexpr int SK = 97, SQ = TR + 1, NC = DP / 16, RI = TR / 16;
    float* Kt = smf;                 // [DP][97]
    float* Vt = Kt + DP * SK;
    float* Qt = Vt + DP * SK;        // [DP][TR+1]
    float* dOt = Qt + DP * SQ;
    float* Ps = dOt + DP * SQ;       // [TR][97]
    float* Ss = Ps + TR * SK;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * TR, d = p.d, N = p.N, L = p.L;
    const T* kvb = (const T*)p.kv + (long)b * L * p.ldkv;
    load_tile_t<T>(Kt, SK, kvb + p.koff + h * d, p.ldkv, 96, d, DP, L);
    load_tile_t<T>(Vt, SK, kvb + p.voff + h * d, p.ldkv, 96, d, DP, L);
    load_tile_t<T>(Qt, SQ, (const T*)p.q + ((long)b * N + q0) * p.ldq + h * d, p.ldq, TR, d, DP, TR);
    load_tile_t<T>(dOt, SQ, (const T*)p.dout + ((long)b * N + q0) * p.lddo + h * d, p.lddo, TR, d, DP, TR);
    __syncthreads();
    float s[RI][6], dp[RI][6];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
    for (int k = 0; k < DP; ++k) {
        float a[RI], a2[RI], bb[6], b2[6];
#pragma unroll
        for (int i = 0; i < RI; ++i) { a[i] = Qt[k * SQ + ty + 16 * i]; a2[i] = dOt[k * SQ + ty + 16 * i]; }
#pragma unroll
        for (int j = 0; j < 6; ++j) { bb[j] = Kt[k * SK + tx + 16 * j]; b2[j] = Vt[k * SK + tx + 16 * j]; }
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) { s[i][j] = fmaf(a[i], bb[j], s[i][j]); dp[i][j] = fmaf(a2[i], b2[j], dp[i][j]); }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 6; ++j) { s[i][j] = (tx + 16 * j < L) ? s[i][j] * p.scale : -INFINITY; mx = fmaxf(mx, s[i][j]); }
        mx = half_warp_max(mx);
        float e[6], sum = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) { e[j] = (tx + 16 * j < L) ? __expf(s[i][j] - mx) : 0.f; sum += e[j]; }
        sum = half_warp_sum(sum);
        const float inv = 1.0f / sum;
        float dd = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) { e[j] *= inv; dd += e[j] * dp[i][j]; }
        dd = half_warp_sum(dd);  // rowsum(P * dP) == rowsum(dO * O)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            Ps[(ty + 16 * i) * SK + tx + 16 * j] = e[j];
            Ss[(ty + 16 * i) * SK + tx + 16 * j] = e[j] * (dp[i][j] - dd) * p.scale;
        }
    }
    __syncthreads();
    // dQ[row ty+16i][col tx+16c] = sum_key dS K
    {
        float acc[RI][NC];
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;
        for (int key = 0; key < L; ++key) {
            float a[RI], bb[NC];
#pragma unroll
            for (int i = 0; i < RI; ++i) a[i] = Ss[(ty + 16 * i) * SK + key];
#pragma unroll
            for (int c = 0; c < NC; ++c) bb[c] = Kt[(tx + 16 * c) * SK + key];
#pragma unroll
            for (int i = 0; i < RI; ++i)
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[i][c] = fmaf(a[i], bb[c], acc[i][c]);
        }
        if (p.dq) {
            T* dqg = (T*)p.dq + ((long)b * N + q0) * p.lddq + h * d;
#pragma unroll
            for (int i = 0; i < RI; ++i)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int col = tx + 16 * c;
                    if (col < d) dqg[(long)(ty + 16 * i) * p.lddq + col] = from_f<T>(acc[i][c]);
                }
        }
    }
    // partial dK / dV [key ty+16i (i<6)][col tx+16c]
    {
        float dk[6][NC], dv[6][NC];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int c = 0; c < NC; ++c) { dk[i][c] = 0.f; dv[i][c] = 0.f; }
        for (int r = 0; r < TR; ++r) {
            float ps[6], ds[6], qv[NC], dov[NC];
#pragma unroll
            for (int i = 0; i < 6; ++i) { ps[i] = Ps[r * SK + ty + 16 * i]; ds[i] = Ss[r * SK + ty + 16 * i]; }
#pragma unroll
            for (int c = 0; c < NC; ++c) { qv[c] = Qt[(tx + 16 * c) * SQ + r]; dov[c] = dOt[(tx + 16 * c) * SQ + r]; }
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int c = 0; c < NC; ++c) { dk[i][c] = fmaf(ds[i], qv[c], dk[i][c]); dv[i][c] = fmaf(ps[i], dov[c], dv[i][c]); }
        }
        float* pg = p.part + (((long)b * gridDim.x + blockIdx.x) * L) * (2L * p.C);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int key = ty + 16 * i;
            if (key >= L) continue;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int col = tx + 16 * c;
                if (col < d) {
                    pg[(long)key * 2 * p.C + h * d + col] = dk[i][c];
                    pg[(long)key * 2 * p.C + p.C + h * d + col] = dv[i][c];
                }
            }
        }
    }
}

// dkv[b*L+key][kv_off + col2] = sum over tiles (fixed order) of part[b][tile][key][col2], col2 in [0, 2C)
__global__ void cross_bwd_fold_k(const float* __restrict__ part, float* __restrict__ dkv, int tiles, int L, int C2, long ld_dkv,
                                 int kv_off, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over B*L*C2
    if (i >= total) return;
    const int col = (int)(i % C2);
    const long r = i / C2;  // b*L + key
    const long b = r / L, key = r % L;
    float a = 0.f;
    for (int t = 0; t < tiles; ++t) a += part[((b * tiles + t) * L + key) * C2 + col];
    dkv[r * ld_dkv + kv_off + col] = a;
}


// ---------------------------------------------------------------------------------------------------------------------
// helpers of the GEMM-based self-attention backward (see etai_unet::self_attention_bwd_gemm)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void attn_pack_heads_k(const T* __restrict__ src, long ld, int N, int heads, int d, int DP, T* __restrict__ dst,
                                  T* __restrict__ dstT, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over heads * N * DP
    if (i >= total) return;
    const int c = (int)(i % DP);
    const long r = (i / DP) % N;
    const int h = (int)(i / ((long)DP * N));
    const T v = c < d ? src[r * ld + h * d + c] : from_f<T>(0.f);
    dst[i] = v;
    if (dstT) dstT[((long)h * DP + c) * N + r] = v;
}
template <typename T>
__global__ void attn_unpack_heads_k(const T* __restrict__ src, int N, int heads, int d, int DP, T* __restrict__ dst, long ld,
                                    long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over heads * N * d
    if (i >= total) return;
    const int c = (int)(i % d);
    const long r = (i / d) % N;
    const int h = (int)(i / ((long)d * N));
    dst[r * ld + h * d + c] = src[((long)h * N + r) * DP + c];
}
template <typename T>
__global__ void attn_rowdot_k(const T* __restrict__ a, long lda, const T* __restrict__ b, long ldb, int N, int heads, int d,
                              float* __restrict__ out) {
    const long w = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per (head, row)
    const int lane = threadIdx.x & 31;
    if (w >= (long)heads * N) return;
    const int h = (int)(w / N);
    const long r = w % N;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc += to_f<T>(a[r * lda + h * d + c]) * to_f<T>(b[r * ldb + h * d + c]);
    acc = warp_sum(acc);
    if (lane == 0) out[w] = acc;
}
template <typename T, int MODE>
__global__ void attn_bwd_elementwise_k(T* __restrict__ x, const T* __restrict__ p, const float* __restrict__ vec, int cols,
                                       float scale, long nvec) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over rows * cols / 8
    if (i >= nvec) return;
    const long e = i * 8;
    const int row = (int)(e / cols), col = (int)(e % cols);
    float v[8], pv[8];
    load8<T>(x + e, v);
    if (MODE != 1) load8<T>(p + e, pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (MODE == 0) v[j] = pv[j] * (v[j] - vec[row]) * scale;
        else if (MODE == 1) v[j] = __expf(v[j] * scale - vec[col + j]);
        else v[j] = pv[j] * (v[j] - vec[col + j]) * scale;
    }
    store8<T>(x + e, v);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------------
void transpose_2d(const void* in, void* out, int R, int Cc, int dtype, cudaStream_t s) {
    dim3 grid(cdiv(Cc, 32), cdiv(R, 32)), block(32, 8);
    ETAI_DISPATCH_DTYPE(dtype, T, (transpose_k<T><<<grid, block, 0, s>>>((const T*)in, (T*)out, R, Cc)));
    KERNEL_CHECK();
}
void conv_weight_flip(const void* in, void* out, int O, int I, int dtype, cudaStream_t s) {
    long total = (long)O * 9 * I;
    ETAI_DISPATCH_DTYPE(dtype, T, (conv_flip_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)in, (T*)out, O, I, total)));
    KERNEL_CHECK();
}
void zero_stuff2x(const void* in, void* out, int B, int Ho, int Wo, int C, int dtype, cudaStream_t s) {
    ETAI_DISPATCH_DTYPE(dtype, T, {
        constexpr int V = 16 / sizeof(T);
        ETAI_CHECK(C % V == 0, ETAI_ERR_ARG, "zero_stuff2x: C must be a multiple of 16 bytes");
        long total = (long)B * 4 * Ho * Wo * (C / V);
        zero_stuff2x_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)in, (T*)out, Ho, Wo, C, total);
    });
    KERNEL_CHECK();
}
void upsample2x_bwd(const void* dout, void* din, int B, int H, int W, int C, int dtype, cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0, ETAI_ERR_ARG, "upsample2x_bwd: C%8");
    long total = (long)B * H * W * (C / 8);
    ETAI_DISPATCH_DTYPE(dtype, T, (upsample2x_bwd_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)dout, (T*)din, H, W, C, total)));
    KERNEL_CHECK();
}
void slice_cols(const void* in, long ld, int off, int C, void* out, bool accumulate, long rows, int dtype, cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0 && off % 8 == 0 && ld % 8 == 0, ETAI_ERR_ARG, "slice_cols: 8-element alignment");
    long total = rows * (C / 8);
    ETAI_DISPATCH_DTYPE(dtype, T, (slice_cols_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)in, ld, off, C, (T*)out,
                                                                                 accumulate ? 1 : 0, total)));
    KERNEL_CHECK();
}
void scale_by_device_scalar(const void* in, void* out, const float* s_dev, bool invert, long n, int dtype, cudaStream_t s) {
    ETAI_DISPATCH_DTYPE(dtype, T, (scale_k<T><<<cdiv(n, 256), 256, 0, s>>>((const T*)in, (T*)out, s_dev, invert ? 1 : 0, n)));
    KERNEL_CHECK();
}
void absmax_scale(const float* x, long n, float target, float* s_dev, cudaStream_t s) {
    absmax_scale_k<<<1, 1024, 0, s>>>(x, n, target, s_dev);
    KERNEL_CHECK();
}
void geglu_fwd(const void* u, void* y, long M, int F, int dtype, cudaStream_t s) {
    ETAI_CHECK(F % 4 == 0, ETAI_ERR_ARG, "geglu: F%4");
    long total = M * F / 4;
    ETAI_DISPATCH_DTYPE(dtype, T, (geglu_fwd_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)u, (T*)y, total)));
    KERNEL_CHECK();
}
void geglu_bwd(const void* u, const void* dy, void* du, long M, int F, int dtype, cudaStream_t s) {
    ETAI_CHECK(F % 4 == 0, ETAI_ERR_ARG, "geglu: F%4");
    long total = M * F / 4;
    ETAI_DISPATCH_DTYPE(dtype, T, (geglu_bwd_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)u, (const T*)dy, (T*)du, total)));
    KERNEL_CHECK();
}
size_t groupnorm_bwd_workspace_bytes(int B, int groups) { return (size_t)2 * B * groups * 32 * 2 * sizeof(double); }

void groupnorm_bwd(const void* x, const void* dy, const void* gamma, const void* beta, void* dx, int B, long HW, int C,
                   int groups, float eps, bool silu, int dtype, void* ws, cudaStream_t s) {
    ETAI_CHECK(C % groups == 0 && ws != nullptr, ETAI_ERR_ARG, "groupnorm_bwd: C%groups, workspace");
    int chunks = (int)((HW * (C / groups) + 8191) / 8192);  // ~8k elements per CTA
    if (chunks > 32) chunks = 32;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, groups, B);
    ETAI_DISPATCH_DTYPE(dtype, T, {
        gn_bwd_k<T, 0><<<grid, 256, 0, s>>>((const T*)x, (const T*)dy, (const T*)gamma, (const T*)beta, (T*)dx, (double*)ws, HW, C,
                                           groups, eps, silu ? 1 : 0);
        gn_bwd_k<T, 1><<<grid, 256, 0, s>>>((const T*)x, (const T*)dy, (const T*)gamma, (const T*)beta, (T*)dx, (double*)ws, HW, C,
                                           groups, eps, silu ? 1 : 0);
        gn_bwd_k<T, 2><<<grid, 256, 0, s>>>((const T*)x, (const T*)dy, (const T*)gamma, (const T*)beta, (T*)dx, (double*)ws, HW, C,
                                           groups, eps, silu ? 1 : 0);
    });
    KERNEL_CHECK();
}
void layernorm_bwd(const void* x, const void* dy, const void* gamma, void* dx, long M, int C, float eps, int dtype,
                   cudaStream_t s) {
    ETAI_DISPATCH_DTYPE(dtype, T, (ln_bwd_k<T><<<cdiv(M, 8), 256, 0, s>>>((const T*)x, (const T*)dy, (const T*)gamma, (T*)dx, M, C,
                                                                        eps)));
    KERNEL_CHECK();
}


void attn_pack_heads(const void* src, long ld, int N, int heads, int d, int DP, void* dst, void* dstT, int dtype, cudaStream_t s) {
    long total = (long)heads * N * DP;
    ETAI_DISPATCH_DTYPE(dtype, T, (attn_pack_heads_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)src, ld, N, heads, d, DP, (T*)dst,
                                                                                     (T*)dstT, total)));
    KERNEL_CHECK();
}
void attn_unpack_heads(const void* src, int N, int heads, int d, int DP, void* dst, long ld, int dtype, cudaStream_t s) {
    long total = (long)heads * N * d;
    ETAI_DISPATCH_DTYPE(dtype, T, (attn_unpack_heads_k<T><<<cdiv(total, 256), 256, 0, s>>>((const T*)src, N, heads, d, DP, (T*)dst, ld,
                                                                                       total)));
    KERNEL_CHECK();
}
void attn_rowdot(const void* a, long lda, const void* b, long ldb, int N, int heads, int d, float* out, int dtype, cudaStream_t s) {
    long warps = (long)heads * N;
    ETAI_DISPATCH_DTYPE(dtype, T, (attn_rowdot_k<T><<<cdiv(warps, 8), 256, 0, s>>>((const T*)a, lda, (const T*)b, ldb, N, heads, d, out)));
    KERNEL_CHECK();
}
void attn_bwd_elementwise(void* x, const void* p, const float* vec, int rows, int cols, float scale, int mode, int dtype,
                          cudaStream_t s) {
    ETAI_CHECK(cols % 8 == 0 && mode >= 0 && mode <= 2, ETAI_ERR_ARG, "attn_bwd_elementwise: cols%8, mode");
    long nvec = (long)rows * cols / 8;
    ETAI_DISPATCH_DTYPE(dtype, T, {
        if (mode == 0) attn_bwd_elementwise_k<T, 0><<<cdiv(nvec, 256), 256, 0, s>>>((T*)x, (const T*)p, vec, cols, scale, nvec);
        else if (mode == 1) attn_bwd_elementwise_k<T, 1><<<cdiv(nvec, 256), 256, 0, s>>>((T*)x, (const T*)p, vec, cols, scale, nvec);
        else attn_bwd_elementwise_k<T, 2><<<cdiv(nvec, 256), 256, 0, s>>>((T*)x, (const T*)p, vec, cols, scale, nvec);
    });
    KERNEL_CHECK();
}

static int pad_head_dim(int d) { return d <= 48 ? 48 : d <= 80 ? 80 : d <= 160 ? 160 : -1; }

void attention_bwd(const SelfAttnBwdArgs& a, cudaStream_t s) {
    const int DP = pad_head_dim(a.d);
    ETAI_CHECK(DP > 0 && a.N % 64 == 0, ETAI_ERR_UNSUPPORTED, "attention_bwd: head dim <= 160 and N % 64 == 0");
    AttnBwdParams p;
    p.q = a.q; p.k = a.k; p.v = a.v; p.o = a.o; p.dout = a.dout; p.dq = a.dq; p.dk = a.dk; p.dv = a.dv;
    p.lse = a.lse; p.dsum = a.dsum; p.N = a.N; p.d = a.d;
    p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldo = a.ldo; p.lddo = a.lddo; p.lddq = a.lddq; p.lddk = a.lddk; p.lddv = a.lddv;
    p.scale = a.scale;
    dim3 grid(a.N / 64, a.heads, a.B);
    const size_t sm_q = ((size_t)4 * DP * 65 + 64 * 65) * sizeof(float);
    const size_t sm_kv = ((size_t)4 * DP * 65 + 2 * 64 * 65 + 128) * sizeof(float);
#define LAUNCH_ATT(T, DPV)                                                                                              \
    do {                                                                                                                \
        CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_k<T, DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_q));   \
        CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_k<T, DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_kv)); \
        attn_bwd_dq_k<T, DPV><<<grid, 256, sm_q, s>>>(p);                                                                \
        attn_bwd_dkv_k<T, DPV><<<grid, 256, sm_kv, s>>>(p);                                                              \
    } while (0)
    ETAI_DISPATCH_DTYPE(a.dtype, T, {
        if (DP == 48) LAUNCH_ATT(T, 48);
        else if (DP == 80) LAUNCH_ATT(T, 80);
        else LAUNCH_ATT(T, 160);
    });
#undef LAUNCH_ATT
    KERNEL_CHECK();
}

size_t cross_attention_bwd_partial_bytes(int B, int N, int L, int C, int d) {
    const int TR = pad_head_dim(d) == 160 ? 32 : 64;
    return (size_t)B * (N / TR) * L * 2 * C * sizeof(float);
}

void cross_attention_bwd(const CrossAttnBwdArgs& a, cudaStream_t s) {
    const int DP = pad_head_dim(a.d);
    ETAI_CHECK(DP > 0 && a.L <= 96, ETAI_ERR_UNSUPPORTED, "cross_attention_bwd: head dim <= 160, at most 96 keys");
    const int TR = DP == 160 ? 32 : 64;
    ETAI_CHECK(a.N % TR == 0, ETAI_ERR_UNSUPPORTED, "cross_attention_bwd: N must be a multiple of the query tile");
    CrossBwdParams p;
    p.q = a.q; p.kv = a.kv; p.dout = a.dout; p.dq = a.dq; p.part = a.part;
    p.N = a.N; p.L = a.L; p.d = a.d; p.C = a.heads * a.d;
    p.ldq = a.ldq; p.ldkv = a.ldkv; p.lddo = a.lddo; p.lddq = a.lddq; p.koff = a.koff; p.voff = a.voff; p.scale = a.scale;
    dim3 grid(a.N / TR, a.heads, a.B);
    const size_t sm = ((size_t)2 * DP * 97 + 2 * DP * (TR + 1) + 2 * TR * 97) * sizeof(float);
#define LAUNCH_X(T, DPV, TRV)                                                                                          \
    do {                                                                                                               \
        CUDA_CHECK(cudaFuncSetAttribute(cross_attn_bwd_k<T, DPV, TRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        cross_attn_bwd_k<T, DPV, TRV><<<grid, 256, sm, s>>>(p);                                                         \
    } while (0)
    ETAI_DISPATCH_DTYPE(a.dtype, T, {
        if (DP == 48) LAUNCH_X(T, 48, 64);
        else if (DP == 80) LAUNCH_X(T, 80, 64);
        else LAUNCH_X(T, 160, 32);
    });
#undef LAUNCH_X
    KERNEL_CHECK();
    const int C2 = 2 * p.C;
    long total = (long)a.B * a.L * C2;
    cross_bwd_fold_k<<<cdiv(total, 256), 256, 0, s>>>(a.part, a.dkv, a.N / TR, a.L, C2, a.ld_dkv, a.kv_off, total);
    KERNEL_CHECK();
}

}  // namespace etai
