// fp32-accumulate SIMT GEMM / implicit-GEMM conv3x3 ("parity mode" back end, and the small odd layers of the
// 16-bit engine: conv_in with Cin=4 and conv_out with Cout=4).  C[M,N] = A[M,K] * W[N,K]^T + epilogue.
//
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-staged double buffering.
// This is the exactness reference on the device for the tcgen05 kernels (gemm_tc.cu) and the engine's fp32 path
// (north_star: "latents must match within 1e-3 max-abs per step in fp32"); its roofline is the fp32 FMA pipe.
#include "ops.cuh"

namespace etai {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256;

struct Params {
    const void* A;
    const void* W;
    void* C;
    const void* bias;
    const float* rowbias;
    const void* residual;
    long M;
    int N, K;
    long lda, ldc, ldr, ldrb, rows_per_group;
    int geglu;
    int H, Wd, Cin, stride, Ho, Wo, pad;
};

// A-tile loader: each thread owns row (tid/2) and 8 consecutive k at (tid%2)*8
template <typename T, bool CONV>
__device__ __forceinline__ void load_a(const Params& p, long m0, int k0, int tid, float (&v)[8]) {
    long m = m0 + (tid >> 1);
    int k = k0 + (tid & 1) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (m >= p.M || k >= p.K) return;
    const T* A = reinterpret_cast<const T*>(p.A);
    if constexpr (!CONV) {
        if (k + 8 <= p.K && (p.lda % 8) == 0) {
            load8<T>(A + m * p.lda + k, v);
        } else {
            for (int j = 0; j < 8 && k + j < p.K; ++j) v[j] = to_f<T>(A[m * p.lda + k + j]);
        }
    } else {
        int ox = (int)(m % p.Wo), oy = (int)((m / p.Wo) % p.Ho);
        long b = m / ((long)p.Wo * p.Ho);
        if ((p.Cin % 8) == 0) {
            int tap = k / p.Cin, c = k % p.Cin;
            int iy = oy * p.stride + tap / 3 - p.pad, ix = ox * p.stride + tap % 3 - p.pad;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.Wd) load8<T>(A + ((b * p.H + iy) * p.Wd + ix) * (long)p.Cin + c, v);
        } else {
            for (int j = 0; j < 8 && k + j < p.K; ++j) {
                int kk = k + j, tap = kk / p.Cin, c = kk % p.Cin;
                int iy = oy * p.stride + tap / 3 - p.pad, ix = ox * p.stride + tap % 3 - p.pad;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.Wd)
                    v[j] = to_f<T>(A[((b * p.H + iy) * p.Wd + ix) * (long)p.Cin + c]);
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ void load_w(const Params& p, int n0, int k0, int tid, float (&v)[8]) {
    int n = n0 + (tid >> 1);
    int k = k0 + (tid & 1) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (n >= p.N || k >= p.K) return;
    const T* W = reinterpret_cast<const T*>(p.W);
    if (k + 8 <= p.K && (p.K % 8) == 0) {
        load8<T>(W + (long)n * p.K + k, v);
    } else {
        for (int j = 0; j < 8 && k + j < p.K; ++j) v[j] = to_f<T>(W[(long)n * p.K + k + j]);
    }
}

template <typename T, bool CONV>
__global__ void __launch_bounds__(THREADS) gemm_simt_k(Params p) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];
    int tid = threadIdx.x;
    long m0 = (long)blockIdx.y * BM;
    int n0 = blockIdx.x * BN;
    int ty = tid / 16, tx = tid % 16;  // 16x16 thread grid; thread owns rows ty*4+{0..3}, 64+ty*4+{0..3}; cols likewise

    // Two-level accumulation: `acc` sums at most 8 k-tiles (128 products) before it is folded into `tot`.  A single fp32
    // running sum over K up to 23 040 drifts by ~sqrt(K) ulp per layer, which the 50 denoising steps of the parity tests
    // amplify (error x ~15 from x_T to x_0); blocked summation keeps the parity path at the accuracy of the CPU BLAS.
    float acc[8][8], tot[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }

    float ra[8], rw[8];
    int nk = (p.K + BK - 1) / BK;
    load_a<T, CONV>(p, m0, 0, tid, ra);
    load_w<T>(p, n0, 0, tid, rw);
    {
        int r = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { As[0][kb + j][r] = ra[j]; Ws[0][kb + j][r] = rw[j]; }
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        int cur = kt & 1;
        if (kt + 1 < nk) {
            load_a<T, CONV>(p, m0, (kt + 1) * BK, tid, ra);
            load_w<T>(p, n0, (kt + 1) * BK, tid, rw);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Ws[cur][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Ws[cur][kk][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            int r = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { As[cur ^ 1][kb + j][r] = ra[j]; Ws[cur ^ 1][kb + j][r] = rw[j]; }
        }
        if ((kt & 7) == 7 || kt + 1 == nk) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
        }
        __syncthreads();
    }

    // epilogue
    T* C = reinterpret_cast<T*>(p.C);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* res = reinterpret_cast<const T*>(p.residual);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        const float* rb = p.rowbias ? p.rowbias + (m / p.rows_per_group) * p.ldrb : nullptr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int n = n0 + h * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = tot[i][h * 4 + j];
                if (n + j < p.N) {
                    if (bias) v[j] += to_f<T>(bias[n + j]);
                    if (rb) v[j] += rb[n + j];
                }
            }
            if (p.geglu) {
                // columns (2j, 2j+1) = (value, gate); output column n/2
                if (n + 3 < p.N) {
                    T o0 = from_f<T>(v[0] * gelu_f(v[1])), o1 = from_f<T>(v[2] * gelu_f(v[3]));
                    C[m * p.ldc + n / 2] = o0;
                    C[m * p.ldc + n / 2 + 1] = o1;
                }
            } else {
                if (n + 3 < p.N && (p.ldc % 4) == 0 && (!res || (p.ldr % 4) == 0)) {
                    if (res) {
                        float r4[4];
                        load4<T>(res + m * p.ldr + n, r4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] += r4[j];
                    }
                    store4<T>(C + m * p.ldc + n, v);
                } else {
                    for (int j = 0; j < 4 && n + j < p.N; ++j) {
                        float o = v[j];
                        if (res) o += to_f<T>(res[m * p.ldr + n + j]);
                        C[m * p.ldc + n + j] = from_f<T>(o);
                    }
                }
            }
        }
    }
}

}  // namespace

void gemm_simt(const GemmArgs& a, cudaStream_t s) {
    ETAI_CHECK(a.M > 0 && a.N > 0 && a.K > 0, ETAI_ERR_ARG, "gemm: empty problem");
    ETAI_CHECK(!a.geglu || a.N % 4 == 0, ETAI_ERR_ARG, "gemm: geglu needs N%4==0");
    Params p;
    p.A = a.A; p.W = a.W; p.C = a.C; p.bias = a.bias; p.rowbias = (const float*)a.rowbias; p.residual = a.residual;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.lda = a.lda; p.ldc = a.ldc; p.ldr = a.ldr; p.ldrb = a.ldrb;
    p.rows_per_group = a.rows_per_group > 0 ? a.rows_per_group : 1;
    p.geglu = a.geglu;
    p.H = a.H; p.Wd = a.Wd; p.Cin = a.Cin; p.stride = a.stride; p.Ho = a.Ho; p.Wo = a.Wo; p.pad = a.pad;
    dim3 grid(cdiv(a.N, BN), cdiv(a.M, BM));
    ETAI_DISPATCH_DTYPE(a.dtype, T, {
        if (a.conv) gemm_simt_k<T, true><<<grid, THREADS, 0, s>>>(p);
        else gemm_simt_k<T, false><<<grid, THREADS, 0, s>>>(p);
    });
    KERNEL_CHECK();
}

}  // namespace etai
