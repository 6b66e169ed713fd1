// Small kernels used only by the VAE and the CLIP text tower (outside the diffusion loops: 1 encode + 2 decodes + 4 text
// encodes per edit, modules/inversion/diffusion_inversion.py:183-247 of the reference):
//   softmax_rows      in-place softmax(scale * x) over the rows of the materialised [N,N] logits of the VAE's single-head
//                     d=512 attention (the two products around it are ordinary GEMMs on the tensor cores)
//   clip_embed        token embedding gather + position embedding
//   quick_gelu        x * sigmoid(1.702 x)  (CLIP ViT-L/14 "quick_gelu")
//   clip_attention    causal 77-token self-attention, 12 heads of 64: one CTA per (batch row, head), K/V in shared memory
#include "ops.cuh"

namespace etai {

namespace {

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = is_max ? -INFINITY : 0.f;
    for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];  // fixed order: bit-reproducible
    __syncthreads();
    return r;
}

// one CTA per row; the row is kept in registers (n <= 256 threads * 8 * PER)
template <typename T, int PER>
__global__ void __launch_bounds__(256) softmax_rows_k(T* __restrict__ x, int n, float scale, float* __restrict__ lse) {
    __shared__ float sm[8];
    T* row = x + (long)blockIdx.x * n;
    float v[PER][8];
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
        const int c = (p * 256 + threadIdx.x) * 8;
        if (c < n) {
            load8<T>(row + c, v[p]);
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[p][j] *= scale; mx = fmaxf(mx, v[p][j]); }
        }
    }
    mx = block_reduce(mx, true, sm);
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
        const int c = (p * 256 + threadIdx.x) * 8;
        if (c < n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[p][j] = __expf(v[p][j] - mx); sum += v[p][j]; }
        }
    }
    sum = block_reduce(sum, false, sm);
    if (lse && threadIdx.x == 0) lse[blockIdx.x] = mx + __logf(sum);  // log-sum-exp of the scaled row (attention backward)
    const float inv = 1.0f / sum;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
        const int c = (p * 256 + threadIdx.x) * 8;
        if (c < n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[p][j] *= inv;
            store8<T>(row + c, v[p]);
        }
    }
}

template <typename T>
__global__ void clip_embed_k(const int* __restrict__ ids, const T* __restrict__ tok, const T* __restrict__ pos, T* __restrict__ out,
                             int L, int C, int vocab, long total) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over rows * C/8
    if (i >= total) return;
    const int cv = (int)(i % (C / 8));
    const long r = i / (C / 8);
    int id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    float a[8], b[8];
    load8<T>(tok + (long)id * C + cv * 8, a);
    load8<T>(pos + (long)(r % L) * C + cv * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    store8<T>(out + r * C + cv * 8, a);
}

template <typename T>
__global__ void quick_gelu_k(T* __restrict__ x, long nvec) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    float a[8];
    load8<T>(x + i * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = a[j] / (1.0f + __expf(-1.702f * a[j]));
    store8<T>(x + i * 8, a);
}

// qkv: [B*L, 3C] (q | k | v), head h at column h*D; out: [B*L, C].  grid (heads, B, query chunks), 256 threads = 8 warps, warp w
// takes query rows w, w+8, ... of its chunk; lane j owns keys j, j+32, j+64 for the scores and output columns j, j+32 for PV.
template <typename T, int D>
__global__ void __launch_bounds__(256) clip_attention_k(const T* __restrict__ qkv, T* __restrict__ out, int L, int C, float scale) {
    extern __shared__ float smf[];
    float* Ks = smf;                   // [L][D+1]
    float* Vs = Ks + L * (D + 1);      // [L][D]
    float* Ps = Vs + L * D;            // [8 warps][L padded to 96]
    const int h = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* base = qkv + (long)b * L * 3 * C + h * D;
    for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
        const int r = i / D, c = i % D;
        Ks[r * (D + 1) + c] = to_f<T>(base[(long)r * 3 * C + C + c]);
        Vs[r * D + c] = to_f<T>(base[(long)r * 3 * C + 2 * C + c]);
    }
    __syncthreads();
    float* P = Ps + warp * 96;
    const int qc = (L + gridDim.z - 1) / gridDim.z, q_lo = blockIdx.z * qc, q_hi = q_lo + qc < L ? q_lo + qc : L;
    for (int i = q_lo + warp; i < q_hi; i += 8) {
        float q[D / 32];  // this lane's slice of the query is not enough for a dot product: broadcast through shuffles
#pragma unroll
        for (int t = 0; t < D / 32; ++t) q[t] = to_f<T>(base[(long)i * 3 * C + t * 32 + lane]) * scale;
        float sc[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const int j = kk * 32 + lane;
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < D / 32; ++t)
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const float qv = __shfl_sync(0xffffffffu, q[t], l);
                    if (j < L) acc = fmaf(qv, Ks[j * (D + 1) + t * 32 + l], acc);
                }
            sc[kk] = (j < L && j <= i) ? acc : -INFINITY;  // causal mask (CLIPTextTransformer: keys j <= i)
        }
        float mx = warp_max(fmaxf(sc[0], fmaxf(sc[1], sc[2])));
        float e[3], sum = 0.f;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) { e[kk] = sc[kk] == -INFINITY ? 0.f : __expf(sc[kk] - mx); sum += e[kk]; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) P[kk * 32 + lane] = e[kk] * inv;
        __syncwarp();
        float o[D / 32];
#pragma unroll
        for (int t = 0; t < D / 32; ++t) o[t] = 0.f;
        for (int j = 0; j <= i; ++j) {
            const float pj = P[j];
#pragma unroll
            for (int t = 0; t < D / 32; ++t) o[t] = fmaf(pj, Vs[j * D + t * 32 + lane], o[t]);
        }
#pragma unroll
        for (int t = 0; t < D / 32; ++t) out[((long)b * L + i) * C + h * D + t * 32 + lane] = from_f<T>(o[t]);
        __syncwarp();
    }
}

}  // namespace

void softmax_rows(void* x, long rows, int n, float scale, int dtype, cudaStream_t s, float* lse) {
    ETAI_CHECK(n % 8 == 0 && n <= 256 * 8 * 4, ETAI_ERR_ARG, "softmax_rows: n%8==0 and n<=8192");
    ETAI_CHECK(rows >= 1 && rows < (1L << 31), ETAI_ERR_ARG, "softmax_rows: rows");
    ETAI_DISPATCH_DTYPE(dtype, T, {
        if (n <= 2048) softmax_rows_k<T, 1><<<(unsigned)rows, 256, 0, s>>>((T*)x, n, scale, lse);
        else if (n <= 4096) softmax_rows_k<T, 2><<<(unsigned)rows, 256, 0, s>>>((T*)x, n, scale, lse);
        else softmax_rows_k<T, 4><<<(unsigned)rows, 256, 0, s>>>((T*)x, n, scale, lse);
    });
    KERNEL_CHECK();
}

void clip_embed(const int* ids, const void* tok, const void* pos, void* out, long rows, int L, int C, int vocab, int dtype,
                cudaStream_t s) {
    ETAI_CHECK(C % 8 == 0, ETAI_ERR_ARG, "clip_embed: C%8");
    long total = rows * (C / 8);
    ETAI_DISPATCH_DTYPE(dtype, T, (clip_embed_k<T><<<cdiv(total, 256), 256, 0, s>>>(ids, (const T*)tok, (const T*)pos, (T*)out, L, C,
                                                                                 vocab, total)));
    KERNEL_CHECK();
}

void quick_gelu(void* x, long n, int dtype, cudaStream_t s) {
    ETAI_CHECK(n % 8 == 0, ETAI_ERR_ARG, "quick_gelu: n%8");
    ETAI_DISPATCH_DTYPE(dtype, T, (quick_gelu_k<T><<<cdiv(n / 8, 256), 256, 0, s>>>((T*)x, n / 8)));
    KERNEL_CHECK();
}

void clip_attention(const void* qkv, void* out, int B, int L, int heads, int d, float scale, int dtype, cudaStream_t s) {
    ETAI_CHECK(d == 64 && L <= 96, ETAI_ERR_UNSUPPORTED, "clip_attention: head dim 64, at most 96 tokens");
    const int C = heads * d;
    size_t smem = ((size_t)L * (d + 1) + (size_t)L * d + 8 * 96) * sizeof(float);
    ETAI_DISPATCH_DTYPE(dtype, T, {
        auto k = clip_attention_k<T, 64>;
        CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        k<<<dim3(heads, B, 5), 256, smem, s>>>((const T*)qkv, (T*)out, L, C, scale);  // 5 chunks of <= 16 queries: 2 per warp
    });
    KERNEL_CHECK();
}

}  // namespace etai
