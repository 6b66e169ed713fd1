// SIMT flash-style self-attention (fp32 math) with per-row (q,k,v) source remap.
//
// out[r] = softmax(Q[qrow[r]] K[krow[r]]^T * scale) V[vrow[r]]
// The remap is how prompt-to-prompt self-attention replacement (ptp.py:194-200), MasaCtrl mutual self-attention
// (masactrl.py:56-72) and PnP q/k injection (pnp_utils.py:76-88) are expressed: no probability matrix is ever
// materialised (the reference materialises sim and attn of [B*8,4096,4096], ptp_utils.py:238-247).
//
// CTA = 128 threads, 64 queries x one (row, head); K/V streamed in 64-key tiles through shared memory with an online
// softmax.  Parity/back-up path; the 16-bit engine uses attn_tc.cu.
#include "ops.cuh"

namespace etai {

namespace {

constexpr int BQ = 64, BKV = 64, ATHREADS = 128;

template <typename T>
__global__ void __launch_bounds__(ATHREADS) attn_simt_k(SelfAttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int d = a.d, pitch = d + 4;
    float* Qs = smem;                  // [BQ][pitch]
    float* Ks = Qs + BQ * pitch;       // [BKV][pitch]
    float* Vs = Ks + BKV * pitch;      // [BKV][pitch]
    float* Ps = Vs + BKV * pitch;      // [BQ][BKV+4]
    constexpr int PP = BKV + 4;

    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * BQ;
    const int head = blockIdx.y, row = blockIdx.z;
    const T* Q = reinterpret_cast<const T*>(a.q) + (long)a.map.q[row] * a.Nq * a.ldq + head * d;
    const T* K = reinterpret_cast<const T*>(a.k) + (long)a.map.k[row] * a.Nk * a.ldk + head * d;
    const T* V = reinterpret_cast<const T*>(a.v) + (long)a.map.v[row] * a.Nk * a.ldv + head * d;
    T* O = reinterpret_cast<T*>(a.out) + (long)row * a.Nq * a.ldo + head * d;

    const int dv = d / 8;  // 8-element vectors per row
    for (int i = tid; i < BQ * dv; i += ATHREADS) {
        int r = i / dv, c = (i % dv) * 8;
        float v[8];
        if (q0 + r < a.Nq) load8<T>(Q + (long)(q0 + r) * a.ldq + c, v);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) Qs[r * pitch + c + j] = v[j] * a.scale;
    }

    const int ty = tid / 8, tx = tid % 8;  // rows i*16+ty (i<4), key columns j*8+tx (j<8), out columns j*8+tx
    const int ncol = d / 8;                // output columns per thread (5, 10, 20 for d = 40, 80, 160)
    float o[4][20];
    float mrow[4], lrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mrow[i] = -INFINITY;
        lrow[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 20; ++j) o[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < a.Nk; k0 += BKV) {
        __syncthreads();  // previous tile fully consumed (also orders the Q fill on the first pass)
        for (int i = tid; i < BKV * dv; i += ATHREADS) {
            int r = i / dv, c = (i % dv) * 8;
            float kv[8], vv[8];
            if (k0 + r < a.Nk) {
                load8<T>(K + (long)(k0 + r) * a.ldk + c, kv);
                load8<T>(V + (long)(k0 + r) * a.ldv + c, vv);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) kv[j] = vv[j] = 0.f;
            }
            *reinterpret_cast<float4*>(&Ks[r * pitch + c]) = make_float4(kv[0], kv[1], kv[2], kv[3]);
            *reinterpret_cast<float4*>(&Ks[r * pitch + c + 4]) = make_float4(kv[4], kv[5], kv[6], kv[7]);
            *reinterpret_cast<float4*>(&Vs[r * pitch + c]) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            *reinterpret_cast<float4*>(&Vs[r * pitch + c + 4]) = make_float4(vv[4], vv[5], vv[6], vv[7]);
        }
        __syncthreads();

        // S = Q K^T  (4 rows x 8 keys per thread)
        float s[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
        for (int c = 0; c < d; c += 4) {
            float4 qv[4], kv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[(i * 16 + ty) * pitch + c]);
#pragma unroll
            for (int j = 0; j < 8; ++j) kv[j] = *reinterpret_cast<const float4*>(&Ks[(j * 8 + tx) * pitch + c]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
                    s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
                    s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
                    s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
                }
        }
        // online softmax; the 8 lanes sharing ty are consecutive lanes of one warp
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (k0 + j * 8 + tx >= a.Nk) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            float mnew = fmaxf(mrow[i], mx);
            float alpha = expf(mrow[i] - mnew);  // exp(-inf) = 0 on the first tile
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float pv = expf(s[i][j] - mnew);
                Ps[(i * 16 + ty) * PP + j * 8 + tx] = pv;
                sum += pv;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            lrow[i] = lrow[i] * alpha + sum;
            mrow[i] = mnew;
#pragma unroll
            for (int j = 0; j < 20; ++j)
                if (j < ncol) o[i][j] *= alpha;
        }
        __syncthreads();
        // O += P V
        for (int key = 0; key < BKV; ++key) {
            float pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = Ps[(i * 16 + ty) * PP + key];
#pragma unroll
            for (int j = 0; j < 20; ++j) {
                if (j < ncol) {
                    float vv = Vs[key * pitch + j * 8 + tx];
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i][j] = fmaf(pv[i], vv, o[i][j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int q = q0 + i * 16 + ty;
        if (q >= a.Nq) continue;
        float inv = 1.f / lrow[i];
#pragma unroll
        for (int j = 0; j < 20; ++j)
            if (j < ncol) O[(long)q * a.ldo + j * 8 + tx] = from_f<T>(o[i][j] * inv);
    }
}

}  // namespace

void attention_simt(const SelfAttnArgs& a, cudaStream_t s) {
    ETAI_CHECK(a.d % 8 == 0 && a.d <= 160, ETAI_ERR_ARG, "attention: head dim must be a multiple of 8 and <= 160");
    ETAI_CHECK(a.B <= ETAI_MAX_ROWS, ETAI_ERR_ARG, "attention: too many rows");
    size_t smem = ((size_t)(BQ + 2 * BKV) * (a.d + 4) + (size_t)BQ * (BKV + 4)) * sizeof(float);
    dim3 grid(cdiv(a.Nq, BQ), a.heads, a.B);
    ETAI_DISPATCH_DTYPE(a.dtype, T, {
        CUDA_CHECK(cudaFuncSetAttribute(attn_simt_k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_simt_k<T><<<grid, ATHREADS, smem, s>>>(a);
    });
    KERNEL_CHECK();
}

}  // namespace etai
