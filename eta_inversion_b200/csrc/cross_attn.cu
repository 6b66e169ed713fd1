// Cross-attention over the text context (L = 77 keys) with prompt-to-prompt control fused in.
//
// Because a whole probability row (77 values) fits in three registers per lane, the controller semantics of the
// reference are applied between softmax and P.V inside the kernel instead of on a materialised [B*8, N, 77] tensor:
//   * AttentionStore            (ptp.py:150-171): acc[slot][pix][w] += sum_heads P'[pix][w]      (post-edit P)
//   * AttentionReplace/Refine/Reweight (ptp.py:205-211,234-274):
//         F[n]  = eq[n] * ( a[n] * sum_w P_base[w] M[w][n] + (1-a[n]) * P_tgt[n] )
//         P'[n] = alpha[n] * F[n] + (1-alpha[n]) * P_tgt[n]                    (no renormalisation)
// Work split: CTA = (block of QB queries) x (group); a group is a plain UNet row or a (base,target) pair.  The CTA
// loops over heads so that per-pixel sums over heads are accumulated by their single owner warp in a fixed order
// (deterministic, no atomics).  K_h/V_h of the current row are staged in shared memory in storage precision.
#include "ops.cuh"

namespace etai {

namespace {

constexpr int QB = 64, CWARPS = 8, CTHREADS = CWARPS * 32, LMAX = 80;  // L <= 80 keys (3 per lane); 77 text tokens

template <typename T>
__global__ void __launch_bounds__(CTHREADS) cross_attn_k(CrossAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int d = a.d, L = a.L;
    const int kp = d + (sizeof(T) == 4 ? 1 : 2);  // padded K row pitch (elements): odd number of 32-bit words
    T* Ksm = reinterpret_cast<T*>(smraw);                         // [L][kp]
    T* Vsm = Ksm + (size_t)LMAX * kp;                             // [L][kp]
    float* Pbase = reinterpret_cast<float*>(Vsm + (size_t)LMAX * kp);  // [QB][LMAX] base-row probabilities (edit)
    float* Acc = Pbase + QB * LMAX;                               // [2][QB][LMAX] store accumulators (base,tgt)
    float* Map = Acc + 2 * QB * LMAX;                             // [L][L] mapper (edit)
    float* Wq = Map + LMAX * LMAX;                                // [CWARPS][160] query scratch
    float* Wp = Wq + CWARPS * 160;                                // [CWARPS][LMAX] probability scratch

    const CrossGroup g = a.groups[blockIdx.y];
    const int q0 = blockIdx.x * QB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool edit = g.tgt >= 0 && a.mapper != nullptr;
    const int nrows = g.tgt >= 0 ? 2 : 1;

    float al[3] = {0, 0, 0}, eq[3] = {1, 1, 1}, ba[3] = {1, 1, 1};
    if (edit) {
        for (int i = threadIdx.x; i < L * L; i += CTHREADS) Map[(i / L) * LMAX + i % L] = a.mapper[(long)g.pair * L * L + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            int n = lane + 32 * j;
            if (n < L) {
                al[j] = a.alpha_step[g.pair * L + n];
                eq[j] = a.equalizer[g.pair * L + n];
                ba[j] = a.blend_a[g.pair * L + n];
            }
        }
    }
    const bool st0 = a.store && g.store_base >= 0, st1 = a.store && g.store_tgt >= 0;
    if (st0 || st1)
        for (int i = threadIdx.x; i < 2 * QB * LMAX; i += CTHREADS) Acc[i] = 0.f;

    float* wq = Wq + warp * 160;
    float* wp = Wp + warp * LMAX;

    for (int head = 0; head < a.heads; ++head) {
        for (int ri = 0; ri < nrows; ++ri) {
            const int row = ri == 0 ? g.base : g.tgt;
            __syncthreads();  // previous K/V fully consumed; Map/Acc init visible
            {
                const T* Kg = reinterpret_cast<const T*>(a.kv) + (long)row * L * a.ldkv + a.koff + head * d;
                const T* Vg = reinterpret_cast<const T*>(a.kv) + (long)row * L * a.ldkv + a.voff + head * d;
                int dv = d / 8;
                for (int i = threadIdx.x; i < L * dv; i += CTHREADS) {
                    int r = i / dv, c = (i % dv) * 8;
                    Pack<T, 8> kk = *reinterpret_cast<const Pack<T, 8>*>(Kg + (long)r * a.ldkv + c);
                    Pack<T, 8> vv = *reinterpret_cast<const Pack<T, 8>*>(Vg + (long)r * a.ldkv + c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { Ksm[r * kp + c + j] = kk.v[j]; Vsm[r * kp + c + j] = vv.v[j]; }
                }
            }
            __syncthreads();
            const T* Qg = reinterpret_cast<const T*>(a.q) + (long)row * a.N * a.ldq + head * d;
            T* Og = reinterpret_cast<T*>(a.out) + (long)row * a.N * a.ldo + head * d;
            const bool do_edit = edit && ri == 1;
            const bool do_store = ri == 0 ? st0 : st1;
            for (int qi = warp; qi < QB; qi += CWARPS) {
                const int q = q0 + qi;
                if (q >= a.N) break;
                for (int c = lane; c < d; c += 32) wq[c] = to_f<T>(Qg[(long)q * a.ldq + c]) * a.scale;
                __syncwarp();
                float p[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int key = lane + 32 * j;
                    float sacc = -INFINITY;
                    if (key < L) {
                        sacc = 0.f;
                        const T* kr = Ksm + key * kp;
                        for (int c = 0; c < d; ++c) sacc = fmaf(wq[c], to_f<T>(kr[c]), sacc);
                    }
                    p[j] = sacc;
                }
                float mx = warp_max(fmaxf(p[0], fmaxf(p[1], p[2])));
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    p[j] = (lane + 32 * j < L) ? expf(p[j] - mx) : 0.f;
                    sum += p[j];
                }
                float inv = 1.f / warp_sum(sum);
#pragma unroll
                for (int j = 0; j < 3; ++j) p[j] *= inv;

                if (edit && ri == 0) {  // remember base probabilities for the target pass
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (lane + 32 * j < L) Pbase[qi * LMAX + lane + 32 * j] = p[j];
                }
                if (do_edit) {
                    const float* pb = Pbase + qi * LMAX;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        int n = lane + 32 * j;
                        if (n < L && al[j] != 0.f) {
                            float m = 0.f;
                            for (int w = 0; w < L; ++w) m = fmaf(pb[w], Map[w * LMAX + n], m);
                            float f = eq[j] * (ba[j] * m + (1.f - ba[j]) * p[j]);
                            p[j] = al[j] * f + (1.f - al[j]) * p[j];
                        }
                    }
                }
                if (do_store) {
                    float* ac = Acc + (ri * QB + qi) * LMAX;
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (lane + 32 * j < L) ac[lane + 32 * j] += p[j];
                }
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (lane + 32 * j < L) wp[lane + 32 * j] = p[j];
                __syncwarp();
                for (int c = lane; c < d; c += 32) {
                    float o = 0.f;
                    for (int key = 0; key < L; ++key) o = fmaf(wp[key], to_f<T>(Vsm[key * kp + c]), o);
                    Og[(long)q * a.ldo + c] = from_f<T>(o);
                }
                __syncwarp();
            }
        }
    }
    if (st0 || st1) {
        __syncthreads();
        for (int ri = 0; ri < nrows; ++ri) {
            int slot = ri == 0 ? g.store_base : g.store_tgt;
            if (slot < 0) continue;
            for (int i = threadIdx.x; i < QB * L; i += CTHREADS) {
                int qi = i / L, w = i % L;
                if (q0 + qi < a.N) a.store[((long)slot * a.N + q0 + qi) * L + w] += Acc[(ri * QB + qi) * LMAX + w];
            }
        }
    }
}

template <typename T>
size_t cross_smem(int d) {
    int kp = d + (sizeof(T) == 4 ? 1 : 2);
    return (size_t)2 * LMAX * kp * sizeof(T) + sizeof(float) * ((size_t)QB * LMAX + 2 * QB * LMAX + LMAX * LMAX + CWARPS * 160 + CWARPS * LMAX);
}

}  // namespace

void cross_attention(const CrossAttnArgs& a, cudaStream_t s) {
    ETAI_CHECK(a.L <= LMAX && a.d % 8 == 0 && a.d <= 160, ETAI_ERR_ARG, "cross attention: L<=80, d%8==0, d<=160");
    ETAI_CHECK(a.n_groups > 0 && a.n_groups <= ETAI_MAX_ROWS, ETAI_ERR_ARG, "cross attention: bad group count");
    dim3 grid(cdiv(a.N, QB), a.n_groups);
    ETAI_DISPATCH_DTYPE(a.dtype, T, {
        size_t smem = cross_smem<T>(a.d);
        CUDA_CHECK(cudaFuncSetAttribute(cross_attn_k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cross_attn_k<T><<<grid, CTHREADS, smem, s>>>(a);
    });
    KERNEL_CHECK();
}

}  // namespace etai
