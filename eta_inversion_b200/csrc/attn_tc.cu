// tcgen05 flash self-attention for sm_100a (f16 / bf16 operands, fp32 softmax and accumulation).
//
//   out[r] = softmax(Q[qrow[r]] K[krow[r]]^T * scale) V[vrow[r]]          per (row r, head h)
//
// CTA = 128 query rows of one (row, head); K/V are streamed in tiles of 128 keys.
//   warp 0      : TMA producer.  Q, K, V tiles are fetched straight out of the fused [B,N,3C] qkv buffer through 4-D
//                 tensor maps {d, heads, N, B}; the box is 64 wide along d, so for d = 40 / 80 / 160 the tail of the last
//                 64-column atom is out of bounds and TMA zero-fills it (no padded copies of q/k/v exist).
//   warp 1      : single-thread MMA issuer.   S_j = Q K_j^T   (UMMA 128 x 128 x 16, both operands K-major)
//                                             O_j = P_j V_j   (UMMA 128 x d_pad x 16, A = P from smem (K-major),
//                                                              B = V tile as it landed: MN-major, 128B swizzle)
//   warps 2..5  : one thread per query row.  Reads its S row from TMEM (tcgen05.ld), online softmax in registers,
//                 writes P (16-bit) into shared memory in the swizzled K-major layout, reads O_j back from TMEM and
//                 folds it into the fp32 running output with the deferred rescale factor.
// S is double buffered in TMEM (2 x 128 columns) so QK^T of tile j+1 overlaps the softmax of tile j; O_j uses
// d_pad columns at column 256.  The (q,k,v) row remap carries PtP self-replacement / MasaCtrl / PnP (see etai.h).
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "ops.cuh"
#include "tc_common.cuh"

namespace etai {

namespace {

using namespace tc;

constexpr int BQ = 128, BKV = 128, AT_THREADS = 192, ATOM_BYTES = 128 * 128;  // one [128 rows x 128 B] swizzle-atom tile

struct AtParams {
    void* out;
    int Nq, Nk, heads, d;
    long ldo;
    float scale_log2e;
    RowMap map;
    int fmt;
};

// MN-major (N contiguous) B-operand tile, 128-byte swizzle: rows = K index at 128 B pitch, 8-row groups SBO = 1024 B,
// 64-element N atoms LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16_bmn(int fmt, int M, int N) {
    return make_idesc_f16(fmt, M, N) | (1u << 16);  // B operand MN-major
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T, int ATOMS, int STAGES>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_k(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
          const __grid_constant__ CUtensorMap tmV, AtParams p) {
    constexpr int OP_BYTES = ATOMS * ATOM_BYTES;          // Q, or one K tile, or one V tile
    constexpr int KV_BYTES = 2 * OP_BYTES;
    constexpr int P_OFF = OP_BYTES + STAGES * KV_BYTES;
    constexpr int BAR_OFF = P_OFF + 2 * ATOM_BYTES;
    constexpr int DPAD = ATOMS == 1 ? 48 : ATOMS == 2 ? 80 : 160;  // UMMA N of the PV product (d = 40 / 80 / 160)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = q_full + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;  // [2]
    uint64_t* p_full = s_full + 2;
    uint64_t* o_full = p_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, head = blockIdx.y, row = blockIdx.z;
    const int ntiles = (p.Nk + BKV - 1) / BKV;
    const int ksteps = (p.d + 15) / 16;  // 16-wide K steps of QK^T actually needed (the rest of the atom is zero)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 256;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            mbar_expect_tx(q_full, OP_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) tma_load_4d(smem + a * ATOM_BYTES, &tmQ, q_full, a * 64, head, q0, p.map.q[row]);
            for (int j = 0; j < ntiles; ++j) {
                int s = j % STAGES;
                uint32_t ph = (j / STAGES) & 1;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_expect_tx(&kv_full[s], KV_BYTES);
                unsigned char* kb = smem + OP_BYTES + s * KV_BYTES;
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) {
                    tma_load_4d(kb + a * ATOM_BYTES, &tmK, &kv_full[s], a * 64, head, j * BKV, p.map.k[row]);
                    tma_load_4d(kb + OP_BYTES + a * ATOM_BYTES, &tmV, &kv_full[s], a * 64, head, j * BKV, p.map.v[row]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc_s = make_idesc_f16(p.fmt, BQ, BKV);
            const uint32_t idesc_o = make_idesc_f16_bmn(p.fmt, BQ, DPAD);
            const uint32_t q_addr = smem_u32(smem);
            const uint32_t p_addr = smem_u32(smem + P_OFF);
            auto issue_s = [&](int j) {
                int s = j % STAGES;
                mbar_wait(&kv_full[s], (j / STAGES) & 1);
                tc_fence_after();
                uint32_t k_addr = smem_u32(smem + OP_BYTES + s * KV_BYTES);
                for (int k = 0; k < ksteps; ++k) {
                    uint32_t off = (uint32_t)(k >> 2) * ATOM_BYTES + (uint32_t)(k & 3) * 32;
                    umma_f16(tmem_base + (uint32_t)(j & 1) * 128, make_smem_desc_sw128(q_addr + off),
                             make_smem_desc_sw128(k_addr + off), idesc_s, k != 0);
                }
                umma_commit(&s_full[j & 1]);
            };
            auto issue_o = [&](int j) {
                int s = j % STAGES;
                mbar_wait(p_full, j & 1);
                tc_fence_after();
                uint32_t v_addr = smem_u32(smem + OP_BYTES + s * KV_BYTES + OP_BYTES);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {
                    uint32_t a_off = (uint32_t)(k >> 2) * ATOM_BYTES + (uint32_t)(k & 3) * 32;  // P: K-major, 2 atoms
                    uint32_t b_off = (uint32_t)k * 16 * 128;                                     // V: 16 key rows
                    umma_f16(tmem_o, make_smem_desc_sw128(p_addr + a_off),
                             make_smem_desc_mn_sw128(v_addr + b_off, ATOM_BYTES), idesc_o, k != 0);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(o_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < ntiles; ++j) {
                if (STAGES >= 2 && j + 1 < ntiles) issue_s(j + 1);  // overlaps the softmax of tile j
                issue_o(j);
                if (STAGES < 2 && j + 1 < ntiles) issue_s(j + 1);
            }
        }
    } else {
        // ===== softmax + output accumulation: thread = one query row =====
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        unsigned char* p_row = smem + P_OFF + r * 128;
        float o_acc[DPAD];
#pragma unroll
        for (int i = 0; i < DPAD; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, alpha_pending = 1.f;

        auto fold_o = [&](int j) {  // o_acc = o_acc * alpha + O_j
            mbar_wait(o_full, j & 1);
            tc_fence_after();
            constexpr int N32 = DPAD / 32;
#pragma unroll
            for (int b = 0; b < N32; ++b) {
                float v[32];
                tmem_ld32(tmem_o + lane_addr + b * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[b * 32 + i] = fmaf(o_acc[b * 32 + i], alpha_pending, v[i]);
            }
            if constexpr (DPAD % 32 != 0) {
                uint32_t rr[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]),
                      "=r"(rr[7]), "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]),
                      "=r"(rr[14]), "=r"(rr[15])
                    : "r"(tmem_o + lane_addr + N32 * 32)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    o_acc[N32 * 32 + i] = fmaf(o_acc[N32 * 32 + i], alpha_pending, __uint_as_float(rr[i]));
            }
        };

        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t s_addr = tmem_base + (uint32_t)(j & 1) * 128 + lane_addr;
            // pass 1: row max of this tile
            float mx = -INFINITY;
            const int kbase = j * BKV;
#pragma unroll 1
            for (int c0 = 0; c0 < BKV; c0 += 32) {
                float v[32];
                tmem_ld32(s_addr + c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (kbase + c0 + i < p.Nk) mx = fmaxf(mx, v[i]);
            }
            const float m_new = fmaxf(m_run, mx);
            const float alpha = exp2f((m_run - m_new) * p.scale_log2e);  // first tile: exp2(-inf) = 0
            const float mb = m_new * p.scale_log2e;
            if (j > 0) fold_o(j - 1);  // PV_{j-1} finished reading P and writing O: safe to overwrite P below
            alpha_pending = alpha;
            // pass 2: probabilities -> shared memory (K-major, 128B swizzle) + row sum
            float sum = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < BKV; c0 += 32) {
                float v[32];
                tmem_ld32(s_addr + c0, v);
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (kbase + c0 + i < p.Nk) ? exp2f(fmaf(v[i], p.scale_log2e, -mb)) : 0.f;
                    float p1 = (kbase + c0 + i + 1 < p.Nk) ? exp2f(fmaf(v[i + 1], p.scale_log2e, -mb)) : 0.f;
                    sum += p0 + p1;
                    if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
                        __half2 h = __floats2half2_rn(p0, p1);
                        packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
                    } else {
                        __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                        packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
                    }
                }
                // 32 keys = 64 B = 4 chunks of 16 B inside atom (c0 / 64), chunk index ((c0 % 64) / 8 + q) ^ (r & 7)
                unsigned char* atom = p_row + (c0 >> 6) * ATOM_BYTES;
                int cbase = (c0 & 63) >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int chunk = (cbase + q) ^ (r & 7);
                    *reinterpret_cast<uint4*>(atom + chunk * 16) =
                        make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                }
            }
            l_run = l_run * alpha + sum;
            m_run = m_new;
            tc_fence_before();   // order the TMEM reads of S_j / O_{j-1} before the arrive
            fence_async_smem();  // make the P stores visible to the tensor-core (async) proxy
            mbar_arrive(p_full);
        }
        fold_o(ntiles - 1);
        const int q = q0 + r;
        if (q < p.Nq) {
            const float inv = 1.f / l_run;
            T* dst = reinterpret_cast<T*>(p.out) + ((long)row * p.Nq + q) * p.ldo + head * p.d;
            constexpr int D = ATOMS == 1 ? 40 : ATOMS == 2 ? 80 : 160;
#pragma unroll
            for (int c = 0; c < D; c += 8) {
                float o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o8[i] = o_acc[c + i] * inv;
                store8<T>(dst + c, o8);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <typename T, int ATOMS, int STAGES>
void launch_attn(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AtParams& p, dim3 grid, cudaStream_t s) {
    constexpr int SMEM = ATOMS * ATOM_BYTES + STAGES * 2 * ATOMS * ATOM_BYTES + 2 * ATOM_BYTES + 256 + 1024;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(attn_tc_k<T, ATOMS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    attn_tc_k<T, ATOMS, STAGES><<<grid, AT_THREADS, SMEM, s>>>(q, k, v, p);
    KERNEL_CHECK();
}


// ---------------------------------------------------------------------------------------------------------------
// Two-query-tile ("ping-pong") variant for the large layers (d = 40 / 80, N >= 256): one CTA owns 256 query rows as
// two 128-row tiles, each with its own softmax warpgroup, S accumulator, P buffer and O accumulator.  The single MMA
// thread alternates between the tiles, so the tensor core computes S/PV of one tile while the other tile's warpgroup
// is in its softmax, and every K/V tile is fetched once for 256 queries.
// ---------------------------------------------------------------------------------------------------------------
constexpr int AT2_THREADS = 32 * 10;  // warp 0 TMA, warp 1 MMA, warps 2..5 softmax tile 0, warps 6..9 softmax tile 1

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <typename T, int ATOMS, int STAGES>
__global__ void __launch_bounds__(AT2_THREADS, 1)
attn_tc2_k(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AtParams p) {
    constexpr int OP_BYTES = ATOMS * ATOM_BYTES;
    constexpr int KV_BYTES = 2 * OP_BYTES;
    constexpr int KV_OFF = 2 * OP_BYTES;                      // after Q0, Q1
    constexpr int P_OFF = KV_OFF + STAGES * KV_BYTES;         // P0, P1: 2 atoms each
    constexpr int BAR_OFF = P_OFF + 4 * ATOM_BYTES;
    constexpr int D = ATOMS == 1 ? 40 : 80;
    // UMMA N of the PV product: d plus one extra column that the softmax threads set to 1.0 in the V tile, so the tensor
    // core also produces the softmax row sums (fp32, from exactly the 16-bit P it multiplies) in O[:, D]
    constexpr int DPAD = ATOMS == 1 ? 48 : 96;
    constexpr int O_STRIDE = ATOMS == 1 ? 64 : 96;            // TMEM columns between O0 and O1 (at column 256)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = q_full + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;  // [2] per query tile
    uint64_t* p_full = s_full + 2;         // [2]
    uint64_t* o_full = p_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256, head = blockIdx.y, row = blockIdx.z;
    const int ntiles = (p.Nk + BKV - 1) / BKV;
    constexpr int ksteps = (D + 15) / 16;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 128); mbar_init(&o_full[g], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * OP_BYTES);
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int a = 0; a < ATOMS; ++a)
                    tma_load_4d(smem + g * OP_BYTES + a * ATOM_BYTES, &tmQ, q_full, a * 64, head, q0 + g * 128, p.map.q[row]);
            for (int j = 0; j < ntiles; ++j) {
                int s = j % STAGES;
                mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], KV_BYTES);
                unsigned char* kb = smem + KV_OFF + s * KV_BYTES;
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) {
                    tma_load_4d(kb + a * ATOM_BYTES, &tmK, &kv_full[s], a * 64, head, j * BKV, p.map.k[row]);
                    tma_load_4d(kb + OP_BYTES + a * ATOM_BYTES, &tmV, &kv_full[s], a * 64, head, j * BKV, p.map.v[row]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_f16(p.fmt, BQ, BKV);
            const uint32_t idesc_o = make_idesc_f16_bmn(p.fmt, BQ, DPAD);
            auto issue_s = [&](int g, int j) {
                int s = j % STAGES;
                uint32_t q_addr = smem_u32(smem + g * OP_BYTES);
                uint32_t k_addr = smem_u32(smem + KV_OFF + s * KV_BYTES);
#pragma unroll
                for (int k = 0; k < ksteps; ++k) {
                    uint32_t off = (uint32_t)(k >> 2) * ATOM_BYTES + (uint32_t)(k & 3) * 32;
                    umma_f16(tmem_base + (uint32_t)g * 128, make_smem_desc_sw128(q_addr + off),
                             make_smem_desc_sw128(k_addr + off), idesc_s, k != 0);
                }
                umma_commit(&s_full[g]);
            };
            auto issue_o = [&](int g, int j) {
                int s = j % STAGES;
                mbar_wait(&p_full[g], j & 1);
                tc_fence_after();
                uint32_t p_addr = smem_u32(smem + P_OFF + g * 2 * ATOM_BYTES);
                uint32_t v_addr = smem_u32(smem + KV_OFF + s * KV_BYTES + OP_BYTES);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {
                    uint32_t a_off = (uint32_t)(k >> 2) * ATOM_BYTES + (uint32_t)(k & 3) * 32;
                    umma_f16(tmem_base + 256 + (uint32_t)g * O_STRIDE, make_smem_desc_sw128(p_addr + a_off),
                             make_smem_desc_mn_sw128(v_addr + (uint32_t)k * 16 * 128, ATOM_BYTES), idesc_o, k != 0);
                }
                if (g == 1) umma_commit(&kv_empty[s]);  // both query tiles are done with K_j / V_j
                umma_commit(&o_full[g]);
            };
            auto wait_kv = [&](int j) {
                mbar_wait(&kv_full[j % STAGES], (j / STAGES) & 1);
                tc_fence_after();
            };
            mbar_wait(q_full, 0);
            wait_kv(0);
            issue_s(0, 0);
            issue_s(1, 0);
            for (int j = 0; j < ntiles; ++j) {
                const bool more = j + 1 < ntiles;
                if (STAGES >= 2) {
                    issue_o(0, j);
                    if (more) { wait_kv(j + 1); issue_s(0, j + 1); }
                    issue_o(1, j);
                    if (more) issue_s(1, j + 1);
                } else {
                    issue_o(0, j);
                    issue_o(1, j);
                    if (more) { wait_kv(j + 1); issue_s(0, j + 1); issue_s(1, j + 1); }
                }
            }
        }
    } else {
        const int g = (warp - 2) >> 2;      // query tile / softmax warpgroup
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t s_addr = tmem_base + (uint32_t)g * 128 + lane_addr;
        const uint32_t o_addr = tmem_base + 256 + (uint32_t)g * O_STRIDE + lane_addr;
        unsigned char* p_row = smem + P_OFF + g * 2 * ATOM_BYTES + r * 128;
        float o_acc[DPAD];
#pragma unroll
        for (int i = 0; i < DPAD; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, alpha_pending = 1.f;

        auto fold_o = [&](int j) {
            mbar_wait(&o_full[g], j & 1);
            tc_fence_after();
            constexpr int N32 = DPAD / 32;
#pragma unroll
            for (int b = 0; b < N32; ++b) {
                float v[32];
                tmem_ld32(o_addr + b * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[b * 32 + i] = fmaf(o_acc[b * 32 + i], alpha_pending, v[i]);
            }
            if constexpr (DPAD % 32 != 0) {
                uint32_t rr[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]),
                      "=r"(rr[7]), "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]),
                      "=r"(rr[14]), "=r"(rr[15])
                    : "r"(o_addr + N32 * 32)
                    : "memory");
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    o_acc[N32 * 32 + i] = fmaf(o_acc[N32 * 32 + i], alpha_pending, __uint_as_float(rr[i]));
            }
        };

        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(&s_full[g], j & 1);
            tc_fence_after();
            const int kbase = j * BKV;
            const bool full_tile = kbase + BKV <= p.Nk;
            // pass 1: row max (two batches of 64 columns, one TMEM wait each)
            float mx = -INFINITY;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t a[32], b[32];
                tmem_ld32_nowait(s_addr + h * 64, a);
                tmem_ld32_nowait(s_addr + h * 64 + 32, b);
                tmem_wait_ld();
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(a[i]), __uint_as_float(b[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (kbase + h * 64 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(a[i]));
                        if (kbase + h * 64 + 32 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(b[i]));
                    }
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float alpha = ex2_approx((m_run - m_new) * p.scale_log2e);
            const float mb = m_new * p.scale_log2e;
            if (j > 0) fold_o(j - 1);
            alpha_pending = alpha;
            // pass 2: P = 2^(s*c - m*c) with the packed 16-bit MUFU (two exponentials per instruction), written straight to
            // shared memory in the K-major / 128B-swizzled A-operand layout
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t a[32], b[32];
                tmem_ld32_nowait(s_addr + h * 64, a);
                tmem_ld32_nowait(s_addr + h * 64 + 32, b);
                tmem_wait_ld();
                unsigned char* atom = p_row + h * ATOM_BYTES;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t packed[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float x0 = fmaf(__uint_as_float(half == 0 ? a[i] : b[i]), p.scale_log2e, -mb);
                        float x1 = fmaf(__uint_as_float(half == 0 ? a[i + 1] : b[i + 1]), p.scale_log2e, -mb);
                        if (!full_tile) {
                            if (kbase + h * 64 + half * 32 + i >= p.Nk) x0 = -INFINITY;
                            if (kbase + h * 64 + half * 32 + i + 1 >= p.Nk) x1 = -INFINITY;
                        }
                        uint32_t pk;  // (the packed ex2.approx.f16x2 form compiles to two MUFU ops plus a PRMT: no gain)
                        const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
                        if constexpr (std::is_same<T, __half>::value)
                            asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(pk) : "f"(e0), "f"(e1));
                        else
                            asm("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(pk) : "f"(e0), "f"(e1));
                        packed[i >> 1] = pk;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int chunk = (half * 4 + q) ^ (r & 7);
                        *reinterpret_cast<uint4*>(atom + chunk * 16) =
                            make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    }
                }
            }
            m_run = m_new;
            if (g == 0) {  // the "ones" column D of this stage's V tile (row = key r): O[:, D] = sum_j P[:, j]
                unsigned char* vt = smem + KV_OFF + (j % STAGES) * KV_BYTES + OP_BYTES + (D >> 6) * ATOM_BYTES;
                const int chunk = ((D & 63) >> 3) ^ (r & 7);
                *reinterpret_cast<uint16_t*>(vt + r * 128 + chunk * 16 + (D & 7) * 2) =
                    std::is_same<T, __half>::value ? (uint16_t)0x3C00 : (uint16_t)0x3F80;
            }
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(&p_full[g]);
        }
        fold_o(ntiles - 1);
        const int q = q0 + g * 128 + r;
        if (q < p.Nq) {
            const float inv = 1.f / o_acc[D];
            T* dst = reinterpret_cast<T*>(p.out) + ((long)row * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll
            for (int c = 0; c < D; c += 8) {
                float o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o8[i] = o_acc[c + i] * inv;
                store8<T>(dst + c, o8);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <typename T, int ATOMS, int STAGES>
void launch_attn2(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AtParams& p, dim3 grid, cudaStream_t s) {
    constexpr int SMEM = 2 * ATOMS * ATOM_BYTES + STAGES * 2 * ATOMS * ATOM_BYTES + 4 * ATOM_BYTES + 256 + 1024;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_k<T, ATOMS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    attn_tc2_k<T, ATOMS, STAGES><<<grid, AT2_THREADS, SMEM, s>>>(q, k, v, p);
    KERNEL_CHECK();
}


constexpr int AT3_THREADS = 384;  // warpgroup 0: warp 0 TMA, warps 1-2 UMMA issue, warp 3 idle; warpgroups 1, 2: softmax of query tiles 0, 1

// ---------------------------------------------------------------------------------------------------------------
// d = 40 flash self-attention (64x64 layers), v5.  Profile of the earlier variants (ncu source view, B200): the softmax
// warps spent 40-60 % of their time in mbarrier waits -- on S_j when S is single-buffered, on O_{j-1} when one thread
// issues the UMMAs of both query tiles (head-of-line blocking: tile 0's PV waits behind tile 1's P).  This version removes
// every wait from the steady state:
//   * 64-key tiles, S double-buffered per query tile (S_{j+2} is issued as soon as P_j is published);
//   * P goes straight from the softmax registers to TMEM (tcgen05.st) and feeds PV as the TMEM A operand of a TS-mode
//     UMMA: no shared-memory round trip, no 4 KB A-tile read per k-step; double-buffered (P_j may be written while
//     PV_{j-1} still reads P_{j-1}; PV_{j-2} is complete because S_j, issued after it by the same thread, is);
//   * O accumulates in TMEM across tiles and is rescaled lazily: the running maximum is only raised when it grew by
//     more than 2^8 (P stays <= 256, exact in fp16/bf16 relative precision; the row sum shares the stale maximum
//     through the "ones" column, so the result is the same softmax), so the O read-modify-write is off the common path;
//   * one UMMA-issuing thread per query tile with descriptors reduced to one add per k-step.
// TMEM: S[g][b] 64 columns at g*128 + b*64, O[g] at 256 + g*64, P[g][b] 32 columns at 384 + g*64 + b*32.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// UMMA with descriptors given as (low word, shared high word); ACC is compile-time so no predicate has to be computed
template <bool ACC>
__device__ __forceinline__ void umma_lohi(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
    if constexpr (ACC)
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.eq.u32 p, %4, %4;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc) : "memory");
    else
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.ne.u32 p, %4, %4;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc) : "memory");
}

// TS-mode UMMA: A operand in TMEM (lane = row, 32-bit column = two consecutive K elements), B from shared memory
template <bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t hi, uint32_t idesc) {
    if constexpr (ACC)
        asm volatile(
            "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.eq.u32 p, %4, %4;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
            ::"r"(d), "r"(a_tmem), "r"(blo), "r"(hi), "r"(idesc) : "memory");
    else
        asm volatile(
            "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.ne.u32 p, %4, %4;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
            ::"r"(d), "r"(a_tmem), "r"(blo), "r"(hi), "r"(idesc) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(AT3_THREADS, 1)
attn_d40_k(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AtParams p) {
    constexpr int STAGES = 4, D = 40, DPAD = 48, BK4 = 64;
    constexpr int Q_BYTES = ATOM_BYTES;               // [128 q][64]
    constexpr int KT_BYTES = BK4 * 128;                // one K or V tile: [64 keys][128 B]
    constexpr int KV_BYTES = 2 * KT_BYTES;
    constexpr int KV_OFF = 2 * Q_BYTES, BAR_OFF = KV_OFF + STAGES * KV_BYTES;
    constexpr float RESCALE_LOG2 = 8.f;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* kv_full = q_full + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;  // [g][b]
    uint64_t* p_full = s_full + 4;         // [g][b]
    uint64_t* o_full = p_full + 4;         // [g]: one phase per PV_j (only the lazy rescale waits on it)
    uint64_t* o_done = o_full + 2;         // [g]: last PV complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256, head = blockIdx.y, row = blockIdx.z;
    const int ntiles = (p.Nk + BK4 - 1) / BK4;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }
        for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 128);
        for (int g = 0; g < 2; ++g) { mbar_init(&o_full[g], 1); mbar_init(&o_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            mbar_expect_tx(q_full, 2 * Q_BYTES);
            tma_load_4d(smem, &tmQ, q_full, 0, head, q0, p.map.q[row]);
            tma_load_4d(smem + Q_BYTES, &tmQ, q_full, 0, head, q0 + 128, p.map.q[row]);
            for (int j = 0; j < ntiles; ++j) {
                int s = j % STAGES;
                mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], KV_BYTES);
                unsigned char* kb = smem + KV_OFF + s * KV_BYTES;
                tma_load_4d(kb, &tmK, &kv_full[s], 0, head, j * BK4, p.map.k[row]);
                tma_load_4d(kb + KT_BYTES, &tmV, &kv_full[s], 0, head, j * BK4, p.map.v[row]);
            }
        } else if ((warp == 1 || warp == 2) && lane == 0) {
            // ===== UMMA issuer of query tile g: S_j = Q_g K_j^T (3 k-steps of 16), O_g += P_j V_j (4 k-steps) =====
            const int g = warp - 1;
            const uint32_t idesc_s = make_idesc_f16(p.fmt, BQ, BK4);
            const uint32_t idesc_o = make_idesc_f16_bmn(p.fmt, BQ, DPAD);
            const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B | version | SWIZZLE_128B
            const uint32_t lo_k = 1u << 16;                                        // K-major: LBO unused (1)
            const uint32_t lo_mn = (uint32_t)(KT_BYTES >> 4) << 16;                // MN-major V: single 64-wide N atom
            const uint32_t q_lo = lo_k | ((smem_u32(smem + g * Q_BYTES) & 0x3FFFF) >> 4);
            const uint32_t kv_lo = (smem_u32(smem + KV_OFF) & 0x3FFFF) >> 4;
            const uint32_t tS = tmem_base + (uint32_t)g * 128, tO = tmem_base + 256 + (uint32_t)g * 64;
            auto issue_s = [&](int j) {
                mbar_wait(&kv_full[j % STAGES], (j / STAGES) & 1);
                tc_fence_after();
                const uint32_t k_lo = lo_k | (kv_lo + (uint32_t)(j % STAGES) * (KV_BYTES >> 4));
                const uint32_t d = tS + (uint32_t)(j & 1) * 64;
                umma_lohi<false>(d, q_lo, k_lo, hi, idesc_s);
                umma_lohi<true>(d, q_lo + 2, k_lo + 2, hi, idesc_s);
                umma_lohi<true>(d, q_lo + 4, k_lo + 4, hi, idesc_s);
                umma_commit(&s_full[g * 2 + (j & 1)]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (ntiles > 1) issue_s(1);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % STAGES;
                mbar_wait(&p_full[g * 2 + (j & 1)], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t v_lo = lo_mn | (kv_lo + (uint32_t)s * (KV_BYTES >> 4) + (KT_BYTES >> 4));
                // P_j lives in TMEM (A operand of a TS-mode UMMA: no shared-memory read of the 128 x 16 A tile per k-step)
                const uint32_t tP = tmem_base + 384 + (uint32_t)g * 64 + (uint32_t)(j & 1) * 32;  // 16 keys = 8 columns
                if (j == 0) umma_ts<false>(tO, tP, v_lo, hi, idesc_o);
                else umma_ts<true>(tO, tP, v_lo, hi, idesc_o);
                umma_ts<true>(tO, tP + 8, v_lo + 128, hi, idesc_o);   // +16 key rows of V
                umma_ts<true>(tO, tP + 16, v_lo + 256, hi, idesc_o);
                umma_ts<true>(tO, tP + 24, v_lo + 384, hi, idesc_o);
                umma_commit(&kv_empty[s]);  // count 2: both issuers are done with K_j / V_j
                umma_commit(&o_full[g]);
                if (j == ntiles - 1) umma_commit(&o_done[g]);
                if (j + 2 < ntiles) issue_s(j + 2);  // S buffer (j & 1) was consumed before P_j was published
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int g = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t o_addr = tmem_base + 256 + (uint32_t)g * 64 + lane_addr;
        const uint32_t ones_addr = smem_u32(smem + KV_OFF + KT_BYTES + r * 128 + (((D >> 3) ^ (r & 7)) * 16) + (D & 7) * 2);
        const uint16_t one = std::is_same<T, __half>::value ? (uint16_t)0x3C00 : (uint16_t)0x3F80;
        float m_run = -INFINITY;
        const float c = p.scale_log2e;

        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(&s_full[g * 2 + (j & 1)], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t s_addr = tmem_base + (uint32_t)g * 128 + (uint32_t)(j & 1) * 64 + lane_addr;
            const int kbase = j * BK4;
            uint32_t sv[64];
            {
                uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
                uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
                tmem_ld32_nowait(s_addr, s0);
                tmem_ld32_nowait(s_addr + 32, s1);
                tmem_wait_ld();
            }
            if (kbase + BK4 > p.Nk) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (kbase + i >= p.Nk) sv[i] = 0xff800000u;  // -inf
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mx[k] = fmaxf(mx[k], fmaxf(__uint_as_float(sv[i + 2 * k]), __uint_as_float(sv[i + 2 * k + 1])));
            }
            const float m_new = fmaxf(fmaxf(m_run, fmaxf(mx[0], mx[1])), fmaxf(mx[2], mx[3]));
            if (j == 0) {
                m_run = m_new;  // O is written (not accumulated) by PV_0
            } else {
                const bool grow = (m_new - m_run) * c > RESCALE_LOG2;
                if (__any_sync(0xffffffffu, grow)) {  // rare after the first tiles: rescale this warp's 32 O rows in TMEM
                    const float alpha = grow ? ex2_approx((m_run - m_new) * c) : 1.f;
                    if (grow) m_run = m_new;
                    mbar_wait(&o_full[g], (j - 1) & 1);  // PV_{j-1} complete; PV_j is not issued before our p_full arrive
                    tc_fence_after();
                    uint32_t o0[32], o1[16];
                    tmem_ld32_nowait(o_addr, o0);
                    tmem_ld16_nowait(o_addr + 32, o1);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
#pragma unroll
                    for (int i = 0; i < 16; ++i) o1[i] = __float_as_uint(__uint_as_float(o1[i]) * alpha);
                    tmem_st32(o_addr, o0);
                    tmem_st16(o_addr + 32, o1);
                    tmem_wait_st();
                }
            }
            const float mb = m_run * c;
            // P_j = 2^(s*c - m*c) -> 16-bit, swizzled A-operand tile, buffer j & 1
            uint32_t pr[32];  // the row's 64 probabilities, packed pairs in key order
#pragma unroll
            for (int blk = 0; blk < 8; ++blk) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float e0 = ex2_approx(fmaf(__uint_as_float(sv[blk * 8 + 2 * e]), c, -mb));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(sv[blk * 8 + 2 * e + 1]), c, -mb));
                    if constexpr (std::is_same<T, __half>::value)
                        asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(pr[blk * 4 + e]) : "f"(e0), "f"(e1));
                    else
                        asm("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(pr[blk * 4 + e]) : "f"(e0), "f"(e1));
                }
            }
            // lane = query row, 32-bit column = key pair: exactly the A-operand layout of a TS-mode UMMA
            tmem_st32(tmem_base + 384 + (uint32_t)g * 64 + (uint32_t)(j & 1) * 32 + lane_addr, pr);
            tmem_wait_st();
            // "ones" column D of this stage's V tile (row = key r): O[:, D] accumulates sum_j P[:, j].  Both query tiles
            // write the same value, so neither issuer depends on the other tile's warpgroup.
            if (r < BK4)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(ones_addr + (uint32_t)(j % STAGES) * KV_BYTES), "h"(one) : "memory");
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(&p_full[g * 2 + (j & 1)]);
        }
        // (o_full cannot be used here: PV of the last TWO tiles may be pending, which a parity wait cannot tell apart)
        mbar_wait(&o_done[g], 0);
        tc_fence_after();
        uint32_t o0[32], o1[16];
        tmem_ld32_nowait(o_addr, o0);
        tmem_ld16_nowait(o_addr + 32, o1);
        tmem_wait_ld();
        const int q = q0 + g * 128 + r;
        if (q < p.Nq) {
            const float inv = 1.f / __uint_as_float(o1[D - 32]);
            T* dst = reinterpret_cast<T*>(p.out) + ((long)row * p.Nq + q) * p.ldo + head * D;
#pragma unroll
            for (int cc = 0; cc < D; cc += 8) {
                float o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o8[i] = __uint_as_float(cc + i < 32 ? o0[(cc + i) & 31] : o1[(cc + i) & 15]) * inv;
                store8<T>(dst + cc, o8);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <typename T>
void launch_attn_d40(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AtParams& p, dim3 grid, cudaStream_t s) {
    constexpr int SMEM = 2 * ATOM_BYTES + 4 * 2 * 64 * 128 + 256 + 1024;  // Q tiles + 4 K/V stages + barriers
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(attn_d40_k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    attn_d40_k<T><<<grid, AT3_THREADS, SMEM, s>>>(q, k, v, p);
    KERNEL_CHECK();
}

CUtensorMap head_tmap(const void* base, int dtype, int d, int heads, int N, int B, long ld, int box_rows = 128) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * N};
    uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    return make_tmap_16bit(base, dtype, 4, dims, str, box);
}

}  // namespace

bool attention_tc_supported(const SelfAttnArgs& a) {
    if (a.dtype != ETAI_F16 && a.dtype != ETAI_BF16) return false;
    if (a.d != 40 && a.d != 80 && a.d != 160) return false;  // UMMA N of the PV product is compiled per head dim
    if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.ldo % 8) return false;
    if (a.Nq < 1 || a.Nk < 1 || a.B > ETAI_MAX_ROWS) return false;
    auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return aligned(a.q) && aligned(a.k) && aligned(a.v) && aligned(a.out);
}

void attention_tc(const SelfAttnArgs& a, cudaStream_t s) {
    ETAI_CHECK(attention_tc_supported(a), ETAI_ERR_UNSUPPORTED, "attention_tc: unsupported problem");
    AtParams p;
    memset(&p, 0, sizeof(p));
    p.out = a.out; p.Nq = a.Nq; p.Nk = a.Nk; p.heads = a.heads; p.d = a.d; p.ldo = a.ldo;
    p.scale_log2e = a.scale * 1.4426950408889634f;
    p.map = a.map;
    p.fmt = a.dtype == ETAI_BF16 ? 1 : 0;
    CUtensorMap tq = head_tmap(a.q, a.dtype, a.d, a.heads, a.Nq, a.B, a.ldq);
    CUtensorMap tk = head_tmap(a.k, a.dtype, a.d, a.heads, a.Nk, a.B, a.ldk);
    CUtensorMap tv = head_tmap(a.v, a.dtype, a.d, a.heads, a.Nk, a.B, a.ldv);
    dim3 grid(cdiv(a.Nq, BQ), a.heads, a.B);
    const bool two = a.d != 160 && a.Nq >= 256;  // ping-pong kernel: 256 query rows per CTA
    if (two) grid.x = cdiv(a.Nq, 256);
    const bool d40 = two && a.d == 40;
    if (d40) {  // 64-key tiles
        tk = head_tmap(a.k, a.dtype, a.d, a.heads, a.Nk, a.B, a.ldk, 64);
        tv = head_tmap(a.v, a.dtype, a.d, a.heads, a.Nk, a.B, a.ldv, 64);
    }
#define LAUNCH(T)                                                              \
    do {                                                                       \
        if (d40) launch_attn_d40<T>(tq, tk, tv, p, grid, s);                   \
        else if (two) launch_attn2<T, 2, 1>(tq, tk, tv, p, grid, s);           \
        else if (a.d == 40) launch_attn<T, 1, 2>(tq, tk, tv, p, grid, s);      \
        else if (a.d == 80) launch_attn<T, 2, 2>(tq, tk, tv, p, grid, s);      \
        else launch_attn<T, 3, 1>(tq, tk, tv, p, grid, s);                     \
    } while (0)
    if (a.dtype == ETAI_F16) LAUNCH(__half);
    else LAUNCH(__nv_bfloat16);
#undef LAUNCH
}

}  // namespace etai
