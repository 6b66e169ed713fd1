// tcgen05 / TMEM / TMA GEMM and implicit-GEMM conv3x3 for sm_100a (f16 / bf16 in, fp32 accumulate in TMEM).
//
//   C[M,N] = A[M,K] * W[N,K]^T  + bias[N] + bias2[N](fp32) + residual[M,N]      (or GEGLU epilogue)
//
// A is either a dense K-major matrix (2-D tensor map) or, for the 3x3 / pad 1 / stride 1 convolution, the NHWC image
// itself read through a 4-D tensor map {C, W, H, B}: for every filter tap the producer shifts the box origin by
// (kx-1, ky-1) and TMA zero-fills the halo, so im2col is never materialised.  One 128-row M tile is a box of
// bw x bh x bb = 128 output pixels (64x2, 32x4, 16x8 or 8x8x2 images), which lands in shared memory in exactly the
// K-major / 128B-swizzled layout the UMMA descriptor expects.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> fused epilogue -> shared-memory panels -> TMA store).  smem ring of
// STAGES x (A 128x64 + B BNx64).  UMMA shape 128 x UN x 16 (cta_group::1), UN in {128, 160}; a 320-wide tile is two
// 160-wide accumulators.  Roofline: tensor pipe (see DESIGN.md section 5 for what the per-role clock trace shows).
#include <algorithm>
#include <mutex>
#include <vector>
#include <cstdlib>
#include <cstring>
#include "ops.cuh"
#include "tc_common.cuh"

namespace etai {

namespace tc {

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (!fn) throw Error(ETAI_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    return fn;
}

CUtensorMap make_tmap_16bit(const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, bool swizzle128) {
    CUtensorMap m;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t b[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; b[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = get_encode_tiled()(&m, dtype == ETAI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                    (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, b, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(ETAI_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return m;
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int EPI_WARPS = 8, GT_THREADS = 32 * (2 + EPI_WARPS);  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct TcParams {
    void* C;
    const void* bias;      // [N] storage dtype or null
    const float* bias2;    // [N] fp32 or null (time-embedding projection)
    const void* residual;  // [M,ldr] or null
    float* partial;        // split-K: fp32 [splits][M][N] workspace (epilogue deferred to splitk_reduce_k)
    long M;
    int N;
    long ldc, ldr;
    int geglu;
    int num_kb;            // K blocks of 64 (whole problem)
    int splits, kb_per_split;
    int tiles_m, tiles_n;
    // conv geometry
    int cin_blocks;        // Cin / 64
    int Himg, Wimg;        // output == input spatial size (stride 1)
    int fmt;               // 0 f16, 1 bf16
    long long* trace;      // debugging aid (ETAI_GEMM_TRACE=1): per-CTA, per-tile clock64() stamps of the three roles
};
constexpr int TRACE_TILES = 64, TRACE_SLOTS = 8;

// ---- epilogue helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t (&r)[32]) {  // no wait: pair with tmem_wait_ld()
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
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {  // two fp32 -> packed 16-bit pair (lo in bits 0..15)
    uint32_t d;
    if constexpr (std::is_same<T, __half>::value) asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(d) : "f"(lo), "f"(hi));
    else asm("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(d) : "f"(lo), "f"(hi));
    return d;
}
template <typename T>
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {  // packed 16-bit add
    uint32_t d;
    if constexpr (std::is_same<T, __half>::value) asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// erf-form GELU for 16-bit outputs: Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the output rounding) with
// one rcp and one ex2 instead of erff's two-range polynomial (the GEGLU epilogue is issue-bound on the K = 320 layers)
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
    const float erf_abs = fmaf(-poly, e, 1.0f);
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// SCHED 0: interleaved schedule, every k-block stage holds an A and a B box.
// SCHED 1 ("B-stationary", K <= 320 projections): the CTA keeps its 160 x K slice of the weights resident in shared memory
//          (BRES_KB k-blocks, loaded once) and walks a run of M tiles; the ring carries A boxes only, so the SM's TMA unit
//          moves 128 instead of 288 operand rows per k-block.
// EB = number of epilogue staging buffers (2: the conversion of tile i+1 does not wait for the TMA store of tile i to have
//          read the panels; costs one ring stage).
template <int BN, int SCHED = 0, int EB = 1>
struct Smem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = SCHED == 1 ? A_BYTES : A_BYTES + B_BYTES;
    static constexpr int BRES_KB = 5;                                      // K <= 320
    static constexpr int BRES_BYTES = SCHED == 1 ? BRES_KB * B_BYTES : 0;
    // Epilogue staging: one pass of PW accumulator columns of all 128 rows, as column panels of 64 16-bit columns
    // ([128 rows][128 B], 128B-swizzled, so the row-per-lane writes are conflict free) plus a narrow tail panel; it leaves
    // through ONE TMA store per panel.  (Round 1 used a [32 x 32] box per warp and chunk: 20 store instructions per
    // 128 x 160 tile cost the SM's TMA unit as much time as the tile's operand loads -- measured with ETAI_GEMM_TRACE.)
    static constexpr int PW = BN == 320 ? 160 : BN;            // accumulator columns per pass (a 320-wide tile takes two)
    static constexpr int PASSES = BN / PW;
    static constexpr int EPI_BYTES = (PW / 64) * 16384 + (PW % 64 ? 8192 : 0);
    static constexpr int EPI_BUFS = EB;
    static constexpr int FREE = 227 * 1024 - EB * EPI_BYTES - BRES_BYTES - 2048;
    static constexpr int STAGES = FREE / STAGE_BYTES > 6 ? 6 : FREE / STAGE_BYTES;
    static constexpr int BRES_OFF = STAGES * STAGE_BYTES;
    static constexpr int EPI_OFF = BRES_OFF + BRES_BYTES;
    static constexpr int BAR_OFF = EPI_OFF + EB * EPI_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // barriers + slack for manual 1024-B alignment
    // BN = 320 is issued as two UMMAs of N = 160 per k-step into two accumulators that sit side by side in TMEM (columns
    // [0,160) and [160,320)), so the epilogue sees one 320-column tile.  It is single-buffered (2 x 320 > 512 columns).
    static constexpr int UN = BN == 320 ? 160 : BN;            // UMMA N
    static constexpr int NACC = BN / UN;                       // UMMAs per k-step
    static constexpr int NBUF = BN == 320 ? 1 : 2;             // accumulator buffers in TMEM
    static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;   // TMEM columns between the accumulator buffers (NBUF = 2)
    static constexpr int TMEM_COLS = BN == 320 ? 512 : 2 * ACC_STRIDE;
    static_assert(STAGES >= 2, "smem ring too short");
    static_assert(EB == 1 || PASSES == 1, "two staging buffers alternate per tile");
    static_assert(SCHED == 0 || NACC == 1, "B-stationary schedule: one UMMA per k-step");
    static_assert((2 * STAGES + 4 + BRES_KB) * 8 + 4 <= 256, "barrier area");
};

// Static schedule of one CTA.  SCHED 0: item = blockIdx.x + li * gridDim.x over (split, m_tile, n_tile) with n fastest, so
// the CTAs running at the same time share A panels in L2.  SCHED 1: gridDim.x = groups * tiles_n; CTA c owns N slice
// c % tiles_n for its whole life and walks the M tiles of group c / tiles_n (the tiles_n CTAs of a group walk the same
// A panels at the same time).
template <int SCHED>
struct Walk {
    int count, n_fixed, m_begin;
    __device__ __forceinline__ explicit Walk(const TcParams& p) {
        if (SCHED == 1) {
            const int ngroups = gridDim.x / p.tiles_n, g = blockIdx.x / p.tiles_n;
            n_fixed = blockIdx.x % p.tiles_n;
            m_begin = (int)((long)g * p.tiles_m / ngroups);
            count = (int)((long)(g + 1) * p.tiles_m / ngroups) - m_begin;
        } else {
            const int items = p.tiles_m * p.tiles_n * p.splits;
            n_fixed = m_begin = 0;
            count = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        }
    }
    __device__ __forceinline__ void at(const TcParams& p, int li, int& split, int& m_tile, int& n_tile) const {
        if (SCHED == 1) {
            split = 0; m_tile = m_begin + li; n_tile = n_fixed;
        } else {
            const int item = blockIdx.x + li * gridDim.x, tiles = p.tiles_m * p.tiles_n, tile = item % tiles;
            split = item / tiles; n_tile = tile % p.tiles_n; m_tile = tile / p.tiles_n;
        }
    }
};

// Persistent kernel: grid = min(#work items, #SMs); every role walks the same static schedule (Walk<SCHED>).
// The smem ring runs across items, the accumulator is double buffered in TMEM so the epilogue of item i overlaps
// the main loop of item i+1.
template <typename T, int BN, bool CONV, int SCHED = 0, int EB = 1>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
          const __grid_constant__ TcParams p) {
    using S = Smem<BN, SCHED, EB>;
    static_assert(!(CONV && SCHED == 1), "the B-stationary schedule serves dense K <= 320 problems");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty = full + S::STAGES;
    uint64_t* acc_full = empty + S::STAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint64_t* b_full = acc_empty + 2;         // [BRES_KB] (SCHED 1: resident weight slice, filled once)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + S::BRES_KB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Walk<SCHED> walk(p);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmC);
        prefetch_tmap(&tmC2);
        for (int s = 0; s < S::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < S::NBUF; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS); }
        if (SCHED == 1)
            for (int k = 0; k < S::BRES_KB; ++k) mbar_init(&b_full[k], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int kc = 0;  // k blocks issued so far by this CTA (ring position)
            for (int tli = 0; tli < walk.count; ++tli) {
                int split, m_tile, n_tile;
                walk.at(p, tli, split, m_tile, n_tile);
                const int n0 = n_tile * BN;
                const long m0 = (long)m_tile * BM;
                int b0 = 0, oy0 = 0, ox0 = 0;
                if (CONV) {
                    long hw = (long)p.Himg * p.Wimg;
                    b0 = (int)(m0 / hw);
                    oy0 = (int)((m0 % hw) / p.Wimg);
                    ox0 = (int)((m0 % hw) % p.Wimg);  // non-zero only for images wider than one tile (VAE: 256, 512)
                }
                const int kb0 = split * p.kb_per_split;
                const int kb1 = kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb;
                if (p.trace && tli < TRACE_TILES) p.trace[((long)blockIdx.x * TRACE_TILES + tli) * TRACE_SLOTS + 0] = clock64();
                for (int kb = kb0; kb < kb1; ++kb, ++kc) {
                    if (SCHED == 1 && tli == 0) {  // the weight slice of this CTA, k-block by k-block ahead of the first tile's A boxes
                        mbar_expect_tx(&b_full[kb], S::B_BYTES);
                        tma_load_2d(smem + S::BRES_OFF + kb * S::B_BYTES, &tmB, &b_full[kb], kb * BK, n0);
                    }
                    int s = kc % S::STAGES;
                    uint32_t ph = (kc / S::STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], S::STAGE_BYTES);
                    unsigned char* sa = smem + s * S::STAGE_BYTES;
                    unsigned char* sb = sa + S::A_BYTES;
                    if (CONV) {
                        int tap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
                        tma_load_4d(sa, &tmA, &full[s], cb * BK, ox0 + tap % 3 - 1, oy0 + tap / 3 - 1, b0);
                    } else {
                        tma_load_2d(sa, &tmA, &full[s], kb * BK, (int)m0);
                    }
                    if (SCHED == 0) {
#pragma unroll
                        for (int j = 0; j < S::NACC; ++j)  // a TMA box has at most 256 rows: one box per UMMA-N slice of the tile
                            tma_load_2d(sb + j * S::UN * 128, &tmB, &full[s], kb * BK, n0 + j * S::UN);
                    }
                }
                if (p.trace && tli < TRACE_TILES) p.trace[((long)blockIdx.x * TRACE_TILES + tli) * TRACE_SLOTS + 1] = clock64();
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_idesc_f16(p.fmt, BM, S::UN);
            int kc = 0;
            for (int li = 0; li < walk.count; ++li) {
                int split, m_tile, n_tile;
                walk.at(p, li, split, m_tile, n_tile);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb;
                const int buf = li % S::NBUF;
                long long* tr = (p.trace && li < TRACE_TILES) ? p.trace + ((long)blockIdx.x * TRACE_TILES + li) * TRACE_SLOTS : nullptr;
                if (tr) tr[2] = clock64();
                mbar_wait(&acc_empty[buf], ((li / S::NBUF) & 1) ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                if (tr) tr[3] = clock64();
                const uint32_t tacc = tmem_base + (uint32_t)buf * S::ACC_STRIDE;
                for (int kb = kb0; kb < kb1; ++kb, ++kc) {
                    int s = kc % S::STAGES;
                    uint32_t ph = (kc / S::STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    if (SCHED == 1 && li == 0) mbar_wait(&b_full[kb], 0);
                    tc_fence_after();
                    if (tr && kb == kb0) tr[4] = clock64();
                    uint32_t sa = smem_u32(smem + s * S::STAGE_BYTES);
                    uint64_t da = make_smem_desc_sw128(sa);
                    uint64_t db = make_smem_desc_sw128(SCHED == 1 ? smem_u32(smem + S::BRES_OFF + kb * S::B_BYTES) : sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 elements (32 B) along K inside the 128-B swizzle atom: +2 in the (addr>>4) field;
                        // accumulator j takes rows [j*UN, (j+1)*UN) of the B stage (UN * 128 B further on)
#pragma unroll
                        for (int j = 0; j < S::NACC; ++j)
                            umma_f16(tacc + (uint32_t)(j * S::UN), da + (uint64_t)(2 * k),
                                     db + (uint64_t)(j * (S::UN * 128 >> 4) + 2 * k), idesc, (kb > kb0) || (k > 0));
                    }
                    umma_commit(&empty[s]);  // frees the smem stage once these MMAs retire
                }
                umma_commit(&acc_full[buf]);
                if (tr) tr[5] = clock64();
            }
        }
    } else {
        // ===== epilogue: warps 2..9 (256 threads).  TMEM lane quarter = warp % 4 = 32 rows of the tile, one row per lane;
        // the two warps of a quarter alternate 32-column chunks.  Per pass: wait until the staging panels are free, every
        // warp converts its chunks (TMEM -> registers -> bias / residual / GEGLU -> 16 bit) into the panels, one named
        // barrier, then one thread issues one TMA store per panel (rows / columns beyond M / N are clipped by the map).
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int trow = quarter * 32 + lane;  // row inside the tile == TMEM lane
        const T* bias = reinterpret_cast<const T*>(p.bias);
        const T* res = reinterpret_cast<const T*>(p.residual);
        const bool leader = warp == 2 && lane == 0;
        constexpr int PW = S::PW;
        for (int li = 0; li < walk.count; ++li) {
            int split, m_tile, n_tile;
            walk.at(p, li, split, m_tile, n_tile);
            const int n0 = n_tile * BN;
            const long m0 = (long)m_tile * BM;
            // staging panels of this tile (EB = 2: tiles alternate between two sets)
            const uint32_t epi = smem_u32(smem + S::EPI_OFF) + (uint32_t)(EB == 2 ? (li & 1) * S::EPI_BYTES : 0);
            const int buf = li % S::NBUF;
            const long m = m0 + trow;
            const bool row_ok = m < p.M;
            const bool lean = !p.partial && !p.bias2;
            const T* brow = bias ? bias + n0 : nullptr;
            const T* rrow = (res && row_ok && !p.geglu) ? res + m * p.ldr + n0 : nullptr;
            // bias / residual of a chunk are requested one chunk ahead -- the first chunk's before the accumulator is even
            // complete -- so their L2 / HBM latency is off the epilogue's critical path
            uint4 b4n[4], r4n[4];
            auto prefetch = [&](int c0) {
                if (n0 + c0 < p.N) {
                    if (brow) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) b4n[j] = ldg128(brow + c0 + 8 * j);
                    }
                    if (rrow) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) r4n[j] = ldg128(rrow + c0 + 8 * j);
                    }
                }
            };
            // stage 16-byte piece `piece` (of the pass's output row) of this lane's row
            auto stage16 = [&](int piece, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
                const int panel = piece >> 3, pp = piece & 7;           // 8 pieces = 64 columns = one 128-byte panel row
                constexpr int FULL = (PW / 64);                         // number of full (swizzled) panels when not GEGLU
                const int full = p.geglu ? (PW / 2) / 64 : FULL;
                uint32_t addr;
                if (panel < full) addr = epi + (uint32_t)panel * 16384u + (uint32_t)trow * 128u + (uint32_t)((pp ^ (trow & 7)) << 4);
                else {  // tail panel: dense rows of 64 B (32 columns) or, for GEGLU at PW = 160, 32 B (16 columns)
                    const uint32_t rowb = p.geglu ? 32u : 64u;
                    addr = epi + (uint32_t)full * 16384u + (uint32_t)trow * rowb + (uint32_t)(pp << 4);
                }
                sts128(addr, a, b, c, d);
            };
            if (lean) prefetch(half * 32);
            mbar_wait(&acc_full[buf], (li / S::NBUF) & 1);
            tc_fence_after();
            long long* tr = (p.trace && li < TRACE_TILES && leader)
                                ? p.trace + ((long)blockIdx.x * TRACE_TILES + li) * TRACE_SLOTS : nullptr;
            if (tr) tr[6] = clock64();
            const uint32_t tacc = tmem_base + (uint32_t)buf * S::ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int ps = 0; ps < S::PASSES; ++ps) {
                const int hh = (ps & 1) ? half ^ 1 : half;  // the second pass swaps roles: 3 + 2 and 2 + 3 chunks per warp
                if (!p.partial) {
                    if (ps > 0 && lean) prefetch(ps * PW + hh * 32);
                    if (leader) {  // the stores that last used these panels have read them
                        if (EB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
#pragma unroll 1
                for (int cc = hh * 32; cc < PW; cc += 64) {
                    const int c0 = ps * PW + cc;            // accumulator column inside the tile
                    const int n = n0 + c0;
                    const bool col_ok = n < p.N;            // warp-uniform
                    if (lean) {
                        // ---- lean path (the bulk of the UNet's GEMMs): operands stay packed (one F2FP per pair, packed adds)
                        uint32_t r[32];
                        tmem_ld32_raw(tacc + (uint32_t)c0, r);  // warp-collective
                        uint4 b4[4], r4[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { b4[j] = b4n[j]; r4[j] = r4n[j]; }
                        if (cc + 64 < PW) prefetch(c0 + 64);
                        tmem_wait_ld();
                        if (!col_ok) continue;
                        if (p.geglu) {
                            // (value, gate) column pairs; bias in fp32 (it feeds the nonlinearity)
                            uint32_t h[8];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float bf[8];
                                if (brow) {
                                    const Pack<T, 8>& pk = *reinterpret_cast<const Pack<T, 8>*>(&b4[j]);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) bf[i] = to_f<T>(pk.v[i]);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) bf[i] = 0.f;
                                }
                                float o[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    o[i] = (__uint_as_float(r[8 * j + 2 * i]) + bf[2 * i]) *
                                           gelu_fast(__uint_as_float(r[8 * j + 2 * i + 1]) + bf[2 * i + 1]);
                                h[2 * j] = pack2<T>(o[0], o[1]);
                                h[2 * j + 1] = pack2<T>(o[2], o[3]);
                            }
                            const int piece = cc >> 4;  // 16 output columns = 2 pieces
                            stage16(piece, h[0], h[1], h[2], h[3]);
                            stage16(piece + 1, h[4], h[5], h[6], h[7]);
                        } else {
                            uint32_t h[16];
                            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                                // bf16 keeps 8 mantissa bits: round ONCE, after bias and residual were added in fp32 (a bf16 ->
                                // fp32 unpack is a shift / mask).  Three packed roundings per layer cost ~3 dB of image PSNR.
                                float f[32];
#pragma unroll
                                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
                                auto add_packed = [&](const uint4* q4) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const uint32_t w[4] = {q4[j].x, q4[j].y, q4[j].z, q4[j].w};
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            f[8 * j + 2 * e] += __uint_as_float(w[e] << 16);
                                            f[8 * j + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
                                        }
                                    }
                                };
                                if (brow) add_packed(b4);
                                if (rrow) add_packed(r4);
#pragma unroll
                                for (int i = 0; i < 16; ++i) h[i] = pack2<T>(f[2 * i], f[2 * i + 1]);
                            } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) h[i] = pack2<T>(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                            if (brow) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    h[4 * j] = add2<T>(h[4 * j], b4[j].x); h[4 * j + 1] = add2<T>(h[4 * j + 1], b4[j].y);
                                    h[4 * j + 2] = add2<T>(h[4 * j + 2], b4[j].z); h[4 * j + 3] = add2<T>(h[4 * j + 3], b4[j].w);
                                }
                            }
                            if (rrow) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    h[4 * j] = add2<T>(h[4 * j], r4[j].x); h[4 * j + 1] = add2<T>(h[4 * j + 1], r4[j].y);
                                    h[4 * j + 2] = add2<T>(h[4 * j + 2], r4[j].z); h[4 * j + 3] = add2<T>(h[4 * j + 3], r4[j].w);
                                }
                            }
                            }
                            const int piece = cc >> 3;  // 32 output columns = 4 pieces
#pragma unroll
                            for (int j = 0; j < 4; ++j) stage16(piece + j, h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                        }
                        continue;
                    }
                    // ---- generic path: split-K partials, fp32 time-embedding bias (resnet conv1) ----
                    float v[32];
                    tmem_ld32(tacc + (uint32_t)c0, v);  // warp-collective
                    if (!col_ok) continue;
                    if (p.partial) {  // split-K: raw fp32 partial sums, epilogue happens in splitk_reduce_k
                        if (row_ok) {
                            float* dst = p.partial + ((long)split * p.M + m) * p.N + n;
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                        continue;
                    }
                    if (bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float b8[8];
                            load8<T>(bias + n + j, b8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[j + i] += b8[i];
                        }
                    }
                    if (p.bias2) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 b4 = *reinterpret_cast<const float4*>(p.bias2 + n + j);
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    }
                    if (res && row_ok && !p.geglu) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float r8[8];
                            load8<T>(res + m * p.ldr + n + j, r8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[j + i] += r8[i];
                        }
                    }
                    if (p.geglu) {
                        uint32_t h[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            h[i] = pack2<T>(v[4 * i] * gelu_f(v[4 * i + 1]), v[4 * i + 2] * gelu_f(v[4 * i + 3]));
                        const int piece = cc >> 4;
                        stage16(piece, h[0], h[1], h[2], h[3]);
                        stage16(piece + 1, h[4], h[5], h[6], h[7]);
                    } else {
                        uint32_t h[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) h[i] = pack2<T>(v[2 * i], v[2 * i + 1]);
                        const int piece = cc >> 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j) stage16(piece + j, h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                    }
                }
                if (ps == S::PASSES - 1) {
                    tc_fence_before();  // TMEM reads of this accumulator are done
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (!p.partial) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // panel writes -> visible to the TMA unit
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (leader) {
                        const int pwo = p.geglu ? PW / 2 : PW;                    // output columns of this pass
                        const int ncol0 = (p.geglu ? n0 / 2 : n0) + ps * pwo;
                        const int nout = p.geglu ? p.N / 2 : p.N;
                        const int full = pwo / 64;
                        for (int k = 0; k < full; ++k)
                            if (ncol0 + k * 64 < nout)
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                                 reinterpret_cast<uint64_t>(&tmC)), "r"(epi + (uint32_t)k * 16384u),
                                             "r"(ncol0 + k * 64), "r"((int)m0) : "memory");
                        if ((pwo & 63) && ncol0 + full * 64 < nout)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&tmC2)), "r"(epi + (uint32_t)full * 16384u),
                                         "r"(ncol0 + full * 64), "r"((int)m0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            if (tr) tr[7] = clock64();
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// split-K second pass: C = epilogue(sum_s partial[s]) in a fixed order (deterministic)
template <typename T>
__global__ void splitk_reduce_k(const float* __restrict__ partial, int splits, TcParams p) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;  // over M * N / 4
    long total = p.M * p.N / 4;
    if (i >= total) return;
    long m = (i * 4) / p.N;
    int n = (int)((i * 4) % p.N);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        float4 v = *reinterpret_cast<const float4*>(partial + ((long)s * p.M + m) * p.N + n);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* res = reinterpret_cast<const T*>(p.residual);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (bias) v[j] += to_f<T>(bias[n + j]);
        if (p.bias2) v[j] += p.bias2[n + j];
    }
    T* C = reinterpret_cast<T*>(p.C);
    if (p.geglu) {
        C[m * p.ldc + n / 2] = from_f<T>(v[0] * gelu_f(v[1]));
        C[m * p.ldc + n / 2 + 1] = from_f<T>(v[2] * gelu_f(v[3]));
    } else {
        if (res) {
            float r4[4];
            load4<T>(res + m * p.ldr + n, r4);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += r4[j];
        }
        store4<T>(C + m * p.ldc + n, v);
    }
}

bool gemm_tc_bn320_disabled() {
    static const bool v = [] { const char* e = getenv("ETAI_GEMM_NO_BN320"); return e && e[0] == '1'; }();
    return v;
}

// ETAI_GEMM_VARIANT is read on every call (a getenv is ~100 ns against a >= 5 us launch; graph replays do not come here)
int gemm_tc_variant() {
    const char* e = getenv("ETAI_GEMM_VARIANT");
    return e ? atoi(e) : 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// B-stationary schedule: grid = groups x tiles_n CTAs, every group walks its own run of M tiles
int bstat_groups(const TcParams& p) { return num_sms() / p.tiles_n; }

template <typename T, int BN, bool CONV, int SCHED = 0, int EB = 1>
void launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2, const TcParams& p,
            cudaStream_t s) {
    using S = Smem<BN, SCHED, EB>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_k<T, BN, CONV, SCHED, EB>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured = true;
    }
    int items = p.tiles_m * p.tiles_n * p.splits;
    int grid = items < num_sms() ? items : num_sms();
    if (SCHED == 1) grid = bstat_groups(p) * p.tiles_n;
    static const bool trace = [] { const char* e = getenv("ETAI_GEMM_TRACE"); return e && e[0] == '1'; }();
    if (trace) {  // debugging aid: synchronous, prints the mean per-tile timeline of CTA 0 and of all CTAs
        TcParams q = p;
        size_t n = (size_t)grid * TRACE_TILES * TRACE_SLOTS;
        CUDA_CHECK(cudaMalloc((void**)&q.trace, n * sizeof(long long)));
        CUDA_CHECK(cudaMemset(q.trace, 0, n * sizeof(long long)));
        gemm_tc_k<T, BN, CONV, SCHED, EB><<<grid, GT_THREADS, S::TOTAL, s>>>(tmA, tmB, tmC, tmC2, q);
        CUDA_CHECK(cudaStreamSynchronize(s));
        std::vector<long long> h(n);
        CUDA_CHECK(cudaMemcpy(h.data(), q.trace, n * sizeof(long long), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaFree(q.trace));
        double sum[6] = {0}; long cnt = 0;
        for (int c = 0; c < grid; ++c) {
            int nt = (items - c + grid - 1) / grid;
            if (SCHED == 1) {
                const int ng = bstat_groups(p), g = c / p.tiles_n;
                nt = (int)((long)(g + 1) * p.tiles_m / ng) - (int)((long)g * p.tiles_m / ng);
            }
            if (nt > TRACE_TILES) nt = TRACE_TILES;
            for (int t = 1; t + 1 < nt; ++t) {  // steady state: skip the first and the last tile
                const long long* r = &h[((size_t)c * TRACE_TILES + t) * TRACE_SLOTS];
                const long long* rn = r + TRACE_SLOTS;
                sum[0] += (double)(rn[2] - r[2]);   // tile period at the MMA thread
                sum[1] += (double)(r[3] - r[2]);    // wait for the accumulator buffer
                sum[2] += (double)(r[4] - r[3]);    // wait for the first k-block
                sum[3] += (double)(r[5] - r[4]);    // issue of the main loop (first k-block landed -> last commit)
                sum[4] += (double)(r[7] - r[6]);    // epilogue (accumulator complete -> buffer released)
                sum[5] += (double)(r[1] - r[0]);    // producer: first empty-wait -> last TMA issue of the tile
                ++cnt;
            }
        }
        if (cnt)
            printf("gemm_tc trace BN=%d conv=%d sched=%d epi_bufs=%d stages=%d M=%ld N=%d kb=%d geglu=%d tiles/CTA=%.1f | cycles per tile: "
                   "period %.0f, acc wait %.0f, first-kb wait %.0f, mainloop issue %.0f, epilogue %.0f, producer %.0f\n", BN,
                   (int)CONV, SCHED, EB, S::STAGES, p.M, p.N, p.num_kb, p.geglu, (double)items / grid, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, sum[4] / cnt, sum[5] / cnt);
        return;
    }
    gemm_tc_k<T, BN, CONV, SCHED, EB><<<grid, GT_THREADS, S::TOTAL, s>>>(tmA, tmB, tmC, tmC2, p);
    KERNEL_CHECK();
    if (p.partial) {
        long total = p.M * p.N / 4;
        splitk_reduce_k<T><<<cdiv(total, 256), 256, 0, s>>>(p.partial, p.splits, p);
        KERNEL_CHECK();
    }
}

}  // namespace

size_t gemm_tc_splitk_workspace_bytes(long M, int N) { return (size_t)16 * M * N * sizeof(float); }

bool gemm_tc_supported(const GemmArgs& a) {
    if (a.dtype != ETAI_F16 && a.dtype != ETAI_BF16) return false;
    if (a.N % 32 != 0 || a.M < 1) return false;
    if (a.rowbias && a.rows_per_group < a.M) return false;  // only a single shared fp32 bias vector is fused
    if (a.geglu && a.N % 64 != 0) return false;
    if (a.conv) {
        if (a.Cin % 64 != 0) return false;
        if (a.stride == 1) {
            int W = a.Wd, H = a.H;
            if (a.pad != 1) return false;
            if (W > 128) return W % 128 == 0;  // one tile = 128 consecutive pixels of one image row
            if (128 % W != 0) return false;
            int bh = 128 / W < H ? 128 / W : H;
            if (H % bh != 0) return false;
            return true;
        }
        return a.stride == 2;  // im2col + dense
    }
    return a.K % 64 == 0 && a.lda % 8 == 0 && a.lda >= a.K;
}

void gemm_tc(const GemmArgs& a0, void* ws, size_t ws_bytes, cudaStream_t s) {
    ETAI_CHECK(gemm_tc_supported(a0), ETAI_ERR_UNSUPPORTED, "gemm_tc: unsupported problem");
    GemmArgs a = a0;
    size_t ws_used = 0;
    if (a.conv && a.stride == 2) {
        // the three downsample convs (0.7% of UNet FLOPs): gather to [M, 9*Cin] once, then a dense GEMM
        size_t need = (size_t)a.M * 9 * a.Cin * 2;
        ETAI_CHECK(ws && ws_bytes >= need, ETAI_ERR_ARG, "gemm_tc: stride-2 conv needs an im2col workspace");
        im2col3x3(a.A, ws, a.B, a.H, a.Wd, a.Cin, 2, a.pad, a.Ho, a.Wo, a.dtype, s);
        a.A = ws; a.conv = 0; a.lda = 9L * a.Cin; a.K = 9 * a.Cin;
        ws_used = (need + 255) & ~size_t(255);
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.C = a.C; p.bias = a.bias; p.bias2 = (const float*)a.rowbias; p.residual = a.residual;
    p.M = a.M; p.N = a.N; p.ldc = a.ldc; p.ldr = a.ldr; p.geglu = a.geglu;
    p.fmt = a.dtype == ETAI_BF16 ? 1 : 0;
    // Tile width.  The main loop of a 128 x 160 tile costs ~600 cycles per 64-wide k-block against 320 cycles of tensor-pipe
    // work, and that does not change when TMA moves 128 instead of 288 operand rows per k-block (the B-stationary schedule,
    // profiles/r02_gemm_role_trace.txt): operand delivery is not what paces it.  What fits the measurements is the shared-
    // memory port: an SS-mode UMMA fetches A and B from shared memory for every instruction (9 KB per 128 x 160 x 16, ~96 cycles
    // measured in isolation, profiles/r01_ubench_umma_latency.txt) while TMA writes and the epilogue's staging traffic use the
    // same 128 B/clk.  A 128 x 320 tile (two accumulators) amortises better per column and is used when it still fills the
    // SMs (or split-K will).
    p.tiles_m = cdiv(a.M, BM);
    int BN = (a.N % 160 == 0) ? 160 : 128;
    if (a.N % 320 == 0 && !gemm_tc_bn320_disabled()) {
        // Cost model in SM cycles per CTA, fitted to profiles/r02_bench_ops_all.txt (M = 4096..65536, K = 320..5120): a 128 x 160
        // tile costs ~590 cycles per k-block + ~3000 per tile; a 128 x 320 tile ~930 per k-block + ~9900 per tile
        // (its accumulators are single-buffered, so the epilogue is exposed).  Per unit of work the wide tile wins from
        // K ~ 1000 on -- every conv3x3 (K >= 2880) and the K >= 1280 projections -- unless wave quantisation says otherwise.
        const long t320 = (long)p.tiles_m * (a.N / 320), t160 = 2 * t320;
        const long nk = (a.conv ? 9 * a.Cin : a.K) / BK;
        auto waves = [&](long tiles) { return (long)cdiv(tiles, num_sms()); };
        const long c160 = waves(t160) * (590 * nk + 3000), c320 = waves(t320) * (930 * nk + 9900);
        const bool splitk320 = t320 * 2 <= num_sms() && nk >= 36 && ws != nullptr;
        static const bool force = [] { const char* e = getenv("ETAI_GEMM_FORCE_BN320"); return e && e[0] == '1'; }();
        if (force || splitk320 || c320 < c160) BN = 320;
    }
    p.tiles_n = cdiv(a.N, BN);

    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
        uint64_t str[1] = {(uint64_t)a.K * 2};
        uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN == 320 ? 160 : BN)};
        tmB = make_tmap_16bit(a.W, a.dtype, 2, dims, str, box);
    }
    if (a.conv) {
        int W = a.Wd, H = a.H;
        int bw = W < 128 ? W : 128, bh = W >= 128 ? 1 : (128 / W < H ? 128 / W : H), bb = 128 / (bw * bh);
        uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)W, (uint64_t)H, (uint64_t)a.B};
        uint64_t str[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.Cin * 2 * W, (uint64_t)a.Cin * 2 * W * H};
        uint32_t box[4] = {(uint32_t)BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
        tmA = make_tmap_16bit(a.A, a.dtype, 4, dims, str, box);
        p.cin_blocks = a.Cin / BK; p.num_kb = 9 * p.cin_blocks;
        p.Himg = H; p.Wimg = W;
    } else {
        uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
        uint64_t str[1] = {(uint64_t)a.lda * 2};
        uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
        tmA = make_tmap_16bit(a.A, a.dtype, 2, dims, str, box);
        p.num_kb = a.K / BK;
    }
    // C leaves in column panels: 64-column panels ([128 rows][128 B], 128B swizzle) through tmC and, when the pass width is
    // not a multiple of 64 output columns, a narrow dense tail panel (32 columns, or 16 for GEGLU) through tmC2
    CUtensorMap tmC, tmC2;
    {
        const int n_out = a.geglu ? a.N / 2 : a.N;
        uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)a.M};
        uint64_t str[1] = {(uint64_t)a.ldc * 2};
        uint32_t box[2] = {64, 128};
        tmC = make_tmap_16bit(a.C, a.dtype, 2, dims, str, box, /*swizzle128=*/true);
        uint32_t box2[2] = {(uint32_t)(a.geglu ? 16 : 32), 128};
        tmC2 = make_tmap_16bit(a.C, a.dtype, 2, dims, str, box2, /*swizzle128=*/false);
    }
    // split-K for the low-resolution layers (few output tiles, K up to 23040): fill the SMs with K slices, fp32
    // partials in the workspace, fixed-order reduction + epilogue in a second launch
    p.splits = 1;
    p.kb_per_split = p.num_kb;
    const int tiles = p.tiles_m * p.tiles_n;
    if (tiles * 2 <= num_sms() && p.num_kb >= 36 && ws != nullptr) {
        int want = num_sms() / tiles;
        if (want > 16) want = 16;
        if (want > p.num_kb / 4) want = p.num_kb / 4;
        size_t avail = ws_bytes > ws_used ? ws_bytes - ws_used : 0;
        while (want > 1 && (size_t)want * a.M * a.N * sizeof(float) > avail) --want;
        if (want > 1) {
            p.kb_per_split = cdiv(p.num_kb, want);
            p.splits = cdiv(p.num_kb, p.kb_per_split);
            p.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ws_used);
        }
    }
    // Kernel variant (ETAI_GEMM_VARIANT, bit mask; results are bit-identical across variants, only the schedule differs):
    //   bit 0: two epilogue staging buffers for the 128 / 160-wide tiles;
    //   bit 1: B-stationary schedule for the dense K <= 320 projections at N % 160 == 0 when every group gets M tiles.
    // Both are opt-in experiments: measured on a B200 (profiles/r02_gemm_variants_ab.txt) they are bit-identical to the
    // default on every production shape and worth 0-7 % per shape, ~1 % over the UNet's dense GEMMs -- the default stays.
    const int variant = gemm_tc_variant();
    const bool eb2 = (variant & 1) && BN != 320;
    const bool bstat = (variant & 2) && BN == 160 && !a.conv && p.splits == 1 && p.num_kb <= 5 &&
                       p.tiles_n * 2 <= num_sms() && p.tiles_m >= 2 * bstat_groups(p);
#define LAUNCH(T)                                                            \
    do {                                                                     \
        if (BN == 320) {                                                     \
            if (a.conv) launch<T, 320, true>(tmA, tmB, tmC, tmC2, p, s);           \
            else launch<T, 320, false>(tmA, tmB, tmC, tmC2, p, s);                 \
        } else if (BN == 160) {                                              \
            if (bstat) launch<T, 160, false, 1, 1>(tmA, tmB, tmC, tmC2, p, s);     \
            else if (eb2) {                                                  \
                if (a.conv) launch<T, 160, true, 0, 2>(tmA, tmB, tmC, tmC2, p, s);  \
                else launch<T, 160, false, 0, 2>(tmA, tmB, tmC, tmC2, p, s);        \
            } else if (a.conv) launch<T, 160, true>(tmA, tmB, tmC, tmC2, p, s);     \
            else launch<T, 160, false>(tmA, tmB, tmC, tmC2, p, s);                 \
        } else if (eb2) {                                                    \
            if (a.conv) launch<T, 128, true, 0, 2>(tmA, tmB, tmC, tmC2, p, s);      \
            else launch<T, 128, false, 0, 2>(tmA, tmB, tmC, tmC2, p, s);            \
        } else {                                                             \
            if (a.conv) launch<T, 128, true>(tmA, tmB, tmC, tmC2, p, s);           \
            else launch<T, 128, false>(tmA, tmB, tmC, tmC2, p, s);                 \
        }                                                                    \
    } while (0)
    if (a.dtype == ETAI_F16) LAUNCH(__half);
    else LAUNCH(__nv_bfloat16);
#undef LAUNCH
}

}  // namespace etai
