// Micro-benchmark (sm_100a): how fast can ONE SM ingest operand tiles through TMA, and what does a TMA store cost?
// Every CTA (grid = #SMs, persistent) runs a producer thread that keeps `stages` tile loads in flight into a shared-memory
// ring and a consumer thread that frees a stage as soon as it has landed -- the GEMM pipeline without the MMA.
//   load test : per "k-block" one box of RA rows + one box of RB rows, each row 128 B (64 halves, SWIZZLE_128B), read
//               from an L2-resident matrix with a row pitch of `ld` halves (ld = 64: dense lines, ld = 320: one line per
//               pitch like a [M, 320] activation matrix).  Reports bytes / clk / SM and clk per 128-byte row request.
//   store test: one warp stores [32 rows x W bytes] boxes (W = 64 or 128) from shared memory with cp.async.bulk.tensor,
//               two in flight (bulk_group), like the GEMM epilogue.  Reports clk per box and per row.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda && ./tma_rate
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

constexpr int MAX_STAGES = 8;

__global__ void __launch_bounds__(64, 1) load_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                               int RA, int RB, int stages, int iters, int rows_total, int kcols,
                                               long long* cycles_out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    const int stage_bytes = (RA + RB) * 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        uint32_t rng = blockIdx.x * 2654435761u + 12345u;
        for (int i = 0; i < iters; ++i) {
            int s = i % stages;
            mbar_wait(&empty[s], ((i / stages) & 1) ^ 1);
            mbar_expect(&full[s], stage_bytes);
            rng = rng * 1664525u + 1013904223u;
            int ra = (int)((rng >> 8) % (uint32_t)(rows_total - RA)) & ~7;
            rng = rng * 1664525u + 1013904223u;
            int rb = (int)((rng >> 8) % (uint32_t)(rows_total - RB)) & ~7;
            int kc = (i % kcols) * 64;
            tma_load_2d(smem + s * stage_bytes, &tmA, &full[s], kc, ra);
            if (RB > 0) tma_load_2d(smem + s * stage_bytes + RA * 128, &tmB, &full[s], kc, rb);
        }
    } else if (threadIdx.x == 32) {
        for (int i = 0; i < iters; ++i) {
            int s = i % stages;
            mbar_wait(&full[s], (i / stages) & 1);
            mbar_arrive(&empty[s]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(32, 1) store_k(const __grid_constant__ CUtensorMap tmC, int box_bytes, int iters, int rows_total,
                                                int ncols_boxes, long long* cycles_out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 2 * 32 * 128 / 4; i += 32) reinterpret_cast<uint32_t*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        uint32_t rng = blockIdx.x * 2654435761u + 777u;
        for (int i = 0; i < iters; ++i) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            rng = rng * 1664525u + 1013904223u;
            int row = (int)((rng >> 8) % (uint32_t)(rows_total - 32)) & ~31;
            int col = (i % ncols_boxes) * (box_bytes / 2);
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(smem + (i & 1) * 32 * 128)), "r"(col), "r"(row) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows,
                            bool swizzle) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows}, str[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const int rows_total = 65536;
    long long* cyc;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    std::vector<long long> h(sms);
    CK(cudaFuncSetAttribute(load_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    printf("# load test: %d SMs, matrix rows %d, L2 resident\n", sms, rows_total);
    printf("ld_halves RA RB stages  bytes/clk/SM  clk/row-request\n");
    for (int ld : {64, 320, 2880}) {
        void* buf;
        size_t bytes = (size_t)rows_total * ld * 2;
        if (bytes > (96u << 20)) bytes = 96u << 20;      // keep it inside L2
        int rows = (int)(bytes / ((size_t)ld * 2));
        CK(cudaMalloc(&buf, (size_t)rows * ld * 2));
        CK(cudaMemset(buf, 1, (size_t)rows * ld * 2));
        for (int cfg = 0; cfg < 6; ++cfg) {
            const int RAs[6] = {128, 128, 128, 128, 256, 64}, RBs[6] = {160, 80, 128, 0, 0, 0};
            for (int stages : {2, 4, 5, 8}) {
                int RA = RAs[cfg], RB = RBs[cfg];
                if ((RA + RB) * 128 * stages > 200 * 1024) continue;
                CUtensorMap ta = make_map(enc, buf, ld, rows, ld, 64, RA, true);
                CUtensorMap tb = make_map(enc, buf, ld, rows, ld, 64, RB > 0 ? RB : 8, true);
                const int iters = 2000;
                for (int rep = 0; rep < 2; ++rep) {
                    load_k<<<sms, 64, (RA + RB) * 128 * stages + 1024>>>(ta, tb, RA, RB, stages, iters, rows, ld / 64, cyc);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
                double mean = 0;
                for (long long v : h) mean += (double)v;
                mean /= sms;
                printf("%5d %4d %4d %3d   %8.1f   %6.2f\n", ld, RA, RB, stages, (double)iters * (RA + RB) * 128 / mean,
                       mean / ((double)iters * (RA + RB)));
            }
        }
        CK(cudaFree(buf));
    }
    // ---- stores ----
    CK(cudaFuncSetAttribute(store_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
    printf("# store test: one warp per SM, boxes of 32 rows, 2 in flight\n");
    printf("ld_halves box_bytes swizzle  clk/box  clk/row\n");
    for (int ld : {320, 2560}) {
        void* buf;
        CK(cudaMalloc(&buf, (size_t)rows_total * ld * 2));
        for (int bb : {64, 128}) {
            for (int sw = 0; sw < 2; ++sw) {
                if (sw && bb != 128) continue;
                CUtensorMap tc = make_map(enc, buf, ld, rows_total, ld, bb / 2, 32, sw != 0);
                const int iters = 4000;
                for (int rep = 0; rep < 2; ++rep) {
                    store_k<<<sms, 32, 12 * 1024>>>(tc, bb, iters, rows_total, ld * 2 / bb, cyc);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
                double mean = 0;
                for (long long v : h) mean += (double)v;
                mean /= sms;
                printf("%5d %5d %3d   %8.1f  %6.2f\n", ld, bb, sw, mean / iters, mean / iters / 32);
            }
        }
        CK(cudaFree(buf));
    }
    return 0;
}
