// Shared helpers for the etai engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

#include "../../include/etai.h"

namespace etai {

// ---- error plumbing: C++ exceptions inside, integer codes at the C ABI ---------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);

#define ETAI_CHECK(cond, code, msg)                                                                 \
    do {                                                                                            \
        if (!(cond)) throw ::etai::Error((code), std::string(msg) + " [" #cond "] at " __FILE__ ":" + \
                                                     std::to_string(__LINE__));                     \
    } while (0)

#define CUDA_CHECK(x)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess)                                                                      \
            throw ::etai::Error(ETAI_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e_) +     \
                                                   " in " #x " at " __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

inline size_t dtype_size(int dt) { return dt == ETAI_F32 ? 4 : 2; }

// ---- scalar conversion --------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T (4 floats or 8 halves)
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T, int N>
struct alignas(sizeof(T) * N) Pack { T v[N]; };

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&out)[8]) {
    if constexpr (sizeof(T) == 4) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
        out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else {
        Pack<T, 8> v = *reinterpret_cast<const Pack<T, 8>*>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = to_f<T>(v.v[i]);
    }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&in)[8]) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(in[4], in[5], in[6], in[7]);
    } else {
        Pack<T, 8> v;
#pragma unroll
        for (int i = 0; i < 8; ++i) v.v[i] = from_f<T>(in[i]);
        *reinterpret_cast<Pack<T, 8>*>(p) = v;
    }
}
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&out)[4]) {
    Pack<T, 4> v = *reinterpret_cast<const Pack<T, 4>*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = to_f<T>(v.v[i]);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&in)[4]) {
    Pack<T, 4> v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.v[i] = from_f<T>(in[i]);
    *reinterpret_cast<Pack<T, 4>*>(p) = v;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }  // fp32 parity path: accurate exp
// SiLU for 16-bit outputs: x*sigmoid(x) = h + h*tanh(h), h = x/2 -- one MUFU op and two FMA-pipe ops instead of
// ex2 + a full-precision division (~20 instructions); tanh.approx.f32 has a relative error of 2^-11, the size of the
// fp16 output rounding.  The normalisation kernels are issue/MUFU-bound on B200 with the exact form (HBM delivers a 16-bit
// element pair every ~0.7 cycles per SM).  fp32 storage (parity mode) keeps silu_f.
__device__ __forceinline__ float silu_fast(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
template <typename T> __device__ __forceinline__ float silu_for(float x) {
    if constexpr (sizeof(T) == 2) return silu_fast(x);
    else return silu_f(x);
}
// exact erf GELU (diffusers GEGLU uses F.gelu default = erf form)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// dispatch on dtype enum -> template type
#define ETAI_DISPATCH_DTYPE(dt, T, ...)                                          \
    do {                                                                         \
        if ((dt) == ETAI_F32) { using T = float; __VA_ARGS__; }                  \
        else if ((dt) == ETAI_F16) { using T = __half; __VA_ARGS__; }            \
        else if ((dt) == ETAI_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }    \
        else throw ::etai::Error(ETAI_ERR_ARG, "bad dtype");                     \
    } while (0)

}  // namespace etai
