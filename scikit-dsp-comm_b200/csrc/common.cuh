// common.cuh -- shared host/device helpers for libb200dsp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/b200dsp.h"

namespace b200dsp {

// ---- thread-local error string + launch counter (host) --------------------------------
void set_error(const char *fmt, ...);
void count_launch();

#define B200_CHECK_CUDA(expr)                                                            \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::b200dsp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                    \
            return B200DSP_E_CUDA;                                                       \
        }                                                                                \
    } while (0)

#define B200_CHECK_LAUNCH(name)                                                          \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        ::b200dsp::count_launch();                                                       \
        if (_e != cudaSuccess) {                                                         \
            ::b200dsp::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return B200DSP_E_CUDA;                                                       \
        }                                                                                \
    } while (0)

// Opt a kernel into >48 KB of dynamic shared memory (idempotent, cheap).
template <typename K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

constexpr size_t kMaxSmemPerBlock = 227 * 1024;  // sm_100: 227 KB usable per CTA

// ---- sample traits ------------------------------------------------------------------
template <typename S> struct Sample;
template <> struct Sample<float>   { using C = float;  static constexpr int NCH = 1; };
template <> struct Sample<float2>  { using C = float;  static constexpr int NCH = 2; };
template <> struct Sample<double>  { using C = double; static constexpr int NCH = 1; };
template <> struct Sample<double2> { using C = double; static constexpr int NCH = 2; };

__device__ __forceinline__ float   zero_of(float)   { return 0.f; }
__device__ __forceinline__ double  zero_of(double)  { return 0.0; }
__device__ __forceinline__ float2  zero_of(float2)  { return make_float2(0.f, 0.f); }
__device__ __forceinline__ double2 zero_of(double2) { return make_double2(0.0, 0.0); }

// acc += t * x  (real tap times real / complex sample)
__device__ __forceinline__ void tap_fma(float &a, float t, float x) { a = fmaf(t, x, a); }
__device__ __forceinline__ void tap_fma(double &a, double t, double x) { a = fma(t, x, a); }
__device__ __forceinline__ void tap_fma(float2 &a, float t, float2 x) {
    a.x = fmaf(t, x.x, a.x);
    a.y = fmaf(t, x.y, a.y);
}
__device__ __forceinline__ void tap_fma(double2 &a, double t, double2 x) {
    a.x = fma(t, x.x, a.x);
    a.y = fma(t, x.y, a.y);
}

__device__ __forceinline__ float   scale_of(float x, float s)     { return x * s; }
__device__ __forceinline__ double  scale_of(double x, double s)   { return x * s; }
__device__ __forceinline__ float2  scale_of(float2 x, float s)    { return make_float2(x.x * s, x.y * s); }
__device__ __forceinline__ double2 scale_of(double2 x, double s)  { return make_double2(x.x * s, x.y * s); }

}  // namespace b200dsp
