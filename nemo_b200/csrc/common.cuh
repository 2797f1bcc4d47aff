// Common helpers for the nemo_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define NB_OK 0
#define NB_ERR_ARG -1
#define NB_ERR_CUDA -2
#define NB_ERR_UNSUPPORTED -3
#define NB_ERR_OVERFLOW -4

#define NB_PREC_F32 0
#define NB_PREC_F64 1

#define NB_MAX_FREQ 4
#define NB_MAX_TAPER 64

void nb_set_error(const char* fmt, ...);

#define NB_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            nb_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return NB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define NB_CHECK_LAUNCH()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            nb_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return NB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define NB_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            nb_set_error(__VA_ARGS__);                                                        \
            return (code);                                                                    \
        }                                                                                     \
    } while (0)

template <typename T> struct V2;
template <> struct V2<float> { typedef float2 type; };
template <> struct V2<double> { typedef double2 type; };

template <typename T> __device__ __forceinline__ typename V2<T>::type mk2(T a, T b) {
    typename V2<T>::type r;
    r.x = a;
    r.y = b;
    return r;
}

// cosine taper of enmap.apod (first sample 0, last sample exactly 1), filled on the host in fp64
struct NbTaper {
    double a[NB_MAX_TAPER];
    int n;
};

template <typename T> __device__ __forceinline__ T taper_at(const NbTaper& tp, int i, int len) {
    // value multiplying sample i of an axis of length len
    T v = (T)1;
    if (i < tp.n) v *= (T)tp.a[i];
    if (i >= len - tp.n) v *= (T)tp.a[len - 1 - i];
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// reflect boundary of scipy.ndimage (mode='reflect': d c b a | a b c d | d c b a)
__host__ __device__ __forceinline__ int reflect_idx(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;   // interior: the common case, no division
    if (n == 1) return 0;
    int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return (i < n) ? i : (p - 1 - i);
}

int nb_sm_count();

// Per-kernel CUDA-event timing (nb_prof_enable / nb_prof_report of the C ABI).  Off by default: the macros then cost
// one load and a branch.  Every launch site is wrapped as { NB_PROF_BEGIN(st); kernel<<<..., st>>>(...); NB_PROF_END("name", st); }
extern int g_nb_prof_on;
extern unsigned long long g_nb_launches;  // kernels launched by this library since load (nb_launch_count)
void nb_prof_begin_(cudaStream_t st);
void nb_prof_end_(const char* name, cudaStream_t st);
#define NB_PROF_BEGIN(st)                 \
    do {                                  \
        if (g_nb_prof_on) nb_prof_begin_(st); \
    } while (0)
#define NB_PROF_END(name, st)                 \
    do {                                      \
        __atomic_add_fetch(&g_nb_launches, 1ULL, __ATOMIC_RELAXED); \
        if (g_nb_prof_on) nb_prof_end_(name, st); \
    } while (0)
