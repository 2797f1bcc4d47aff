// Gaussian kernel weights as scipy.ndimage builds them, passed to kernels by value.
#pragma once
#include "common.cuh"
#include <math.h>

#define NB_GAUSS_MAXR 127
struct GaussW {
    double w[2 * NB_GAUSS_MAXR + 1];
    int r;
};

static inline int gauss_radius(double sigma) { return (int)(4.0 * sigma + 0.5); }

static inline int make_gauss(double sigma, GaussW* g) {
    // scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius), radius = int(truncate * sigma + 0.5), truncate = 4
    int r = gauss_radius(sigma);
    NB_REQUIRE(r <= NB_GAUSS_MAXR, NB_ERR_UNSUPPORTED, "gaussian radius %d too large (max %d)", r, NB_GAUSS_MAXR);
    g->r = r;
    double sum = 0;
    for (int i = -r; i <= r; ++i) {
        g->w[i + r] = exp(-0.5 / (sigma * sigma) * (double)i * (double)i);
        sum += g->w[i + r];
    }
    for (int i = 0; i <= 2 * r; ++i) g->w[i] /= sum;
    return NB_OK;
}


#ifdef __CUDACC__
// Wide kernels (radius > NB_GAUSS_MAXR, e.g. the background-subtraction scale of a real-space filter): the 2 r + 1 weights do
// not fit a kernel argument, so every CTA tabulates them in shared memory from sigma, the way scipy does (exp, then divide by
// the sum; the sum is taken in index order here, pairwise in numpy: a last-digit difference).
__device__ __forceinline__ void gauss_weights_smem(double* s_w, double sigma, int r) {
    const double a = -0.5 / (sigma * sigma);
    for (int i = threadIdx.x; i <= 2 * r; i += blockDim.x) {
        const double d = (double)(i - r);
        s_w[i] = exp(a * d * d);
    }
    __syncthreads();
    double sum = 0;
    for (int i = 0; i <= 2 * r; ++i) sum += s_w[i];
    __syncthreads();
    for (int i = threadIdx.x; i <= 2 * r; i += blockDim.x) s_w[i] /= sum;
    __syncthreads();
}
#endif
