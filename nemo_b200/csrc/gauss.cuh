// Gaussian kernel weights as scipy.ndimage builds them, passed to kernels by value.
#pragma once
#include "common.cuh"
#include <math.h>

#define NB_GAUSS_MAXR 127
struct GaussW {
    double w[2 * NB_GAUSS_MAXR + 1];
    int r;
};

static inline int make_gauss(double sigma, GaussW* g) {
    // scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius), radius = int(truncate * sigma + 0.5), truncate = 4
    int r = (int)(4.0 * sigma + 0.5);
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

