// Bicubic interpolating-spline evaluation at one point by one warp (K11; also used by K5).
//
// scipy's RectBivariateSpline(kx=3, ky=3, s=0) is the interpolating cubic spline on integer knots.  Away from
// the edges it equals cardinal cubic B-spline interpolation:  value = sum_{a,b} By[a] Bx[b] c[iy-1+a, ix-1+b]
// with c = (h (x) h) * img and h[n] = sqrt(3) (sqrt(3) - 2)^|n| the exact inverse of the B-spline sampling
// filter (1, 4, 1)/6.  h is truncated at |n| <= 30 (0.268^30 ~ 7e-18), which folds the prefilter and the 4x4
// B-spline stencil into one separable 64 x 64 window:  value = ry^T S rx.
#pragma once
#include "common.cuh"

#define NB_SPL_R 30
#define NB_SPL_WIN (2 * NB_SPL_R + 4)  // 64

__device__ __forceinline__ int mirror_idx(int i, int n) {
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return (i < n) ? i : (p - i);
}

__device__ __forceinline__ void bspline_w(double f, double* w) {
    const double f2 = f * f, f3 = f2 * f, omf = 1.0 - f;
    w[0] = omf * omf * omf / 6.0;
    w[1] = (3.0 * f3 - 6.0 * f2 + 4.0) / 6.0;
    w[2] = (-3.0 * f3 + 3.0 * f2 + 3.0 * f + 1.0) / 6.0;
    w[3] = f3 / 6.0;
}

// combined 1-D response at window position i (0..63), i.e. image index i0 - 1 - R + i
__device__ __forceinline__ double spl_resp(const double* w, int i) {
    const double z1 = -0.26794919243112270647;  // sqrt(3) - 2
    const double s3 = 1.73205080756887729353;
    double r = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int d = i - (NB_SPL_R + a);
        if (d < 0) d = -d;
        r += w[a] * s3 * pow(z1, (double)d);
    }
    return r;
}

// All 32 lanes must call; every lane returns the value.
template <typename T>
__device__ double spline_eval_warp(const T* __restrict__ img, int H, int W, double y, double x, double sign) {
    const int lane = threadIdx.x & 31;
    int iy = (int)floor(y), ix = (int)floor(x);
    double wy[4], wx[4];
    bspline_w(y - (double)iy, wy);
    bspline_w(x - (double)ix, wx);
    // each lane owns two window columns
    const int j0 = lane, j1 = lane + 32;
    const double rx0 = spl_resp(wx, j0), rx1 = spl_resp(wx, j1);
    const int c0 = mirror_idx(ix - 1 - NB_SPL_R + j0, W), c1 = mirror_idx(ix - 1 - NB_SPL_R + j1, W);
    // lane i also computes ry for rows i and i+32, broadcast by shuffle
    const double ryA = spl_resp(wy, lane), ryB = spl_resp(wy, lane + 32);
    double acc = 0;
    for (int i = 0; i < NB_SPL_WIN; ++i) {
        const double ry = __shfl_sync(0xffffffffu, (i < 32) ? ryA : ryB, i & 31);
        const int r = mirror_idx(iy - 1 - NB_SPL_R + i, H);
        const T* row = img + (size_t)r * W;
        acc += ry * (rx0 * (double)row[c0] + rx1 * (double)row[c1]);
    }
    acc = warp_sum(acc);
    return sign * acc;
}
