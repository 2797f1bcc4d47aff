// Second-order moments of the detected objects for findObjects' SExtractor-style ellipses (measureShapes = True,
// nemo/photometry.py:126-178): for every label of a segmentation map the corner (xMin, yMin) of its bounding box and, with
// xs = x - xMin, ys = y - yMin and d the (sign-flipped when the map is inverted) S/N of the pixel, the six sums
//     sum d, sum xs d, sum ys d, sum xs^2 d, sum ys^2 d, sum xs ys d
// The reference shifts the coordinates by the box corner before it weights them (:134-142), so the sums are taken about the same
// origin here: one sweep finds the corners (integer atomicMin), a second one adds the moments (fp64 atomics; objects are a
// few dozen pixels and well under one per cent of the map, so contention does not matter).  The host turns the sums into
// theta, A, B, e exactly as the reference does (nemo_b200/photometry.py ellipsesFromMoments).
#include "nemo_b200_internal.h"
#include <limits.h>

#define SH_THREADS 256
#define SH_COLS 8      // doubles per table row: 6 sums, xMin, yMin

__global__ void shapes_init_kernel(double* __restrict__ table, int* __restrict__ corner, long long nRows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) return;
#pragma unroll
    for (int k = 0; k < SH_COLS; ++k) table[i * SH_COLS + k] = 0.0;
    corner[2 * i] = INT_MAX;
    corner[2 * i + 1] = INT_MAX;
}

__global__ void __launch_bounds__(SH_THREADS) shapes_corner_kernel(const int* __restrict__ labels, int* __restrict__ corner, int H, int W, int maxLabels) {
    const int HW = H * W;
    const int* lab = labels + (size_t)blockIdx.y * HW;
    int* cor = corner + (size_t)blockIdx.y * (maxLabels + 1) * 2;
    for (int p = blockIdx.x * SH_THREADS + threadIdx.x; p < HW; p += gridDim.x * SH_THREADS) {
        const int l = lab[p];
        if (l <= 0 || l > maxLabels) continue;
        const int y = p / W, x = p - y * W;
        atomicMin(cor + 2 * l, x);
        atomicMin(cor + 2 * l + 1, y);
    }
}

template <typename T>
__global__ void __launch_bounds__(SH_THREADS) shapes_moments_kernel(const T* __restrict__ map, double sign, const int* __restrict__ labels,
                                                                    const int* __restrict__ corner, double* __restrict__ table, int H, int W,
                                                                    int maxLabels) {
    const int HW = H * W;
    const T* m = map + (size_t)blockIdx.y * HW;
    const int* lab = labels + (size_t)blockIdx.y * HW;
    const int* cor = corner + (size_t)blockIdx.y * (maxLabels + 1) * 2;
    double* tab = table + (size_t)blockIdx.y * (maxLabels + 1) * SH_COLS;
    for (int p = blockIdx.x * SH_THREADS + threadIdx.x; p < HW; p += gridDim.x * SH_THREADS) {
        const int l = lab[p];
        if (l <= 0 || l > maxLabels) continue;
        const int y = p / W, x = p - y * W;
        const double xs = (double)(x - cor[2 * l]), ys = (double)(y - cor[2 * l + 1]);
        const double d = sign * (double)m[p];
        double* row = tab + (size_t)l * SH_COLS;
        atomicAdd(row + 0, d);
        atomicAdd(row + 1, xs * d);
        atomicAdd(row + 2, ys * d);
        atomicAdd(row + 3, (xs * xs) * d);
        atomicAdd(row + 4, (ys * ys) * d);
        atomicAdd(row + 5, (xs * ys) * d);
    }
}

__global__ void shapes_final_kernel(double* __restrict__ table, const int* __restrict__ corner, long long nRows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) return;
    table[i * SH_COLS + 6] = corner[2 * i] == INT_MAX ? -1.0 : (double)corner[2 * i];
    table[i * SH_COLS + 7] = corner[2 * i + 1] == INT_MAX ? -1.0 : (double)corner[2 * i + 1];
}

extern "C" int nb_label_shapes(int prec, const void* map, double sign, const int* labels, double* table, int* scratch, int B, int H, int W,
                               int maxLabels, cudaStream_t st) {
    NB_REQUIRE(map && labels && table && scratch && B > 0 && maxLabels >= 0 && H > 0 && W > 0, NB_ERR_ARG, "nb_label_shapes: bad arguments");
    NB_REQUIRE(prec == NB_PREC_F64 || prec == NB_PREC_F32, NB_ERR_ARG, "nb_label_shapes: bad precision %d", prec);
    const long long nRows = (long long)B * (maxLabels + 1);
    const unsigned rowBlocks = (unsigned)((nRows + 255) / 256);
    { NB_PROF_BEGIN(st); shapes_init_kernel<<<rowBlocks, 256, 0, st>>>(table, scratch, nRows); NB_PROF_END("shapes_init_kernel", st); }
    NB_CHECK_LAUNCH();
    int gx = (H * W + SH_THREADS * 8 - 1) / (SH_THREADS * 8);
    dim3 grid(gx < 1 ? 1 : gx, B);
    { NB_PROF_BEGIN(st); shapes_corner_kernel<<<grid, SH_THREADS, 0, st>>>(labels, scratch, H, W, maxLabels); NB_PROF_END("shapes_corner_kernel", st); }
    NB_CHECK_LAUNCH();
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); shapes_moments_kernel<double><<<grid, SH_THREADS, 0, st>>>((const double*)map, sign, labels, scratch, table, H, W, maxLabels); NB_PROF_END("shapes_moments_kernel", st); }
    else
        { NB_PROF_BEGIN(st); shapes_moments_kernel<float><<<grid, SH_THREADS, 0, st>>>((const float*)map, sign, labels, scratch, table, H, W, maxLabels); NB_PROF_END("shapes_moments_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); shapes_final_kernel<<<rowBlocks, 256, 0, st>>>(table, scratch, nRows); NB_PROF_END("shapes_final_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}
