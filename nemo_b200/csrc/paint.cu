// Model-image painter: maps.makeModelImage of the reference (nemo/maps.py:1734-1899) for catalogs of point sources
// or clusters -- for every object a radial profile evaluated on the distance map maps.makeDegreesDistanceMap (:2414-2471)
// fills in a box around it, times the object's amplitude, summed over the objects in catalog order.
//
// One CTA per 32 x 8 pixel block of an image walks the image's object list (staged through shared memory) and keeps, for
// each of its pixels, the running sum over the objects whose box holds the pixel: every pixel is written once, coalesced,
// no atomics, and the order of the additions is the catalog's -- with the reference's own operations (distance = sqrt of the
// sum of two products, numpy.interp's slope * (x - x_lo) + y_lo, all rounded separately) the result is the reference's bit
// for bit.  The pixel window and, for clusters, the beam are applied afterwards by the FFT passes (K8 / K4).
#include "nemo_b200_internal.h"
#include <math.h>

#define PT_TX 32
#define PT_TY 8
#define PT_OBJ 64          // objects staged per round
#define PT_FIELDS 10       // x0, y0, xPixScale, yPixScale, amplitude, minX, maxX, minY, maxY, profile index

// numpy.interp (compiled_base.c arr_interp) for 0 <= r: largest j with xp[j] <= r by bisection; the table value itself at the
// last node or on a node; 0 beyond the table (interp1d's fill_value, signals.py:614)
__device__ __forceinline__ double pt_interp(const double* __restrict__ xp, const double* __restrict__ fp, int n, double r) {
    if (!(r <= xp[n - 1]) || r < xp[0]) return 0.0;
    int lo = 0, hi = n - 1;          // xp[lo] <= r <= xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= r) lo = mid; else hi = mid;
    }
    if (xp[hi] <= r) lo = hi;
    if (lo == n - 1 || xp[lo] == r) return fp[lo];
    const double slope = __ddiv_rn(__dsub_rn(fp[lo + 1], fp[lo]), __dsub_rn(xp[lo + 1], xp[lo]));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(r, xp[lo])), fp[lo]);
}

template <typename T>
__global__ void __launch_bounds__(PT_TX * PT_TY) paint_objects_kernel(const T* __restrict__ base, T* __restrict__ out, int H, int W,
                                                                      const double* __restrict__ obj, const int* __restrict__ objStart,
                                                                      const double* __restrict__ profR, const double* __restrict__ profV,
                                                                      int nR) {
    __shared__ double s_obj[PT_OBJ][PT_FIELDS];
    __shared__ int s_hit[PT_OBJ];
    const int img = blockIdx.z;
    const int tx = threadIdx.x % PT_TX, ty = threadIdx.x / PT_TX;
    const int x = blockIdx.x * PT_TX + tx, y = blockIdx.y * PT_TY + ty;
    const int bx0 = blockIdx.x * PT_TX, by0 = blockIdx.y * PT_TY;
    const bool inside = x < W && y < H;
    double acc = 0.0;
    const int first = objStart[img], last = objStart[img + 1];
    for (int o0 = first; o0 < last; o0 += PT_OBJ) {
        const int m = min(PT_OBJ, last - o0);
        __syncthreads();
        for (int e = threadIdx.x; e < m * PT_FIELDS; e += PT_TX * PT_TY) s_obj[e / PT_FIELDS][e % PT_FIELDS] = obj[(size_t)o0 * PT_FIELDS + e];
        __syncthreads();
        if (threadIdx.x < m) {       // does the object's box [minX, maxX) x [minY, maxY) meet this block of pixels?
            const double* q = s_obj[threadIdx.x];
            s_hit[threadIdx.x] = (q[5] < bx0 + PT_TX && q[6] > bx0 && q[7] < by0 + PT_TY && q[8] > by0) ? 1 : 0;
        }
        __syncthreads();
        if (!inside) continue;
        for (int k = 0; k < m; ++k) {
            if (!s_hit[k]) continue;
            const double* q = s_obj[k];
            if (x < q[5] || x >= q[6] || y < q[7] || y >= q[8]) continue;
            const double xDeg = __dmul_rn(__dsub_rn((double)x, q[0]), q[2]);
            const double yDeg = __dmul_rn(__dsub_rn((double)y, q[1]), q[3]);
            const double r = sqrt(__dadd_rn(__dmul_rn(yDeg, yDeg), __dmul_rn(xDeg, xDeg)));
            const int p = (int)q[9];
            const double v = pt_interp(profR + (size_t)p * nR, profV + (size_t)p * nR, nR, r);
            acc = __dadd_rn(acc, __dmul_rn(v, q[4]));
        }
    }
    if (inside) {
        const size_t off = ((size_t)img * H + y) * W + x;
        out[off] = base ? (T)__dadd_rn((double)base[off], acc) : (T)acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) add_maps_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n,
                                                       long long nB) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i % nB];
}

extern "C" {

int nb_paint_objects(int prec, const void* base, void* out, int nImg, int H, int W, const double* obj, const int* objStart,
                     const double* profR, const double* profV, int nProf, int nR, cudaStream_t st) {
    NB_REQUIRE(out && obj && objStart && profR && profV && nImg > 0 && H > 0 && W > 0 && nProf > 0 && nR >= 2, NB_ERR_ARG,
               "nb_paint_objects: bad arguments");
    NB_REQUIRE(nImg <= 65535 && (H + PT_TY - 1) / PT_TY <= 65535, NB_ERR_UNSUPPORTED, "nb_paint_objects: too many images or rows");
    dim3 grid((W + PT_TX - 1) / PT_TX, (H + PT_TY - 1) / PT_TY, nImg);
    if (prec == NB_PREC_F64) {
        { NB_PROF_BEGIN(st); paint_objects_kernel<double><<<grid, PT_TX * PT_TY, 0, st>>>((const double*)base, (double*)out, H, W, obj, objStart, profR, profV, nR); NB_PROF_END("paint_objects_kernel", st); }
    } else if (prec == NB_PREC_F32) {
        { NB_PROF_BEGIN(st); paint_objects_kernel<float><<<grid, PT_TX * PT_TY, 0, st>>>((const float*)base, (float*)out, H, W, obj, objStart, profR, profV, nR); NB_PROF_END("paint_objects_kernel", st); }
    } else {
        nb_set_error("nb_paint_objects: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_add_maps(int prec, const void* a, const void* b, void* out, long long n, long long nB, cudaStream_t st) {
    NB_REQUIRE(a && b && out && n > 0 && nB > 0 && n % nB == 0, NB_ERR_ARG, "nb_add_maps: bad arguments");
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (prec == NB_PREC_F64) {
        { NB_PROF_BEGIN(st); add_maps_kernel<double><<<blocks, 256, 0, st>>>((const double*)a, (const double*)b, (double*)out, n, nB); NB_PROF_END("add_maps_kernel", st); }
    } else if (prec == NB_PREC_F32) {
        { NB_PROF_BEGIN(st); add_maps_kernel<float><<<blocks, 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, n, nB); NB_PROF_END("add_maps_kernel", st); }
    } else {
        nb_set_error("nb_add_maps: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

}  // extern "C"
