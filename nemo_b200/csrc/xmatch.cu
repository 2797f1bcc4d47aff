// Cross-scale / cross-tile catalog merge on the device: the decision of catalogs.makeOptimalCatalog
// (nemo/catalogs.py:171-183) and of the duplicate removal after the all-gather (nemo/pipelines.py:329-331, catalogs.py:520-565).
//
// For every row c (a "centre") the reference measures astCoords.calcAngSepDeg(c, all rows) -- a tangent-plane separation about c,
// not symmetric in its arguments -- takes the rows closer than the cross-match radius (c itself included) and, when there is more
// than one, flags all of them except the one with the highest S/N (np.argmax: the lowest row index among equals).  A row survives
// when no centre flags it.  The rule is local: rows further than 1.01 radii from the centre in declination cannot match, so the
// rows are first binned by declination (bins a little wider than that window: count, scan, scatter -- a cell list) and a centre
// only looks at its own bin and the two next to it.  One WARP per centre: the lanes share the candidate rows, each keeps the best
// row of its share (higher S/N, lower row index among equals -- the order inside a bin therefore does not matter), a shuffle
// reduction picks the group's best, and a second walk flags the others (plain byte stores: every writer writes 1).  A step's
// 2 x 10^4 rows take tens of microseconds; the all-pairs walk this replaces took 5.9 ms, at the end of the step where nothing
// overlaps it.
// Centres [firstCentre, firstCentre + numCentres) only: ranks of a multi-GPU run split the centres and OR their masks.
#include "nemo_b200_internal.h"
#include <math.h>

#define XM_THREADS 256
#define XM_MAX_BINS 65536

struct XmBins {
    double width;      // degrees of declination per bin (>= 1.001 x the match window)
    int numBins;
};

static XmBins xm_bins(double radiusDeg) {
    XmBins b;
    b.width = radiusDeg * 1.01 * 1.001;
    if (b.width < 180.0 / (XM_MAX_BINS - 2)) b.width = 180.0 / (XM_MAX_BINS - 2);
    b.numBins = (int)ceil(180.0 / b.width) + 2;
    return b;
}

__device__ __forceinline__ int xm_bin_of(double decDeg, double width, int numBins) {
    const double v = floor((decDeg + 90.0) / width);
    return !(v > 0.0) ? 0 : (v >= (double)(numBins - 1) ? numBins - 1 : (int)v);      // NaN -> bin 0 (it never matches anything)
}

__global__ void xmatch_count_kernel(const double* __restrict__ dec, int n, double width, int numBins, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(counts + xm_bin_of(dec[i], width, numBins), 1);
}

// exclusive scan of counts [numBins] into start [numBins + 1] and a copy into cursor [numBins] (one block)
__global__ void __launch_bounds__(1024) xmatch_scan_kernel(const int* __restrict__ counts, int numBins, int* __restrict__ start, int* __restrict__ cursor) {
    __shared__ int s_part[1024];
    const int per = (numBins + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, numBins), hi = min(lo + per, numBins);
    int sum = 0;
    for (int b = lo; b < hi; ++b) sum += counts[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int b = lo; b < hi; ++b) {
        start[b] = run;
        cursor[b] = run;
        run += counts[b];
    }
    if (threadIdx.x == 1023) start[numBins] = s_part[1023];
}

__global__ void xmatch_scatter_kernel(const double* __restrict__ dec, int n, double width, int numBins, int* __restrict__ cursor, int* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(cursor + xm_bin_of(dec[i], width, numBins), 1)] = i;
}

__global__ void __launch_bounds__(XM_THREADS) xmatch_keep_best_kernel(const double* __restrict__ ra, const double* __restrict__ dec,
                                                                      const double* __restrict__ snr, int n, int firstCentre, int numCentres,
                                                                      double radiusDeg, double width, int numBins, const int* __restrict__ start,
                                                                      const int* __restrict__ order, unsigned char* __restrict__ toRemove) {
    const double D2R = 0.017453292519943295, R2D = 57.29577951308232;      // np.radians / np.degrees multiply by these
    const int lane = threadIdx.x & 31;
    const int c = firstCentre + (int)((blockIdx.x * (unsigned)XM_THREADS + threadIdx.x) >> 5);
    if (c >= firstCentre + numCentres || c >= n) return;      // whole warps leave together
    const double dec0 = dec[c], a0 = ra[c] * D2R;
    double sd0, cd0;
    sincos(dec0 * D2R, &sd0, &cd0);
    const double window = radiusDeg * 1.01;
    const int b0 = xm_bin_of(dec0, width, numBins);
    const int kLo = start[max(b0 - 1, 0)], kHi = start[min(b0 + 2, numBins)];
    // separation of row j from this warp's centre (nemo_b200/wcs.py calcAngSepDeg, same operations)
    auto matches = [&](int j) -> bool {
        const double d = dec[j];
        if (fabs(d - dec0) > window) return false;
        double sd, cd, sda, cda;
        sincos(d * D2R, &sd, &cd);
        sincos(ra[j] * D2R - a0, &sda, &cda);
        const double cosC = sd * sd0 + cd * cd0 * cda;
        const double x = cd0 * sda / cosC;
        const double y = (cd * sd0 - sd * cd0 * cda) / cosC;
        return sqrt(x * x + y * y) * R2D < radiusDeg;
    };
    int best = 0x7fffffff, members = 0;
    double bestSNR = 0;
    for (int k = kLo + lane; k < kHi; k += 32) {
        const int j = order[k];
        if (!matches(j)) continue;
        ++members;
        const double s = snr[j];
        if (best == 0x7fffffff || s > bestSNR || (s == bestSNR && j < best)) {      // np.argmax: the first maximum wins
            best = j;
            bestSNR = s;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, off);
        const double os = __shfl_xor_sync(0xffffffffu, bestSNR, off);
        members += __shfl_xor_sync(0xffffffffu, members, off);
        if (ob != 0x7fffffff && (best == 0x7fffffff || os > bestSNR || (os == bestSNR && ob < best))) {
            best = ob;
            bestSNR = os;
        }
    }
    if (members < 2) return;
    for (int k = kLo + lane; k < kHi; k += 32) {
        const int j = order[k];
        if (j != best && matches(j)) toRemove[j] = 1;
    }
}

extern "C" long long nb_xmatch_scratch_ints(int n, double radiusDeg) {
    if (n < 0 || !(radiusDeg > 0)) return 0;
    const XmBins b = xm_bins(radiusDeg);
    return (long long)3 * (b.numBins + 1) + n;
}

extern "C" int nb_xmatch_keep_best(const double* raDeg, const double* decDeg, const double* snr, int n, int firstCentre, int numCentres,
                                   double radiusDeg, unsigned char* toRemove, int* scratch, cudaStream_t st) {
    NB_REQUIRE(raDeg && decDeg && snr && toRemove && scratch && n >= 0 && firstCentre >= 0 && numCentres >= 0, NB_ERR_ARG, "nb_xmatch_keep_best: bad arguments");
    NB_REQUIRE(radiusDeg > 0 && radiusDeg < 10.0, NB_ERR_ARG, "nb_xmatch_keep_best: radius %g deg out of range", radiusDeg);
    if (n == 0 || numCentres == 0) return NB_OK;
    const XmBins b = xm_bins(radiusDeg);
    int* counts = scratch;
    int* start = counts + (b.numBins + 1);
    int* cursor = start + (b.numBins + 1);
    int* order = cursor + (b.numBins + 1);
    NB_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)b.numBins, st));
    const int rowBlocks = (n + 255) / 256;
    { NB_PROF_BEGIN(st); xmatch_count_kernel<<<rowBlocks, 256, 0, st>>>(decDeg, n, b.width, b.numBins, counts); NB_PROF_END("xmatch_count_kernel", st); }
    { NB_PROF_BEGIN(st); xmatch_scan_kernel<<<1, 1024, 0, st>>>(counts, b.numBins, start, cursor); NB_PROF_END("xmatch_scan_kernel", st); }
    { NB_PROF_BEGIN(st); xmatch_scatter_kernel<<<rowBlocks, 256, 0, st>>>(decDeg, n, b.width, b.numBins, cursor, order); NB_PROF_END("xmatch_scatter_kernel", st); }
    const int centres = min(numCentres, n - min(firstCentre, n));
    if (centres <= 0) return NB_OK;
    const int blocks = (int)(((long long)centres * 32 + XM_THREADS - 1) / XM_THREADS);
    { NB_PROF_BEGIN(st); xmatch_keep_best_kernel<<<blocks, XM_THREADS, 0, st>>>(raDeg, decDeg, snr, n, firstCentre, numCentres, radiusDeg, b.width, b.numBins, start, order, toRemove); NB_PROF_END("xmatch_keep_best_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}
