// Cross-scale / cross-tile catalog merge on the device: the decision of catalogs.makeOptimalCatalog
// (nemo/catalogs.py:171-183) and of the duplicate removal after the all-gather (nemo/pipelines.py:329-331, catalogs.py:520-565).
//
// For every row c (a "centre") the reference measures astCoords.calcAngSepDeg(c, all rows) -- a tangent-plane separation about c,
// not symmetric in its arguments -- takes the rows closer than the cross-match radius (c itself included) and, when there is more
// than one, flags all of them except the one with the highest S/N (np.argmax: the lowest row index among equals).  A row survives
// when no centre flags it.  The rule is local and the catalogs are small (10^4 - 10^5 rows), so it is evaluated directly: one
// thread per centre walks all rows, staged through shared memory in tiles with their sines and cosines; a declination window
// rejects almost every pair after one subtraction and the full fp64 separation is computed for the handful that remain.  The
// first walk finds the group's best row, the second flags the others (plain byte stores: every writer writes 1).
// Centres [firstCentre, firstCentre + numCentres) only: ranks of a multi-GPU run split the centres and OR their masks.
#include "nemo_b200_internal.h"
#include <math.h>

#define XM_THREADS 128
#define XM_TILE 128

__global__ void __launch_bounds__(XM_THREADS) xmatch_keep_best_kernel(const double* __restrict__ ra, const double* __restrict__ dec,
                                                                      const double* __restrict__ snr, int n, int firstCentre, int numCentres,
                                                                      double radiusDeg, unsigned char* __restrict__ toRemove) {
    __shared__ double s_a[XM_TILE], s_sd[XM_TILE], s_cd[XM_TILE], s_dec[XM_TILE], s_snr[XM_TILE];
    const double D2R = 0.017453292519943295, R2D = 57.29577951308232;      // np.radians / np.degrees multiply by these
    const int c = firstCentre + blockIdx.x * XM_THREADS + threadIdx.x;
    const bool active = c < firstCentre + numCentres && c < n;
    double a0 = 0, sd0 = 0, cd0 = 1, dec0 = 0;
    if (active) {
        a0 = ra[c] * D2R;
        dec0 = dec[c];
        sincos(dec0 * D2R, &sd0, &cd0);
    }
    const double window = radiusDeg * 1.01;
    // separation of row j of the current tile from this thread's centre (nemo_b200/wcs.py calcAngSepDeg, same operations)
    auto sepDeg = [&](int j) -> double {
        double sda, cda;
        sincos(s_a[j] - a0, &sda, &cda);
        const double cosC = s_sd[j] * sd0 + s_cd[j] * cd0 * cda;
        const double x = cd0 * sda / cosC;
        const double y = (s_cd[j] * sd0 - s_sd[j] * cd0 * cda) / cosC;
        return sqrt(x * x + y * y) * R2D;
    };
    int best = -1, members = 0;
    double bestSNR = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int t0 = 0; t0 < n; t0 += XM_TILE) {
            __syncthreads();
            const int j = t0 + threadIdx.x;
            if (threadIdx.x < XM_TILE && j < n) {
                const double d = dec[j];
                s_a[threadIdx.x] = ra[j] * D2R;
                s_dec[threadIdx.x] = d;
                sincos(d * D2R, &s_sd[threadIdx.x], &s_cd[threadIdx.x]);
                s_snr[threadIdx.x] = snr[j];
            }
            __syncthreads();
            if (!active || (pass == 1 && members < 2)) continue;
            const int m = min(XM_TILE, n - t0);
            for (int k = 0; k < m; ++k) {
                if (fabs(s_dec[k] - dec0) > window) continue;
                if (!(sepDeg(k) < radiusDeg)) continue;
                if (pass == 0) {
                    ++members;
                    if (best < 0 || s_snr[k] > bestSNR) {     // rows come in ascending order: the first maximum wins
                        best = t0 + k;
                        bestSNR = s_snr[k];
                    }
                } else if (t0 + k != best) {
                    toRemove[t0 + k] = 1;
                }
            }
        }
    }
}

extern "C" int nb_xmatch_keep_best(const double* raDeg, const double* decDeg, const double* snr, int n, int firstCentre, int numCentres,
                                   double radiusDeg, unsigned char* toRemove, cudaStream_t st) {
    NB_REQUIRE(raDeg && decDeg && snr && toRemove && n >= 0 && firstCentre >= 0 && numCentres >= 0, NB_ERR_ARG, "nb_xmatch_keep_best: bad arguments");
    NB_REQUIRE(radiusDeg > 0 && radiusDeg < 10.0, NB_ERR_ARG, "nb_xmatch_keep_best: radius %g deg out of range", radiusDeg);
    if (n == 0 || numCentres == 0) return NB_OK;
    const int blocks = (numCentres + XM_THREADS - 1) / XM_THREADS;
    { NB_PROF_BEGIN(st); xmatch_keep_best_kernel<<<blocks, XM_THREADS, 0, st>>>(raDeg, decDeg, snr, n, firstCentre, numCentres, radiusDeg, toRemove); NB_PROF_END("xmatch_keep_best_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}
