// Noise / signal-to-noise kernels (K6, K7 of SURVEY.md section 2a).
//   nb_median_weights : np.median over frequencies of the weight maps                 filters.py:360-363
//   nb_cell_rms       : MapFilter.makeNoiseMap grid branch, 3-sigma clipped std       filters.py:409-481
//   nb_snr_mask       : S/N, rank_filter edge trim, survey / apod masks               filters.py:716-758, 1112-1161
//   nb_snr_from_rms   : S/N from a cached RMS map                                     pipelines.py:216-228
#include "nemo_b200_internal.h"
#include <math.h>
#include <algorithm>
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) median_weights_kernel(const T* __restrict__ w, T* __restrict__ out, int F, long long HW) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const T* base = w + (size_t)blockIdx.y * F * HW + p;
    T v[NB_MAX_FREQ];
    for (int f = 0; f < F; ++f) v[f] = base[(size_t)f * HW];
    // insertion sort of <= 4 values
    for (int i = 1; i < F; ++i) {
        const T key = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > key) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = key;
    }
    T m;
    if (F & 1)
        m = v[F / 2];
    else
        m = (T)(((double)v[F / 2 - 1] + (double)v[F / 2]) * 0.5);  // np.median: mean of the middle two
    out[(size_t)blockIdx.y * HW + p] = m;
}

// ------------------------------------------------------------------------------------------------------
// K6.  One CTA per (cell, image).
#define CR_THREADS 512
#define CR_WARPS (CR_THREADS / 32)
#define CR_MAXWARPS 32   // scratch of the block reductions below is sized for up to 1024 threads

#include "noise_cells.cuh"

// block-wide sum of two doubles; result valid in every thread
__device__ __forceinline__ void block_sum2(double& a, double& b, double* s_red) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
        s_red[wid] = a;
        s_red[CR_MAXWARPS + wid] = b;
    }
    __syncthreads();
    double ta = 0, tb = 0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) {
        ta += s_red[i];
        tb += s_red[CR_MAXWARPS + i];
    }
    a = ta;
    b = tb;
}

__device__ __forceinline__ void block_minmax(double& mn, double& mx, double* s_red) {
    mn = warp_min(mn);
    mx = warp_max(mx);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
        s_red[wid] = mn;
        s_red[CR_MAXWARPS + wid] = mx;
    }
    __syncthreads();
    double ta = INFINITY, tb = -INFINITY;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) {
        ta = fmin(ta, s_red[i]);
        tb = fmax(tb, s_red[CR_MAXWARPS + i]);
    }
    mn = ta;
    mx = tb;
}

__device__ __forceinline__ unsigned long long enc_w(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dec_w(unsigned long long e) {
    const unsigned long long b = (e & 0x8000000000000000ULL) ? (e & 0x7fffffffffffffffULL) : ~e;
    return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) weight among the good pixels of the cell: MSB-first radix select, 8 bits per pass
template <typename T>
__device__ double select_kth(const T* __restrict__ m, const T* __restrict__ wm, int W, CellGeom g, long long k, int* s_hist,
                             unsigned long long* s_sel) {
    unsigned long long prefix = 0, maskbits = 0;
    const int cw = g.x1 - g.x0;
    const int n = (g.y1 - g.y0) * cw;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int r = e / cw, c = e - r * cw;
            const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
            if (m[off] != (T)0) {
                const unsigned long long key = enc_w((double)wm[off]);
                if ((key & maskbits) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & 255ULL)], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long kk = (long long)s_sel[1];
            if (pass == 0) kk = k;
            int d = 0;
            for (; d < 255; ++d) {
                if (kk < s_hist[d]) break;
                kk -= s_hist[d];
            }
            s_sel[0] = prefix | ((unsigned long long)d << shift);
            s_sel[1] = (unsigned long long)kk;
        }
        __syncthreads();
        prefix = s_sel[0];
        maskbits |= 0xFFULL << shift;
    }
    __syncthreads();
    return dec_w(prefix);
}

// k-th smallest (0-based) |value| among the good pixels of the cell whose weight lies in [wlo, whi): the same radix select on
// the bit pattern of |v| (non-negative doubles order like their bits) -- the 'percentile' RMS estimator (filters.py:464-465)
// SIGNED = false: key |v - shiftBy| (0: the percentile estimator; the median: the median absolute deviation of the biweight
// estimator); SIGNED = true: key v - shiftBy itself through the order-preserving map enc_w (the median)
template <typename T, bool SIGNED = false>
__device__ double select_kth_abs(const T* __restrict__ m, const T* __restrict__ wm, int W, CellGeom g, long long k, double wlo, double whi,
                                 int* s_hist, unsigned long long* s_sel, double shiftBy = 0.0) {
    unsigned long long prefix = 0, maskbits = 0;
    const int cw = g.x1 - g.x0;
    const int n = (g.y1 - g.y0) * cw;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int r = e / cw, c = e - r * cw;
            const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
            const double v = (double)m[off], w = (double)wm[off];
            if (v != 0.0 && w >= wlo && w < whi) {
                const unsigned long long key = SIGNED ? enc_w(v - shiftBy) : (unsigned long long)__double_as_longlong(fabs(v - shiftBy));
                if ((key & maskbits) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & 255ULL)], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long kk = (long long)s_sel[1];
            if (pass == 0) kk = k;
            int d = 0;
            for (; d < 255; ++d) {
                if (kk < s_hist[d]) break;
                kk -= s_hist[d];
            }
            s_sel[0] = prefix | ((unsigned long long)d << shift);
            s_sel[1] = (unsigned long long)kk;
        }
        __syncthreads();
        prefix = s_sel[0];
        maskbits |= 0xFFULL << shift;
    }
    __syncthreads();
    return SIGNED ? dec_w(prefix) : __longlong_as_double((long long)prefix);
}

// np.median of the keys above: the middle order statistic, or the mean of the two middle ones
template <typename T, bool SIGNED>
__device__ double median_of_bin(const T* __restrict__ m, const T* __restrict__ wm, int W, CellGeom g, long long n, double wlo, double whi,
                                int* s_hist, unsigned long long* s_sel, double off) {
    if (n & 1LL) return select_kth_abs<T, SIGNED>(m, wm, W, g, n / 2, wlo, whi, s_hist, s_sel, off);
    const double a = select_kth_abs<T, SIGNED>(m, wm, W, g, n / 2 - 1, wlo, whi, s_hist, s_sel, off);
    const double b = select_kth_abs<T, SIGNED>(m, wm, W, g, n / 2, wlo, whi, s_hist, s_sel, off);
    return (a + b) / 2.0;
}

template <typename T>
__global__ void __launch_bounds__(CR_THREADS) cell_rms_stream_kernel(const T* __restrict__ map, const T* __restrict__ wmed, int S, int H,
                                                              int W, const Chunks ch, int nx,
                                                              int ov, int numBins, double* __restrict__ table,
                                                              double* __restrict__ edges, int estimator) {
    __shared__ double s_red[2 * CR_MAXWARPS];
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_sel[2];
    extern __shared__ double s_edges[];  // numBins + 1
    const int cell = blockIdx.x, img = blockIdx.y;
    const int ci = cell / nx, ck = cell % nx;
    const CellGeom g = cell_geom(ch, ci, ck, ov, H, W);
    const T* m = map + (size_t)img * H * W;
    const T* wm = wmed + (size_t)(img / S) * H * W;
    const int cw = g.x1 - g.x0;
    const int n = (g.y1 - g.y0) * cw;
    double* tab = table + ((size_t)img * gridDim.x + cell) * numBins;
    double* edg = edges + ((size_t)img * gridDim.x + cell) * (numBins + 1);

    // good pixels: count, weight range
    double cnt = 0, dummy = 0, wmin = INFINITY, wmax = -INFINITY;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int r = e / cw, c = e - r * cw;
        const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
        if (m[off] != (T)0) {
            cnt += 1.0;
            const double w = (double)wm[off];
            wmin = fmin(wmin, w);
            wmax = fmax(wmax, w);
        }
    }
    block_sum2(cnt, dummy, s_red);
    block_minmax(wmin, wmax, s_red);
    if (cnt == 0) {  // filters.py:446-447
        for (int b = threadIdx.x; b < numBins; b += blockDim.x) tab[b] = 0.0;
        for (int b = threadIdx.x; b <= numBins; b += blockDim.x) edg[b] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const long long nGood = (long long)cnt;
    // bin edges: np.percentile(weights[good], p) for p = b * 100 / numBins, then max + 1e-6  (filters.py:450-454)
    if (threadIdx.x == 0) {
        s_edges[0] = wmin;
        s_edges[numBins] = wmax + 1e-6;
    }
    for (int b = 1; b < numBins; ++b) {
        if (wmin == wmax) {  // constant weights in the cell: every percentile is that constant
            if (threadIdx.x == 0) s_edges[b] = wmin;
            continue;
        }
        const double p = (double)b * (100.0 / (double)numBins);  // np.arange(0, 100, 100/numBins)[b]
        const double q = p / 100.0;
        const double nn = (double)nGood;
        const double vi = nn * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;  // numpy's virtual index, method 'linear'
        const double fl = floor(vi);
        long long lo = (long long)fl, hi = lo + 1;
        if (lo < 0) lo = 0;
        if (lo > nGood - 1) lo = nGood - 1;
        if (hi < 0) hi = 0;
        if (hi > nGood - 1) hi = nGood - 1;
        const double gmm = vi - fl;
        const double a = select_kth<T>(m, wm, W, g, lo, s_hist, s_sel);
        const double bb = (hi == lo) ? a : select_kth<T>(m, wm, W, g, hi, s_hist, s_sel);
        const double diff = bb - a;
        double res = a + diff * gmm;           // numpy _lerp
        if (gmm >= 0.5) res = bb - diff * (1.0 - gmm);
        if (gmm == 0.0) res = a;
        if (threadIdx.x == 0) s_edges[b] = res;
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= numBins; b += blockDim.x) edg[b] = s_edges[b];

    for (int b = 0; b < numBins; ++b) {
        const double lo = s_edges[b], hi = s_edges[b + 1];
        // mean and std of the bin's good pixels
        double c0 = 0, s0 = 0;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int r = e / cw, c = e - r * cw;
            const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
            const double v = (double)m[off];
            const double w = (double)wm[off];
            if (v != 0.0 && w >= lo && w < hi) {
                c0 += 1.0;
                s0 += v;
            }
        }
        block_sum2(c0, s0, s_red);
        double rms = 0.0;
        if (estimator == 1) {
            // np.percentile(abs(binValues), 68.3) (filters.py:464-465); an empty bin gives nan there, which paints nothing
            if (c0 > 0) {
                const double q = 68.3 / 100.0;
                const double vi = c0 * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;   // numpy's virtual index, method 'linear'
                const double fl = floor(vi);
                long long klo = (long long)fl, khi = klo + 1;
                const long long nn = (long long)c0;
                klo = klo < 0 ? 0 : (klo > nn - 1 ? nn - 1 : klo);
                khi = khi < 0 ? 0 : (khi > nn - 1 ? nn - 1 : khi);
                const double gmm = vi - fl;
                const double a = select_kth_abs<T>(m, wm, W, g, klo, lo, hi, s_hist, s_sel);
                const double bb = (khi == klo) ? a : select_kth_abs<T>(m, wm, W, g, khi, lo, hi, s_hist, s_sel);
                const double diff = bb - a;
                rms = a + diff * gmm;           // numpy _lerp
                if (gmm >= 0.5) rms = bb - diff * (1.0 - gmm);
                if (gmm == 0.0) rms = a;
            }
        } else if (estimator == 2) {
            // astropy.stats.biweight_scale(binValues, c = 9.0, modify_sample_size = True) for bins of at least 10 pixels, else 0
            // (filters.py:461-465): M = median, MAD = median |v - M|, u = (v - M) / (c MAD), over |u| < 1:
            // n' = count, f1 = sum (v - M)^2 (1 - u^2)^4, f2 = sum (1 - u^2) (1 - 5 u^2); scale = sqrt(n' f1) / |f2|; 0 when MAD = 0
            if (c0 >= 10) {
                const long long nn = (long long)c0;
                const double M = median_of_bin<T, true>(m, wm, W, g, nn, lo, hi, s_hist, s_sel, 0.0);
                const double mad = median_of_bin<T, false>(m, wm, W, g, nn, lo, hi, s_hist, s_sel, M);
                if (mad > 0.0) {
                    double cntU = 0, f1 = 0, f2 = 0, dmy = 0;
                    const double cm = 9.0 * mad;
                    for (int e = threadIdx.x; e < n; e += blockDim.x) {
                        const int r = e / cw, c = e - r * cw;
                        const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
                        const double v = (double)m[off];
                        const double w = (double)wm[off];
                        if (v != 0.0 && w >= lo && w < hi) {
                            const double d = v - M;
                            double u = d / cm;
                            if (fabs(u) < 1.0) {
                                u = u * u;
                                const double t = 1.0 - u;
                                cntU += 1.0;
                                f1 += d * d * (t * t) * (t * t);
                                f2 += t * (1.0 - 5.0 * u);
                            }
                        }
                    }
                    block_sum2(cntU, f1, s_red);
                    block_sum2(f2, dmy, s_red);
                    const double af2 = fabs(f2);
                    rms = sqrt(cntU * f1 / (af2 * af2));
                }
            }
        } else if (c0 > 0) {
            double mean = s0 / c0;
            double q0 = 0, d0 = 0;
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                const int r = e / cw, c = e - r * cw;
                const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
                const double v = (double)m[off];
                const double w = (double)wm[off];
                if (v != 0.0 && w >= lo && w < hi) {
                    const double d = v - mean;
                    q0 += d * d;
                }
            }
            block_sum2(q0, d0, s_red);
            rms = sqrt(q0 / c0);
            double prevC = -1.0, prevS = 0.0;
            for (int it = 0; it < 10; ++it) {  // filters.py:473-477
                const double thr = fabs(mean + 3.0 * rms);
                double c1 = 0, s1 = 0;
                for (int e = threadIdx.x; e < n; e += blockDim.x) {
                    const int r = e / cw, c = e - r * cw;
                    const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
                    const double v = (double)m[off];
                    const double w = (double)wm[off];
                    if (v != 0.0 && w >= lo && w < hi && fabs(v) < thr) {
                        c1 += 1.0;
                        s1 += v;
                    }
                }
                block_sum2(c1, s1, s_red);
                if (c1 == 0) continue;  // "if mask.sum() > 0": nothing changes, and nothing will in later passes
                mean = s1 / c1;
                double q1 = 0, d1 = 0;
                for (int e = threadIdx.x; e < n; e += blockDim.x) {
                    const int r = e / cw, c = e - r * cw;
                    const size_t off = (size_t)(g.y0 + r) * W + g.x0 + c;
                    const double v = (double)m[off];
                    const double w = (double)wm[off];
                    if (v != 0.0 && w >= lo && w < hi && fabs(v) < thr) {
                        const double d = v - mean;
                        q1 += d * d;
                    }
                }
                block_sum2(q1, d1, s_red);
                const double newRms = sqrt(q1 / c1);
                // fixed point: the same threshold selects the same samples from here on
                const bool fixed = (newRms == rms) && (c1 == prevC) && (s1 == prevS);
                rms = newRms;
                prevC = c1;
                prevS = s1;
                if (fixed) break;
            }
        }
        if (threadIdx.x == 0) tab[b] = rms;
    }
}

// ------------------------------------------------------------------------------------------------------
// K6, on-chip version.  One CTA per (cell, image) loads the cell ONCE; every later pass (order statistics of the
// weights, the two-pass mean / std and the ten clipping iterations of every weight bin) runs on data held in
// registers (OC_PPT slots per thread) and, for what does not fit there, in shared memory.  The slots first hold
// the order-preserving integer key of each good pixel's weight (for the percentile bin edges), then the pixel
// value.  Each warp owns a contiguous run of the cell's pixels (lane-interleaved, so loads coalesce): weights
// vary smoothly over the sky, so a warp usually meets only a few of the weight bins and skips the others.
// Cells larger than the on-chip capacity use cell_rms_stream_kernel above.
#ifndef OC_THREADS
#define OC_THREADS 512
#endif
#define OC_WARPS (OC_THREADS / 32)
#ifndef OC_PPT
#define OC_PPT 24         // register slots per thread (a multiple of OC_LB)
#endif
#ifndef OC_LB
#define OC_LB 8           // loads issued back to back per thread
#endif
#define OC_TGX 4          // warp tiles across the cell
#define OC_TGY (OC_WARPS / OC_TGX)
#define OC_GROUPS 16      // order statistics resolved together by the fallback radix select (one warp each)
#define OC_MAXBINS 64
#define OC_MAXTARGETS (2 * (OC_MAXBINS - 1))
#define OC_L1BITS 11
#define OC_L1BUCKETS (1 << OC_L1BITS)
#define OC_CANDCAP 2048   // candidate keys gathered from the buckets that hold a wanted rank
#define OC_SEGCAP 512     // largest bucket resolved by rank counting

template <typename T> struct KeyOf;
template <> struct KeyOf<double> {
    typedef unsigned long long type;
    static constexpr int BITS = 64;
    static __device__ __forceinline__ type enc(double v) { return enc_w(v); }
    static __device__ __forceinline__ double dec(type k) { return dec_w(k); }
    static __device__ __forceinline__ type fromValue(double v) { return (type)__double_as_longlong(v); }
    static __device__ __forceinline__ double toValue(type k) { return __longlong_as_double((long long)k); }
};
template <> struct KeyOf<float> {
    typedef unsigned int type;
    static constexpr int BITS = 32;
    static __device__ __forceinline__ type enc(float v) {
        const unsigned int b = __float_as_uint(v);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    static __device__ __forceinline__ double dec(type e) {
        const unsigned int b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
        return (double)__uint_as_float(b);
    }
    static __device__ __forceinline__ type fromValue(float v) { return __float_as_uint(v); }
    static __device__ __forceinline__ double toValue(type k) { return (double)__uint_as_float(k); }
};

struct OcBinState {   // per weight bin, shared memory
    double mean, rms, thr, pivot, prevC, prevS;
    int done;
    int pad;
};

#define OC_SPT 32          // most shared-memory slots per lane the all-shared variant keeps track of in registers

// shared-memory carve-up shared by host (size) and device (pointers).  Hybrid variant: the slots region holds the keys, then
// the values, of the pixels that do not fit the register slots.  All-shared variant (allShared): the slots region is sized for
// every pixel of the cell (valueSlots); while the weights are being ranked its first smemSlots entries hold keys and the
// selection scratch sits behind them, afterwards the whole region holds the values sorted by (warp, bin).
template <typename KT> struct OcLayout {
    size_t red, edges, state, warpBins, trank, tval, tuniq, usize, uoff, ucount, umin, umax, ubucket, hist, mark, cand, slots, wstart, wrun, bins, total;
    __host__ __device__ OcLayout(int numBins, int smemSlots, int valueSlots) {
        const bool allShared = valueSlots > 0;
        size_t o = 0;
        red = o;      o += sizeof(double) * 2 * OC_WARPS * numBins * 3;
        edges = o;    o += sizeof(double) * (numBins + 1);
        state = o;    o += sizeof(OcBinState) * OC_WARPS * numBins;
        warpBins = o; o += sizeof(unsigned long long) * OC_WARPS;
        wstart = o;   o += sizeof(int) * OC_WARPS * (numBins + 1);
        wrun = o;     o += sizeof(int) * OC_WARPS * (numBins + 1);
        o = (o + 15) & ~(size_t)15;
        slots = o;
        const size_t slotsEnd = o + sizeof(KT) * (size_t)(allShared ? valueSlots : smemSlots);
        o += sizeof(KT) * (size_t)smemSlots;           // the scratch follows the keys
        o = (o + 15) & ~(size_t)15;
        trank = o;    o += sizeof(long long) * OC_MAXTARGETS;
        tval = o;     o += sizeof(unsigned long long) * OC_MAXTARGETS;
        umin = o;     o += sizeof(unsigned long long) * OC_MAXTARGETS;
        umax = o;     o += sizeof(unsigned long long) * OC_MAXTARGETS;
        cand = o;     o += sizeof(KT) * OC_CANDCAP;
        tuniq = o;    o += sizeof(int) * OC_MAXTARGETS;
        usize = o;    o += sizeof(int) * OC_MAXTARGETS;
        uoff = o;     o += sizeof(int) * OC_MAXTARGETS;
        ucount = o;   o += sizeof(int) * OC_MAXTARGETS;
        ubucket = o;  o += sizeof(int) * OC_MAXTARGETS;
        hist = o;     o += sizeof(int) * (OC_L1BUCKETS > OC_GROUPS * 256 ? OC_L1BUCKETS : OC_GROUPS * 256);
        mark = o;     o += sizeof(short) * OC_L1BUCKETS;
        bins = o;     o += allShared ? 0 : (size_t)smemSlots;
        if (o < slotsEnd) o = slotsEnd;
        total = (o + 15) & ~(size_t)15;
    }
};

template <typename T, bool AS>
__global__ void __launch_bounds__(OC_THREADS, 1) cell_rms_onchip_kernel(const T* __restrict__ map, const T* __restrict__ wmed, int S,
                                                                        int H, int W, const Chunks ch, int nx, int ov, int numBins,
                                                                        double* __restrict__ table, double* __restrict__ edges,
                                                                        int smemSlots, int valueSlots) {
    typedef KeyOf<T> K;
    typedef typename K::type key_t;
    constexpr key_t SENT = ~(key_t)0;
    extern __shared__ __align__(16) unsigned char oc_smem[];
    const OcLayout<key_t> lay(numBins, smemSlots, AS ? valueSlots : 0);
    double* s_red = reinterpret_cast<double*>(oc_smem + lay.red);
    double* s_edges = reinterpret_cast<double*>(oc_smem + lay.edges);
    OcBinState* s_state = reinterpret_cast<OcBinState*>(oc_smem + lay.state);
    unsigned long long* s_warpBins = reinterpret_cast<unsigned long long*>(oc_smem + lay.warpBins);
    long long* s_trank = reinterpret_cast<long long*>(oc_smem + lay.trank);
    unsigned long long* s_tval = reinterpret_cast<unsigned long long*>(oc_smem + lay.tval);
    unsigned long long* s_umin = reinterpret_cast<unsigned long long*>(oc_smem + lay.umin);
    unsigned long long* s_umax = reinterpret_cast<unsigned long long*>(oc_smem + lay.umax);
    key_t* s_cand = reinterpret_cast<key_t*>(oc_smem + lay.cand);
    key_t* s_slots = reinterpret_cast<key_t*>(oc_smem + lay.slots);
    int* s_tuniq = reinterpret_cast<int*>(oc_smem + lay.tuniq);
    int* s_usize = reinterpret_cast<int*>(oc_smem + lay.usize);
    int* s_uoff = reinterpret_cast<int*>(oc_smem + lay.uoff);
    int* s_ucount = reinterpret_cast<int*>(oc_smem + lay.ucount);
    int* s_ubucket = reinterpret_cast<int*>(oc_smem + lay.ubucket);
    int* s_hist = reinterpret_cast<int*>(oc_smem + lay.hist);
    short* s_mark = reinterpret_cast<short*>(oc_smem + lay.mark);
    int* s_wstart = reinterpret_cast<int*>(oc_smem + lay.wstart);
    int* s_wrun = reinterpret_cast<int*>(oc_smem + lay.wrun);
    unsigned char* s_bins = oc_smem + lay.bins;
    __shared__ double s_scr[2 * CR_MAXWARPS];
    __shared__ int s_misc[4];   // [0] number of unique target buckets, [1] fallback flag, [2] groups (fallback)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cell = blockIdx.x, img = blockIdx.y;
    const CellGeom g = cell_geom(ch, cell / nx, cell % nx, ov, H, W);
    const T* m = map + (size_t)img * H * W;
    const T* wm = wmed + (size_t)(img / S) * H * W;
    const int cw = g.x1 - g.x0;
    (void)cw;
    // the cell is cut into 4 x 4 warp tiles of tr x tc pixels; slot i = j * 32 + lane of a warp is pixel (i / tc, i % tc)
    // of its tile: j < OC_PPT in registers, the others in the warp's shared-memory run
    const int Rn = g.y1 - g.y0;
    const int tr = (Rn + OC_TGY - 1) / OC_TGY, tc = (cw + OC_TGX - 1) / OC_TGX;
    const int JT = (tr * tc + 31) >> 5;                      // slots per lane
    const int spw = smemSlots / OC_WARPS;                    // shared-memory slots per warp
    const int ty0 = (wid / OC_TGX) * tr, tx0 = (wid % OC_TGX) * tc;
    key_t* wSlots = s_slots + (size_t)wid * spw;             // this warp's run (keys; values too in the hybrid variant)
    unsigned char* wBins = s_bins + (size_t)wid * spw;
    double* tab = table + ((size_t)img * gridDim.x + cell) * numBins;
    double* edg = edges + ((size_t)img * gridDim.x + cell) * (numBins + 1);
    // global offset of this lane's slot j (pixel i = j * 32 + lane of the warp's tile, row i / tc, column i % tc), -1 outside the
    // cell.  The division is a multiply-high by ceil(2^32 / tc), exact while i * tc < 2^32 (tiles hold < 2^18 pixels).
    const unsigned int tcMagic = tc > 1 ? 0xFFFFFFFFu / (unsigned int)tc + 1u : 0u;
    const int tileBase = (g.y0 + ty0) * W + g.x0 + tx0;
    const int rowsLeft = min(tr, Rn - ty0), colsLeft = cw - tx0;     // rows / columns of the tile that lie inside the cell
    auto offsetOf = [&](int j) -> int {
        const unsigned int i = (unsigned int)(j * 32 + lane);
        const int r = tc > 1 ? (int)__umulhi(i, tcMagic) : (int)i;
        const int c = (int)i - r * tc;
        return (r < rowsLeft && c < colsLeft) ? tileBase + r * W + c : -1;
    };

    key_t reg[OC_PPT];
    // ---- load: weight keys of the good pixels (loads issued eight slots at a time, none depends on another) ----
    double cnt = 0, dummy = 0, wmin = INFINITY, wmax = -INFINITY;
    auto makeKey = [&](int off, T mv, T wv) -> key_t {
        if (off < 0 || mv == (T)0) return SENT;
        cnt += 1.0;
        wmin = fmin(wmin, (double)wv);
        wmax = fmax(wmax, (double)wv);
        return K::enc(wv);
    };
#pragma unroll
    for (int j0 = 0; j0 < OC_PPT; j0 += OC_LB) {
        int off[OC_LB];
        T mv[OC_LB], wv[OC_LB];
#pragma unroll
        for (int u = 0; u < OC_LB; ++u) {
            off[u] = offsetOf(j0 + u);
            mv[u] = (off[u] >= 0) ? m[off[u]] : (T)0;
            wv[u] = (off[u] >= 0) ? wm[off[u]] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < OC_LB; ++u) reg[j0 + u] = makeKey(off[u], mv[u], wv[u]);
    }
    for (int j0 = OC_PPT; j0 < JT; j0 += OC_LB) {
        int off[OC_LB];
        T mv[OC_LB], wv[OC_LB];
#pragma unroll
        for (int u = 0; u < OC_LB; ++u) {
            off[u] = offsetOf(j0 + u);
            mv[u] = (off[u] >= 0) ? m[off[u]] : (T)0;
            wv[u] = (off[u] >= 0) ? wm[off[u]] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < OC_LB; ++u)
            if (j0 + u < JT) wSlots[(j0 + u - OC_PPT) * 32 + lane] = makeKey(off[u], mv[u], wv[u]);
    }
    block_sum2(cnt, dummy, s_scr);
    block_minmax(wmin, wmax, s_scr);
    if (cnt == 0) {  // filters.py:446-447
        for (int b = tid; b < numBins; b += OC_THREADS) tab[b] = 0.0;
        for (int b = tid; b <= numBins; b += OC_THREADS) edg[b] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const long long nGood = (long long)cnt;
#if defined(OC_EXP_STOP) && OC_EXP_STOP == 1
    if (nGood >= 0) { if (tid == 0) tab[0] = (double)nGood; return; }
#endif

    // ---- bin edges: np.percentile(weights[good], b * 100 / numBins), max + 1e-6  (filters.py:450-454) -----------
    if (tid == 0) {
        s_edges[0] = wmin;
        s_edges[numBins] = wmax + 1e-6;
    }
    if (numBins > 1 && wmin == wmax) {  // constant weights: every percentile is that constant
        for (int b = 1 + tid; b < numBins; b += OC_THREADS) s_edges[b] = wmin;
    } else if (numBins > 1) {
        const key_t kmin = K::enc((T)wmin), kmax = K::enc((T)wmax);
        const int nTargets = 2 * (numBins - 1);   // ranks lo_b, hi_b for b = 1..numBins-1, ascending
        // numpy's virtual index of percentile p (method 'linear'): ranks floor(vi) and floor(vi) + 1
        auto virtualIndex = [&](int b) -> double {
            const double p = (double)b * (100.0 / (double)numBins);  // np.arange(0, 100, 100/numBins)[b]
            const double q = p / 100.0;
            const double nn = (double)nGood;
            return nn * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
        };
        for (int t = tid; t < nTargets; t += OC_THREADS) {
            const double vi = virtualIndex(1 + t / 2);
            long long lo = (long long)floor(vi), hi = lo + 1;
            lo = max(0LL, min(lo, nGood - 1));
            hi = max(0LL, min(hi, nGood - 1));
            s_trank[t] = (t & 1) ? hi : lo;
        }
        // -- level 1: one histogram of the top OC_L1BITS bits of (key - kmin) ---------------------------------------
        int nbits = 0;
        {
            key_t x = kmax - kmin;
            while (x) {
                ++nbits;
                x >>= 1;
            }
        }
        const int sh1 = nbits > OC_L1BITS ? nbits - OC_L1BITS : 0;
        for (int i = tid; i < OC_L1BUCKETS; i += OC_THREADS) {
            s_hist[i] = 0;
            s_mark[i] = -1;
        }
        if (tid == 0) s_misc[1] = 0;
        __syncthreads();
        auto count1 = [&](key_t key) {
            if (key != SENT) atomicAdd(&s_hist[(int)((key - kmin) >> sh1)], 1);
        };
#pragma unroll
        for (int j = 0; j < OC_PPT; ++j) count1(reg[j]);
        for (int j = OC_PPT; j < JT; ++j) count1(wSlots[(j - OC_PPT) * 32 + lane]);
        __syncthreads();
        // exclusive scan of the 2048 buckets (4 per thread), in place
        {
            constexpr int PER = OC_L1BUCKETS / OC_THREADS;
            int loc[PER], tot = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                loc[i] = s_hist[tid * PER + i];
                tot += loc[i];
            }
            int inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            int* s_wsum = reinterpret_cast<int*>(s_scr);
            __syncthreads();
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            int base = 0;
            for (int w = 0; w < wid; ++w) base += s_wsum[w];
            int run = base + inc - tot;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                s_hist[tid * PER + i] = run;
                run += loc[i];
            }
        }
        __syncthreads();
        // bucket of every target (largest bucket whose exclusive count <= rank), then the distinct buckets
        for (int t = tid; t < nTargets; t += OC_THREADS) {
            const int rank = (int)s_trank[t];
            int lo = 0, hi = OC_L1BUCKETS - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_hist[mid] <= rank) lo = mid; else hi = mid - 1;
            }
            s_tuniq[t] = lo;                       // bucket for now
            s_trank[t] = rank - s_hist[lo];        // rank inside the bucket
        }
        __syncthreads();
        if (tid == 0) {
            int nu = 0, off = 0, bad = 0;
            for (int t = 0; t < nTargets; ++t) {
                const int bkt = s_tuniq[t];
                if (nu == 0 || s_ubucket[nu - 1] != bkt) {   // targets ascend, so equal buckets are adjacent
                    const int size = ((bkt + 1 < OC_L1BUCKETS) ? s_hist[bkt + 1] : (int)nGood) - s_hist[bkt];
                    s_ubucket[nu] = bkt;
                    s_usize[nu] = size;
                    s_ucount[nu] = 0;
                    s_umin[nu] = ~0ULL;
                    s_umax[nu] = 0ULL;
                    if (size <= OC_SEGCAP && off + size <= OC_CANDCAP) {
                        s_uoff[nu] = off;
                        off += size;
                    } else {
                        s_uoff[nu] = -1;           // too large to gather: resolved only if all its keys are equal
                    }
                    s_mark[bkt] = (short)nu;
                    ++nu;
                }
                s_tuniq[t] = nu - 1;
            }
            s_misc[0] = nu;
            (void)bad;
        }
        __syncthreads();
        // -- level 2: gather the keys of the wanted buckets (or their min / max when the bucket is too large) ------
        auto gather = [&](key_t key) {
            if (key == SENT) return;
            const int u = s_mark[(int)((key - kmin) >> sh1)];
            if (u < 0) return;
            const int off = s_uoff[u];
            if (off >= 0) {
                s_cand[off + atomicAdd(&s_ucount[u], 1)] = key;
            } else {
                atomicMin(&s_umin[u], (unsigned long long)key);
                atomicMax(&s_umax[u], (unsigned long long)key);
            }
        };
#pragma unroll
        for (int j = 0; j < OC_PPT; ++j) gather(reg[j]);
        for (int j = OC_PPT; j < JT; ++j) gather(wSlots[(j - OC_PPT) * 32 + lane]);
        __syncthreads();
        // -- one warp per target: rank counting inside its bucket ----------------------------------------------------
        for (int t = wid; t < nTargets; t += OC_WARPS) {
            const int u = s_tuniq[t];
            const int off = s_uoff[u], msz = s_usize[u];
            const int r = (int)s_trank[t];
            if (off < 0) {
                if (lane == 0) {
                    if (s_umin[u] == s_umax[u]) s_tval[t] = s_umin[u];
                    else s_misc[1] = 1;            // a large bucket of distinct keys: fall back to the radix select
                }
                continue;
            }
            for (int i = lane; i < msz; i += 32) {
                const key_t ki = s_cand[off + i];
                int c = 0;
                for (int j = 0; j < msz; ++j) {
                    const key_t kj = s_cand[off + j];
                    c += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
                }
                if (c == r) s_tval[t] = (unsigned long long)ki;
            }
        }
        __syncthreads();
        if (s_misc[1]) {
            // ---- fallback: MSB-first radix select, 8 bits per pass, OC_GROUPS targets at a time ------------------------
            int nb2 = 0;
            {
                key_t x = kmin ^ kmax;
                while (x) {
                    ++nb2;
                    x >>= 1;
                }
            }
            const int npass = (nb2 + 7) / 8;
            const key_t lowMask = (8 * npass >= K::BITS) ? SENT : (((key_t)1 << (8 * npass)) - 1);
            unsigned long long* s_tprefix = s_umin;   // reuse
            int* s_tgroup = s_ucount;
            for (int t0 = 0; t0 < nTargets; t0 += OC_GROUPS) {
                const int nt = min(OC_GROUPS, nTargets - t0);
                __syncthreads();
                if (tid < nt) {
                    const int t = t0 + tid;
                    const double vi = virtualIndex(1 + t / 2);
                    long long lo = (long long)floor(vi), hi = lo + 1;
                    lo = max(0LL, min(lo, nGood - 1));
                    hi = max(0LL, min(hi, nGood - 1));
                    s_tprefix[tid] = (unsigned long long)(kmin & ~lowMask);
                    s_trank[tid] = (t & 1) ? hi : lo;   // s_trank[0..nt) is scratch from here on
                }
                key_t maskbits = ~lowMask;
                for (int pass = 0; pass < npass; ++pass) {
                    const int shift = 8 * (npass - 1 - pass);
                    __syncthreads();
                    if (tid == 0) {
                        int ng = 0;
                        for (int t = 0; t < nt; ++t) {
                            if (t == 0 || s_tprefix[t] != s_tprefix[t - 1]) ++ng;
                            s_tgroup[t] = ng - 1;
                        }
                    }
                    for (int i = tid; i < OC_GROUPS * 256; i += OC_THREADS) s_hist[i] = 0;
                    __syncthreads();
                    auto count = [&](key_t key) {
                        if (key == SENT) return;
                        const unsigned long long pre = (unsigned long long)(key & maskbits);
                        int loT = 0, hiT = nt - 1;
                        while (loT < hiT) {
                            const int mid = (loT + hiT) >> 1;
                            if (s_tprefix[mid] < pre) loT = mid + 1; else hiT = mid;
                        }
                        if (s_tprefix[loT] == pre) atomicAdd(&s_hist[s_tgroup[loT] * 256 + (int)((key >> shift) & (key_t)255)], 1);
                    };
#pragma unroll
                    for (int j = 0; j < OC_PPT; ++j) count(reg[j]);
                    for (int j = OC_PPT; j < JT; ++j) count(wSlots[(j - OC_PPT) * 32 + lane]);
                    __syncthreads();
                    if (wid < nt) {
                        const int* hst = s_hist + s_tgroup[wid] * 256;
                        int loc[8], tot = 0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            loc[i] = hst[lane * 8 + i];
                            tot += loc[i];
                        }
                        int inc = tot;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += t;
                        }
                        const long long kk = s_trank[wid];
                        const long long before = (long long)(inc - tot);
                        if ((kk >= before) && (kk < before + tot)) {
                            long long r = kk - before;
                            int d = 0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (d == i && r >= loc[i]) {
                                    r -= loc[i];
                                    d = i + 1;
                                }
                            }
                            if (d > 7) d = 7;
                            s_tprefix[wid] |= ((unsigned long long)(lane * 8 + d)) << shift;
                            s_trank[wid] = r;
                        }
                    }
                    maskbits |= (key_t)255 << shift;
                }
                __syncthreads();
                if (tid < nt) s_tval[t0 + tid] = s_tprefix[tid];
            }
            __syncthreads();
        }
        // pairs (lo, hi) -> interpolated percentiles
        for (int b = 1 + tid; b < numBins; b += OC_THREADS) {
            const double a = K::dec((key_t)s_tval[2 * (b - 1)]);
            const double bb = K::dec((key_t)s_tval[2 * (b - 1) + 1]);
            const double vi = virtualIndex(b);
            const double gmm = vi - floor(vi);
            const double diff = bb - a;
            double res = a + diff * gmm;  // numpy _lerp
            if (gmm >= 0.5) res = bb - diff * (1.0 - gmm);
            if (gmm == 0.0) res = a;
            s_edges[b] = res;
        }
    }
    __syncthreads();
    for (int b = tid; b <= numBins; b += OC_THREADS) edg[b] = s_edges[b];
#if defined(OC_EXP_STOP) && OC_EXP_STOP == 2
    if (nGood >= 0) return;
#endif

    // ---- bin index of every good pixel, then the slots take the pixel values -----------------------------------
    // edges ascend and edges[0] <= w < edges[numBins] for every good pixel, so the bin with edges[b] <= w < edges[b + 1]
    // (filters.py:458) is found by bisection
    auto binOf = [&](key_t key) -> unsigned int {
        if (key == SENT) return 255u;
        const double w = K::dec(key);
        int lo = 0, hi = numBins;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (w >= s_edges[mid]) lo = mid; else hi = mid;
        }
        return (unsigned int)lo;
    };
    int* wStart = s_wstart + wid * (numBins + 1);
    int* wRun = s_wrun + wid * (numBins + 1);
    for (int b = lane; b <= numBins; b += 32) {
        wStart[b] = 0;
        wRun[b] = 0;
    }
    __syncwarp();
    unsigned long long present = 0ULL;
    unsigned int rbin[(OC_PPT + 3) / 4];
#pragma unroll
    for (int j4 = 0; j4 < (OC_PPT + 3) / 4; ++j4) rbin[j4] = 0xffffffffu;
    // per-warp, per-bin counting and placement: one slot of every lane at a time; lanes that share a bin take consecutive places
    auto countBin = [&](unsigned int bin) {
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 255u && lane == (__ffs(peers) - 1)) wStart[bin + 1] += __popc(peers);
        __syncwarp();
    };
    auto placeOf = [&](unsigned int bin) -> int {
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (bin != 255u && lane == leader) {
            base = wRun[bin];
            wRun[bin] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        __syncwarp();
        return base + __popc(peers & ((1u << lane) - 1u));
    };
    auto prefixBins = [&]() {
        if (lane == 0) {
            int run = 0;
            for (int b = 0; b < numBins; ++b) {
                const int c = wStart[b + 1];
                wStart[b] = run;
                wRun[b] = run;
                run += c;
            }
            wStart[numBins] = run;
        }
        __syncwarp();
    };
    if constexpr (AS) {
        // all-shared variant: bins of every slot into registers, counts, then (the keys and the selection scratch being dead
        // for the whole block) every value goes to its place in the warp's run of the slots region, sorted by bin
        unsigned int sbin[OC_SPT / 4];
#pragma unroll
        for (int j4 = 0; j4 < OC_SPT / 4; ++j4) sbin[j4] = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < OC_PPT; ++j) {
            const unsigned int bin = binOf(reg[j]);
            rbin[j >> 2] = (rbin[j >> 2] & ~(255u << (8 * (j & 3)))) | (bin << (8 * (j & 3)));
            countBin(bin);
        }
#pragma unroll
        for (int jj = 0; jj < OC_SPT; ++jj) {
            if (OC_PPT + jj < JT) {   // uniform over the warp
                const unsigned int bin = binOf(wSlots[jj * 32 + lane]);
                sbin[jj >> 2] = (sbin[jj >> 2] & ~(255u << (8 * (jj & 3)))) | (bin << (8 * (jj & 3)));
                countBin(bin);
            }
        }
        prefixBins();
        __syncthreads();
        key_t* vRun = s_slots + (size_t)wid * (valueSlots / OC_WARPS);
#pragma unroll
        for (int j0 = 0; j0 < OC_PPT; j0 += OC_LB) {
            T mv[OC_LB];
#pragma unroll
            for (int u = 0; u < OC_LB; ++u) {
                const int off = offsetOf(j0 + u);
                mv[u] = (off >= 0) ? m[off] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < OC_LB; ++u) {
                const int j = j0 + u;
                const unsigned int bin = (rbin[j >> 2] >> (8 * (j & 3))) & 255u;
                const int pos = placeOf(bin);
                if (bin != 255u) vRun[pos] = K::fromValue(mv[u]);
            }
        }
#pragma unroll
        for (int j0 = 0; j0 < OC_SPT; j0 += OC_LB) {
            if (OC_PPT + j0 < JT) {
                T mv[OC_LB];
#pragma unroll
                for (int u = 0; u < OC_LB; ++u) {
                    const int off = offsetOf(OC_PPT + j0 + u);
                    mv[u] = (off >= 0) ? m[off] : (T)0;
                }
#pragma unroll
                for (int u = 0; u < OC_LB; ++u) {
                    const int jj = j0 + u;
                    if (jj < OC_SPT && OC_PPT + jj < JT) {
                        const unsigned int bin = (sbin[jj >> 2] >> (8 * (jj & 3))) & 255u;
                        const int pos = placeOf(bin);
                        if (bin != 255u) vRun[pos] = K::fromValue(mv[u]);
                    }
                }
            }
        }
        wSlots = vRun;     // the sweeps below read the sorted run
        __syncwarp();
    } else {
        // hybrid variant.  Register slots: bins packed four per word, values loaded in place of the keys
#pragma unroll
        for (int j0 = 0; j0 < OC_PPT; j0 += OC_LB) {
            T mv[OC_LB];
#pragma unroll
            for (int u = 0; u < OC_LB; ++u) {
                const int off = offsetOf(j0 + u);
                mv[u] = (off >= 0) ? m[off] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < OC_LB; ++u) {
                const int j = j0 + u;
                const unsigned int bin = binOf(reg[j]);
                rbin[j >> 2] = (rbin[j >> 2] & ~(255u << (8 * (j & 3)))) | (bin << (8 * (j & 3)));
                if (bin != 255u) present |= 1ULL << bin;
                reg[j] = K::fromValue(bin != 255u ? mv[u] : (T)0);
            }
        }
        // shared-memory slots: count per bin, then rewrite the run sorted by bin so that a sweep of bin b reads only its own
        for (int j = OC_PPT; j < JT; ++j) {
            const unsigned int bin = binOf(wSlots[(j - OC_PPT) * 32 + lane]);
            wBins[(j - OC_PPT) * 32 + lane] = (unsigned char)bin;
            countBin(bin);
        }
        prefixBins();
        for (int j0 = OC_PPT; j0 < JT; j0 += OC_LB) {
            T mv[OC_LB];
#pragma unroll
            for (int u = 0; u < OC_LB; ++u) {
                const int off = offsetOf(j0 + u);
                mv[u] = (off >= 0) ? m[off] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < OC_LB; ++u) {
                if (j0 + u >= JT) break;   // uniform over the warp
                const unsigned int bin = wBins[(j0 + u - OC_PPT) * 32 + lane];
                const int pos = placeOf(bin);
                if (bin != 255u) wSlots[pos] = K::fromValue(mv[u]);
            }
        }
        __syncwarp();
    }
    for (int b = lane; b < numBins; b += 32)
        if (wStart[b + 1] > wStart[b]) present |= 1ULL << b;
    {
        unsigned int plo = (unsigned int)present, phi = (unsigned int)(present >> 32);
        plo = __reduce_or_sync(0xffffffffu, plo);
        phi = __reduce_or_sync(0xffffffffu, phi);
        present = ((unsigned long long)phi << 32) | plo;   // bins any lane of this warp holds
        if (lane == 0) s_warpBins[wid] = present;
    }
    // every warp keeps its own copy of the per-bin state and updates it with the same arithmetic, so one block barrier
    // per sweep (publishing the partial sums, double-buffered) is all the synchronisation the iterations need
    OcBinState* wState = s_state + (size_t)wid * numBins;
    for (int b = lane; b < numBins; b += 32) {
        OcBinState st;
        st.mean = 0.0; st.rms = 0.0; st.thr = INFINITY; st.pivot = 0.0; st.prevC = -1.0; st.prevS = 0.0; st.done = 0; st.pad = 0;
        wState[b] = st;
    }
    __syncthreads();   // s_warpBins, and the scratch the selection used, are settled

    int nbp = 1;                             // bins per round of the reduction below (a power of two), lanes per bin, lane's group
    while (nbp < numBins && nbp < 32) nbp <<= 1;
    const int lpb = 32 / nbp, sub = lane / nbp;
    // ---- sweeps: -2 mean, -1 std about the mean, 0..9 clipping iterations (filters.py:466-477) -----------------
#ifdef OC_EXP_NOSWEEP
    for (int it = -2; it < -2; ++it) {
#else
    for (int it = -2; it < 10; ++it) {
#endif
        double* red = s_red + (size_t)(it & 1) * OC_WARPS * numBins * 3;
        unsigned long long todo = present;
        while (todo) {
            const int b = __ffsll((long long)todo) - 1;
            todo &= todo - 1;
            if (wState[b].done) continue;
            const double thr = wState[b].thr, piv = wState[b].pivot;
            double c0 = 0, s0 = 0, q0 = 0, c1 = 0, s1 = 0, q1 = 0;
            if constexpr (!AS) {
#pragma unroll
                for (int j = 0; j < OC_PPT; ++j) {
                    const unsigned int bj = (rbin[j >> 2] >> (8 * (j & 3))) & 255u;
                    const double v = K::toValue(reg[j]);
                    if (bj == (unsigned int)b && fabs(v) < thr) {
                        const double d = v - piv;
                        if (j & 1) {
                            c1 += 1.0; s1 += v; q1 += d * d;
                        } else {
                            c0 += 1.0; s0 += v; q0 += d * d;
                        }
                    }
                }
            }
            {
                const int iEnd = wStart[b + 1];
                int i = wStart[b] + lane;
                for (; i + 32 < iEnd; i += 64) {   // two independent chains
                    const double v = K::toValue(wSlots[i]), v2 = K::toValue(wSlots[i + 32]);
                    if (fabs(v) < thr) {
                        const double d = v - piv;
                        c0 += 1.0; s0 += v; q0 += d * d;
                    }
                    if (fabs(v2) < thr) {
                        const double d = v2 - piv;
                        c1 += 1.0; s1 += v2; q1 += d * d;
                    }
                }
                if (i < iEnd) {
                    const double v = K::toValue(wSlots[i]);
                    if (fabs(v) < thr) {
                        const double d = v - piv;
                        c0 += 1.0; s0 += v; q0 += d * d;
                    }
                }
            }
            c0 = warp_sum(c0 + c1);
            s0 = warp_sum(s0 + s1);
            q0 = warp_sum(q0 + q1);
            if (lane == 0) {
                double* r = red + ((size_t)wid * numBins + b) * 3;
                r[0] = c0; r[1] = s0; r[2] = q0;
            }
        }
        __syncthreads();
        // per-bin totals: the lanes of a warp share the bins and the warps' partial sums between them (nbp bins x lpb groups of
        // warps), then combine by butterfly -- the same additions in every warp, so every copy of the state stays bit-identical
        bool myDone = true;
        for (int b0 = 0; b0 < numBins; b0 += 32) {
            const int b = b0 + (lane & (nbp - 1));
            const bool live = b < numBins && !wState[b].done;
            double c = 0, s = 0, q = 0;
            if (live) {
                for (int w = sub; w < OC_WARPS; w += lpb) {
                    if ((s_warpBins[w] >> b) & 1ULL) {
                        const double* r = red + ((size_t)w * numBins + b) * 3;
                        c += r[0]; s += r[1]; q += r[2];
                    }
                }
            }
            for (int o = nbp; o < 32; o <<= 1) {
                c += __shfl_xor_sync(0xffffffffu, c, o);
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (live && sub == 0) {
                OcBinState st = wState[b];
                if (it == -2) {
                    if (c == 0) {
                        st.rms = 0.0;
                        st.done = 1;
                    } else {
                        st.mean = s / c;
                        st.pivot = st.mean;
                    }
                } else if (it == -1) {
                    st.rms = sqrt(q / c);
                    st.thr = fabs(st.mean + 3.0 * st.rms);
                } else if (c == 0) {
                    st.done = 1;   // "if mask.sum() > 0": nothing changes now, and nothing will in later passes
                } else {
                    const double mean = s / c;
                    const double dm = mean - st.pivot;
                    double var = q / c - dm * dm;
                    if (var < 0) var = 0;
                    // same sample as in the previous pass: a fixed point, the remaining passes change nothing
                    const bool fixed = (c == st.prevC) && (s == st.prevS);
                    st.mean = mean;
                    st.rms = sqrt(var);
                    st.pivot = mean;
                    st.thr = fabs(mean + 3.0 * st.rms);
                    st.prevC = c;
                    st.prevS = s;
                    if (fixed) st.done = 1;
                }
                wState[b] = st;
                myDone = myDone && (st.done != 0);
            }
        }
        __syncwarp();
        if (__all_sync(0xffffffffu, myDone)) break;   // the same decision in every warp
    }
    if (wid == 0)
        for (int b = lane; b < numBins; b += 32) tab[b] = wState[b].rms;
}

// the RMS a pixel takes: that of the last cell (raster order) that covers it, has rms > 0 and whose weight bin holds its weight w
// (filters.py:480-481: later cells overwrite earlier ones where they overlap); 0 when there is none
// candidate cells along one axis for coordinate v: the chunk that holds it and every neighbour whose margin reaches it
__device__ __forceinline__ void paint_axis_range(int v, const int* __restrict__ c, int n, int ov, int L, int& lo, int& hi) {
    int ci = min((int)(((long long)v * n) / L), n - 1);
    while (ci > 0 && v < c[ci]) --ci;
    while (ci < n - 1 && v >= c[ci + 1]) ++ci;
    lo = hi = ci;
    while (lo > 0 && v < min(c[lo] + ov, L)) --lo;
    while (hi < n - 1 && v >= max(c[hi + 1] - ov, 0)) ++hi;
}

// the value cell (i, k) of image img paints at weight w, or 0 when it paints nothing there.  The weight bins of a cell are
// disjoint and ascending, so the reference's "last bin that holds w" (filters.py:455-481) is the only bin that holds it.
__device__ __forceinline__ double paint_cell_value(double w, int cell, int img, int nCells, int numBins,
                                                   const double* __restrict__ table, const double* __restrict__ edges) {
    const double* edg = edges + ((size_t)img * nCells + cell) * (numBins + 1);
    const double e0 = edg[0];
    if (isnan(e0) || !(w >= e0) || !(w < edg[numBins])) return 0.0;
    int lo = 0, hi = numBins;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (w >= edg[mid]) lo = mid; else hi = mid;
    }
    const double v = table[((size_t)img * nCells + cell) * numBins + lo];
    return v > 0.0 ? v : 0.0;
}

__device__ __forceinline__ double paint_value_in(double w, int x, int y, int img, const Chunks& ch, int nx, int ov, int numBins,
                                                 const double* __restrict__ table, const double* __restrict__ edges, int H, int W,
                                                 int nCells, int iLo, int iHi, int kLo, int kHi) {
    // candidates are visited last-written first (raster order of the cells)
    for (int i = iHi; i >= iLo; --i) {
        if (y < max(ch.y[i] - ov, 0) || y >= min(ch.y[i + 1] + ov, H)) continue;
        for (int k = kHi; k >= kLo; --k) {
            if (x < max(ch.x[k] - ov, 0) || x >= min(ch.x[k + 1] + ov, W)) continue;
            const double v = paint_cell_value(w, i * nx + k, img, nCells, numBins, table, edges);
            if (v > 0.0) return v;
        }
    }
    return 0.0;
}

__device__ __forceinline__ double paint_value(double w, int x, int y, int img, const Chunks& ch, int ny, int nx, int ov, int numBins,
                                              const double* __restrict__ table, const double* __restrict__ edges, int H, int W) {
    int iLo, iHi, kLo, kHi;
    paint_axis_range(y, ch.y, ny, ov, H, iLo, iHi);
    paint_axis_range(x, ch.x, nx, ov, W, kLo, kHi);
    return paint_value_in(w, x, y, img, ch, nx, ov, numBins, table, edges, H, W, ny * nx, iLo, iHi, kLo, kHi);
}

// every pixel takes the RMS of the last cell (raster order) that covers it, has rms > 0 and whose weight bin holds it
template <typename T>
__global__ void __launch_bounds__(256) paint_rms_kernel(const T* __restrict__ wmed, int S, int H, int W, const Chunks ch,
                                                        int ny, int nx, int ov, int numBins,
                                                        const double* __restrict__ table, const double* __restrict__ edges,
                                                        T* __restrict__ rms) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x >= W) return;
    const double w = (double)wmed[((size_t)(img / S) * H + y) * W + x];
    const double val = paint_value(w, x, y, img, ch, ny, nx, ov, numBins, table, edges, H, W);
    rms[((size_t)img * H + y) * W + x] = (T)val;
}

// ------------------------------------------------------------------------------------------------------
#define ER_ROWS 8
// K7.  rank_filter(a, 0, size=(T,T)) > 0  <=>  no zero of a inside the T x T window, whose offsets along each axis are
// [-T/2, T-1-T/2] with scipy's reflect boundary.  Done separably: row pass with a prefix sum, then a sliding column sum
// fused with the final elementwise step.
template <typename T>
__global__ void __launch_bounds__(256) edge_rows_kernel(const T* __restrict__ map, const unsigned char* __restrict__ ps, int S, int H,
                                                        int W, int trim, unsigned char* __restrict__ rowAny,
                                                        const int* __restrict__ zeroFlags) {
    extern __shared__ int s_pre[];  // W + 1 prefix counts of zeros
    __shared__ int s_part[256];
    const int img = blockIdx.y;
    if (zeroFlags && zeroFlags[img] == 0) return;      // the kernel that wrote this image saw no zero: the column pass will not read rowAny
    // (ER_ROWS rows per CTA: a hundred thousand one-row CTAs that only return still cost 54 us per launch)
    for (int y = blockIdx.x * ER_ROWS; y < min((int)(blockIdx.x + 1) * ER_ROWS, H); ++y) {
    __syncthreads();      // the previous row's prefix counts have been read
    const T* m = map + ((size_t)img * H + y) * W;
    const unsigned char* p = ps + ((size_t)(img / S) * H + y) * W;
    const int per = (W + blockDim.x - 1) / blockDim.x;
    const int b0 = threadIdx.x * per, b1 = min(b0 + per, W);
    int local = 0;
    for (int x = b0; x < b1; ++x) {
        const T a = m[x] + (T)(1 - (int)p[x]);
        const int z = (a == (T)0) ? 1 : 0;
        local += z;
        s_pre[x + 1] = local;  // local inclusive, offset added below
    }
    // exclusive scan of the per-thread counts over the block: shuffles inside a warp, the warp totals through shared memory (a
    // single thread walking all 256 partial counts cost more than the loads of the row)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_part[wid] = inc;
    if (threadIdx.x == 0) s_pre[0] = 0;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid; ++w) base += s_part[w];
    const int offv = base + inc - local;
    for (int x = b0; x < b1; ++x) s_pre[x + 1] += offv;
    __syncthreads();
    const int lo = -(trim / 2), hi = trim - 1 - trim / 2;
    unsigned char* out = rowAny + ((size_t)img * H + y) * W;
    // zeros among the virtual positions [0, n) of the reflected row (d c b a | a b c d | d c b a ...), period 2 W; a window
    // wider than the row folds more than once, as scipy's 'reflect' does
    auto ext = [&](int n) -> int {
        const int full = n / (2 * W), r = n - full * 2 * W;
        const int part = (r <= W) ? s_pre[r] : s_pre[W] + (s_pre[W] - s_pre[2 * W - r]);
        return full * 2 * s_pre[W] + part;
    };
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const int a = x + lo, b = x + hi;
        int cnt;
        if (a >= 0 && b < W) cnt = s_pre[b + 1] - s_pre[a];       // window inside the row: no reflection, no divisions
        else if (a >= 0) cnt = ext(b + 1) - ext(a);
        else if (b < 0) cnt = ext(-a) - ext(-b - 1);     // position -k mirrors position k - 1
        else cnt = ext(-a) + ext(b + 1);
        out[x] = cnt > 0 ? 1 : 0;
    }
    }
}

// A thread walks one column for `seg` rows.  With an edge trim its window count starts from a walk over the `trim` rows around
// the segment's first row (240 rows for the reference's default of three noise-grid steps), so short segments spend more
// loads on that than on their pixels: 64-row segments 1.81 ms per 128 images of 840 x 1440, 420-row segments 1.32 ms, and 1.27 ms
// with two rows in flight instead of four (whole columns, 1.44 ms, leave too few CTAs).  ec_seg_rows picks about 420.
#define EC_SEG_MAX 1024
#ifndef EC_UNR
#define EC_UNR 2      // rows whose loads are issued together
#endif
static int ec_seg_rows(int H) {
    int nSeg = (H + 210) / 420;
    if (nSeg < 1) nSeg = 1;
    int seg = (H + nSeg - 1) / nSeg;
    return seg > EC_SEG_MAX ? EC_SEG_MAX : seg;
}
struct PaintArgs {   // what edge_cols_final_kernel<T, true> needs to take the RMS straight from the cell table
    const void* wmed;
    const double* table;
    const double* edges;
    int ny, nx, ov, numBins;
};

template <typename T, bool PAINT>
__global__ void __launch_bounds__(128) edge_cols_final_kernel(T* map, T* rms, const unsigned char* __restrict__ ps,
                                                              const unsigned char* __restrict__ survey, int S, int H, int W, int trim,
                                                              int apodPix, int realSpace, const unsigned char* __restrict__ rowAny,
                                                              T* sn, unsigned char* __restrict__ maskOut, PaintArgs pa, const Chunks ch,
                                                              const int seg, const int* __restrict__ zeroFlags) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int ys = blockIdx.y * seg, img = blockIdx.z;
    // an image whose producer reported that map + (1 - psMask) is nowhere zero has edgeCheck = 1 everywhere: no window to count
    if (zeroFlags && zeroFlags[img] == 0) trim = 0;
    // candidate cell rows of every row of this segment: the same for all columns, so found once per block (the search divides)
    __shared__ short s_iLo[EC_SEG_MAX], s_iHi[EC_SEG_MAX];
    if constexpr (PAINT) {
        for (int r = threadIdx.x; r < seg; r += blockDim.x) {
            int a = 0, b = 0;
            if (ys + r < H) paint_axis_range(ys + r, ch.y, pa.ny, pa.ov, H, a, b);
            s_iLo[r] = (short)a;
            s_iHi[r] = (short)b;
        }
        __syncthreads();
    }
    if (x >= W) return;
    const size_t ib = (size_t)img * H * W, tb = (size_t)(img / S) * H * W;
    const unsigned char* ra = rowAny + ib;
    const int lo = -(trim / 2), hi = trim - 1 - trim / 2;
    int cnt = 0;
    if (trim > 0) {
        for (int d = lo; d <= hi; ++d) cnt += ra[(size_t)reflect_idx(ys + d, H) * W + x];
    }
    const int ye = min(ys + seg, H);
    int kLo = 0, kHi = 0;
    if constexpr (PAINT) paint_axis_range(x, ch.x, pa.nx, pa.ov, W, kLo, kHi);   // the same for every row of this column
    // the last-written candidate cell (iHi, kHi) always covers the pixel and nearly always paints it; while the rows of this
    // column stay in the same cell and the weight stays in the same bin of it, the value is the one found for the row above
    int cacheI = -1;
    double cacheLo = 0.0, cacheHi = 0.0, cacheVal = 0.0;
    // rows are taken EC_UNR at a time with all their loads issued before any store: map and rms are read and written in place,
    // so the compiler must keep every load behind the previous row's stores and a row-by-row loop pays a full memory latency
    // per row
    for (int yb = ys; yb < ye; yb += EC_UNR) {
        T m0v[EC_UNR], wv[EC_UNR], rv[EC_UNR];
        unsigned char psv8[EC_UNR], sv8[EC_UNR], raIn[EC_UNR], raOut[EC_UNR];
#pragma unroll
        for (int u = 0; u < EC_UNR; ++u) {
            const int y = yb + u;
            const bool ok = y < ye;
            const size_t o = (size_t)(ok ? y : ys) * W + x;
            m0v[u] = map[ib + o];
            if constexpr (PAINT) {
                wv[u] = reinterpret_cast<const T*>(pa.wmed)[tb + o];
                rv[u] = (T)0;
            } else {
                wv[u] = (T)0;
                rv[u] = rms[ib + o];
            }
            psv8[u] = ps[tb + o];
            sv8[u] = survey[tb + o];
            raIn[u] = 0;
            raOut[u] = 0;
            if (trim > 0 && ok && y + 1 < ye) {   // the rows that enter / leave the window when it slides down from row y
                raIn[u] = ra[(size_t)reflect_idx(y + 1 + hi, H) * W + x];
                raOut[u] = ra[(size_t)reflect_idx(y + lo, H) * W + x];
            }
        }
#pragma unroll
        for (int u = 0; u < EC_UNR; ++u) {
            const int y = yb + u;
            if (y >= ye) break;
            const size_t o = (size_t)y * W + x;
            const T m0 = m0v[u];
            T r0;
            if constexpr (PAINT) {
                const double w = (double)wv[u];
                const int iLo = s_iLo[y - ys], iHi = s_iHi[y - ys];
                if (iHi == cacheI && w >= cacheLo && w < cacheHi) {
                    r0 = (T)cacheVal;
                } else {
                    const int nCells = pa.ny * pa.nx, cell = iHi * pa.nx + kHi;
                    const double* edg = pa.edges + ((size_t)img * nCells + cell) * (pa.numBins + 1);
                    const double e0 = edg[0];
                    double val = 0.0;
                    cacheI = -1;
                    if (!isnan(e0) && w >= e0 && w < edg[pa.numBins]) {
                        int lob = 0, hib = pa.numBins;
                        while (hib - lob > 1) {
                            const int mid = (lob + hib) >> 1;
                            if (w >= edg[mid]) lob = mid; else hib = mid;
                        }
                        val = pa.table[((size_t)img * nCells + cell) * pa.numBins + lob];
                        if (val > 0.0) {
                            cacheI = iHi;
                            cacheLo = edg[lob];
                            cacheHi = edg[lob + 1];
                            cacheVal = val;
                        }
                    }
                    if (!(val > 0.0))   // the first candidate paints nothing here: the general search over all candidates
                        val = paint_value_in(w, x, y, img, ch, pa.nx, pa.ov, pa.numBins, pa.table, pa.edges, H, W, nCells, iLo, iHi, kLo, kHi);
                    r0 = (T)val;
                }
            } else {
                r0 = rv[u];
            }
            const T edge = (trim > 0 && cnt > 0) ? (T)0 : (T)1;
            const T psv = (T)psv8[u], sv = (T)sv8[u];
            T snv = (r0 > (T)0) ? m0 / r0 : m0;
            T d = m0 * edge;
            const T sm1 = edge * sv * psv;
            if (!realSpace) d = d * sm1;
            const bool apodOk = (apodPix <= 0) || (x >= apodPix - 1 && x <= W - apodPix && y >= apodPix - 1 && y <= H - apodPix);
            const T mo = apodOk ? sm1 : (T)0;
            snv = snv * mo;
            if (snv != snv) snv = (T)0;
            map[ib + o] = d;
            sn[ib + o] = snv;
            if (!PAINT || rms) rms[ib + o] = r0 * mo;      // (painting variant: the RMS map itself is optional)
            maskOut[ib + o] = (mo != (T)0) ? 1 : 0;
            cnt += (int)raIn[u] - (int)raOut[u];      // slide the window down by one row
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) snr_from_rms_kernel(const T* __restrict__ map, const T* __restrict__ rms, T* __restrict__ sn,
                                                           long long HW, int S) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const size_t o = (size_t)blockIdx.y * HW + p;
    const T r = rms[(size_t)(blockIdx.y / S) * HW + p];
    const T m = map[o];
    sn[o] = (r > (T)0) ? m / r : m;
}

// ------------------------------------------------------------------------------------------------------
template <typename T>
static int cell_rms_t(const void* map, const void* wmed, int nImg, int S, int H, int W, const int* yChunks, int ny,
                      const int* xChunks, int nx, int ov, int numBins, void* rms, double* table, double* edges, cudaStream_t st,
                      int estimator = 0) {
    NB_REQUIRE(ny <= NB_MAX_CHUNKS && nx <= NB_MAX_CHUNKS, NB_ERR_UNSUPPORTED, "nb_cell_rms: more than %d noise cells per axis", NB_MAX_CHUNKS);
    Chunks ch;
    memset(&ch, 0, sizeof(ch));
    for (int i = 0; i <= ny; ++i) ch.y[i] = yChunks[i];
    for (int i = 0; i <= nx; ++i) ch.x[i] = xChunks[i];
    static const bool noOneBin = getenv("NEMO_B200_NO_ONEBIN") != nullptr;      // the on-chip kernel for one bin as well (comparison runs)
    if (numBins == 1 && estimator == 0 && !noOneBin && nb_onebin_supported(H, W, yChunks, ny, xChunks, nx, ov)) {
        const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
        int rc = nb_cell_rms_onebin(prec, map, wmed, nImg, S, H, W, ch, ny, nx, ov, table, edges, st);
        if (rc != NB_OK) return rc;
        if (rms) {
            { NB_PROF_BEGIN(st); paint_rms_kernel<T><<<dim3((W + 255) / 256, H, nImg), 256, 0, st>>>((const T*)wmed, S, H, W, ch, ny, nx, ov, numBins, table,
                                                                                 edges, (T*)rms); NB_PROF_END("paint_rms_kernel", st); }
            NB_CHECK_LAUNCH();
        }
        return NB_OK;
    }
    // largest cell (with its overlap margins) decides between the on-chip kernel and the streaming one
    int maxH = 0, maxW = 0;
    for (int i = 0; i < ny; ++i) maxH = std::max(maxH, std::min(ch.y[i + 1] + ov, H) - std::max(ch.y[i] - ov, 0));
    for (int k = 0; k < nx; ++k) maxW = std::max(maxW, std::min(ch.x[k + 1] + ov, W) - std::max(ch.x[k] - ov, 0));
    const long long maxCell = (long long)maxH * maxW;
    // 4 x 4 warp tiles; OC_PPT * 32 pixels of a tile live in registers, the rest in shared memory
    const long long tilePix = (long long)((maxH + OC_TGY - 1) / OC_TGY) * ((maxW + OC_TGX - 1) / OC_TGX);
    const long long CH = ((tilePix + 31) / 32) * 32;
    const long long perWarp = CH > (long long)OC_PPT * 32 ? CH - (long long)OC_PPT * 32 : 0;
    const long long smemSlots = perWarp * OC_WARPS;
    const bool fits = maxCell < (1LL << 22) && numBins <= OC_MAXBINS && estimator == 0;    // other estimators: streaming kernel
    typedef typename KeyOf<T>::type KT;
    // preferred: every value in shared memory, sorted by (warp, bin); else registers + shared memory; else streaming
    const long long valueSlots = CH * OC_WARPS;
    const size_t asBytes = (fits && perWarp / 32 <= OC_SPT) ? OcLayout<KT>(numBins, (int)smemSlots, (int)valueSlots).total : (size_t)1 << 30;
    const size_t ocBytes = fits ? OcLayout<KT>(numBins, (int)smemSlots, 0).total : (size_t)1 << 30;
    if (getenv("NB_DEBUG")) fprintf(stderr, "nb_cell_rms: maxCell %lld CH %lld perWarp %lld asBytes %zu ocBytes %zu fits %d\n", maxCell, CH, perWarp, asBytes, ocBytes, (int)fits);
    if (asBytes <= 227 * 1024) {
        NB_CHECK_CUDA(cudaFuncSetAttribute(cell_rms_onchip_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asBytes));
        { NB_PROF_BEGIN(st); cell_rms_onchip_kernel<T, true><<<dim3(ny * nx, nImg), OC_THREADS, asBytes, st>>>(
            (const T*)map, (const T*)wmed, S, H, W, ch, nx, ov, numBins, table, edges, (int)smemSlots, (int)valueSlots); NB_PROF_END("cell_rms_onchip_kernel", st); }
    } else if (ocBytes <= 227 * 1024) {
        NB_CHECK_CUDA(cudaFuncSetAttribute(cell_rms_onchip_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ocBytes));
        { NB_PROF_BEGIN(st); cell_rms_onchip_kernel<T, false><<<dim3(ny * nx, nImg), OC_THREADS, ocBytes, st>>>(
            (const T*)map, (const T*)wmed, S, H, W, ch, nx, ov, numBins, table, edges, (int)smemSlots, 0); NB_PROF_END("cell_rms_hybrid_kernel", st); }
    } else {
        { NB_PROF_BEGIN(st); cell_rms_stream_kernel<T><<<dim3(ny * nx, nImg), CR_THREADS, sizeof(double) * (numBins + 1), st>>>(
            (const T*)map, (const T*)wmed, S, H, W, ch, nx, ov, numBins, table, edges, estimator); NB_PROF_END("cell_rms_stream_kernel", st); }
    }
    NB_CHECK_LAUNCH();
    if (rms) {
        { NB_PROF_BEGIN(st); paint_rms_kernel<T><<<dim3((W + 255) / 256, H, nImg), 256, 0, st>>>((const T*)wmed, S, H, W, ch, ny, nx, ov, numBins, table,
                                                                             edges, (T*)rms); NB_PROF_END("paint_rms_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    return NB_OK;
}

template <typename T>
static int snr_mask_t(void* map, void* rms, const unsigned char* ps, const unsigned char* survey, int nImg, int S, int H, int W,
                      int trim, int apodPix, int realSpace, void* sn, unsigned char* maskOut, unsigned char* scratch, cudaStream_t st,
                      const int* zeroFlags, const PaintArgs* pa = nullptr, const Chunks* chp = nullptr) {
    if (trim > 0) {
        { NB_PROF_BEGIN(st); edge_rows_kernel<T><<<dim3((H + ER_ROWS - 1) / ER_ROWS, nImg), 256, sizeof(int) * (W + 1), st>>>((const T*)map, ps, S, H, W, trim, scratch, zeroFlags); NB_PROF_END("edge_rows_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    const int seg = ec_seg_rows(H);
    if (pa) {
        { NB_PROF_BEGIN(st); edge_cols_final_kernel<T, true><<<dim3((W + 127) / 128, (H + seg - 1) / seg, nImg), 128, 0, st>>>(
            (T*)map, (T*)rms, ps, survey, S, H, W, trim, apodPix, realSpace, scratch, (T*)sn, maskOut, *pa, *chp, seg, zeroFlags); NB_PROF_END("edge_cols_final_paint_kernel", st); }
    } else {
        PaintArgs none = {nullptr, nullptr, nullptr, 0, 0, 0, 0};
        Chunks ch0;
        memset(&ch0, 0, sizeof(ch0));
        { NB_PROF_BEGIN(st); edge_cols_final_kernel<T, false><<<dim3((W + 127) / 128, (H + seg - 1) / seg, nImg), 128, 0, st>>>(
            (T*)map, (T*)rms, ps, survey, S, H, W, trim, apodPix, realSpace, scratch, (T*)sn, maskOut, none, ch0, seg, zeroFlags); NB_PROF_END("edge_cols_final_kernel", st); }
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

__global__ void __launch_bounds__(256) widen_f32_kernel(const float4* __restrict__ in, double4* __restrict__ out, long long n4,
                                                        const float* __restrict__ tailIn, double* __restrict__ tailOut, int nTail) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 v = in[i];
        double4 o;
        o.x = (double)v.x; o.y = (double)v.y; o.z = (double)v.z; o.w = (double)v.w;
        out[i] = o;
    }
    if (i < nTail) tailOut[i] = (double)tailIn[i];
}

extern "C" {

int nb_widen_f32(const float* in, double* out, long long n, cudaStream_t st) {
    NB_REQUIRE(in && out && n > 0, NB_ERR_ARG, "nb_widen_f32: bad arguments");
    NB_REQUIRE(((size_t)in % 16) == 0 && ((size_t)out % 16) == 0, NB_ERR_ARG, "nb_widen_f32: buffers must be 16-byte aligned");
    const long long n4 = n / 4;
    const int nTail = (int)(n - 4 * n4);
    const long long threads = n4 > nTail ? n4 : nTail;
    { NB_PROF_BEGIN(st); widen_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const float4*)in, (double4*)out, n4, in + 4 * n4,
                                                                                           out + 4 * n4, nTail); NB_PROF_END("widen_f32_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_median_weights(int prec, const void* weights, void* out, int B, int F, int H, int W, cudaStream_t st) {
    NB_REQUIRE(weights && out && B > 0 && F >= 1 && F <= NB_MAX_FREQ, NB_ERR_ARG, "nb_median_weights: bad arguments");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); median_weights_kernel<double><<<grid, 256, 0, st>>>((const double*)weights, (double*)out, F, HW); NB_PROF_END("median_weights_kernel", st); }
    else if (prec == NB_PREC_F32)
        { NB_PROF_BEGIN(st); median_weights_kernel<float><<<grid, 256, 0, st>>>((const float*)weights, (float*)out, F, HW); NB_PROF_END("median_weights_kernel", st); }
    else {
        nb_set_error("nb_median_weights: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_cell_rms(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, const int* yChunks, int ny,
                const int* xChunks, int nx, int overlapPix, int numBins, void* rms, double* table, double* edges,
                cudaStream_t st) {
    NB_REQUIRE(map && wmed && yChunks && xChunks && table && edges && nImg > 0 && S > 0 && ny > 0 && nx > 0, NB_ERR_ARG,
               "nb_cell_rms: bad arguments");
    NB_REQUIRE(numBins >= 1 && numBins <= 64, NB_ERR_ARG, "nb_cell_rms: numBins %d out of range (1..64)", numBins);
    if (prec == NB_PREC_F64)
        return cell_rms_t<double>(map, wmed, nImg, S, H, W, yChunks, ny, xChunks, nx, overlapPix, numBins, rms, table, edges, st);
    if (prec == NB_PREC_F32)
        return cell_rms_t<float>(map, wmed, nImg, S, H, W, yChunks, ny, xChunks, nx, overlapPix, numBins, rms, table, edges, st);
    nb_set_error("nb_cell_rms: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_cell_rms_est(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, const int* yChunks, int ny,
                    const int* xChunks, int nx, int overlapPix, int numBins, int estimator, void* rms, double* table, double* edges,
                    cudaStream_t st) {
    NB_REQUIRE(map && wmed && yChunks && xChunks && table && edges && nImg > 0 && S > 0 && ny > 0 && nx > 0, NB_ERR_ARG,
               "nb_cell_rms_est: bad arguments");
    NB_REQUIRE(numBins >= 1 && numBins <= 64, NB_ERR_ARG, "nb_cell_rms_est: numBins %d out of range (1..64)", numBins);
    NB_REQUIRE(estimator >= 0 && estimator <= 2, NB_ERR_UNSUPPORTED, "nb_cell_rms_est: estimator %d not implemented (0 clip, 1 percentile, 2 biweight)", estimator);
    if (prec == NB_PREC_F64)
        return cell_rms_t<double>(map, wmed, nImg, S, H, W, yChunks, ny, xChunks, nx, overlapPix, numBins, rms, table, edges, st, estimator);
    if (prec == NB_PREC_F32)
        return cell_rms_t<float>(map, wmed, nImg, S, H, W, yChunks, ny, xChunks, nx, overlapPix, numBins, rms, table, edges, st, estimator);
    nb_set_error("nb_cell_rms_est: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_snr_mask(int prec, void* map, void* rms, const unsigned char* psMask, const unsigned char* surveyMask, int nImg, int S,
                int H, int W, int trimPix, int apodPix, int realSpace, void* sn, unsigned char* maskOut, unsigned char* scratch,
                const int* zeroFlags, cudaStream_t st) {
    NB_REQUIRE(map && rms && psMask && surveyMask && sn && maskOut && nImg > 0 && S > 0, NB_ERR_ARG, "nb_snr_mask: bad arguments");
    NB_REQUIRE(trimPix == 0 || scratch, NB_ERR_ARG, "nb_snr_mask: scratch required when trimPix > 0");
    NB_REQUIRE(trimPix >= 0, NB_ERR_ARG, "nb_snr_mask: negative trimPix %d", trimPix);
    NB_REQUIRE(W <= 8192, NB_ERR_UNSUPPORTED, "nb_snr_mask: tile too wide");
    if (prec == NB_PREC_F64)
        return snr_mask_t<double>(map, rms, psMask, surveyMask, nImg, S, H, W, trimPix, apodPix, realSpace, sn, maskOut, scratch, st, zeroFlags);
    if (prec == NB_PREC_F32)
        return snr_mask_t<float>(map, rms, psMask, surveyMask, nImg, S, H, W, trimPix, apodPix, realSpace, sn, maskOut, scratch, st, zeroFlags);
    nb_set_error("nb_snr_mask: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_snr_mask_cells(int prec, void* map, const void* wmed, const double* table, const double* edges, const int* yChunks, int ny,
                      const int* xChunks, int nx, int overlapPix, int numBins, const unsigned char* psMask,
                      const unsigned char* surveyMask, int nImg, int S, int H, int W, int trimPix, int apodPix, int realSpace, void* sn,
                      void* rmsOut, unsigned char* maskOut, unsigned char* scratch, const int* zeroFlags, cudaStream_t st) {
    NB_REQUIRE(map && wmed && table && edges && yChunks && xChunks && psMask && surveyMask && sn && maskOut && nImg > 0 && S > 0,
               NB_ERR_ARG, "nb_snr_mask_cells: bad arguments");      // rmsOut may be NULL: the painted RMS map is then not written
    NB_REQUIRE(trimPix == 0 || scratch, NB_ERR_ARG, "nb_snr_mask_cells: scratch required when trimPix > 0");
    NB_REQUIRE(trimPix >= 0, NB_ERR_ARG, "nb_snr_mask_cells: negative trimPix %d", trimPix);
    NB_REQUIRE(W <= 8192, NB_ERR_UNSUPPORTED, "nb_snr_mask_cells: tile too wide");
    NB_REQUIRE(ny > 0 && nx > 0 && ny <= NB_MAX_CHUNKS && nx <= NB_MAX_CHUNKS, NB_ERR_UNSUPPORTED, "nb_snr_mask_cells: bad noise grid");
    Chunks ch;
    memset(&ch, 0, sizeof(ch));
    for (int i = 0; i <= ny; ++i) ch.y[i] = yChunks[i];
    for (int i = 0; i <= nx; ++i) ch.x[i] = xChunks[i];
    PaintArgs pa = {wmed, table, edges, ny, nx, overlapPix, numBins};
    if (prec == NB_PREC_F64)
        return snr_mask_t<double>(map, rmsOut, psMask, surveyMask, nImg, S, H, W, trimPix, apodPix, realSpace, sn, maskOut, scratch, st, zeroFlags, &pa, &ch);
    if (prec == NB_PREC_F32)
        return snr_mask_t<float>(map, rmsOut, psMask, surveyMask, nImg, S, H, W, trimPix, apodPix, realSpace, sn, maskOut, scratch, st, zeroFlags, &pa, &ch);
    nb_set_error("nb_snr_mask_cells: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_snr_from_rms(int prec, const void* map, const void* rms, void* sn, int nImg, int S, int H, int W, cudaStream_t st) {
    NB_REQUIRE(map && rms && sn && nImg > 0 && S > 0, NB_ERR_ARG, "nb_snr_from_rms: bad arguments");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), nImg);
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); snr_from_rms_kernel<double><<<grid, 256, 0, st>>>((const double*)map, (const double*)rms, (double*)sn, HW, S); NB_PROF_END("snr_from_rms_kernel", st); }
    else if (prec == NB_PREC_F32)
        { NB_PROF_BEGIN(st); snr_from_rms_kernel<float><<<grid, 256, 0, st>>>((const float*)map, (const float*)rms, (float*)sn, HW, S); NB_PROF_END("snr_from_rms_kernel", st); }
    else {
        nb_set_error("nb_snr_from_rms: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

}  // extern "C"
