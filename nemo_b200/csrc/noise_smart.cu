// MapFilter.makeNoiseMap, noiseGridArcmin = "smart" (nemo/filters.py:366-407): the noise level is measured in whole-map bins
// of similar weight instead of grid cells.
//
//   binEdges = np.linspace(medWeights.min(), medWeights.max(), numBins)          -> numBins - 1 bins
//   weightMask_i = (medWeights > binEdges[i]) & (medWeights < binEdges[i + 1])   (strict on both sides: a pixel whose weight
//                                                                                 equals an edge belongs to no bin)
//   3-sigma clipped standard deviation of the non-zero map values of the bin (mean / std, then ten passes of
//   |v| < |mean + 3 sigma|), written to every pixel of the bin when it is > 0.
//
// A bin is a scattered set of up to H * W pixels, so nothing is kept on chip: every pass streams map and weights once
// (16 bytes per pixel, HBM-bound) and accumulates (count, sum, sum of squares about a pivot) per bin with warp-aggregated
// fp64 atomics; a one-block kernel per image turns the sums into the next threshold.  28 small launches per call.
#include "nemo_b200_internal.h"
#include <math.h>
#include <algorithm>

#define SM_MAXBINS 1024
#define SM_THREADS 256
#define SM_PIX_PER_THREAD 8

struct SmartState {      // per (image, bin)
    double mean, rms, thr, pivot, prevC, prevS;
    int done, pad;
};

__device__ __forceinline__ unsigned long long sm_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double sm_unkey(unsigned long long e) {
    const unsigned long long b = (e & 0x8000000000000000ULL) ? (e & 0x7fffffffffffffffULL) : ~e;
    return __longlong_as_double((long long)b);
}

__global__ void smart_init_kernel(unsigned long long* __restrict__ mm, int nTiles, SmartState* __restrict__ state, double* __restrict__ accum,
                                  int* __restrict__ imgDone, int nImg, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nTiles) {
        mm[2 * i] = ~0ULL;
        mm[2 * i + 1] = 0ULL;
    }
    if (i < nImg) imgDone[i] = 0;
    if (i < nImg * nb) {
        SmartState st;
        st.mean = 0; st.rms = 0; st.thr = INFINITY; st.pivot = 0; st.prevC = -1; st.prevS = 0; st.done = 0; st.pad = 0;
        state[i] = st;
        accum[3 * i] = 0; accum[3 * i + 1] = 0; accum[3 * i + 2] = 0;
    }
}

template <typename T>
__global__ void __launch_bounds__(SM_THREADS) smart_minmax_kernel(const T* __restrict__ wmed, long long HW, unsigned long long* __restrict__ mm) {
    const T* w = wmed + (size_t)blockIdx.y * HW;
    unsigned long long lo = ~0ULL, hi = 0ULL;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = sm_key((double)w[p]);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[2 * blockIdx.y], lo);
        atomicMax(&mm[2 * blockIdx.y + 1], hi);
    }
}

// np.linspace(start, stop, num): y = arange(num) * step + start with step = (stop - start) / (num - 1) (multiply, then add: two
// roundings), y = arange(num) / (num - 1) * (stop - start) + start when the step underflows to 0, last element = stop exactly
__global__ void smart_edges_kernel(const unsigned long long* __restrict__ mm, double* __restrict__ edges, int numBins) {
    const int tile = blockIdx.x;
    const double start = sm_unkey(mm[2 * tile]), stop = sm_unkey(mm[2 * tile + 1]);
    const double div = (double)(numBins - 1);
    const double delta = __dsub_rn(stop, start);
    const double step = numBins > 1 ? __ddiv_rn(delta, div) : 0.0;
    for (int i = threadIdx.x; i < numBins; i += blockDim.x) {
        double y;
        if (numBins == 1) y = start;
        else if (step != 0.0) y = __dadd_rn(__dmul_rn((double)i, step), start);
        else y = __dadd_rn(__dmul_rn(__ddiv_rn((double)i, div), delta), start);
        if (numBins > 1 && i == numBins - 1) y = stop;
        edges[(size_t)tile * numBins + i] = y;
    }
}

// bin of weight w: the i with edges[i] < w < edges[i + 1], or -1
__device__ __forceinline__ int smart_bin(double w, const double* __restrict__ s_edges, int numBins) {
    int lo = 0, hi = numBins;             // k = number of edges < w, by bisection (edges ascend, possibly with repeats)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_edges[mid] < w) lo = mid + 1; else hi = mid;
    }
    const int k = lo;
    if (k < 1 || k > numBins - 1) return -1;
    return (s_edges[k] != w) ? k - 1 : -1;   // edges[k] >= w by construction; equality puts the pixel on an edge
}

// one pass over an image: phase 0 count and sum of the non-zero values of every bin, phase 1 squared deviations about the
// mean, phase 2 count / sum / squares about the pivot of the values with |v| < thr
template <typename T>
__global__ void __launch_bounds__(SM_THREADS) smart_sweep_kernel(const T* __restrict__ map, const T* __restrict__ wmed, int S, long long HW,
                                                                 int numBins, const double* __restrict__ edges,
                                                                 const SmartState* __restrict__ state, double* __restrict__ accum,
                                                                 const int* __restrict__ imgDone, int phase) {
    extern __shared__ double s_sm[];      // edges [numBins], thr [nb], pivot [nb], done flags as doubles [nb]
    const int img = blockIdx.y, nb = numBins - 1;
    if (imgDone[img]) return;
    double* s_edges = s_sm;
    double* s_thr = s_sm + numBins;
    double* s_piv = s_thr + nb;
    double* s_done = s_piv + nb;
    for (int i = threadIdx.x; i < numBins; i += blockDim.x) s_edges[i] = edges[(size_t)(img / S) * numBins + i];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const SmartState st = state[(size_t)img * nb + i];
        s_thr[i] = st.thr;
        s_piv[i] = phase == 0 ? 0.0 : (phase == 1 ? st.mean : st.pivot);
        s_done[i] = st.done ? 1.0 : 0.0;
    }
    __syncthreads();
    const T* m = map + (size_t)img * HW;
    const T* wm = wmed + (size_t)(img / S) * HW;
    double* acc = accum + (size_t)img * nb * 3;
    const int lane = threadIdx.x & 31;
    const long long base = ((long long)blockIdx.x * SM_PIX_PER_THREAD) * blockDim.x;
#pragma unroll 2
    for (int u = 0; u < SM_PIX_PER_THREAD; ++u) {
        const long long p = base + (long long)u * blockDim.x + threadIdx.x;
        int bin = -1;
        double c = 0, s = 0, q = 0;
        if (p < HW) {
            const double v = (double)m[p];
            const int b = smart_bin((double)wm[p], s_edges, numBins);
            if (b >= 0 && v != 0.0 && s_done[b] == 0.0) {
                const double d = v - s_piv[b];
                if (phase == 0) {
                    bin = b; c = 1.0; s = v;
                } else if (phase == 1) {
                    bin = b; q = d * d;
                } else if (fabs(v) < s_thr[b]) {
                    bin = b; c = 1.0; s = d; q = d * d;
                }
            }
        }
        // lanes of the same bin combine before the atomics: weights vary smoothly, a warp meets one or two bins
        unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int b = __shfl_sync(0xffffffffu, bin, leader);
            const bool mine = bin == b;
            const unsigned grp = __ballot_sync(0xffffffffu, mine);
            double cc = mine ? c : 0.0, ss = mine ? s : 0.0, qq = mine ? q : 0.0;
            for (int o = 16; o > 0; o >>= 1) {
                cc += __shfl_xor_sync(0xffffffffu, cc, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
                qq += __shfl_xor_sync(0xffffffffu, qq, o);
            }
            if (lane == leader) {
                if (phase != 1) {
                    atomicAdd(&acc[3 * b], cc);
                    atomicAdd(&acc[3 * b + 1], ss);
                }
                if (phase != 0) atomicAdd(&acc[3 * b + 2], qq);
            }
            todo &= ~grp;
        }
    }
}

// sums of a pass -> next state of every bin of an image (the same rules as the grid kernel, filters.py:393-403)
__global__ void smart_update_kernel(SmartState* __restrict__ state, double* __restrict__ accum, int* __restrict__ imgDone, int nb, int phase) {
    const int img = blockIdx.x;
    __shared__ int s_live;
    if (threadIdx.x == 0) s_live = 0;
    __syncthreads();
    if (imgDone[img]) return;
    int live = 0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        SmartState st = state[(size_t)img * nb + b];
        double* a = accum + ((size_t)img * nb + b) * 3;
        if (!st.done) {
            const double c = a[0], s = a[1], q = a[2];
            if (phase == 0) {
                if (c == 0) {
                    st.rms = 0.0;      // "np.not_equal(chunkValues, 0).sum() != 0" fails
                    st.done = 1;
                } else {
                    st.mean = s / c;
                    st.prevC = c;      // the count of phase 0 serves phase 1
                }
                a[0] = 0; a[1] = 0;
            } else if (phase == 1) {
                st.rms = sqrt(q / st.prevC);
                st.thr = fabs(st.mean + 3.0 * st.rms);
                st.pivot = st.mean;
                st.prevC = -1.0;
                a[2] = 0;
            } else {
                if (c == 0) {
                    st.done = 1;       // "if mask.sum() > 0" fails: nothing changes now or later
                } else {
                    const double dm = s / c;
                    double var = q / c - dm * dm;
                    if (var < 0) var = 0;
                    const bool fixed = (c == st.prevC) && (s + c * st.pivot == st.prevS);
                    st.mean = st.pivot + dm;
                    st.rms = sqrt(var);
                    st.thr = fabs(st.mean + 3.0 * st.rms);
                    st.prevC = c;
                    st.prevS = s + c * st.pivot;
                    st.pivot = st.mean;
                    if (fixed) st.done = 1;
                }
                a[0] = 0; a[1] = 0; a[2] = 0;
            }
            state[(size_t)img * nb + b] = st;
        }
        if (!st.done) live = 1;
    }
    if (live) atomicOr(&s_live, 1);
    __syncthreads();
    if (threadIdx.x == 0 && !s_live) imgDone[img] = 1;
}

template <typename T>
__global__ void __launch_bounds__(SM_THREADS) smart_paint_kernel(const T* __restrict__ wmed, int S, long long HW, int numBins,
                                                                 const double* __restrict__ edges, const SmartState* __restrict__ state,
                                                                 T* __restrict__ rms, double* __restrict__ table) {
    extern __shared__ double s_sm[];
    const int img = blockIdx.y, nb = numBins - 1;
    double* s_edges = s_sm;
    double* s_rms = s_sm + numBins;
    for (int i = threadIdx.x; i < numBins; i += blockDim.x) s_edges[i] = edges[(size_t)(img / S) * numBins + i];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const double r = state[(size_t)img * nb + i].rms;
        s_rms[i] = r;
        if (blockIdx.x == 0 && table) table[(size_t)img * nb + i] = r;
    }
    __syncthreads();
    const T* wm = wmed + (size_t)(img / S) * HW;
    T* out = rms + (size_t)img * HW;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
        const int b = smart_bin((double)wm[p], s_edges, numBins);
        const double r = b >= 0 ? s_rms[b] : 0.0;
        out[p] = (T)(r > 0.0 ? r : 0.0);
    }
}

template <typename T>
static int smart_rms_t(const void* map, const void* wmed, int nImg, int S, int H, int W, int numBins, void* rms, double* table,
                       double* edges, void* scratch, cudaStream_t st) {
    const long long HW = (long long)H * W;
    const int nTiles = nImg / S, nb = numBins - 1;
    // scratch: [2 nTiles] u64 min / max keys, [nImg nb] states, [3 nImg nb] sums, [nImg] done flags
    unsigned long long* mm = (unsigned long long*)scratch;
    SmartState* state = (SmartState*)(mm + 2 * nTiles);
    double* accum = (double*)(state + (size_t)nImg * (nb > 0 ? nb : 1));
    int* imgDone = (int*)(accum + 3 * (size_t)nImg * (nb > 0 ? nb : 1));
    const int nInit = std::max(std::max(nTiles, nImg), nImg * nb);
    { NB_PROF_BEGIN(st); smart_init_kernel<<<(nInit + 255) / 256, 256, 0, st>>>(mm, nTiles, state, accum, imgDone, nImg, nb); NB_PROF_END("smart_init_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); smart_minmax_kernel<T><<<dim3(148 * 2, nTiles), SM_THREADS, 0, st>>>((const T*)wmed, HW, mm); NB_PROF_END("smart_minmax_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); smart_edges_kernel<<<nTiles, 128, 0, st>>>(mm, edges, numBins); NB_PROF_END("smart_edges_kernel", st); }
    NB_CHECK_LAUNCH();
    const size_t smem = sizeof(double) * (numBins + 3 * (size_t)(nb > 0 ? nb : 1));
    const unsigned blocks = (unsigned)((HW + (long long)SM_THREADS * SM_PIX_PER_THREAD - 1) / ((long long)SM_THREADS * SM_PIX_PER_THREAD));
    if (nb > 0) {
        for (int pass = 0; pass < 12; ++pass) {
            const int phase = pass < 2 ? pass : 2;
            { NB_PROF_BEGIN(st); smart_sweep_kernel<T><<<dim3(blocks, nImg), SM_THREADS, smem, st>>>((const T*)map, (const T*)wmed, S, HW, numBins, edges, state,
                                                                                                 accum, imgDone, phase); NB_PROF_END("smart_sweep_kernel", st); }
            NB_CHECK_LAUNCH();
            { NB_PROF_BEGIN(st); smart_update_kernel<<<nImg, 128, 0, st>>>(state, accum, imgDone, nb, phase); NB_PROF_END("smart_update_kernel", st); }
            NB_CHECK_LAUNCH();
        }
    }
    { NB_PROF_BEGIN(st); smart_paint_kernel<T><<<dim3(148 * 4, nImg), SM_THREADS, smem, st>>>((const T*)wmed, S, HW, numBins, edges, state, (T*)rms, table); NB_PROF_END("smart_paint_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}


// ------------------------------------------------------------------------------------------------------
// completeness.getRMSTab (nemo/completeness.py:981-999): map area per distinct noise level.  The levels of an RMS map are the
// values its cells / bins painted, so the caller passes them sorted (cand [nImg, K], padded with +inf) and every pixel finds
// its own by bisection; a warp's 32 neighbouring pixels share one or two levels, so they combine before the atomics.
// area[img, k] += rowArea[img / S, y] * mask[p]; total[img] = masked area of the image; miss[img] counts non-zero RMS values
// that are not in the list (none, when the list came from the noise tables).
template <typename T>
__global__ void __launch_bounds__(256) rms_area_kernel(const T* __restrict__ rms, const unsigned char* __restrict__ mask,
                                                       const double* __restrict__ rowArea, int S, int H, int W,
                                                       const double* __restrict__ cand, int K, double* __restrict__ area,
                                                       double* __restrict__ total, int* __restrict__ miss) {
    extern __shared__ double s_cand[];
    const int img = blockIdx.y, y = blockIdx.x;
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_cand[i] = cand[(size_t)img * K + i];
    __syncthreads();
    const double ra = rowArea[(size_t)(img / S) * H + y];
    const T* r = rms + ((size_t)img * H + y) * W;
    const unsigned char* mk = mask + ((size_t)img * H + y) * W;
    const int lane = threadIdx.x & 31;
    double tot = 0;
    for (int x0 = 0; x0 < W; x0 += blockDim.x) {       // uniform trip count: the shuffles need every lane
        const int x = x0 + threadIdx.x;
        int idx = -1;
        double a = 0;
        if (x < W) {
            const double v = (double)r[x];
            a = ra * (double)mk[x];
            tot += a;
            if (v != 0.0) {
                int lo = 0, hi = K;                    // first candidate >= v
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_cand[mid] < v) lo = mid + 1; else hi = mid;
                }
                if (lo < K && s_cand[lo] == v) idx = lo;
                else atomicAdd(&miss[img], 1);
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, idx >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int k = __shfl_sync(0xffffffffu, idx, leader);
            const bool mine = idx == k;
            const unsigned grp = __ballot_sync(0xffffffffu, mine);
            double aa = mine ? a : 0.0;
            for (int o = 16; o > 0; o >>= 1) aa += __shfl_xor_sync(0xffffffffu, aa, o);
            if (lane == leader) atomicAdd(&area[(size_t)img * K + k], aa);
            todo &= ~grp;
        }
    }
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) atomicAdd(&total[img], tot);
}

extern "C" {

int nb_smart_rms(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, int numBins, void* rms, double* table,
                 double* edges, void* scratch, cudaStream_t st) {
    NB_REQUIRE(map && wmed && rms && edges && scratch && nImg > 0 && S > 0 && (nImg % S) == 0, NB_ERR_ARG, "nb_smart_rms: bad arguments");
    NB_REQUIRE(numBins >= 1 && numBins <= SM_MAXBINS, NB_ERR_ARG, "nb_smart_rms: numNoiseBins %d out of range (1..%d)", numBins, SM_MAXBINS);
    if (prec == NB_PREC_F64) return smart_rms_t<double>(map, wmed, nImg, S, H, W, numBins, rms, table, edges, scratch, st);
    if (prec == NB_PREC_F32) return smart_rms_t<float>(map, wmed, nImg, S, H, W, numBins, rms, table, edges, scratch, st);
    nb_set_error("nb_smart_rms: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_rms_area_table(int prec, const void* rms, const unsigned char* mask, const double* rowArea, int nImg, int S, int H, int W,
                      const double* cand, int K, double* area, double* total, int* miss, cudaStream_t st) {
    NB_REQUIRE(rms && mask && rowArea && cand && area && total && miss && nImg > 0 && S > 0 && K > 0, NB_ERR_ARG, "nb_rms_area_table: bad arguments");
    NB_REQUIRE(K <= 16384, NB_ERR_UNSUPPORTED, "nb_rms_area_table: more than 16384 noise levels per image");
    NB_CHECK_CUDA(cudaMemsetAsync(area, 0, sizeof(double) * (size_t)nImg * K, st));
    NB_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(double) * nImg, st));
    NB_CHECK_CUDA(cudaMemsetAsync(miss, 0, sizeof(int) * nImg, st));
    const size_t smem = sizeof(double) * (size_t)K;
    if (prec == NB_PREC_F64) {
        if (smem > 48 * 1024) NB_CHECK_CUDA(cudaFuncSetAttribute(rms_area_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { NB_PROF_BEGIN(st); rms_area_kernel<double><<<dim3(H, nImg), 256, smem, st>>>((const double*)rms, mask, rowArea, S, H, W, cand, K, area, total, miss); NB_PROF_END("rms_area_kernel", st); }
    } else if (prec == NB_PREC_F32) {
        if (smem > 48 * 1024) NB_CHECK_CUDA(cudaFuncSetAttribute(rms_area_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { NB_PROF_BEGIN(st); rms_area_kernel<float><<<dim3(H, nImg), 256, smem, st>>>((const float*)rms, mask, rowArea, S, H, W, cand, K, area, total, miss); NB_PROF_END("rms_area_kernel", st); }
    } else {
        nb_set_error("nb_rms_area_table: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

}  // extern "C"
