// K6 for the common noise set-up of one weight bin per cell (no numNoiseBins in noiseParams: every cluster-search
// configuration of the reference's examples/, DR5ClusterSearch.yml:80-81): MapFilter.makeNoiseMap, grid branch, default
// 3-sigma-clipped estimator (filters.py:409-481) as two streaming kernels instead of one CTA holding the cell on chip.
//
// With one bin the sample of a cell is simply its non-zero pixels (filters.py:440-447, 455-459: the only bin runs from the
// smallest weight of a good pixel to the largest + 1e-6), and two facts make the cell cheap:
//   * Cells overlap by half a grid step, so along each axis the image splits into ZONES -- the core that belongs to one cell
//     and the margin two neighbouring cells share -- and a cell is a 3 x 3 block of zones.  First-pass moments (count, sum,
//     sum of squares about a pivot) and the weight range of the good pixels are therefore taken once per zone block
//     (zone_moments_kernel: every pixel read once) and a cell adds up its nine blocks.
//   * Every clipping pass selects |v| < |mean + 3 rms| from the FULL sample (filters.py:473-477), so its sample is a prefix of
//     the values in |v| order, and the thresholds of successive passes move only slightly.  cell_clip_kernel (one CTA per
//     cell) sweeps the cell once for the first pass, with t_lo = 3/4 of that pass's threshold: values below t_lo go into
//     running moments, the few per cent above it into a list in shared memory.  Later passes touch the list only (one warp,
//     no block barrier).  A pass whose threshold falls below t_lo, or a list that overflows, sweeps the cell again -- the
//     result never depends on the shortcut.  A pass that selects as many values as the one before selects the same values:
//     the iteration has reached its fixed point and stops.
// Sums are fp64 about a pivot (the image's centre pixel for the zone pass, the cell's first-pass mean afterwards), so
// variance = E[d^2] - E[d]^2 loses nothing that matters for maps whose mean is within ~1e5 standard deviations of the pivot.
#include "nemo_b200_internal.h"
#include "noise_cells.cuh"
#include <math.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#define OB_THREADS 256
#define OB_WARPS (OB_THREADS / 32)
#define OB_CAP 3072            // list entries per cell (shared memory: 24 KB of doubles)
#ifndef OB_TLO
#define OB_TLO 0.65           // t_lo as a fraction of the sweeping pass's threshold
#endif
#ifndef OB_MINB
#define OB_MINB 4
#endif
#ifndef OB_U
#define OB_U 8              // loads in flight per thread of a sweeping pass
#endif
#ifndef OB_ZU
#define OB_ZU 4             // loads of map and weights in flight per thread of the zone pass
#endif
#define OB_ZT 5                // doubles per zone block: count, sum d, sum d^2, min weight, max weight (good pixels)

// zone z of an axis with chunk boundaries c[0..n] and overlap ov: even z = 2 i is the core of cell i, odd z = 2 i + 1 the margin
// cells i and i + 1 share.  Requires every chunk to be at least 2 ov long (checked by the host).
__device__ __forceinline__ int zone_lo(const int* __restrict__ c, int n, int ov, int L, int z) {
    if (z <= 0) return 0;
    if (z >= 2 * n - 1) return L;
    const int i = z >> 1;
    return (z & 1) ? c[i + 1] - ov : c[i] + ov;
}

// loads the compiler may neither drop nor move under a later condition (it would turn "load map, load weight" into "load map,
// test, load weight": two memory latencies in a row)
__device__ __forceinline__ double ob_ldg(const double* q) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(q));
    return v;
}
__device__ __forceinline__ float ob_ldg(const float* q) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(q));
    return v;
}

__device__ __forceinline__ double ob_pivot(double centre) { return (centre - centre == 0.0) ? centre : 0.0; }   // finite, else 0

// One CTA per chunk (i, k) takes the four zone blocks (2 i, 2 i + 1) x (2 k, 2 k + 1): the chunk's core, the margins it shares with
// its right and lower neighbours and the corner between them, one after the other.
template <typename T>
__global__ void __launch_bounds__(OB_THREADS) zone_moments_kernel(const T* __restrict__ map, const T* __restrict__ wmed, int S, int H, int W,
                                                                  const Chunks ch, int ny, int nx, int ov, double* __restrict__ zoneTab) {
    __shared__ double s_red[4][OB_ZT][OB_WARPS];
    const int ck = blockIdx.x, ci = blockIdx.y, img = blockIdx.z;
    const int nZy = 2 * ny - 1, nZx = 2 * nx - 1;
    const T* m = map + (size_t)img * H * W;
    const T* wm = wmed + (size_t)(img / S) * H * W;
    const double p = ob_pivot((double)m[(size_t)(H / 2) * W + W / 2]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
        const int zy = 2 * ci + (sub >> 1), zx = 2 * ck + (sub & 1);
        double cnt = 0, s1 = 0, s2 = 0, wmin = INFINITY, wmax = -INFINITY;
        if (zy < nZy && zx < nZx) {
            const int y0 = zone_lo(ch.y, ny, ov, H, zy), y1 = zone_lo(ch.y, ny, ov, H, zy + 1);
            const int x0 = zone_lo(ch.x, nx, ov, W, zx), x1 = zone_lo(ch.x, nx, ov, W, zx + 1);
            const int bw = x1 - x0, n = (y1 - y0) * bw;
            if (n > 0) {
                // pixel e of the block sits at offset e + (e / bw) (W - bw) from its first pixel: 32-bit arithmetic, the division a
                // multiply-high (exact for e * bw < 2^32); loads are unconditional (index clamped), slots past the end count as 0
                const unsigned int magic = bw > 1 ? 0xFFFFFFFFu / (unsigned int)bw + 1u : 0u;
                const unsigned int skip = (unsigned int)(W - bw), last = (unsigned int)(n - 1);
                const T* mb = m + y0 * W + x0;
                const T* wb = wm + y0 * W + x0;
                int icnt = 0;
                auto take = [&](const T v, const T w) {
                    if (v != (T)0) {          // good pixel (filters.py:444)
                        const double d = (double)v - p, wv = (double)w;
                        ++icnt;
                        s1 += d;
                        s2 += d * d;
                        wmin = wv < wmin ? wv : wmin;
                        wmax = wv > wmax ? wv : wmax;
                    }
                };
                auto offsetOf = [&](const unsigned int e) { return e + (bw > 1 ? __umulhi(e, magic) : e) * skip; };
                constexpr int U = OB_ZU;
                unsigned int e0 = threadIdx.x;
                for (; e0 + (U - 1) * OB_THREADS <= last; e0 += U * OB_THREADS) {     // full batches: U loads of each array in flight
                    T v[U], w[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const unsigned int off = offsetOf(e0 + u * OB_THREADS);
                        v[u] = ob_ldg(mb + off);
                        w[u] = ob_ldg(wb + off);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) take(v[u], w[u]);
                }
                for (; e0 <= last; e0 += OB_THREADS) {      // what is left (thin zone blocks never reach the loop above)
                    const unsigned int off = offsetOf(e0);
                    take(ob_ldg(mb + off), ob_ldg(wb + off));
                }
                cnt = (double)icnt;
            }
        }
        cnt = warp_sum(cnt);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        wmin = warp_min(wmin);
        wmax = warp_max(wmax);
        if (lane == 0) {
            s_red[sub][0][wid] = cnt;
            s_red[sub][1][wid] = s1;
            s_red[sub][2][wid] = s2;
            s_red[sub][3][wid] = wmin;
            s_red[sub][4][wid] = wmax;
        }
    }
    __syncthreads();
    if (threadIdx.x < 4 * OB_ZT) {
        const int sub = threadIdx.x / OB_ZT, q = threadIdx.x % OB_ZT;
        const int zy = 2 * ci + (sub >> 1), zx = 2 * ck + (sub & 1);
        if (zy < nZy && zx < nZx) {
            double acc = s_red[sub][q][0];
            for (int w = 1; w < OB_WARPS; ++w) {
                const double t = s_red[sub][q][w];
                acc = q < 3 ? acc + t : (q == 3 ? fmin(acc, t) : fmax(acc, t));
            }
            zoneTab[(((size_t)img * nZy + zy) * nZx + zx) * OB_ZT + q] = acc;
        }
    }
}

#ifdef OB_STATS
__device__ unsigned long long g_ob_stats[4];     // cells, sweeping passes, list passes, list overflows
#define OB_COUNT(i) do { if (threadIdx.x == 0) atomicAdd(&g_ob_stats[i], 1ULL); } while (0)
#else
#define OB_COUNT(i) do { } while (0)
#endif

template <typename T>
__global__ void __launch_bounds__(OB_THREADS, OB_MINB) cell_clip_kernel(const T* __restrict__ map, int H, int W, const Chunks ch, int ny, int nx, int ov,
                                                               const double* __restrict__ zoneTab, double* __restrict__ table,
                                                               double* __restrict__ edges) {
    __shared__ T s_tail[OB_CAP];
    __shared__ double s_red[OB_WARPS][3];
    __shared__ double s_pub[8];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cell = blockIdx.x, img = blockIdx.y;
    const int ci = cell / nx, ck = cell % nx;
    const CellGeom g = cell_geom(ch, ci, ck, ov, H, W);
    const T* m = map + (size_t)img * H * W;
    double* tab = table + (size_t)img * gridDim.x + cell;
    double* edg = edges + ((size_t)img * gridDim.x + cell) * 2;
    const int nZy = 2 * ny - 1, nZx = 2 * nx - 1;

    // ---- first pass from the zone table: the cell is zones 2 i - 1 .. 2 i + 1 along each axis --------------------------
    if (wid == 0) {
        double cnt = 0, s1 = 0, s2 = 0, wmin = INFINITY, wmax = -INFINITY;
        if (lane < 9) {
            const int zy = 2 * ci - 1 + lane / 3, zx = 2 * ck - 1 + lane % 3;
            if (zy >= 0 && zy < nZy && zx >= 0 && zx < nZx) {
                const double* z = zoneTab + (((size_t)img * nZy + zy) * nZx + zx) * OB_ZT;
                cnt = z[0];
                s1 = z[1];
                s2 = z[2];
                wmin = z[3];
                wmax = z[4];
            }
        }
        cnt = warp_sum(cnt);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        wmin = warp_min(wmin);
        wmax = warp_max(wmax);
        if (lane == 0) {
            s_pub[0] = cnt;
            s_pub[1] = s1;
            s_pub[2] = s2;
            s_pub[3] = wmin;
            s_pub[4] = wmax;
        }
    }
    __syncthreads();
    const double n0 = s_pub[0];
    if (n0 == 0) {      // filters.py:446-447: nothing good in the cell, nothing is painted
        if (tid == 0) {
            tab[0] = 0.0;
            edg[0] = edg[1] = __longlong_as_double(0x7ff8000000000000LL);
        }
        return;
    }
    const double pImg = ob_pivot((double)m[(size_t)(H / 2) * W + W / 2]);
    double mean, rms;
    {
        const double e1 = s_pub[1] / n0, e2 = s_pub[2] / n0;
        double var = e2 - e1 * e1;
        if (var <= e2 * (n0 * 8.9e-16)) var = 0.0;     // within the cancellation error of identical values: np.std gives 0 there
        mean = pImg + e1;
        rms = sqrt(var);
        if (tid == 0) {
            edg[0] = s_pub[3];
            edg[1] = s_pub[4] + 1e-6;                  // filters.py:454
        }
    }
    __syncthreads();        // s_pub is reused below

    const double p = (mean - mean == 0.0) ? mean : 0.0;        // pivot of everything that follows
    const int cw = g.x1 - g.x0, nPix = (g.y1 - g.y0) * cw;
    const unsigned int magic = cw > 1 ? 0xFFFFFFFFu / (unsigned int)cw + 1u : 0u;
    const unsigned int skip = (unsigned int)(W - cw), lastPix = (unsigned int)(nPix - 1);
    const T* m0 = m + g.y0 * W + g.x0;
    OB_COUNT(0);
    double prevN = n0;
    int it = 0;
    bool done = false;
    // moments (count, sum d, sum d^2 about p) of the good values with |v| < t; with LIST also appends the good values that
    // fail the test to s_tail (count in s_cnt, entries past OB_CAP dropped).  Block-wide result in every thread.
    auto sweep = [&](const double t, const bool list, double& nb, double& sb, double& qb) {
        constexpr int U = OB_U;
        int icnt = 0;
        sb = 0;
        qb = 0;
        for (int eb = 0; eb < nPix; eb += U * OB_THREADS) {      // warp-uniform trip count: the append uses warp collectives
            T v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int e = min((unsigned int)(eb + tid + u * OB_THREADS), lastPix);      // clamped, unconditional loads; 32-bit offsets
                const unsigned int r = cw > 1 ? __umulhi(e, magic) : e;
                v[u] = ob_ldg(m0 + (e + r * skip));
            }
            unsigned int tmask = 0;          // which of this thread's U values go to the list
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double dv = (double)v[u];
                const bool good = (eb + tid + u * OB_THREADS < nPix) && v[u] != (T)0;
                const bool below = fabs(dv) < t;
                if (good && below) {
                    const double d = dv - p;
                    ++icnt;
                    sb += d;
                    qb += d * d;
                }
                tmask |= ((good && !below) ? 1u : 0u) << u;
            }
            if (list) {
                // one shared-memory atomic per warp and batch: list slots by a warp scan of the per-thread counts
                const int c = __popc(tmask);
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int q = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += q;
                }
                const int total = __shfl_sync(0xffffffffu, inc, 31);
                if (total) {
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&s_cnt, total);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    int idx = base + inc - c;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if ((tmask >> u) & 1u) {
                            if (idx < OB_CAP) s_tail[idx] = v[u];
                            ++idx;
                        }
                    }
                }
            }
        }
        nb = warp_sum((double)icnt);
        sb = warp_sum(sb);
        qb = warp_sum(qb);
        if (lane == 0) {
            s_red[wid][0] = nb;
            s_red[wid][1] = sb;
            s_red[wid][2] = qb;
        }
        __syncthreads();
        nb = sb = qb = 0;
#pragma unroll
        for (int w = 0; w < OB_WARPS; ++w) {
            nb += s_red[w][0];
            sb += s_red[w][1];
            qb += s_red[w][2];
        }
    };
    // one clipping pass given the moments of its sample (filters.py:474-477); the same in every thread that calls it
    auto update = [&](const double n, const double sd, const double qd) {
        if (n > 0) {                 // "if mask.sum() > 0"
            const double e1 = sd / n, e2 = qd / n;
            double var = e2 - e1 * e1;
            if (var <= e2 * (n * 8.9e-16)) var = 0.0;
            mean = p + e1;
            rms = sqrt(var);
        }
        if (n == prevN || n == 0) done = true;     // same sample as the pass before (or none): the fixed point
        prevN = n > 0 ? n : prevN;
        ++it;
    };
    while (it < 10 && !done) {       // filters.py:473
        const double thr = fabs(mean + 3.0 * rms);
        if (!(thr == thr)) break;    // nan: no pass can select anything, nothing changes any more
        // ---- sweep the cell: moments below t_lo, list of everything from t_lo up ------------------------------------------
        const double tlo = OB_TLO * thr;
        OB_COUNT(1);
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        double nb, sb, qb;
        sweep(tlo, true, nb, sb, qb);
        const int nTail = s_cnt;
        __syncthreads();             // everyone has read s_cnt and s_red
        if (nTail > OB_CAP) {        // list incomplete: this pass directly, the next one sweeps again
            OB_COUNT(3);
            sweep(thr, false, nb, sb, qb);
            __syncthreads();
            update(nb, sb, qb);
            continue;
        }
        // ---- passes that only need the list (one warp, no barrier), while their thresholds stay at or above t_lo ----------
        if (wid == 0) {
            while (it < 10 && !done) {
                const double t = fabs(mean + 3.0 * rms);
                if (!(t == t)) {
                    done = true;
                    break;
                }
                if (t < tlo) break;
                OB_COUNT(2);
                double nt = 0, st = 0, qt = 0;
                for (int e = lane; e < nTail; e += 32) {
                    const double dv = (double)s_tail[e];
                    if (fabs(dv) < t) {
                        const double d = dv - p;
                        nt += 1.0;
                        st += d;
                        qt += d * d;
                    }
                }
                nt = warp_sum(nt);
                st = warp_sum(st);
                qt = warp_sum(qt);
                update(nb + nt, sb + st, qb + qt);
            }
            if (lane == 0) {
                s_pub[0] = mean;
                s_pub[1] = rms;
                s_pub[2] = prevN;
                s_pub[3] = (double)it;
                s_pub[4] = done ? 1.0 : 0.0;
            }
        }
        __syncthreads();
        mean = s_pub[0];
        rms = s_pub[1];
        prevN = s_pub[2];
        it = (int)s_pub[3];
        done = s_pub[4] != 0.0;
        __syncthreads();     // s_pub, s_cnt and the list are rewritten by the next sweeping pass
    }
    if (tid == 0) tab[0] = rms;
}

// per-stream scratch for the zone table.  Buffers only grow, and an outgrown buffer is kept: a captured CUDA graph may still
// hold its address.
static std::mutex g_ob_mutex;
static std::map<std::pair<int, cudaStream_t>, std::pair<double*, size_t>> g_ob_scratch;

static int ob_scratch(cudaStream_t st, size_t bytes, double** out) {
    std::lock_guard<std::mutex> lock(g_ob_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    auto& ent = g_ob_scratch[std::make_pair(dev, st)];
    if (ent.second < bytes) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        NB_REQUIRE(cap == cudaStreamCaptureStatusNone, NB_ERR_CUDA,
                   "nb_cell_rms: scratch for this stream must grow under stream capture (run the same call once outside capture first)");
        double* d = nullptr;
        NB_CHECK_CUDA(cudaMalloc(&d, bytes));
        ent.first = d;
        ent.second = bytes;
    }
    *out = ent.first;
    return NB_OK;
}

bool nb_onebin_supported(int H, int W, const int* yChunks, int ny, const int* xChunks, int nx, int ov) {
    if (ov < 0 || ny < 1 || nx < 1) return false;
    for (int i = 0; i < ny; ++i)
        if (yChunks[i + 1] - yChunks[i] < 2 * ov) return false;
    for (int k = 0; k < nx; ++k)
        if (xChunks[k + 1] - xChunks[k] < 2 * ov) return false;
    if (yChunks[0] != 0 || yChunks[ny] != H || xChunks[0] != 0 || xChunks[nx] != W) return false;
    // the kernels divide a pixel index by the cell width with a multiply-high that is exact while index * width < 2^32
    long long maxH = 0, maxW = 0;
    for (int i = 0; i < ny; ++i) maxH = std::max<long long>(maxH, std::min(yChunks[i + 1] + ov, H) - std::max(yChunks[i] - ov, 0));
    for (int k = 0; k < nx; ++k) maxW = std::max<long long>(maxW, std::min(xChunks[k + 1] + ov, W) - std::max(xChunks[k] - ov, 0));
    return maxH * maxW * maxW < (1LL << 32) && (long long)H * W < (1LL << 31);
}

template <typename T>
static int onebin_t(const T* map, const T* wmed, int nImg, int S, int H, int W, const Chunks& ch, int ny, int nx, int ov, double* table,
                    double* edges, cudaStream_t st) {
    const int nZy = 2 * ny - 1, nZx = 2 * nx - 1;
    NB_REQUIRE(nImg <= 65535, NB_ERR_UNSUPPORTED, "nb_cell_rms: more than 65535 images in one call");
    double* zoneTab = nullptr;
    int rc = ob_scratch(st, sizeof(double) * OB_ZT * (size_t)nImg * nZy * nZx, &zoneTab);
    if (rc != NB_OK) return rc;
    { NB_PROF_BEGIN(st); zone_moments_kernel<T><<<dim3(nx, ny, nImg), OB_THREADS, 0, st>>>(map, wmed, S, H, W, ch, ny, nx, ov, zoneTab); NB_PROF_END("zone_moments_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); cell_clip_kernel<T><<<dim3(ny * nx, nImg), OB_THREADS, 0, st>>>(map, H, W, ch, ny, nx, ov, zoneTab, table, edges); NB_PROF_END("cell_clip_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

#ifdef OB_STATS
extern "C" int nb_onebin_stats(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_ob_stats, sizeof(unsigned long long) * 4);
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_ob_stats, z, sizeof(z));
    }
    return 0;
}
#endif

int nb_cell_rms_onebin(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, const Chunks& ch, int ny, int nx, int ov,
                       double* table, double* edges, cudaStream_t st) {
    if (prec == NB_PREC_F64) return onebin_t<double>((const double*)map, (const double*)wmed, nImg, S, H, W, ch, ny, nx, ov, table, edges, st);
    return onebin_t<float>((const float*)map, (const float*)wmed, nImg, S, H, W, ch, ny, nx, ov, table, edges, st);
}
