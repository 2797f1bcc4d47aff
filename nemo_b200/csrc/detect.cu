// Detection kernels (K9, K10, K11 of SURVEY.md section 2a).
//   nb_label4         : ndimage.label(map > thr), 4-connectivity, scipy's numbering    photometry.py:212-214
//   nb_label_moments  : center_of_mass / sum / maximum_position per label              photometry.py:215-220
//   nb_spline_eval    : RectBivariateSpline(...)(y, x)                                 photometry.py:76-79, 122
#include "nemo_b200_internal.h"
#include "spline.cuh"

// ------------------------------------------------------------------------------------------------------
// K9.  Union-find on the linear pixel index; every root is the smallest index of its component, i.e. the
// component's first pixel in raster order, so ranking the roots in index order reproduces scipy's labels.
#define LBL_BLOCK 1024

// L holds, for foreground pixels, the index *within the image* of a pixel of the same component with a smaller
// or equal index (a parent pointer); -1 for background.
template <typename T>
__global__ void __launch_bounds__(LBL_BLOCK) label_init_img_kernel(const T* __restrict__ map, double thr, double sign,
                                                                   int* __restrict__ L, int HW) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const size_t off = (size_t)blockIdx.y * HW + p;
    const double v = sign * (double)map[off];
    L[off] = (v > thr) ? p : -1;
}

__device__ __forceinline__ int uf_find(const int* L, int a) {
    int r = L[a];
    while (r != a) {
        a = r;
        r = L[a];
    }
    return a;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done = false;
    while (!done) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    }
}

__global__ void __launch_bounds__(LBL_BLOCK) label_merge_kernel(int* Lall, int H, int W) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    int* L = Lall + (size_t)blockIdx.y * HW;
    if (L[p] < 0) return;
    const int x = p % W;
    if (x > 0 && L[p - 1] >= 0) uf_union(L, p, p - 1);
    if (p >= W && L[p - W] >= 0) uf_union(L, p, p - W);
}

// flatten to roots and count roots per block
__global__ void __launch_bounds__(LBL_BLOCK) label_flatten_kernel(int* Lall, int HW, int* __restrict__ blockCounts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    int* L = Lall + (size_t)blockIdx.y * HW;
    int isRoot = 0;
    if (p < HW && L[p] >= 0) {
        const int r = uf_find(L, p);
        isRoot = (r == p);
        if (!isRoot) L[p] = r;  // roots are never rewritten, so concurrent finds stay valid
    }
    const int c = __syncthreads_count(isRoot);
    if (threadIdx.x == 0) blockCounts[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = c;
}

// exclusive scan of the per-block root counts, one CTA per image
__global__ void __launch_bounds__(1024) label_scan_kernel(int* __restrict__ blockCounts, int nBlocks, int* __restrict__ nLabels) {
    int* bc = blockCounts + (size_t)blockIdx.x * nBlocks;
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nBlocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < nBlocks) ? bc[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_w[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        const int warpOff = (wid > 0) ? s_w[wid - 1] : 0;
        if (i < nBlocks) bc[i] = carry + warpOff + inc - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + warpOff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) nLabels[blockIdx.x] = s_carry;
}

// roots receive -(label + 1) (labels start at 1, so codes are <= -2; -1 stays background)
__global__ void __launch_bounds__(LBL_BLOCK) label_rank_kernel(int* __restrict__ Lall, int HW, const int* __restrict__ blockOffsets) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    int* L = Lall + (size_t)blockIdx.y * HW;
    const int isRoot = (p < HW && L[p] == p) ? 1 : 0;
    __shared__ int s_w[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, isRoot);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
        int w = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_w[lane] = w;
    }
    __syncthreads();
    if (isRoot) {
        const int rank = blockOffsets[(size_t)blockIdx.y * gridDim.x + blockIdx.x] + ((wid > 0) ? s_w[wid - 1] : 0) + before;
        L[p] = -(rank + 2);
    }
}

__global__ void __launch_bounds__(LBL_BLOCK) label_final_kernel(const int* __restrict__ Lall, int* __restrict__ labels, int HW) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int* L = Lall + (size_t)blockIdx.y * HW;
    const int v = L[p];
    int lab = 0;
    if (v <= -2)
        lab = -v - 1;
    else if (v >= 0)
        lab = -L[v] - 1;
    labels[(size_t)blockIdx.y * HW + p] = lab;
}

// ------------------------------------------------------------------------------------------------------
// K10.  table rows: {sum v, sum y v, sum x v, count, max v (order-encoded until finalised), first argmax, first pixel, 0}
__device__ __forceinline__ unsigned long long enc_ordered(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dec_ordered(unsigned long long e) {
    const unsigned long long b = (e & 0x8000000000000000ULL) ? (e & 0x7fffffffffffffffULL) : ~e;
    return __longlong_as_double((long long)b);
}

__global__ void moments_init_kernel(double* __restrict__ table, long long nRows) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRows) return;
    double* t = table + r * 8;
    t[0] = t[1] = t[2] = t[3] = 0.0;
    reinterpret_cast<unsigned long long*>(t)[4] = 0ULL;  // below every encoded value
    t[5] = INFINITY;
    t[6] = INFINITY;
    t[7] = 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256) moments_accum_kernel(const T* __restrict__ map, double sign, const int* __restrict__ labels,
                                                            double* __restrict__ table, int H, int W, int maxLabels) {
    const int HW = H * W;
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    const int* lb = labels + (size_t)img * HW;
    double* tab = table + (size_t)img * (maxLabels + 1) * 8;
    // label 0 (background) is accumulated per thread and reduced; objects go straight to atomics
    double s0 = 0, sy0 = 0, sx0 = 0, c0 = 0, mx0 = -INFINITY;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const int l = lb[p];
        const double v = sign * (double)m[p];
        const int y = p / W, x = p - y * W;
        if (l == 0) {
            s0 += v;
            sy0 += v * (double)y;
            sx0 += v * (double)x;
            c0 += 1.0;
            mx0 = fmax(mx0, v);
        } else if (l <= maxLabels) {
            double* t = tab + (size_t)l * 8;
            atomicAdd(t + 0, v);
            atomicAdd(t + 1, v * (double)y);
            atomicAdd(t + 2, v * (double)x);
            atomicAdd(t + 3, 1.0);
            atomicMax(reinterpret_cast<unsigned long long*>(t) + 4, enc_ordered(v));
            atomicMin(reinterpret_cast<unsigned long long*>(t) + 6, (unsigned long long)__double_as_longlong((double)p));
        }
    }
    s0 = warp_sum(s0);
    sy0 = warp_sum(sy0);
    sx0 = warp_sum(sx0);
    c0 = warp_sum(c0);
    mx0 = warp_max(mx0);
    if ((threadIdx.x & 31) == 0 && c0 > 0) {
        atomicAdd(tab + 0, s0);
        atomicAdd(tab + 1, sy0);
        atomicAdd(tab + 2, sx0);
        atomicAdd(tab + 3, c0);
        atomicMax(reinterpret_cast<unsigned long long*>(tab) + 4, enc_ordered(mx0));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) moments_argmax_kernel(const T* __restrict__ map, double sign, const int* __restrict__ labels,
                                                             double* __restrict__ table, int H, int W, int maxLabels) {
    const int HW = H * W;
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    const int* lb = labels + (size_t)img * HW;
    double* tab = table + (size_t)img * (maxLabels + 1) * 8;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const int l = lb[p];
        if (l > maxLabels) continue;
        const double v = sign * (double)m[p];
        double* t = tab + (size_t)l * 8;
        if (enc_ordered(v) == reinterpret_cast<const unsigned long long*>(t)[4])
            atomicMin(reinterpret_cast<unsigned long long*>(t) + 5, (unsigned long long)__double_as_longlong((double)p));
    }
}

__global__ void moments_final_kernel(double* __restrict__ table, long long nRows) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRows) return;
    double* t = table + r * 8;
    const unsigned long long e = reinterpret_cast<unsigned long long*>(t)[4];
    t[4] = (e == 0ULL) ? -INFINITY : dec_ordered(e);
    if (t[3] == 0.0) {
        t[5] = -1.0;
        t[6] = -1.0;
    }
    if (t[5] == INFINITY) t[5] = -1.0;
    if (t[6] == INFINITY) t[6] = -1.0;  // label 0 has no tracked first pixel
}

// The sparse detector knows an image's label count before it takes the moments, so only the rows it will fill are prepared and
// finished (the tables have room for 8192 labels per image: 67 MB per 128 images, of which a step's images use a few dozen rows
// each).  Row 0 (background, accumulated by the dense scan) is prepared before the scan; rows 1 .. min(nLabels, cap) after the
// ranking.  Rows beyond an image's count are never read (the host cuts at the count, nb_ring_flags takes the counts).
__global__ void moments_init_row0_kernel(double* __restrict__ table, int maxLabels, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* t = table + (size_t)b * (maxLabels + 1) * 8;
    t[0] = t[1] = t[2] = t[3] = 0.0;
    reinterpret_cast<unsigned long long*>(t)[4] = 0ULL;
    t[5] = INFINITY;
    t[6] = INFINITY;
    t[7] = 0.0;
}

template <bool FINAL>
__global__ void moments_rows_kernel(double* __restrict__ table, int maxLabels, const int* __restrict__ nLabels) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x + (FINAL ? 0 : 1);
    if (r > min(nLabels[b], maxLabels)) return;
    double* t = table + ((size_t)b * (maxLabels + 1) + r) * 8;
    if (!FINAL) {
        t[0] = t[1] = t[2] = t[3] = 0.0;
        reinterpret_cast<unsigned long long*>(t)[4] = 0ULL;
        t[5] = INFINITY;
        t[6] = INFINITY;
        t[7] = 0.0;
    } else {
        const unsigned long long e = reinterpret_cast<unsigned long long*>(t)[4];
        t[4] = (e == 0ULL) ? -INFINITY : dec_ordered(e);
        if (t[3] == 0.0) {
            t[5] = -1.0;
            t[6] = -1.0;
        }
        if (t[5] == INFINITY) t[5] = -1.0;
        if (t[6] == INFINITY) t[6] = -1.0;
    }
}

// ------------------------------------------------------------------------------------------------------
// K11.  One warp per evaluation point.
template <typename T>
__global__ void __launch_bounds__(128) spline_eval_kernel(const T* __restrict__ img, int H, int W, const int* __restrict__ imgIdx,
                                                          const double* __restrict__ yx, double* __restrict__ out, int K, double sign) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= K) return;  // whole warps exit together
    const int b = imgIdx ? imgIdx[k] : 0;
    const double v = spline_eval_warp<T>(img + (size_t)b * H * W, H, W, yx[2 * k], yx[2 * k + 1], sign);
    if ((threadIdx.x & 31) == 0) out[k] = v;
}

// ------------------------------------------------------------------------------------------------------
extern "C" {

int nb_label4(int prec, const void* map, double threshold, double sign, int* labels, int* nLabels, int* scratch,
              int B, int H, int W, cudaStream_t st) {
    NB_REQUIRE(map && labels && nLabels && scratch && B > 0 && H > 0 && W > 0, NB_ERR_ARG, "nb_label4: bad arguments");
    NB_REQUIRE((long long)H * W < 0x7fffffffLL, NB_ERR_UNSUPPORTED, "nb_label4: image too large");
    const int HW = H * W;
    const int nBlocks = (HW + LBL_BLOCK - 1) / LBL_BLOCK;
    dim3 grid(nBlocks, B);
    // scratch: [B, HW] parent pointers followed by [B, nBlocks] block counts
    int* L = scratch;
    int* blockCounts = scratch + (size_t)B * HW;
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); label_init_img_kernel<double><<<grid, LBL_BLOCK, 0, st>>>((const double*)map, threshold, sign, L, HW); NB_PROF_END("label_init_img_kernel", st); }
    else if (prec == NB_PREC_F32)
        { NB_PROF_BEGIN(st); label_init_img_kernel<float><<<grid, LBL_BLOCK, 0, st>>>((const float*)map, threshold, sign, L, HW); NB_PROF_END("label_init_img_kernel", st); }
    else {
        nb_set_error("nb_label4: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); label_merge_kernel<<<grid, LBL_BLOCK, 0, st>>>(L, H, W); NB_PROF_END("label_merge_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); label_flatten_kernel<<<grid, LBL_BLOCK, 0, st>>>(L, HW, blockCounts); NB_PROF_END("label_flatten_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); label_scan_kernel<<<B, 1024, 0, st>>>(blockCounts, nBlocks, nLabels); NB_PROF_END("label_scan_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); label_rank_kernel<<<grid, LBL_BLOCK, 0, st>>>(L, HW, blockCounts); NB_PROF_END("label_rank_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); label_final_kernel<<<grid, LBL_BLOCK, 0, st>>>(L, labels, HW); NB_PROF_END("label_final_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

long long nb_label4_scratch_ints(int B, int H, int W) {
    const long long HW = (long long)H * W;
    return (long long)B * HW + (long long)B * ((HW + LBL_BLOCK - 1) / LBL_BLOCK);
}

int nb_label_moments(int prec, const void* map, double sign, const int* labels, double* table, int B, int H, int W,
                     int maxLabels, cudaStream_t st) {
    NB_REQUIRE(map && labels && table && B > 0 && maxLabels >= 0, NB_ERR_ARG, "nb_label_moments: bad arguments");
    const long long nRows = (long long)B * (maxLabels + 1);
    { NB_PROF_BEGIN(st); moments_init_kernel<<<(unsigned)((nRows + 255) / 256), 256, 0, st>>>(table, nRows); NB_PROF_END("moments_init_kernel", st); }
    NB_CHECK_LAUNCH();
    const int HW = H * W;
    int gx = (HW + 256 * 8 - 1) / (256 * 8);
    if (gx < 1) gx = 1;
    dim3 grid(gx, B);
    if (prec == NB_PREC_F64) {
        { NB_PROF_BEGIN(st); moments_accum_kernel<double><<<grid, 256, 0, st>>>((const double*)map, sign, labels, table, H, W, maxLabels); NB_PROF_END("moments_accum_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); moments_argmax_kernel<double><<<grid, 256, 0, st>>>((const double*)map, sign, labels, table, H, W, maxLabels); NB_PROF_END("moments_argmax_kernel", st); }
    } else if (prec == NB_PREC_F32) {
        { NB_PROF_BEGIN(st); moments_accum_kernel<float><<<grid, 256, 0, st>>>((const float*)map, sign, labels, table, H, W, maxLabels); NB_PROF_END("moments_accum_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); moments_argmax_kernel<float><<<grid, 256, 0, st>>>((const float*)map, sign, labels, table, H, W, maxLabels); NB_PROF_END("moments_argmax_kernel", st); }
    } else {
        nb_set_error("nb_label_moments: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); moments_final_kernel<<<(unsigned)((nRows + 255) / 256), 256, 0, st>>>(table, nRows); NB_PROF_END("moments_final_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_spline_eval(int prec, const void* img, int B, int H, int W, const int* imgIdx, const double* yx, double* out,
                   int K, double sign, cudaStream_t st) {
    NB_REQUIRE(img && yx && out && B > 0, NB_ERR_ARG, "nb_spline_eval: bad arguments");
    if (K <= 0) return NB_OK;
    const int wpb = 4;
    const unsigned grid = (unsigned)((K + wpb - 1) / wpb);
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); spline_eval_kernel<double><<<grid, 32 * wpb, 0, st>>>((const double*)img, H, W, imgIdx, yx, out, K, sign); NB_PROF_END("spline_eval_kernel", st); }
    else if (prec == NB_PREC_F32)
        { NB_PROF_BEGIN(st); spline_eval_kernel<float><<<grid, 32 * wpb, 0, st>>>((const float*)img, H, W, imgIdx, yx, out, K, sign); NB_PROF_END("spline_eval_kernel", st); }
    else {
        nb_set_error("nb_spline_eval: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------
// K9 + K10 fused for sparse foregrounds (nb_detect).  Above a detection threshold only ~0.1-1 % of an S/N map is
// foreground, so only ONE dense pass is made (threshold, background moments, zeroed label image, compact list of
// the foreground pixels); linking, root finding, ranking, labelling and the per-label moments then run over the
// list.  Numbering is scipy's: roots are the smallest linear index of their component and are ranked in index order.
#define DT_THREADS 256
#define DT_ITEMS 8
#define DT_ROOTCAP 8192

struct DetectScratch {
    int* L;          // [B, HW] parent pointers, touched only at foreground pixels
    int* fgList;     // [B, listCap]
    int* rootList;   // [B, DT_ROOTCAP]
    int* fgCount;    // [B]
    int* rootCount;  // [B]
    int listCap;
};

__host__ __device__ inline long long detect_list_cap(int H, int W) {
    long long c = ((long long)H * W) / 8;
    return c < 4096 ? 4096 : c;
}

__global__ void detect_reset_kernel(int* __restrict__ fgCount, int* __restrict__ rootCount, int* __restrict__ status, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        fgCount[i] = 0;
        rootCount[i] = 0;
        status[i] = 0;
    }
}

template <typename T>
__global__ void __launch_bounds__(DT_THREADS) detect_scan_kernel(const T* __restrict__ map, double thr, double sign,
                                                                 int* __restrict__ labels, DetectScratch sc, double* __restrict__ table,
                                                                 int* __restrict__ status, int H, int W, int maxLabels) {
    const int HW = H * W;
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    int* lab = labels + (size_t)img * HW;
    int* L = sc.L + (size_t)img * HW;
    int* list = sc.fgList + (size_t)img * sc.listCap;
    const int lane = threadIdx.x & 31;
    double s0 = 0, sy0 = 0, sx0 = 0, c0 = 0, mx0 = -INFINITY;
    const int base = blockIdx.x * (DT_THREADS * DT_ITEMS) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < DT_ITEMS; ++k) {
        const int p = base + k * DT_THREADS;
        const bool in = p < HW;
        const double v = in ? sign * (double)m[p] : 0.0;
        const bool fg = in && (v > thr);
        if (in) lab[p] = 0;
        if (in && !fg) {
            const int y = p / W, x = p - y * W;
            s0 += v;
            sy0 += v * (double)y;
            sx0 += v * (double)x;
            c0 += 1.0;
            mx0 = fmax(mx0, v);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, fg);
        if (bal) {
            int start = 0;
            if (lane == 0) start = atomicAdd(&sc.fgCount[img], __popc(bal));
            start = __shfl_sync(0xffffffffu, start, 0);
            if (fg) {
                const int pos = start + __popc(bal & ((1u << lane) - 1u));
                L[p] = p;
                if (pos < sc.listCap) list[pos] = p;
                else status[img] = 1;
            }
        }
    }
    s0 = warp_sum(s0);
    sy0 = warp_sum(sy0);
    sx0 = warp_sum(sx0);
    c0 = warp_sum(c0);
    mx0 = warp_max(mx0);
    __shared__ double s_b[DT_THREADS / 32][5];
    if (lane == 0) {
        double* r = s_b[threadIdx.x >> 5];
        r[0] = s0; r[1] = sy0; r[2] = sx0; r[3] = c0; r[4] = mx0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, d = 0, e = -INFINITY;
        for (int w = 0; w < DT_THREADS / 32; ++w) {
            a += s_b[w][0]; b += s_b[w][1]; c += s_b[w][2]; d += s_b[w][3]; e = fmax(e, s_b[w][4]);
        }
        if (d > 0) {
            double* tab = table + (size_t)img * (maxLabels + 1) * 8;
            atomicAdd(tab + 0, a);
            atomicAdd(tab + 1, b);
            atomicAdd(tab + 2, c);
            atomicAdd(tab + 3, d);
            atomicMax(reinterpret_cast<unsigned long long*>(tab) + 4, enc_ordered(e));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) detect_link_kernel(const T* __restrict__ map, double thr, double sign, DetectScratch sc, int H, int W) {
    const int HW = H * W;
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    int* L = sc.L + (size_t)img * HW;
    const int* list = sc.fgList + (size_t)img * sc.listCap;
    const int n = min(sc.fgCount[img], sc.listCap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i];
        const int x = p % W;
        if (x > 0 && sign * (double)m[p - 1] > thr) uf_union(L, p, p - 1);
        if (p >= W && sign * (double)m[p - W] > thr) uf_union(L, p, p - W);
    }
}

__global__ void __launch_bounds__(256) detect_flatten_kernel(DetectScratch sc, int* __restrict__ status, int HW) {
    const int img = blockIdx.y;
    int* L = sc.L + (size_t)img * HW;
    const int* list = sc.fgList + (size_t)img * sc.listCap;
    int* roots = sc.rootList + (size_t)img * DT_ROOTCAP;
    const int n = min(sc.fgCount[img], sc.listCap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i];
        const int r = uf_find(L, p);
        if (r == p) {
            const int pos = atomicAdd(&sc.rootCount[img], 1);
            if (pos < DT_ROOTCAP) roots[pos] = p;
            else status[img] = 1;
        } else {
            L[p] = r;  // roots are never rewritten, so concurrent finds stay valid
        }
    }
}

// one CTA per image: sort the roots by pixel index (bitonic, shared memory) and store -(rank + 2) at each root
__global__ void __launch_bounds__(1024) detect_rank_kernel(DetectScratch sc, int* __restrict__ nLabels, int HW) {
    __shared__ int s_r[DT_ROOTCAP];
    const int img = blockIdx.x;
    int* L = sc.L + (size_t)img * HW;
    const int* roots = sc.rootList + (size_t)img * DT_ROOTCAP;
    const int R = min(sc.rootCount[img], DT_ROOTCAP);
    int n2 = 1;
    while (n2 < R) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) s_r[i] = (i < R) ? roots[i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = s_r[i], b = s_r[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) {
                        s_r[i] = b;
                        s_r[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < R; i += blockDim.x) L[s_r[i]] = -(i + 2);
    if (threadIdx.x == 0) nLabels[img] = sc.rootCount[img];
}

template <typename T>
__global__ void __launch_bounds__(256) detect_moments_kernel(const T* __restrict__ map, double sign, DetectScratch sc,
                                                             int* __restrict__ labels, double* __restrict__ table, int H, int W,
                                                             int maxLabels) {
    const int HW = H * W;
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    const int* L = sc.L + (size_t)img * HW;
    int* lab = labels ? labels + (size_t)img * HW : nullptr;
    const int* list = sc.fgList + (size_t)img * sc.listCap;
    double* tab = table + (size_t)img * (maxLabels + 1) * 8;
    const int n = min(sc.fgCount[img], sc.listCap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i];
        const int v0 = L[p];
        const int l = (v0 <= -2) ? (-v0 - 1) : (-L[v0] - 1);
        if (l < 1) continue;   // an unranked root: only when the compact structures overflowed (status is set)
        if (labels) lab[p] = l;
        if (l <= maxLabels) {
            const double v = sign * (double)m[p];
            const int y = p / W, x = p - y * W;
            double* t = tab + (size_t)l * 8;
            atomicAdd(t + 0, v);
            atomicAdd(t + 1, v * (double)y);
            atomicAdd(t + 2, v * (double)x);
            atomicAdd(t + 3, 1.0);
            atomicMax(reinterpret_cast<unsigned long long*>(t) + 4, enc_ordered(v));
            atomicMin(reinterpret_cast<unsigned long long*>(t) + 6, (unsigned long long)__double_as_longlong((double)p));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) detect_argmax_kernel(const T* __restrict__ map, double sign, DetectScratch sc,
                                                            double* __restrict__ table, int HW, int maxLabels) {
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    const int* L = sc.L + (size_t)img * HW;
    const int* list = sc.fgList + (size_t)img * sc.listCap;
    double* tab = table + (size_t)img * (maxLabels + 1) * 8;
    const int n = min(sc.fgCount[img], sc.listCap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i];
        const int v0 = L[p];
        const int l = (v0 <= -2) ? (-v0 - 1) : (-L[v0] - 1);
        if (l < 1 || l > maxLabels) continue;
        double* t = tab + (size_t)l * 8;
        if (enc_ordered(sign * (double)m[p]) == reinterpret_cast<const unsigned long long*>(t)[4])
            atomicMin(reinterpret_cast<unsigned long long*>(t) + 5, (unsigned long long)__double_as_longlong((double)p));
    }
}

// first arg-max of the background (label 0), dense; only when the caller asks for maximum positions
template <typename T>
__global__ void __launch_bounds__(256) detect_bg_argmax_kernel(const T* __restrict__ map, double sign, double thr,
                                                               double* __restrict__ table, int HW, int maxLabels) {
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    double* t = table + (size_t)img * (maxLabels + 1) * 8;
    const unsigned long long want = reinterpret_cast<const unsigned long long*>(t)[4];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const double v = sign * (double)m[p];
        if (!(v > thr) && enc_ordered(v) == want)
            atomicMin(reinterpret_cast<unsigned long long*>(t) + 5, (unsigned long long)__double_as_longlong((double)p));
    }
}

template <typename T>
static int detect_t(const T* map, double thr, double sign, int* labels, int* nLabels, int* status, double* table, int maxLabels,
                    int* scratch, int B, int H, int W, int flags, cudaStream_t st) {
    const int HW = H * W;
    DetectScratch sc;
    sc.listCap = (int)detect_list_cap(H, W);
    sc.L = scratch;
    sc.fgList = sc.L + (size_t)B * HW;
    sc.rootList = sc.fgList + (size_t)B * sc.listCap;
    sc.fgCount = sc.rootList + (size_t)B * DT_ROOTCAP;
    sc.rootCount = sc.fgCount + B;
    const long long nRows = (long long)B * (maxLabels + 1);
    { NB_PROF_BEGIN(st); detect_reset_kernel<<<(B + 255) / 256, 256, 0, st>>>(sc.fgCount, sc.rootCount, status, B); NB_PROF_END("detect_reset_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); moments_init_kernel<<<(unsigned)((nRows + 255) / 256), 256, 0, st>>>(table, nRows); NB_PROF_END("moments_init_kernel", st); }
    NB_CHECK_LAUNCH();
    dim3 gridDense((HW + DT_THREADS * DT_ITEMS - 1) / (DT_THREADS * DT_ITEMS), B);
    { NB_PROF_BEGIN(st); detect_scan_kernel<T><<<gridDense, DT_THREADS, 0, st>>>(map, thr, sign, labels, sc, table, status, H, W, maxLabels); NB_PROF_END("detect_scan_kernel", st); }
    NB_CHECK_LAUNCH();
    dim3 gridSparse(64, B);
    { NB_PROF_BEGIN(st); detect_link_kernel<T><<<gridSparse, 256, 0, st>>>(map, thr, sign, sc, H, W); NB_PROF_END("detect_link_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); detect_flatten_kernel<<<gridSparse, 256, 0, st>>>(sc, status, HW); NB_PROF_END("detect_flatten_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); detect_rank_kernel<<<B, 1024, 0, st>>>(sc, nLabels, HW); NB_PROF_END("detect_rank_kernel", st); }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); detect_moments_kernel<T><<<gridSparse, 256, 0, st>>>(map, sign, sc, labels, table, H, W, maxLabels); NB_PROF_END("detect_moments_kernel", st); }
    NB_CHECK_LAUNCH();
    if (flags & 1) {
        { NB_PROF_BEGIN(st); detect_argmax_kernel<T><<<gridSparse, 256, 0, st>>>(map, sign, sc, table, HW, maxLabels); NB_PROF_END("detect_argmax_kernel", st); }
        NB_CHECK_LAUNCH();
        dim3 gridBg((HW + 256 * 8 - 1) / (256 * 8), B);
        { NB_PROF_BEGIN(st); detect_bg_argmax_kernel<T><<<gridBg, 256, 0, st>>>(map, sign, thr, table, HW, maxLabels); NB_PROF_END("detect_bg_argmax_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    { NB_PROF_BEGIN(st); moments_final_kernel<<<(unsigned)((nRows + 255) / 256), 256, 0, st>>>(table, nRows); NB_PROF_END("moments_final_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

// ------------------------------------------------------------------------------------------------------
// findObjects' device work in one call (nb_find_objects): the map is read ONCE for the detection and the ring thresholds,
// both foregrounds are linked / numbered / reduced over their compact lists, and the ring test is evaluated per candidate.
// No dense label image is written unless asked for.
struct DetectPass {
    double thr;
    DetectScratch sc;
    double* table;
    int* status;
    int maxLabels;
};

template <typename T, int NP, bool BG>
__global__ void __launch_bounds__(DT_THREADS) detect_scan_multi_kernel(const T* __restrict__ map, double sign, DetectPass p0,
                                                                       DetectPass p1, int* __restrict__ labels0, int H, int W) {
    const int HW = H * W;
    const int img = blockIdx.y;
    const T* m = map + (size_t)img * HW;
    const int lane = threadIdx.x & 31;
    const DetectPass* ps[2] = {&p0, &p1};
    double acc[NP][5];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
        acc[q][4] = -INFINITY;
    }
    const int base = blockIdx.x * (DT_THREADS * DT_ITEMS) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < DT_ITEMS; ++k) {
        const int p = base + k * DT_THREADS;
        const bool in = p < HW;
        const double v = in ? sign * (double)m[p] : 0.0;
        const int y = BG ? p / W : 0, x = BG ? p - y * W : 0;
        if (labels0 && in) labels0[(size_t)img * HW + p] = 0;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const DetectPass& ps_q = *ps[q];
            const bool fg = in && (v > ps_q.thr);
            if (BG && in && !fg) {
                acc[q][0] += v;
                acc[q][1] += v * (double)y;
                acc[q][2] += v * (double)x;
                acc[q][3] += 1.0;
                acc[q][4] = fmax(acc[q][4], v);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, fg);
            if (bal) {
                int start = 0;
                if (lane == 0) start = atomicAdd(&ps_q.sc.fgCount[img], __popc(bal));
                start = __shfl_sync(0xffffffffu, start, 0);
                if (fg) {
                    const int pos = start + __popc(bal & ((1u << lane) - 1u));
                    ps_q.sc.L[(size_t)img * HW + p] = p;
                    if (pos < ps_q.sc.listCap) ps_q.sc.fgList[(size_t)img * ps_q.sc.listCap + pos] = p;
                    else ps_q.status[img] = 1;
                }
            }
        }
    }
    if constexpr (BG) {   // otherwise label 0 (background) moments are not wanted: its table row stays empty
    __shared__ double s_b[DT_THREADS / 32][NP][5];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const double a = warp_sum(acc[q][0]), b = warp_sum(acc[q][1]), c = warp_sum(acc[q][2]), d = warp_sum(acc[q][3]);
        const double e = warp_max(acc[q][4]);
        if (lane == 0) {
            double* r = s_b[threadIdx.x >> 5][q];
            r[0] = a; r[1] = b; r[2] = c; r[3] = d; r[4] = e;
        }
    }
    __syncthreads();
    if (threadIdx.x < NP) {
        const int q = threadIdx.x;
        double a = 0, b = 0, c = 0, d = 0, e = -INFINITY;
        for (int w = 0; w < DT_THREADS / 32; ++w) {
            a += s_b[w][q][0]; b += s_b[w][q][1]; c += s_b[w][q][2]; d += s_b[w][q][3]; e = fmax(e, s_b[w][q][4]);
        }
        if (d > 0) {
            double* tab = ps[q]->table + (size_t)img * (ps[q]->maxLabels + 1) * 8;
            atomicAdd(tab + 0, a);
            atomicAdd(tab + 1, b);
            atomicAdd(tab + 2, c);
            atomicAdd(tab + 3, d);
            atomicMax(reinterpret_cast<unsigned long long*>(tab) + 4, enc_ordered(e));
        }
    }
    }
}

// label of pixel p at a threshold whose parent array has been ranked (0 = background)
template <typename T>
__device__ __forceinline__ int sparse_label_at(const T* __restrict__ m, const int* __restrict__ L, int p, double thr, double sign) {
    if (!(sign * (double)m[p] > thr)) return 0;
    const int v0 = L[p];
    return (v0 <= -2) ? (-v0 - 1) : (-L[v0] - 1);
}

template <typename T>
__global__ void __launch_bounds__(256) ring_flags_sparse_kernel(const T* __restrict__ map, double sign, const double* __restrict__ detTable,
                                                                const int* __restrict__ detN, int detCap, double ringThr,
                                                                const int* __restrict__ ringL, const double* __restrict__ ringTable,
                                                                int ringCap, int H, int W, double minRingPix, int posMode,
                                                                unsigned char* __restrict__ flags) {
    const int b = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o > detCap) return;
    unsigned char out = 0;
    if (o >= 1 && o <= detN[b]) {
        const double* t = detTable + ((size_t)b * (detCap + 1) + o) * 8;
        double y, x;
        if (posMode == 0) {
            y = t[1] / t[0];
            x = t[2] / t[0];
        } else {
            y = floor(t[5] / (double)W);
            x = t[5] - y * (double)W;
        }
        if (y == y && x == x) {
            const int yi = min(max((int)y, 0), H - 1), xi = min(max((int)x, 0), W - 1);
            const T* m = map + (size_t)b * H * W;
            const int* L = ringL + (size_t)b * H * W;
            const int r = sparse_label_at<T>(m, L, yi * W + xi, ringThr, sign);
            if (r > 0 && r <= ringCap) {
                const double* rt = ringTable + ((size_t)b * (ringCap + 1) + r) * 8;
                if (rt[3] > minRingPix) {
                    const double ry = rt[1] / rt[0], rx = rt[2] / rt[0];
                    const int ryi = min(max((int)ry, 0), H - 1), rxi = min(max((int)rx, 0), W - 1);
                    if (sparse_label_at<T>(m, L, ryi * W + rxi, ringThr, sign) != r) out = 1;
                }
            }
        }
    }
    flags[(size_t)b * (detCap + 1) + o] = out;
}

__global__ void detect_merge_status_kernel(int* __restrict__ status, const int* __restrict__ ringStatus, const int* __restrict__ ringN,
                                           int ringCap, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B && (ringStatus[i] != 0 || ringN[i] > ringCap)) status[i] = 1;
}

static void detect_scratch_carve(DetectScratch& sc, int*& cursor, int B, int H, int W) {
    const size_t HW = (size_t)H * W;
    sc.listCap = (int)detect_list_cap(H, W);
    sc.L = cursor;
    sc.fgList = sc.L + (size_t)B * HW;
    sc.rootList = sc.fgList + (size_t)B * sc.listCap;
    sc.fgCount = sc.rootList + (size_t)B * DT_ROOTCAP;
    sc.rootCount = sc.fgCount + B;
    cursor = sc.rootCount + B;
}

template <typename T>
static int find_objects_t(const T* map, double thr, double ringThr, int haveRing, double sign, double minRingPix, int posMode,
                          int* nLabels, int* status, double* table, unsigned char* ringFlags, int* labels, int maxLabels, int* scratch,
                          int B, int H, int W, int flags, cudaStream_t st) {
    const int HW = H * W;
    int* cursor = scratch;
    DetectPass p[2];
    detect_scratch_carve(p[0].sc, cursor, B, H, W);
    p[0].thr = thr;
    p[0].table = table;
    p[0].status = status;
    p[0].maxLabels = maxLabels;
    int *ringN = nullptr, *ringStatus = nullptr;
    const int ringCap = DT_ROOTCAP;
    if (haveRing) {
        detect_scratch_carve(p[1].sc, cursor, B, H, W);
        ringN = cursor;
        ringStatus = ringN + B;
        cursor = ringStatus + B;
        if ((reinterpret_cast<size_t>(cursor) & 7) != 0) ++cursor;   // the ring table holds doubles
        p[1].thr = ringThr;
        p[1].table = reinterpret_cast<double*>(cursor);
        p[1].status = ringStatus;
        p[1].maxLabels = ringCap;
    } else {
        p[1] = p[0];
    }
    const int NP = haveRing ? 2 : 1;
    for (int q = 0; q < NP; ++q) {
        const long long nRows = (long long)B * (p[q].maxLabels + 1);
        { NB_PROF_BEGIN(st); detect_reset_kernel<<<(B + 255) / 256, 256, 0, st>>>(p[q].sc.fgCount, p[q].sc.rootCount, p[q].status, B); NB_PROF_END("detect_reset_kernel", st); }
        NB_CHECK_LAUNCH();
        (void)nRows;
        { NB_PROF_BEGIN(st); moments_init_row0_kernel<<<(B + 255) / 256, 256, 0, st>>>(p[q].table, p[q].maxLabels, B); NB_PROF_END("moments_init_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    dim3 gridDense((HW + DT_THREADS * DT_ITEMS - 1) / (DT_THREADS * DT_ITEMS), B);
    if (haveRing) {
        { NB_PROF_BEGIN(st); detect_scan_multi_kernel<T, 2, false><<<gridDense, DT_THREADS, 0, st>>>(map, sign, p[0], p[1], labels, H, W); NB_PROF_END("detect_scan_multi_kernel", st); }
    } else {
        { NB_PROF_BEGIN(st); detect_scan_multi_kernel<T, 1, false><<<gridDense, DT_THREADS, 0, st>>>(map, sign, p[0], p[1], labels, H, W); NB_PROF_END("detect_scan_multi_kernel", st); }
    }
    NB_CHECK_LAUNCH();
    dim3 gridSparse(64, B);
    for (int q = 0; q < NP; ++q) {
        int* nl = (q == 0) ? nLabels : ringN;
        { NB_PROF_BEGIN(st); detect_link_kernel<T><<<gridSparse, 256, 0, st>>>(map, p[q].thr, sign, p[q].sc, H, W); NB_PROF_END("detect_link_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); detect_flatten_kernel<<<gridSparse, 256, 0, st>>>(p[q].sc, p[q].status, HW); NB_PROF_END("detect_flatten_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); detect_rank_kernel<<<B, 1024, 0, st>>>(p[q].sc, nl, HW); NB_PROF_END("detect_rank_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); moments_rows_kernel<false><<<dim3((p[q].maxLabels + 255) / 256, B), 256, 0, st>>>(p[q].table, p[q].maxLabels, nl); NB_PROF_END("moments_init_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); detect_moments_kernel<T><<<gridSparse, 256, 0, st>>>(map, sign, p[q].sc, (q == 0) ? labels : nullptr, p[q].table, H, W, p[q].maxLabels); NB_PROF_END("detect_moments_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    if (flags & 1) {
        { NB_PROF_BEGIN(st); detect_argmax_kernel<T><<<gridSparse, 256, 0, st>>>(map, sign, p[0].sc, table, HW, maxLabels); NB_PROF_END("detect_argmax_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    for (int q = 0; q < NP; ++q) {
        { NB_PROF_BEGIN(st); moments_rows_kernel<true><<<dim3((p[q].maxLabels + 1 + 255) / 256, B), 256, 0, st>>>(p[q].table, p[q].maxLabels, (q == 0) ? nLabels : ringN); NB_PROF_END("moments_final_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    if (haveRing) {
        { NB_PROF_BEGIN(st); ring_flags_sparse_kernel<T><<<dim3((maxLabels + 1 + 255) / 256, B), 256, 0, st>>>(map, sign, table, nLabels, maxLabels, ringThr, p[1].sc.L, p[1].table, ringCap, H, W, minRingPix, posMode, ringFlags); NB_PROF_END("ring_flags_sparse_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); detect_merge_status_kernel<<<(B + 255) / 256, 256, 0, st>>>(status, ringStatus, ringN, ringCap, B); NB_PROF_END("detect_merge_status_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    return NB_OK;
}

// Ring removal (photometry.py:60-70, 104-105) on the device: a candidate is dropped when the ring-threshold component under its
// position has more than minRingPix pixels and does not contain its own centre of mass.  One thread per candidate label.
__global__ void __launch_bounds__(256) ring_flags_kernel(const double* __restrict__ detTable, const int* __restrict__ detN, int detCap,
                                                         const int* __restrict__ ringLabels, const double* __restrict__ ringTable,
                                                         int ringCap, int H, int W, double minRingPix, int posMode,
                                                         unsigned char* __restrict__ flags) {
    const int b = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o > detCap) return;
    unsigned char out = 0;
    if (o >= 1 && o <= detN[b]) {
        const double* t = detTable + ((size_t)b * (detCap + 1) + o) * 8;
        double y, x;
        if (posMode == 0) {
            y = t[1] / t[0];
            x = t[2] / t[0];
        } else {
            y = floor(t[5] / (double)W);
            x = t[5] - y * (double)W;
        }
        if (y == y && x == x) {
            const int yi = min(max((int)y, 0), H - 1), xi = min(max((int)x, 0), W - 1);
            const int* rl = ringLabels + (size_t)b * H * W;
            const int r = rl[(size_t)yi * W + xi];
            if (r > 0 && r <= ringCap) {
                const double* rt = ringTable + ((size_t)b * (ringCap + 1) + r) * 8;
                if (rt[3] > minRingPix) {
                    const double ry = rt[1] / rt[0], rx = rt[2] / rt[0];
                    const int ryi = min(max((int)ry, 0), H - 1), rxi = min(max((int)rx, 0), W - 1);
                    if (rl[(size_t)ryi * W + rxi] != r) out = 1;
                }
            }
        }
    }
    flags[(size_t)b * (detCap + 1) + o] = out;
}

extern "C" {

int nb_ring_flags(const double* detTable, const int* detN, int detCap, const int* ringLabels, const double* ringTable, int ringCap,
                  int B, int H, int W, double minRingPix, int posMode, unsigned char* flags, cudaStream_t st) {
    NB_REQUIRE(detTable && detN && ringLabels && ringTable && flags && B > 0 && detCap >= 0 && ringCap >= 0, NB_ERR_ARG,
               "nb_ring_flags: bad arguments");
    { NB_PROF_BEGIN(st); ring_flags_kernel<<<dim3((detCap + 1 + 255) / 256, B), 256, 0, st>>>(detTable, detN, detCap, ringLabels, ringTable, ringCap, H, W, minRingPix, posMode, flags); NB_PROF_END("ring_flags_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

long long nb_find_objects_scratch_ints(int B, int H, int W) {
    const long long HW = (long long)H * W;
    const long long perPass = (long long)B * (HW + detect_list_cap(H, W) + DT_ROOTCAP + 2);
    return 2 * perPass + 2LL * B + 2 + 2LL * B * (DT_ROOTCAP + 1) * 8;
}

int nb_find_objects(int prec, const void* map, double threshold, double ringThreshold, int haveRing, double sign, double minRingPix,
                    int posMode, int* nLabels, int* status, double* table, unsigned char* ringFlags, int* labels, int maxLabels,
                    int* scratch, int B, int H, int W, int flags, cudaStream_t st) {
    NB_REQUIRE(map && nLabels && status && table && scratch && B > 0 && H > 0 && W > 0 && maxLabels >= 0, NB_ERR_ARG,
               "nb_find_objects: bad arguments");
    NB_REQUIRE(!haveRing || ringFlags, NB_ERR_ARG, "nb_find_objects: ringFlags required with a ring threshold");
    NB_REQUIRE((long long)H * W < 0x7fffffffLL, NB_ERR_UNSUPPORTED, "nb_find_objects: image too large");
    if (prec == NB_PREC_F64)
        return find_objects_t<double>((const double*)map, threshold, ringThreshold, haveRing, sign, minRingPix, posMode, nLabels, status,
                                      table, ringFlags, labels, maxLabels, scratch, B, H, W, flags, st);
    if (prec == NB_PREC_F32)
        return find_objects_t<float>((const float*)map, threshold, ringThreshold, haveRing, sign, minRingPix, posMode, nLabels, status,
                                     table, ringFlags, labels, maxLabels, scratch, B, H, W, flags, st);
    nb_set_error("nb_find_objects: bad precision %d", prec);
    return NB_ERR_ARG;
}

long long nb_detect_scratch_ints(int B, int H, int W) {
    const long long HW = (long long)H * W;
    return (long long)B * (HW + detect_list_cap(H, W) + DT_ROOTCAP + 2);
}

int nb_detect(int prec, const void* map, double threshold, double sign, int* labels, int* nLabels, int* status, double* table,
              int maxLabels, int* scratch, int B, int H, int W, int flags, cudaStream_t st) {
    NB_REQUIRE(map && labels && nLabels && status && table && scratch && B > 0 && H > 0 && W > 0 && maxLabels >= 0, NB_ERR_ARG,
               "nb_detect: bad arguments");
    NB_REQUIRE((long long)H * W < 0x7fffffffLL, NB_ERR_UNSUPPORTED, "nb_detect: image too large");
    if (prec == NB_PREC_F64)
        return detect_t<double>((const double*)map, threshold, sign, labels, nLabels, status, table, maxLabels, scratch, B, H, W, flags, st);
    if (prec == NB_PREC_F32)
        return detect_t<float>((const float*)map, threshold, sign, labels, nLabels, status, table, maxLabels, scratch, B, H, W, flags, st);
    nb_set_error("nb_detect: bad precision %d", prec);
    return NB_ERR_ARG;
}

}  // extern "C"
