// Real-space filter kernels (K12, K13 of SURVEY.md section 2a) and small cast utilities.
//   nb_gauss_filter : ndimage.gaussian_filter / maps.subtractBackground       maps.py:1355-1368, 1433-1458
//   nb_conv2d       : ndimage.convolve with the truncated matched-filter kernel, summed over frequency
//                                                                             filters.py:1200-1216
// Accumulation is fp64; storage follows prec.
#include "nemo_b200_internal.h"
#include "gauss.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------------
// K12 row pass: one CTA per row, the row staged in shared memory as fp64
template <typename T, bool WIDE>
__global__ void __launch_bounds__(256) gauss_rows_kernel(const T* in, T* out, int W, GaussW g, double sigma, int rWide) {
    extern __shared__ double s_row[];   // [W] (+ [2 r + 1] weights when WIDE)
    const int r = WIDE ? rWide : g.r;
    const double* wts = g.w;
    if (WIDE) {
        gauss_weights_smem(s_row + W, sigma, r);
        wts = s_row + W;
    }
    const size_t base = (size_t)blockIdx.x * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) s_row[x] = (double)in[base + x];
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        double acc = 0;
        if (x >= r && x + r < W) {
            for (int d = -r; d <= r; ++d) acc += wts[d + r] * s_row[x + d];
        } else {
            for (int d = -r; d <= r; ++d) acc += wts[d + r] * s_row[reflect_idx(x + d, W)];
        }
        out[base + x] = (T)acc;
    }
}

// K12 column pass: 32 columns x 64 rows per CTA with a reflected halo; out = subtract ? ref - smooth : smooth
#define GC_TX 32
#define GC_TY 64
// TX columns per CTA: 32, or fewer when a wide halo would not fit shared memory otherwise
template <typename T, int TX, bool WIDE>
__global__ void __launch_bounds__(256) gauss_cols_kernel(const T* __restrict__ in, const T* ref, T* out, int H, int W, GaussW g,
                                                         int subtract, double sigma, int rWide) {
    extern __shared__ double s_t[];     // [(GC_TY + 2 r)][TX] (+ [2 r + 1] weights when WIDE)
    const int r = WIDE ? rWide : g.r;
    const int rows = GC_TY + 2 * r;
    const double* wts = g.w;
    if (WIDE) {
        gauss_weights_smem(s_t + (size_t)rows * TX, sigma, r);
        wts = s_t + (size_t)rows * TX;
    }
    constexpr int TYS = 256 / TX;       // thread rows
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * GC_TY;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int x = x0 + tx;
    for (int i = ty; i < rows; i += TYS) {
        const int y = reflect_idx(y0 - r + i, H);
        s_t[i * TX + tx] = (x < W) ? (double)in[plane + (size_t)y * W + x] : 0.0;
    }
    __syncthreads();
    if (x >= W) return;
    for (int i = ty; i < GC_TY; i += TYS) {
        const int y = y0 + i;
        if (y >= H) break;
        double acc = 0;
        for (int d = 0; d <= 2 * r; ++d) acc += wts[d] * s_t[(i + d) * TX + tx];
        const size_t o = plane + (size_t)y * W + x;
        // scipy rounds the smoothed map to the array dtype before maps.subtractBackground subtracts it
        out[o] = subtract ? (T)(ref[o] - (T)acc) : (T)acc;
    }
}

// ------------------------------------------------------------------------------------------------------
// K12, register-blocked (radius <= NB_GAUSS_MAXR: the weights are a kernel argument, i.e. constant-bank operands).  The plain
// kernels above read one input and one weight from shared memory per multiply-add; here a thread makes 16 consecutive outputs
// along the filtered axis and, per chunk of 8 taps, reads the 23 inputs they touch once: 23 shared-memory reads per 128 DFMA.
// Same products in the same order (tap 0 .. 2 r), so the results are bit-identical to the plain kernels'.
#define GRB_OUT 16
#define GRB_CH 8

// (the tile in shared memory is padded by CH - 1 entries past the halo: the last chunk reads them, only taps past 2 r would use them)
template <int R_OUT, int CH, typename LoadV>
__device__ __forceinline__ void gauss_rb_core(double (&acc)[R_OUT], const GaussW& g, int r, LoadV loadv) {
    for (int d0 = 0; d0 <= 2 * r; d0 += CH) {
        double v[R_OUT + CH - 1];
#pragma unroll
        for (int q = 0; q < R_OUT + CH - 1; ++q) v[q] = loadv(d0 + q);
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            if (d0 + u <= 2 * r) {
                const double w = g.w[d0 + u];
#pragma unroll
                for (int j = 0; j < R_OUT; ++j) acc[j] += w * v[u + j];
            }
        }
    }
}

// column pass: 32 columns x 128 rows per CTA (8 thread rows x 16 outputs), reflected halo of r rows above and below
#define GCB_TX 32
#define GCB_TY 128
template <typename T>
__global__ void __launch_bounds__(256, 2) gauss_cols_rb_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, GaussW g) {
    extern __shared__ double s_t[];     // [GCB_TY + 2 r + GRB_CH - 1][GCB_TX]
    const int r = g.r;
    const int rows = GCB_TY + 2 * r;
    const int x0 = blockIdx.x * GCB_TX, y0 = blockIdx.y * GCB_TY;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int tx = threadIdx.x % GCB_TX, tg = threadIdx.x / GCB_TX;
    const int x = x0 + tx;
    for (int i = tg; i < rows; i += 256 / GCB_TX) {
        const int y = reflect_idx(y0 - r + i, H);
        s_t[i * GCB_TX + tx] = (x < W) ? (double)in[plane + (size_t)y * W + x] : 0.0;
    }
    __syncthreads();
    if (x >= W) return;
    double acc[GRB_OUT];
#pragma unroll
    for (int j = 0; j < GRB_OUT; ++j) acc[j] = 0;
    const double* base = s_t + (size_t)(tg * GRB_OUT) * GCB_TX + tx;
    gauss_rb_core<GRB_OUT, GRB_CH>(acc, g, r, [&](int i) { return base[i * GCB_TX]; });
#pragma unroll
    for (int j = 0; j < GRB_OUT; ++j) {
        const int y = y0 + tg * GRB_OUT + j;
        if (y < H) out[plane + (size_t)y * W + x] = (T)acc[j];
    }
}

// row pass: 32 rows x 128 columns per CTA; a warp's lanes are the 32 ROWS (odd row pitch: conflict-free), 8 warps x 16 outputs
// along x.  The results go back through shared memory so that the stores are coalesced; out = subtract ? ref - smooth : smooth.
#define GRR_TY 32
#define GRR_TX 128
template <typename T>
__global__ void __launch_bounds__(256, 2) gauss_rows_rb_kernel(const T* __restrict__ in, const T* __restrict__ ref, T* __restrict__ out,
                                                              int H, int W, GaussW g, int subtract) {
    extern __shared__ double s_t[];     // [GRR_TY][LD], LD = GRR_TX + 2 r + GRB_CH - 1 rounded up to odd
    const int r = g.r;
    const int cols = GRR_TX + 2 * r;
    const int LD = (cols + GRB_CH - 1) | 1;
    const int x0 = blockIdx.x * GRR_TX, y0 = blockIdx.y * GRR_TY;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    for (int rr = warp; rr < GRR_TY; rr += 8) {
        const int y = y0 + rr;
        const T* src = in + plane + (size_t)(y < H ? y : H - 1) * W;
        for (int c = lane; c < cols; c += 32) s_t[rr * LD + c] = (double)src[reflect_idx(x0 - r + c, W)];
    }
    __syncthreads();
    double acc[GRB_OUT];
#pragma unroll
    for (int j = 0; j < GRB_OUT; ++j) acc[j] = 0;
    const double* base = s_t + (size_t)lane * LD + warp * GRB_OUT;
    gauss_rb_core<GRB_OUT, GRB_CH>(acc, g, r, [&](int i) { return base[i]; });
    __syncthreads();
#pragma unroll
    for (int j = 0; j < GRB_OUT; ++j) s_t[lane * LD + warp * GRB_OUT + j] = acc[j];
    __syncthreads();
    for (int rr = warp; rr < GRR_TY; rr += 8) {
        const int y = y0 + rr;
        if (y >= H) break;
        for (int c = lane; c < GRR_TX; c += 32) {
            const int x = x0 + c;
            if (x >= W) break;
            const size_t o = plane + (size_t)y * W + x;
            const T sm = (T)s_t[rr * LD + c];
            // scipy rounds the smoothed map to the array dtype before maps.subtractBackground subtracts it
            out[o] = subtract ? (T)(ref[o] - sm) : sm;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K13.  32 x 32 outputs per CTA; input tile with a (kh-1, kw-1) reflected halo and the kernel in shared memory.
// ndimage.convolve: out[y, x] = sum_{a, b} kern[a, b] * in[y + kh/2 - a, x + kw/2 - b]   (odd kh, kw)
#define CV_T 32
template <typename T>
__global__ void __launch_bounds__(256) conv2d_kernel(const T* __restrict__ in, const double* __restrict__ kern, T* __restrict__ out,
                                                     T* __restrict__ planes, int F, int H, int W, int kh, int kw, double scale,
                                                     const double* __restrict__ norm, const unsigned char* __restrict__ mask) {
    extern __shared__ double s_mem[];
    const int th = CV_T + kh - 1, tw = CV_T + kw - 1;
    double* s_in = s_mem;             // [th][tw]
    double* s_k = s_mem + th * tw;    // [kh][kw]
    const int x0 = blockIdx.x * CV_T, y0 = blockIdx.y * CV_T, b = blockIdx.z;
    const int tx = threadIdx.x % CV_T, ty = threadIdx.x / CV_T;  // 32 x 8, four rows per thread
    const int ry = kh / 2, rx = kw / 2;
    double acc[4] = {0, 0, 0, 0};
    for (int f = 0; f < F; ++f) {
        __syncthreads();
        const T* src = in + ((size_t)b * F + f) * H * W;
        for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
            const int r = i / tw, c = i - r * tw;
            s_in[i] = (double)src[(size_t)reflect_idx(y0 - ry + r, H) * W + reflect_idx(x0 - rx + c, W)];
        }
        for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) s_k[i] = kern[(size_t)f * kh * kw + i];
        __syncthreads();
        double pa[4] = {0, 0, 0, 0};
        for (int a = 0; a < kh; ++a) {
            for (int c = 0; c < kw; ++c) {
                const double kv = s_k[a * kw + c];
                // input row (y + ry - a) sits at tile row (ty_out + 2*ry - a); column likewise
                const int col = tx + 2 * rx - c;
#pragma unroll
                for (int j = 0; j < 4; ++j) pa[j] += kv * s_in[(ty + 8 * j + 2 * ry - a) * tw + col];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = y0 + ty + 8 * j, x = x0 + tx;
            if (planes && y < H && x < W) planes[(((size_t)b * F + f) * H + y) * W + x] = (T)pa[j];
            acc[j] += (double)(T)pa[j];  // each plane is an array of the map dtype before the sum over frequency
        }
    }
    double fac = scale;
    if (norm) fac *= norm[b];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + ty + 8 * j, x = x0 + tx;
        if (y < H && x < W) {
            double v = acc[j] * fac;
            if (mask) v *= (double)mask[((size_t)b * H + y) * W + x];
            out[((size_t)b * H + y) * W + x] = (T)v;
        }
    }
}


// ------------------------------------------------------------------------------------------------------
// K13, register-blocked and batched (nb_conv2d_batch): a CTA makes a 32 x 64 output tile, a thread 16 consecutive rows of one
// column.  For a kernel column c the thread reads, per chunk of 8 kernel rows, the 23 input values those taps touch into
// registers once and does 128 multiply-adds on them: 31 shared-memory reads per 128 DFMA (the plain kernel above needs 5 per 4,
// which leaves the fp64 pipe waiting for the shared-memory pipe; with 8 rows per thread, 23 per 64, the two pipes were still
// level: 7.4 TFMA/s).  Output image b reads input image b / inDiv (inDiv = S: the
// S filters of a tile) or b % inMod (inMod = S: templates shared by all tiles), kernel b (kernPerImage) or the common one, and
// mask b / maskDiv.
#define CVB_TX 32
#define CVB_TY 64
#define CVB_RY 16     // outputs per thread
#define CVB_CH 8      // kernel rows per chunk
#define CVB_THREADS (CVB_TX * CVB_TY / CVB_RY)
// KWMAX (15, 31 or 63) fixes the row pitch of the input tile and of the kernel taps in shared memory at compile time, so that
// the 23 + 8 shared-memory reads of a chunk are immediate offsets from one base register; the tile carries CVB_CH - 1 rows of
// padding above (rows that only taps beyond the kernel would read), so no index is ever clamped.  Measured before this: 64 % of
// the instructions of the kernel were not DFMA (fp64 pipe 40 % busy at 56 % issue utilisation).
template <int KWMAX> struct ConvPitch {
    static constexpr int IN = (CVB_TX + KWMAX - 1 + 1) & ~1;     // doubles per tile row
    static constexpr int K = KWMAX + 1;                          // doubles per kernel row
};
template <typename T, int KWMAX>
__global__ void __launch_bounds__(CVB_THREADS, 3) conv2d_rb_kernel(const T* __restrict__ in, const double* __restrict__ kern, T* __restrict__ out,
                                                           T* __restrict__ planes, int F, int H, int W, int kh, int kw, double scale,
                                                           const double* __restrict__ norm, const unsigned char* __restrict__ mask,
                                                           int inDiv, int inMod, int kernPerImage, int maskDiv) {
    extern __shared__ double s_mem[];
    constexpr int PI = ConvPitch<KWMAX>::IN, PK = ConvPitch<KWMAX>::K;
    constexpr int PAD = CVB_CH - 1;
    const int th = CVB_TY + kh - 1, tw = CVB_TX + kw - 1;
    const int khPad = (kh + CVB_CH - 1) / CVB_CH * CVB_CH;
    double* s_in = s_mem;                        // [PAD + th][PI]
    double* s_k = s_mem + (PAD + th) * PI;       // [khPad][PK], rows kh .. khPad - 1 zero
    const int x0 = blockIdx.x * CVB_TX, y0 = blockIdx.y * CVB_TY, b = blockIdx.z;
    const int tx = threadIdx.x % CVB_TX, tg = threadIdx.x / CVB_TX;      // 32 columns x 4 row groups
    const int ry = kh / 2, rx = kw / 2;
    const int bi = inMod > 0 ? b % inMod : b / inDiv;
    const double* kb = kern + (kernPerImage ? (size_t)b * F * kh * kw : (size_t)0);
    double acc[CVB_RY];
#pragma unroll
    for (int j = 0; j < CVB_RY; ++j) acc[j] = 0;
    for (int i = threadIdx.x; i < PAD * PI; i += blockDim.x) s_in[i] = 0.0;
    for (int f = 0; f < F; ++f) {
        __syncthreads();
        const T* src = in + ((size_t)bi * F + f) * H * W;
        for (int r = threadIdx.x / 32; r < th; r += CVB_THREADS / 32) {
            const size_t rowOff = (size_t)reflect_idx(y0 - ry + r, H) * W;
            for (int c = threadIdx.x % 32; c < tw; c += 32) s_in[(PAD + r) * PI + c] = (double)src[rowOff + reflect_idx(x0 - rx + c, W)];
        }
        for (int i = threadIdx.x; i < khPad * PK; i += blockDim.x) {
            const int a = i / PK, c = i - a * PK;
            s_k[i] = (a < kh && c < kw) ? kb[(size_t)f * kh * kw + a * kw + c] : 0.0;
        }
        __syncthreads();
        double pa[CVB_RY];
#pragma unroll
        for (int j = 0; j < CVB_RY; ++j) pa[j] = 0;
        // ndimage.convolve: out[y, x] = sum_{a, c} kern[a, c] in[y + ry - a, x + rx - c]; tile row of output j of this thread
        // for kernel row a: PAD + tg * 16 + j + 2 ry - a
        const double* colBase = s_in + (tg * CVB_RY + 2 * ry) * PI + tx + 2 * rx;
        for (int c = 0; c < kw; ++c) {
            const double* colp = colBase - c;
            const double* kp = s_k + c;
            for (int a0 = 0; a0 < kh; a0 += CVB_CH) {
                // tile rows (PAD +) tg * 16 + 2 ry - a0 - 7 .. + 15
                double v[CVB_RY + CVB_CH - 1];
                const double* vp = colp - a0 * PI;
#pragma unroll
                for (int q = 0; q < CVB_RY + CVB_CH - 1; ++q) v[q] = vp[q * PI];
                const double* kq = kp + a0 * PK;
#pragma unroll
                for (int u = 0; u < CVB_CH; ++u) {
                    if (a0 + u < kh) {
                        const double kv = kq[u * PK];
#pragma unroll
                        for (int j = 0; j < CVB_RY; ++j) pa[j] += kv * v[CVB_CH - 1 - u + j];
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CVB_RY; ++j) {
            const int y = y0 + tg * CVB_RY + j, x = x0 + tx;
            if (planes && y < H && x < W) planes[(((size_t)b * F + f) * H + y) * W + x] = (T)pa[j];
            acc[j] += (double)(T)pa[j];  // each plane is an array of the map dtype before the sum over frequency
        }
    }
    double fac = scale;
    if (norm) fac *= norm[b];
#pragma unroll
    for (int j = 0; j < CVB_RY; ++j) {
        const int y = y0 + tg * CVB_RY + j, x = x0 + tx;
        if (y < H && x < W) {
            double v = acc[j] * fac;
            if (mask) v *= (double)mask[((size_t)(b / maskDiv) * H + y) * W + x];
            out[((size_t)b * H + y) * W + x] = (T)v;
        }
    }
}

// mean over the rows of every [H, Wc] plane (the ky average of a folded filter: row 0 of its inverse transform, which is all
// that MapFilter.makeRealSpaceFilterProfile looks at, filters.py:282-304)
template <typename T>
__global__ void __launch_bounds__(256) mean_rows_kernel(const T* __restrict__ in, double* __restrict__ out, int H, int Wc) {
    const int kx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (kx >= Wc) return;
    const T* p = in + (size_t)b * H * Wc + kx;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int y = 0;
    for (; y + 3 < H; y += 4) {
        s0 += (double)p[(size_t)y * Wc];
        s1 += (double)p[(size_t)(y + 1) * Wc];
        s2 += (double)p[(size_t)(y + 2) * Wc];
        s3 += (double)p[(size_t)(y + 3) * Wc];
    }
    for (; y < H; ++y) s0 += (double)p[(size_t)y * Wc];
    out[(size_t)b * Wc + kx] = ((s0 + s1) + (s2 + s3)) / (double)H;
}

// ------------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (TO)in[i];
}
template <typename TI>
__global__ void __launch_bounds__(256) mask_u8_kernel(const TI* __restrict__ in, unsigned char* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (in[i] != (TI)0) ? 1 : 0;
}

template <typename T>
__global__ void __launch_bounds__(256) sub_kernel(const T* a, const T* b, T* o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] - b[i];
}



template <typename T, bool WIDE>
static int gauss_cols_launch(const T* in, T* out, int B, int H, int W, const GaussW& g, double sigma, int r, cudaStream_t st) {
    const size_t extra = WIDE ? sizeof(double) * (2 * (size_t)r + 1) : 0;
    const int rows = GC_TY + 2 * r;
#define GC_LAUNCH(TX)                                                                                                          \
    {                                                                                                                          \
        const size_t smem = (size_t)rows * TX * sizeof(double) + extra;                                                        \
        if (smem <= 220 * 1024) {                                                                                              \
            if (smem > 48 * 1024)                                                                                              \
                NB_CHECK_CUDA(cudaFuncSetAttribute(gauss_cols_kernel<T, TX, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            dim3 grid((W + TX - 1) / TX, (H + GC_TY - 1) / GC_TY, B);                                                          \
            { NB_PROF_BEGIN(st); gauss_cols_kernel<T, TX, WIDE><<<grid, 256, smem, st>>>(in, nullptr, out, H, W, g, 0, sigma, r); NB_PROF_END("gauss_cols_kernel", st); } \
            NB_CHECK_LAUNCH();                                                                                                 \
            return NB_OK;                                                                                                      \
        }                                                                                                                      \
    }
    GC_LAUNCH(32)
    GC_LAUNCH(16)
    GC_LAUNCH(8)
#undef GC_LAUNCH
    nb_set_error("nb_gauss_filter: smoothing radius %d too large", r);
    return NB_ERR_UNSUPPORTED;
}

template <typename T, bool WIDE>
static int gauss_rows_launch(const T* in, T* out, int B, int H, int W, const GaussW& g, double sigma, int r, cudaStream_t st) {
    const size_t smem = sizeof(double) * ((size_t)W + (WIDE ? 2 * (size_t)r + 1 : 0));
    if (smem > 48 * 1024)
        NB_CHECK_CUDA(cudaFuncSetAttribute(gauss_rows_kernel<T, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { NB_PROF_BEGIN(st); gauss_rows_kernel<T, WIDE><<<(unsigned)((size_t)B * H), 256, smem, st>>>(in, out, W, g, sigma, r); NB_PROF_END("gauss_rows_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T>
static int gauss_filter_t(const T* in, T* tmp, T* out, int B, int H, int W, double sigmaY, double sigmaX, int subtract, cudaStream_t st) {
    // scipy filters axis 0 (y) first, then axis 1 (x); each pass rounds to the array dtype
    const int ry = gauss_radius(sigmaY), rx = gauss_radius(sigmaX);
    NB_REQUIRE(ry <= 4096 && rx <= 4096, NB_ERR_UNSUPPORTED, "nb_gauss_filter: smoothing scale too large");
    GaussW gy, gx;
    gy.r = gx.r = 0;
    int rc;
    // both radii within the weights a kernel argument holds, tiles within shared memory: the register-blocked passes, with the
    // subtraction fused into the row pass
    const size_t smemC = sizeof(double) * (size_t)(GCB_TY + 2 * ry + GRB_CH - 1) * GCB_TX;
    const size_t smemR = sizeof(double) * (size_t)GRR_TY * (size_t)((GRR_TX + 2 * rx + GRB_CH - 1) | 1);
    if (ry <= NB_GAUSS_MAXR && rx <= NB_GAUSS_MAXR && ry >= 1 && rx >= 1 && smemC <= 110 * 1024 && smemR <= 110 * 1024 && out != in &&
        out != tmp && !getenv("NEMO_B200_GAUSS_PLAIN")) {
        if ((rc = make_gauss(sigmaY, &gy)) != NB_OK) return rc;
        if ((rc = make_gauss(sigmaX, &gx)) != NB_OK) return rc;
        if (smemC > 48 * 1024)
            NB_CHECK_CUDA(cudaFuncSetAttribute(gauss_cols_rb_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC));
        if (smemR > 48 * 1024)
            NB_CHECK_CUDA(cudaFuncSetAttribute(gauss_rows_rb_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemR));
        { NB_PROF_BEGIN(st); gauss_cols_rb_kernel<T><<<dim3((W + GCB_TX - 1) / GCB_TX, (H + GCB_TY - 1) / GCB_TY, B), 256, smemC, st>>>(in, tmp, H, W, gy); NB_PROF_END("gauss_cols_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); gauss_rows_rb_kernel<T><<<dim3((W + GRR_TX - 1) / GRR_TX, (H + GRR_TY - 1) / GRR_TY, B), 256, smemR, st>>>(tmp, in, out, H, W, gx, subtract); NB_PROF_END("gauss_rows_kernel", st); }
        NB_CHECK_LAUNCH();
        return NB_OK;
    }
    if (ry <= NB_GAUSS_MAXR) {
        if ((rc = make_gauss(sigmaY, &gy)) != NB_OK) return rc;
        rc = gauss_cols_launch<T, false>(in, tmp, B, H, W, gy, sigmaY, ry, st);
    } else {
        rc = gauss_cols_launch<T, true>(in, tmp, B, H, W, gy, sigmaY, ry, st);
    }
    if (rc != NB_OK) return rc;
    T* rowsOut = subtract ? tmp : out;
    if (rx <= NB_GAUSS_MAXR) {
        if ((rc = make_gauss(sigmaX, &gx)) != NB_OK) return rc;
        rc = gauss_rows_launch<T, false>(tmp, rowsOut, B, H, W, gx, sigmaX, rx, st);
    } else {
        rc = gauss_rows_launch<T, true>(tmp, rowsOut, B, H, W, gx, sigmaX, rx, st);
    }
    if (rc != NB_OK) return rc;
    if (subtract) {
        const long long n = (long long)B * H * W;
        { NB_PROF_BEGIN(st); sub_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, (const T*)tmp, out, n); NB_PROF_END("sub_kernel", st); }
        NB_CHECK_LAUNCH();
    }
    return NB_OK;
}

extern "C" {

int nb_gauss_filter(int prec, const void* in, void* tmp, void* out, int B, int H, int W, double sigmaY, double sigmaX,
                    int subtract, cudaStream_t st) {
    NB_REQUIRE(in && tmp && out && B > 0 && sigmaY > 0 && sigmaX > 0, NB_ERR_ARG, "nb_gauss_filter: bad arguments");
    NB_REQUIRE(W <= 16384, NB_ERR_UNSUPPORTED, "nb_gauss_filter: tile too wide");
    if (prec == NB_PREC_F64) return gauss_filter_t<double>((const double*)in, (double*)tmp, (double*)out, B, H, W, sigmaY, sigmaX, subtract, st);
    if (prec == NB_PREC_F32) return gauss_filter_t<float>((const float*)in, (float*)tmp, (float*)out, B, H, W, sigmaY, sigmaX, subtract, st);
    nb_set_error("nb_gauss_filter: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_conv2d(int prec, const void* in, const double* kern, void* out, void* planes, int B, int F, int H, int W, int kh,
              int kw, double scale, const double* norm, const unsigned char* mask, cudaStream_t st) {
    NB_REQUIRE(in && kern && out && B > 0 && F > 0, NB_ERR_ARG, "nb_conv2d: bad arguments");
    NB_REQUIRE((kh & 1) && (kw & 1) && kh <= 63 && kw <= 63, NB_ERR_UNSUPPORTED, "nb_conv2d: kernel %dx%d not supported (odd sides <= 63)", kh, kw);
    const size_t smem = sizeof(double) * ((size_t)(CV_T + kh - 1) * (CV_T + kw - 1) + (size_t)kh * kw);
    dim3 grid((W + CV_T - 1) / CV_T, (H + CV_T - 1) / CV_T, B);
#define CONV(T)                                                                                                           \
    {                                                                                                                     \
        if (smem > 48 * 1024)                                                                                             \
            NB_CHECK_CUDA(cudaFuncSetAttribute(conv2d_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        { NB_PROF_BEGIN(st); conv2d_kernel<T><<<grid, 256, smem, st>>>((const T*)in, kern, (T*)out, (T*)planes, F, H, W, kh, kw, scale, norm, mask); NB_PROF_END("conv2d_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    if (prec == NB_PREC_F64) CONV(double)
    if (prec == NB_PREC_F32) CONV(float)
#undef CONV
    nb_set_error("nb_conv2d: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_conv2d_batch(int prec, const void* in, const double* kern, void* out, void* planes, int B, int F, int H, int W, int kh,
                    int kw, double scale, const double* norm, const unsigned char* mask, int inDiv, int inMod, int kernPerImage,
                    int maskDiv, cudaStream_t st) {
    NB_REQUIRE(in && kern && out && B > 0 && F > 0, NB_ERR_ARG, "nb_conv2d_batch: bad arguments");
    NB_REQUIRE((kh & 1) && (kw & 1) && kh <= 63 && kw <= 63, NB_ERR_UNSUPPORTED, "nb_conv2d_batch: kernel %dx%d not supported (odd sides <= 63)", kh, kw);
    NB_REQUIRE(kh / 2 < H && kw / 2 < W, NB_ERR_UNSUPPORTED, "nb_conv2d_batch: kernel larger than the map");
    if (inDiv < 1) inDiv = 1;
    if (maskDiv < 1) maskDiv = 1;
    dim3 grid((W + CVB_TX - 1) / CVB_TX, (H + CVB_TY - 1) / CVB_TY, B);
    const int khPad = (kh + CVB_CH - 1) / CVB_CH * CVB_CH;
#define CONVB(T, KWMAX)                                                                                                   \
    {                                                                                                                     \
        const size_t smem = sizeof(double) * ((size_t)(CVB_CH - 1 + CVB_TY + kh - 1) * ConvPitch<KWMAX>::IN + (size_t)khPad * ConvPitch<KWMAX>::K); \
        if (smem > 48 * 1024)                                                                                             \
            NB_CHECK_CUDA(cudaFuncSetAttribute(conv2d_rb_kernel<T, KWMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        { NB_PROF_BEGIN(st); conv2d_rb_kernel<T, KWMAX><<<grid, CVB_THREADS, smem, st>>>((const T*)in, kern, (T*)out, (T*)planes, F, H, W, kh, kw, scale, norm, mask, inDiv, inMod, kernPerImage, maskDiv); NB_PROF_END("conv2d_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
#define CONVB_K(T)                      \
    {                                   \
        if (kw <= 15) CONVB(T, 15)      \
        if (kw <= 31) CONVB(T, 31)      \
        CONVB(T, 63)                    \
    }
    if (prec == NB_PREC_F64) CONVB_K(double)
    if (prec == NB_PREC_F32) CONVB_K(float)
#undef CONVB_K
#undef CONVB
    nb_set_error("nb_conv2d_batch: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_mean_over_rows(int prec, const void* in, double* out, int B, int H, int Wc, cudaStream_t st) {
    NB_REQUIRE(in && out && B > 0 && H > 0 && Wc > 0, NB_ERR_ARG, "nb_mean_over_rows: bad arguments");
    dim3 grid((Wc + 255) / 256, B);
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); mean_rows_kernel<double><<<grid, 256, 0, st>>>((const double*)in, out, H, Wc); NB_PROF_END("mean_rows_kernel", st); }
    else if (prec == NB_PREC_F32)
        { NB_PROF_BEGIN(st); mean_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)in, out, H, Wc); NB_PROF_END("mean_rows_kernel", st); }
    else {
        nb_set_error("nb_mean_over_rows: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_cast(int precIn, const void* in, int precOut, void* out, long long n, cudaStream_t st) {
    NB_REQUIRE(in && out && n >= 0, NB_ERR_ARG, "nb_cast: bad arguments");
    if (n == 0) return NB_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (precIn == NB_PREC_F32 && precOut == NB_PREC_F64)
        { NB_PROF_BEGIN(st); cast_kernel<float, double><<<grid, 256, 0, st>>>((const float*)in, (double*)out, n); NB_PROF_END("cast_kernel", st); }
    else if (precIn == NB_PREC_F64 && precOut == NB_PREC_F32)
        { NB_PROF_BEGIN(st); cast_kernel<double, float><<<grid, 256, 0, st>>>((const double*)in, (float*)out, n); NB_PROF_END("cast_kernel", st); }
    else if (precIn == NB_PREC_F64 && precOut == NB_PREC_F64)
        { NB_PROF_BEGIN(st); cast_kernel<double, double><<<grid, 256, 0, st>>>((const double*)in, (double*)out, n); NB_PROF_END("cast_kernel", st); }
    else if (precIn == NB_PREC_F32 && precOut == NB_PREC_F32)
        { NB_PROF_BEGIN(st); cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, n); NB_PROF_END("cast_kernel", st); }
    else {
        nb_set_error("nb_cast: bad precision");
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_mask_to_u8(int precIn, const void* in, unsigned char* out, long long n, cudaStream_t st) {
    NB_REQUIRE(in && out && n >= 0, NB_ERR_ARG, "nb_mask_to_u8: bad arguments");
    if (n == 0) return NB_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (precIn == NB_PREC_F32)
        { NB_PROF_BEGIN(st); mask_u8_kernel<float><<<grid, 256, 0, st>>>((const float*)in, out, n); NB_PROF_END("mask_u8_kernel", st); }
    else if (precIn == NB_PREC_F64)
        { NB_PROF_BEGIN(st); mask_u8_kernel<double><<<grid, 256, 0, st>>>((const double*)in, out, n); NB_PROF_END("mask_u8_kernel", st); }
    else {
        nb_set_error("nb_mask_to_u8: bad precision");
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

}  // extern "C"
