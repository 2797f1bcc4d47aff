// Noise-grid geometry shared by the K6 / K7 translation units (noise.cu, noise_onebin.cu).
#pragma once
#include "common.cuh"

#define NB_MAX_CHUNKS 128
struct Chunks {  // np.linspace(0, H, n + 1, dtype=int) boundaries, passed by value
    int y[NB_MAX_CHUNKS + 1];
    int x[NB_MAX_CHUNKS + 1];
};

struct CellGeom {
    int y0, y1, x0, x1;
};

// cell (i, k) with its overlap margins, clipped to the map (filters.py:430-443)
__device__ __forceinline__ CellGeom cell_geom(const Chunks& ch, int i, int k, int ov, int H, int W) {
    CellGeom g;
    g.y0 = max(ch.y[i] - ov, 0);
    g.y1 = min(ch.y[i + 1] + ov, H);
    g.x0 = max(ch.x[k] - ov, 0);
    g.x1 = min(ch.x[k + 1] + ov, W);
    return g;
}

// one weight bin per cell, clipped estimator: zone moments + list-based clipping passes (noise_onebin.cu)
bool nb_onebin_supported(int H, int W, const int* yChunks, int ny, const int* xChunks, int nx, int ov);
int nb_cell_rms_onebin(int prec, const void* map, const void* wmed, int nImg, int S, int H, int W, const Chunks& ch, int ny, int nx, int ov,
                       double* table, double* edges, cudaStream_t st);
