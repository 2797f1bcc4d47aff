// Error reporting and device queries for libnemo_b200.so.
#include "nemo_b200_internal.h"
#include <stdarg.h>

static thread_local char g_err[1024] = "";

void nb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nb_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

extern "C" {

const char* nb_last_error(void) { return g_err; }

int nb_version(void) { return 100; }

int nb_device_info(int* smCount, int* ccMajor, int* ccMinor) {
    int dev = 0;
    NB_CHECK_CUDA(cudaGetDevice(&dev));
    if (smCount) NB_CHECK_CUDA(cudaDeviceGetAttribute(smCount, cudaDevAttrMultiProcessorCount, dev));
    if (ccMajor) NB_CHECK_CUDA(cudaDeviceGetAttribute(ccMajor, cudaDevAttrComputeCapabilityMajor, dev));
    if (ccMinor) NB_CHECK_CUDA(cudaDeviceGetAttribute(ccMinor, cudaDevAttrComputeCapabilityMinor, dev));
    return NB_OK;
}

}  // extern "C"
