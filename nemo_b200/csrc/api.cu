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

// ------------------------------------------------------------------------------------------------------
// per-kernel event timing
#include <map>
#include <mutex>
#include <string>
#include <vector>

int g_nb_prof_on = 0;
unsigned long long g_nb_launches = 0;
namespace {
struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
};
std::mutex g_prof_mutex;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t g_prof_pending = nullptr;

cudaEvent_t prof_event() {
    cudaEvent_t e = nullptr;
    if (!g_prof_pool.empty()) {
        e = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else if (cudaEventCreate(&e) != cudaSuccess) {
        e = nullptr;
    }
    return e;
}
}  // namespace

void nb_prof_begin_(cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (g_prof_pending) g_prof_pool.push_back(g_prof_pending);
    g_prof_pending = prof_event();
    if (g_prof_pending) cudaEventRecord(g_prof_pending, st);
}

void nb_prof_end_(const char* name, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (!g_prof_pending) return;
    ProfRec r;
    r.name = name;
    r.e0 = g_prof_pending;
    r.e1 = prof_event();
    g_prof_pending = nullptr;
    if (!r.e1) {
        g_prof_pool.push_back(r.e0);
        return;
    }
    cudaEventRecord(r.e1, st);
    g_prof_recs.push_back(r);
}

// ------------------------------------------------------------------------------------------------------
// tensor maps for the TMA column tiles (tma.cuh): cuTensorMapEncodeTiled through the runtime's driver entry point
#include "tma.cuh"

typedef CUresult (*NbEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int nb_make_tmap_2d(CUtensorMap* map, const void* base, int elemBytes, unsigned long long rows, unsigned long long cols,
                    unsigned long long pitchBytes, unsigned boxRows, unsigned boxCols) {
    static NbEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        NB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, NB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (NbEncodeTiled)fn;
    }
    NB_REQUIRE((elemBytes == 4 || elemBytes == 8) && boxRows >= 1 && boxRows <= 256 && boxCols >= 1 && boxCols <= 256 &&
                   ((size_t)boxCols * elemBytes) % 16 == 0 && pitchBytes % 16 == 0 && ((uintptr_t)base & 15) == 0,
               NB_ERR_ARG, "nb_make_tmap_2d: box %u x %u of %d-byte elements, pitch %llu, base %p cannot be a TMA tile", boxRows, boxCols,
               elemBytes, pitchBytes, base);
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitchBytes};
    const cuuint32_t box[2] = {boxCols, boxRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, elemBytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NB_REQUIRE(r == CUDA_SUCCESS, NB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return NB_OK;
}

extern "C" {

long long nb_launch_count(void) { return (long long)__atomic_load_n(&g_nb_launches, __ATOMIC_RELAXED); }

long long nb_launch_count_add(long long n) {
    return (long long)__atomic_add_fetch(&g_nb_launches, (unsigned long long)n, __ATOMIC_RELAXED);
}

int nb_prof_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& r : g_prof_recs) {
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    g_nb_prof_on = on ? 1 : 0;
    return NB_OK;
}

long long nb_prof_report(char* buf, long long cap) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (cudaDeviceSynchronize() != cudaSuccess) return NB_ERR_CUDA;
    std::map<std::string, std::pair<long long, double>> agg;
    std::vector<std::string> order;
    for (auto& r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        auto it = agg.find(r.name);
        if (it == agg.end()) {
            agg[r.name] = std::make_pair(1LL, (double)ms);
            order.push_back(r.name);
        } else {
            it->second.first += 1;
            it->second.second += (double)ms;
        }
    }
    std::string out;
    char line[256];
    for (auto& n : order) {
        snprintf(line, sizeof(line), "%s,%lld,%.6f\n", n.c_str(), agg[n].first, agg[n].second);
        out += line;
    }
    if (buf && cap > 0) {
        const long long n = (long long)out.size() < cap - 1 ? (long long)out.size() : cap - 1;
        memcpy(buf, out.data(), (size_t)n);
        buf[n] = 0;
    }
    return (long long)out.size();
}

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
