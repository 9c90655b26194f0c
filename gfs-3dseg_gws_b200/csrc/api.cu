// api.cu -- error plumbing and small utilities of the C ABI (include/gfs3d.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <map>
#include <set>
#include <utility>

#include "common.cuh"

namespace gfs {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached[dev] = n;
    }
    return cached[dev];
}

static int g_pdl = -1;   // -1: read GFS3D_PDL on first use
bool pdl_enabled(int cls) {
    if (g_pdl < 0) {
        const char* e = getenv("GFS3D_PDL");
        g_pdl = e ? (atoi(e) & 3) : 3;
    }
    return (g_pdl & cls) != 0;
}

static thread_local bool g_pdl_break = false;
void pdl_break() { g_pdl_break = true; }
bool pdl_take_break() {
    const bool b = g_pdl_break;
    g_pdl_break = false;
    return b;
}

cudaError_t allow_smem(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;   // largest size granted so far
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find({dev, func});
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done[{dev, func}] = bytes;
    return e;
}

}  // namespace gfs

extern "C" int gfs_version(void) { return 100; }
extern "C" const char* gfs_last_error_string(void) { return gfs::g_err; }
extern "C" int gfs_device_sm_count(void) { return gfs::sm_count(); }
extern "C" int gfs_set_pdl(int on) {
    gfs::g_pdl = on & 3;
    return GFS_OK;
}
