// Library-level plumbing: last-error text, device queries, small elementwise helpers.
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"

namespace mwe {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static thread_local cudaEvent_t g_ev_start = nullptr;
static thread_local cudaEvent_t g_ev_stop = nullptr;

void timing_events(cudaEvent_t* start, cudaEvent_t* stop) {
    *start = g_ev_start;
    *stop = g_ev_stop;
}

int device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return dev < 0 ? 0 : (dev >= MWE_MAX_DEVICES ? MWE_MAX_DEVICES - 1 : dev);
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 148;
    }
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        else cudaGetLastError();
        cached_dev = dev;
    }
    return cached;
}

__global__ void divide_kernel(double* __restrict__ buf, int64_t n, double divisor) {
    pdl_wait();
    pdl_launch_dependents();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = buf[i] / divisor;
}

}  // namespace mwe

extern "C" int mwe_set_timing_events(void* start, void* stop) {
    mwe::g_ev_start = static_cast<cudaEvent_t>(start);
    mwe::g_ev_stop = static_cast<cudaEvent_t>(stop);
    return MWE_OK;
}

extern "C" int mwe_abi_version(void) { return MWE_ABI_VERSION; }
extern "C" const char* mwe_last_error(void) { return mwe::g_last_error; }
namespace mwe {
bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("MWE_PDL"); return !(e && atoi(e) == 0); }();   // tuning knob
    return on;
}
}  // namespace mwe

extern "C" int mwe_device_sm_count(void) { return mwe::sm_count(); }

extern "C" int mwe_host_register(void* ptr, size_t bytes) {
    MWE_REQUIRE(ptr != nullptr && bytes > 0, "host_register: empty range");
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();   // not sticky, but it must not surface at the next launch check
        mwe::set_last_error("host_register: %s", cudaGetErrorString(e));
        return MWE_E_CUDA;
    }
    return MWE_OK;
}

extern "C" int mwe_host_unregister(void* ptr) {
    MWE_REQUIRE(ptr != nullptr, "host_unregister: null pointer");
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        mwe::set_last_error("host_unregister: %s", cudaGetErrorString(e));
        return MWE_E_CUDA;
    }
    return MWE_OK;
}

extern "C" int mwe_divide_f64(double* buf, int64_t count, double divisor, void* stream) {
    MWE_REQUIRE(count >= 0, "divide: negative count");
    if (count == 0) return MWE_OK;
    int64_t blocks = (count + 255) / 256;
    int64_t cap = (int64_t)mwe::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    MWE_CHECK_CUDA(mwe::launch_pdl(mwe::divide_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), buf, count, divisor));
    return MWE_OK;
}
