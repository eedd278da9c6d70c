// Shared helpers for the msm_we_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/msm_we_b200.h"

namespace mwe {

void set_last_error(const char* fmt, ...);
int sm_count();
// index of the current device, clamped to [0, MWE_MAX_DEVICES): per-device caches (function attributes, SM count)
static constexpr int MWE_MAX_DEVICES = 64;
int device_slot();
// optional CUDA events recorded around the dominant kernel of the next calls (bench roofline timing)
void timing_events(cudaEvent_t* start, cudaEvent_t* stop);

// Programmatic dependent launch: the kernels of the hot path are short (3-60 us) and run back to back on one
// stream, so launch latency and CTA ramp-up are a visible share of the step.  Launched with launch_pdl(), a kernel
// may become resident while its predecessor is still running; it must call pdl_wait() before it touches anything
// the predecessor wrote (all of them do so first thing), and pdl_launch_dependents() lets ITS successor in.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define MWE_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mwe::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MWE_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define MWE_CHECK_LAUNCH()                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            mwe::set_last_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return MWE_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define MWE_REQUIRE(cond, msg)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            mwe::set_last_error("%s:%d: %s", __FILE__, __LINE__, msg);                         \
            return MWE_E_INVALID;                                                              \
        }                                                                                      \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace.
struct Carver {
    char* base;
    size_t off;
    size_t cap;
    Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

static inline int ceil_log2_u64(uint64_t x) {  // smallest b with 2^b >= x
    int b = 0;
    while (b < 64 && ((uint64_t)1 << b) < x) ++b;
    return b;
}

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// exclusive scan of one int per thread across a block of NW warps (NW <= 32); returns the exclusive prefix and
// writes the block total to *total (same for all threads). scratch: NW + 1 ints of shared memory.
template <int NW>
__device__ __forceinline__ int block_excl_scan(int v, int* scratch, int* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int s = (lane < NW) ? scratch[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (uint32_t)o) si += n;
        }
        if (lane < NW) scratch[lane] = si - s;  // exclusive warp base
        if (lane == NW - 1) scratch[NW] = si;
    }
    __syncthreads();
    int r = scratch[warp] + inc - v;
    *total = scratch[NW];
    __syncthreads();
    return r;
}
__device__ __forceinline__ int block_excl_scan_256(int v, int* scratch, int* total) { return block_excl_scan<8>(v, scratch, total); }

}  // namespace mwe
