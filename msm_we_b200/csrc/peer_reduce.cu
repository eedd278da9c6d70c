// Exchange step of the multi-GPU flux path, done by ONE kernel over NVLink peer memory.
//
// reference: get_fluxMatrix adds the per-iteration matrices of all Ray workers on the driver and divides by the
// iteration count (msm_we/_hamsm/_fluxmatrix.py:311-327, :342).  With WE iterations sharded over ranks (one process
// per GPU) that is an all-reduce(sum) of the dense (n+2)^2 fp64 matrix followed by "/ nI".  The matrix is small
// (2.9 MB at cfg2), so the exchange is latency-bound: NCCL all-reduce + a divide launch cost ~30 us per step.
// Here every rank's partial sum, result buffer and flag words are mapped into every other rank (CUDA IPC), and one
// kernel per rank does: signal "my partial is complete" -> wait for all peers -> reduce ITS slice of the cells by
// reading all partials in RANK ORDER (so the result does not depend on scheduling, unlike a ring/tree all-reduce) ->
// divide -> store the finished cells into every rank's result buffer -> signal / wait "slice delivered".
// Ranks hold contiguous iteration blocks in rank order, so rank order is iteration order.
//
// Flags are monotonically increasing epochs (never reset): flags[r][p] = epoch says "rank p's partial of call
// `epoch` is ready", flags[r][world + p] = epoch says "rank p's slice of call `epoch` has landed in rank r".
#include <string.h>

#include "common.cuh"

namespace mwe {

static constexpr int PR_THREADS = 256;
static constexpr int PR_MAX_WORLD = 16;
static constexpr unsigned long long PR_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // a peer is gone, not late
static constexpr uint32_t PR_POISON = 0xFFFFFFFFu;   // written instead of an epoch by a rank that gave up

struct PeerParams {
    const double* partial[PR_MAX_WORLD];
    double* out[PR_MAX_WORLD];
    uint32_t* flags[PR_MAX_WORLD];
    int rank, world;
    int64_t count;
    double divisor;
    uint32_t epoch;
    unsigned int* cta_counter;   // local, zero before the call, left at zero
    int32_t* err_count;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait until flags[base + p] of THIS rank has reached `epoch` for every peer p (threads 0..world-1 of the CTA).
// Returns false -- for the whole CTA -- when a peer poisoned its flag or did not show up within PR_TIMEOUT_NS: the
// caller must then neither reduce nor deliver (a stale partial would otherwise be summed into everybody's result).
__device__ __forceinline__ bool wait_all(const PeerParams& P, int base) {
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    __syncthreads();
    if ((int)threadIdx.x < P.world) {
        const uint32_t* f = P.flags[P.rank] + base + threadIdx.x;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
            const uint32_t v = ld_acquire_sys(f);
            if (v == PR_POISON) { atomicExch(&s_abort, 1); break; }
            if ((int32_t)(v - P.epoch) >= 0) break;
            if (++spins > 4096u) {
                __nanosleep(200);
                if (t0 == 0) t0 = global_ns();
                else if (global_ns() - t0 > PR_TIMEOUT_NS) { atomicExch(&s_abort, 1); break; }
            }
        }
    }
    __syncthreads();
    return s_abort == 0;
}
// Give up: count the error locally and poison this rank's "ready" and "delivered" words in EVERY rank, so that every
// peer's wait fails too and every rank raises -- nobody returns a matrix built from a stale partial.
__device__ __forceinline__ void poison_all(const PeerParams& P) {
    if (threadIdx.x == 0) atomicAdd(&P.err_count[MWE_ERR_INTERNAL], 1);
    if ((int)threadIdx.x < P.world) {
        st_release_sys(P.flags[threadIdx.x] + P.rank, PR_POISON);
        st_release_sys(P.flags[threadIdx.x] + P.world + P.rank, PR_POISON);
    }
}

__global__ void __launch_bounds__(PR_THREADS) flux_peer_allreduce_kernel(const PeerParams P) {
    pdl_wait();                 // the partial sum is complete (last K3 kernel of this stream)
    pdl_launch_dependents();
    // ---- my partial is complete (earlier kernels of this stream wrote it): tell everyone ----
    if (blockIdx.x == 0 && (int)threadIdx.x < P.world) {
        __threadfence_system();
        st_release_sys(P.flags[threadIdx.x] + P.rank, P.epoch);
    }
    if (!wait_all(P, 0)) {
        poison_all(P);
        return;
    }
    // ---- my slice of the cells: sum in rank order, divide, deliver to every rank ----
    const int64_t lo = P.count * P.rank / P.world, hi = P.count * (P.rank + 1) / P.world;
    const bool divide = P.divisor != 0.0 && P.divisor != 1.0;
    for (int64_t i = lo + (int64_t)blockIdx.x * PR_THREADS + threadIdx.x; i < hi; i += (int64_t)gridDim.x * PR_THREADS) {
        double s = __ldcv(P.partial[0] + i);      // peer memory: never served from a stale cache line
        for (int p = 1; p < P.world; ++p) s = __dadd_rn(s, __ldcv(P.partial[p] + i));
        if (divide) s = __ddiv_rn(s, P.divisor);
        for (int p = 0; p < P.world; ++p) P.out[p][i] = s;
    }
    // ---- delivered: the last CTA of this rank tells everyone, then waits for everyone's slices ----
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(P.cta_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *P.cta_counter = 0;
    __syncthreads();
    if ((int)threadIdx.x < P.world) {
        __threadfence_system();
        st_release_sys(P.flags[threadIdx.x] + P.world + P.rank, P.epoch);
    }
    // the grid (hence the stream) does not complete before every peer's slice is in my result buffer, and
    // before every peer has finished reading my partial (a peer signals only after its reads)
    if (!wait_all(P, P.world)) poison_all(P);
}

}  // namespace mwe

extern "C" int mwe_device_malloc(size_t bytes, void** out) {
    MWE_REQUIRE(out != nullptr && bytes > 0, "device_malloc: bad arguments");
    MWE_CHECK_CUDA(cudaMalloc(out, bytes));
    MWE_CHECK_CUDA(cudaMemset(*out, 0, bytes));
    return MWE_OK;
}

extern "C" int mwe_device_free(void* ptr) {
    if (ptr) MWE_CHECK_CUDA(cudaFree(ptr));
    return MWE_OK;
}

extern "C" int mwe_ipc_export(void* device_ptr, unsigned char* handle64) {
    MWE_REQUIRE(device_ptr && handle64, "ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    MWE_CHECK_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle64, &h, 64);
    return MWE_OK;
}

extern "C" int mwe_ipc_open(const unsigned char* handle64, void** out) {
    MWE_REQUIRE(handle64 && out, "ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        mwe::set_last_error("ipc_open: %s", cudaGetErrorString(e));
        return MWE_E_CUDA;
    }
    return MWE_OK;
}

extern "C" int mwe_ipc_close(void* ptr) {
    if (ptr) MWE_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return MWE_OK;
}

extern "C" int mwe_flux_peer_allreduce_f64(const void* const* partials, void* const* outs, void* const* flags, int rank,
                                           int world, int64_t count, double divisor, uint32_t epoch,
                                           void* cta_counter, int32_t* err_count, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(world >= 1 && world <= PR_MAX_WORLD && rank >= 0 && rank < world, "peer_allreduce: bad rank / world");
    MWE_REQUIRE(partials && outs && flags && cta_counter && err_count && count >= 0, "peer_allreduce: null pointer");
    MWE_REQUIRE(epoch != 0, "peer_allreduce: epochs start at 1");
    PeerParams P;
    for (int p = 0; p < PR_MAX_WORLD; ++p) {
        P.partial[p] = p < world ? static_cast<const double*>(partials[p]) : nullptr;
        P.out[p] = p < world ? static_cast<double*>(outs[p]) : nullptr;
        P.flags[p] = p < world ? static_cast<uint32_t*>(flags[p]) : nullptr;
        MWE_REQUIRE(p >= world || (P.partial[p] && P.out[p] && P.flags[p]), "peer_allreduce: null peer pointer");
    }
    P.rank = rank; P.world = world; P.count = count; P.divisor = divisor; P.epoch = epoch;
    P.cta_counter = static_cast<unsigned int*>(cta_counter);
    P.err_count = err_count;
    // One cell per thread while the grid fits a few waves: a peer load is a ~2 us round trip, so the slice wants
    // as many of them in flight as possible.  (CTAs that are not resident yet only ever wait for flags that the
    // FIRST CTA of every rank sets on entry, so the grid size cannot deadlock the ranks.)
    int64_t grid = (count / world + PR_THREADS - 1) / PR_THREADS;
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
    if (grid < 1) grid = 1;
    MWE_CHECK_CUDA(launch_pdl(flux_peer_allreduce_kernel, dim3((unsigned)grid), dim3(PR_THREADS), 0, static_cast<cudaStream_t>(stream), P));
    return MWE_OK;
}
