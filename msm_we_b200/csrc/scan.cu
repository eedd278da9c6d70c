// Device-wide exclusive scan of int32 (reduce -> scan of block sums -> scan + add), deterministic.
#include "common.cuh"
#include "sort.cuh"

namespace mwe {

static constexpr int SC_THREADS = 256;
static constexpr int SC_ITEMS = 8;
static constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                int64_t* __restrict__ block_sums) {
    __shared__ long long s_part[SC_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    long long s = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + j * SC_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < SC_THREADS / 32; ++w) t += s_part[w];
        block_sums[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of block sums in place (serial chunks of 256, fixed order)
__global__ void __launch_bounds__(SC_THREADS) scan_blocksums_kernel(int64_t* __restrict__ block_sums, int64_t nblocks,
                                                                   int64_t* __restrict__ total_out) {
    __shared__ long long s_warp[SC_THREADS / 32 + 1];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += SC_THREADS) {
        const int64_t i = base + threadIdx.x;
        long long v = (i < nblocks) ? block_sums[i] : 0;
        long long inc = v;
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long nb = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += nb;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long run = 0;
            for (int w = 0; w < SC_THREADS / 32; ++w) {
                long long t = s_warp[w];
                s_warp[w] = run;
                run += t;
            }
            s_warp[SC_THREADS / 32] = run;
        }
        __syncthreads();
        const long long carry = s_carry;
        if (i < nblocks) block_sums[i] = carry + s_warp[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[SC_THREADS / 32];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                               int64_t n, const int64_t* __restrict__ block_sums) {
    __shared__ int scratch[9];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    // blocked arrangement: thread t owns items [t*8, t*8+8) of the tile
    int v[SC_ITEMS];
    int tsum = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + (int64_t)threadIdx.x * SC_ITEMS + j;
        v[j] = (i < n) ? in[i] : 0;
        tsum += v[j];
    }
    int blk_total;
    int excl = block_excl_scan_256(tsum, scratch, &blk_total);
    long long run = block_sums[blockIdx.x] + excl;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        int64_t i = base + (int64_t)threadIdx.x * SC_ITEMS + j;
        if (i < n) out[i] = (int32_t)run;
        run += v[j];
    }
}

// small inputs: one CTA walks the array in tiles, carrying the running total (one launch instead of three)
__global__ void __launch_bounds__(SC_THREADS) scan_small_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                               int64_t n, int64_t* __restrict__ total_out) {
    __shared__ int scratch[9];
    long long carry = 0;
    for (int64_t base = 0; base < n; base += SC_TILE) {
        int v[SC_ITEMS];
        int tsum = 0;
#pragma unroll
        for (int j = 0; j < SC_ITEMS; ++j) {
            const int64_t i = base + (int64_t)threadIdx.x * SC_ITEMS + j;
            v[j] = (i < n) ? in[i] : 0;
            tsum += v[j];
        }
        int blk_total;
        const int excl = block_excl_scan_256(tsum, scratch, &blk_total);
        long long run = carry + excl;
#pragma unroll
        for (int j = 0; j < SC_ITEMS; ++j) {
            const int64_t i = base + (int64_t)threadIdx.x * SC_ITEMS + j;
            if (i < n) out[i] = (int32_t)run;
            run += v[j];
        }
        carry += blk_total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

static constexpr int64_t SC_SMALL_MAX = 1 << 16;

size_t scan_workspace_bytes(int64_t n) {
    int64_t nblocks = (n + SC_TILE - 1) / SC_TILE;
    if (nblocks < 1) nblocks = 1;
    return align_up((size_t)nblocks * sizeof(int64_t), 256) + 256;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int64_t* total_out, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
    if (n <= 0) {
        if (total_out) MWE_CHECK_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), stream));
        return MWE_OK;
    }
    if (ws_bytes < scan_workspace_bytes(n)) {
        set_last_error("scan: workspace too small");
        return MWE_E_WORKSPACE;
    }
    if (n <= SC_SMALL_MAX) {
        scan_small_kernel<<<1, SC_THREADS, 0, stream>>>(in, out, n, total_out);
        MWE_CHECK_LAUNCH();
        return MWE_OK;
    }
    Carver cv(ws, ws_bytes);
    const int64_t nblocks = (n + SC_TILE - 1) / SC_TILE;
    int64_t* block_sums = cv.take<int64_t>((size_t)nblocks);
    scan_reduce_kernel<<<(unsigned)nblocks, SC_THREADS, 0, stream>>>(in, n, block_sums);
    scan_blocksums_kernel<<<1, SC_THREADS, 0, stream>>>(block_sums, nblocks, total_out);
    scan_apply_kernel<<<(unsigned)nblocks, SC_THREADS, 0, stream>>>(in, out, n, block_sums);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

}  // namespace mwe
