// Stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.
//
// Shared by the flux scatter (key = matrix cell, value = transition index) and the centroid
// accumulation (key = cluster label, value = point index).  Stability is what makes both results
// order-deterministic: inside one key the values stay in input order, so the segmented fp64 sums
// that follow always add in the same sequence.
//
// Launches over a FIXED grid of G CTAs, each owning a contiguous span of tiles:
//   hist    : per-CTA digit histogram (first pass only; every scatter counts the next pass's digits)
//   scan    : exclusive scan over [digit][cta]         (one CTA, G*256 counters)
//   scatter : per tile, warp-level match ranking -> tile-local reorder in shared memory ->
//             coalesced run-wise stores                 (reads + writes keys and values)
// HBM-bound integer work: everything is coalesced 128-bit-friendly streaming, the histogram and
// staging live in shared memory, and G is a multiple of the SM count when the input is large.
#include "common.cuh"
#include "sort.cuh"

namespace mwe {

static constexpr int RS_THREADS = 512;      // scatter kernel: 16 warps x 4 items rank a tile in 4 match rounds
static constexpr int RS_ITEMS = 4;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048
static constexpr int RS_HTHREADS = 256;     // histogram / scan kernels: one thread per digit
static constexpr int RS_RADIX = 256;
static constexpr int RS_WARPS = RS_THREADS / 32;
static constexpr int RS_FUSED_SCAN_MAX_G = 640;   // up to here (every grid this file launches) the scatter kernel scans the histograms itself

// Only the first pass runs this kernel: every scatter pass counts the NEXT pass's per-CTA digits while it stores
// (it knows where each element lands), so later passes need no histogram launch.  The later histograms are zeroed
// here, each CTA its own rows.
__global__ void __launch_bounds__(RS_HTHREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t N, int shift,
                                                            int tiles_per_cta, uint32_t* __restrict__ hist, int later_passes,
                                                            size_t hist_stride) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ uint32_t s_hist[RS_RADIX];
    s_hist[threadIdx.x] = 0;
    for (int q = 1; q <= later_passes; ++q) hist[q * hist_stride + (size_t)blockIdx.x * RS_RADIX + threadIdx.x] = 0;
    __syncthreads();
    const int64_t begin = (int64_t)blockIdx.x * tiles_per_cta * RS_TILE;
    int64_t end = begin + (int64_t)tiles_per_cta * RS_TILE;
    if (end > N) end = N;
    for (int64_t i = begin + threadIdx.x; i < end; i += RS_HTHREADS) {
        uint32_t d = (uint32_t)(keys[i] >> shift) & 0xffu;
        atomicAdd(&s_hist[d], 1u);
    }
    __syncthreads();
    hist[(size_t)blockIdx.x * RS_RADIX + threadIdx.x] = s_hist[threadIdx.x];
}

// hist layout [G][256]; afterwards hist[c][d] = global start offset of digit d for CTA c.
__global__ void __launch_bounds__(RS_RADIX) rs_scan_kernel(uint32_t* __restrict__ hist, int G) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ int scratch[9];
    const int d = threadIdx.x;
    uint32_t total = 0;
    for (int c = 0; c < G; ++c) total += hist[(size_t)c * RS_RADIX + d];
    int blk_total;
    uint32_t run = (uint32_t)block_excl_scan_256((int)total, scratch, &blk_total);
    for (int c = 0; c < G; ++c) {
        uint32_t t = hist[(size_t)c * RS_RADIX + d];
        hist[(size_t)c * RS_RADIX + d] = run;
        run += t;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
    rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                      uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t N, int shift,
                      int tiles_per_cta, const uint32_t* __restrict__ offsets, int fused_scan,
                      uint32_t* __restrict__ next_hist) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ uint32_t s_base[RS_RADIX];
    __shared__ uint32_t s_warp_cnt[RS_WARPS][RS_RADIX];
    __shared__ uint32_t s_tile_excl[RS_RADIX];
    __shared__ uint32_t s_tile_total[RS_RADIX];
    __shared__ uint64_t s_keys[RS_TILE];
    __shared__ uint32_t s_vals[RS_TILE];
    __shared__ int scratch[RS_WARPS + 1];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool dthread = tid < RS_RADIX;           // the first 256 threads double as "one thread per digit"
    if (fused_scan) {
        // small grids: every CTA derives its own digit offsets from the raw per-CTA histograms
        // (offsets[c][d] = count) instead of waiting for a separate one-CTA scan kernel
        uint32_t below = 0, total = 0;
        const int G = (int)gridDim.x;
        // 32 independent loads in flight per thread: the loop is pure L2 latency otherwise
        if (dthread)
            for (int c0 = 0; c0 < G; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (c0 + i < G) ? offsets[(size_t)(c0 + i) * RS_RADIX + tid] : 0u;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (c0 + i < (int)blockIdx.x) below += v[i];
                    total += v[i];
                }
            }
        int blk_total;
        const uint32_t digit_base = (uint32_t)block_excl_scan<RS_WARPS>((int)total, scratch, &blk_total);
        if (dthread) s_base[tid] = digit_base + below;
    } else if (dthread) {
        s_base[tid] = offsets[(size_t)blockIdx.x * RS_RADIX + tid];
    }

    const uint32_t span = (uint32_t)tiles_per_cta * RS_TILE;
    const int64_t span_begin = (int64_t)blockIdx.x * span;
    for (int t = 0; t < tiles_per_cta; ++t) {
        const int64_t tile_base = span_begin + (int64_t)t * RS_TILE;
        if (tile_base >= N) break;
        const int tile_count = (int)((N - tile_base < RS_TILE) ? (N - tile_base) : RS_TILE);
        for (uint32_t i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s_warp_cnt[0][0])[i] = 0;
        __syncthreads();

        // warp-striped load: item j of lane l is element warp*256 + j*32 + l of the tile
        uint64_t key[RS_ITEMS];
        uint32_t val[RS_ITEMS];
        uint32_t rank[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int e = warp * (32 * RS_ITEMS) + j * 32 + lane;
            if (e < tile_count) {
                key[j] = keys_in[tile_base + e];
                val[j] = vals_in[tile_base + e];
            } else {
                key[j] = ~0ull;
                val[j] = 0;
            }
        }
        // stable ranking inside the warp, item by item (lower item / lower lane first)
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int e = warp * (32 * RS_ITEMS) + j * 32 + lane;
            const bool valid = e < tile_count;
            const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & 0xffu) : 0x100u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            uint32_t prefix = 0;
            if (valid) prefix = s_warp_cnt[warp][d];
            __syncwarp();
            rank[j] = prefix + __popc(peers & lt_mask);
            if (valid && (peers & lt_mask) == 0) s_warp_cnt[warp][d] = prefix + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // per digit: exclusive prefix over warps, tile total, then exclusive scan over digits
        {
            uint32_t run = 0;
            if (dthread) {
#pragma unroll
                for (int w = 0; w < RS_WARPS; ++w) {
                    uint32_t c = s_warp_cnt[w][tid];
                    s_warp_cnt[w][tid] = run;
                    run += c;
                }
                s_tile_total[tid] = run;
            }
            int blk_total;
            const uint32_t ex = (uint32_t)block_excl_scan<RS_WARPS>((int)run, scratch, &blk_total);
            if (dthread) s_tile_excl[tid] = ex;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int e = warp * (32 * RS_ITEMS) + j * 32 + lane;
            if (e < tile_count) {
                const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
                const uint32_t pos = s_tile_excl[d] + s_warp_cnt[warp][d] + rank[j];
                s_keys[pos] = key[j];
                s_vals[pos] = val[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            const int p = k * RS_THREADS + tid;
            if (p < tile_count) {
                const uint64_t kk = s_keys[p];
                const uint32_t d = (uint32_t)(kk >> shift) & 0xffu;
                const uint32_t g = s_base[d] + ((uint32_t)p - s_tile_excl[d]);
                keys_out[g] = kk;
                vals_out[g] = s_vals[p];
                // per-CTA digit histogram of the next pass: element g belongs to CTA g / span there
                if (next_hist) atomicAdd(&next_hist[(size_t)(g / span) * RS_RADIX + ((uint32_t)(kk >> (shift + 8)) & 0xffu)], 1u);
            }
        }
        __syncthreads();
        if (dthread) s_base[tid] += s_tile_total[tid];
        __syncthreads();
    }
}

static void plan(int64_t N, int* G, int* tiles_per_cta) {
    int64_t tiles = (N + RS_TILE - 1) / RS_TILE;
    if (tiles < 1) tiles = 1;
    int64_t maxG = (int64_t)sm_count() * 4;
    int64_t tpc = (tiles + maxG - 1) / maxG;
    if (tpc < 1) tpc = 1;
    *tiles_per_cta = (int)tpc;
    *G = (int)((tiles + tpc - 1) / tpc);
}

size_t sort_workspace_bytes(int64_t N) {
    if (N < 1) N = 1;
    size_t b = 0;
    b += align_up((size_t)N * sizeof(uint64_t), 256);
    b += align_up((size_t)N * sizeof(uint32_t), 256);
    b += 8 * align_up((size_t)sm_count() * 4 * RS_RADIX * sizeof(uint32_t), 256);   // one histogram per pass
    return b + 1024;
}

int sort_pairs(uint64_t* keys, uint32_t* vals, int64_t N, int key_bits, void* ws, size_t ws_bytes,
               cudaStream_t stream, uint64_t** keys_sorted, uint32_t** vals_sorted) {
    MWE_REQUIRE(N >= 0 && N < ((int64_t)1 << 32), "sort: N must be < 2^32");
    MWE_REQUIRE(key_bits >= 0 && key_bits <= 64, "sort: key_bits out of range");
    *keys_sorted = keys;
    *vals_sorted = vals;
    if (N <= 1 || key_bits == 0) return MWE_OK;
    if (ws_bytes < sort_workspace_bytes(N)) {
        set_last_error("sort: workspace too small (%zu < %zu)", ws_bytes, sort_workspace_bytes(N));
        return MWE_E_WORKSPACE;
    }
    Carver cv(ws, ws_bytes);
    uint64_t* keys_alt = cv.take<uint64_t>((size_t)N);
    uint32_t* vals_alt = cv.take<uint32_t>((size_t)N);
    const size_t hist_stride = align_up((size_t)sm_count() * 4 * RS_RADIX * sizeof(uint32_t), 256) / sizeof(uint32_t);
    uint32_t* hist = cv.take<uint32_t>(8 * hist_stride);
    int G, tpc;
    plan(N, &G, &tpc);
    uint64_t* kin = keys;
    uint32_t* vin = vals;
    uint64_t* kout = keys_alt;
    uint32_t* vout = vals_alt;
    const int passes = (key_bits + 7) / 8;
    const int fused = G <= RS_FUSED_SCAN_MAX_G;
    MWE_CHECK_CUDA(launch_pdl(rs_hist_kernel, dim3(G), dim3(RS_HTHREADS), 0, stream, kin, N, 0, tpc, hist, passes - 1, hist_stride));
    for (int p = 0; p < passes; ++p) {
        const int shift = p * 8;
        uint32_t* hp = hist + (size_t)p * hist_stride;
        if (!fused) MWE_CHECK_CUDA(launch_pdl(rs_scan_kernel, dim3(1), dim3(RS_RADIX), 0, stream, hp, G));
        MWE_CHECK_CUDA(launch_pdl(rs_scatter_kernel, dim3(G), dim3(RS_THREADS), 0, stream, kin, vin, kout, vout, N, shift, tpc, hp,
                                  fused, p + 1 < passes ? hp + hist_stride : nullptr));
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *keys_sorted = kin;
    *vals_sorted = vin;
    return MWE_OK;
}

}  // namespace mwe

extern "C" size_t mwe_sort_workspace_bytes(int64_t N) { return mwe::sort_workspace_bytes(N); }

extern "C" int mwe_sort_pairs_u64_u32(uint64_t* keys, uint32_t* vals, int64_t N, int key_bits, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint64_t* ks;
    uint32_t* vs;
    int rc = mwe::sort_pairs(keys, vals, N, key_bits, workspace, workspace_bytes, s, &ks, &vs);
    if (rc != MWE_OK) return rc;
    if (ks != keys) {
        MWE_CHECK_CUDA(cudaMemcpyAsync(keys, ks, (size_t)N * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        MWE_CHECK_CUDA(cudaMemcpyAsync(vals, vs, (size_t)N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    return MWE_OK;
}
