// K3: weighted transition scatter into the flux matrix, order-deterministic.
//
// reference: FluxMatrixMixin.build_flux_matrix (msm_we/_hamsm/_fluxmatrix.py:97-164: relabel
// end[target]=n+1, start[basis]=n, end[basis]=n, then scipy coo_matrix which sums duplicates in
// input order), the per-iteration dense accumulation of get_fluxMatrix (:232-260) and the
// history-coloured count scatter of NonMarkovModel.fit (msm_we/nmm.py:132-158,
// row = 2*s_prev + colour_prev, col = 2*s_now + colour_now).
//
// One launch sequence covers every transition of every iteration of the shard:
//   keys    : cell = (C*start+c0)*CM + (C*end+c1) after the basis/target overrides; bad labels get
//             a sentinel cell that sorts last and is dropped
//   sort    : stable LSD radix sort by cell (radix_sort.cu), value = transition index, so inside a
//             cell the transitions stay in (iteration, segment) order
//   mark    : gather weights in sorted order; flag the first element of every cell and of every
//             (cell, iteration) group
//   group   : one thread per (cell, iteration) group adds its weights in segment order -- the value
//             scipy's coo_matrix gives that iteration's matrix cell
//   cell    : one thread per cell adds the cell's groups in iteration order onto the running matrix
//             value and performs the single write (dense, or one COO triple)
//             This two-level sum IS the association of the reference's serial path, the chains are
//             bounded by segments-per-iteration and by the iteration count, and there are no
//             floating-point atomics anywhere, so the result does not depend on thread scheduling.
// Integer/HBM-bound: algorithmic bytes per transition = 2*8 (labels) + 8 (weight) [+2 flags +2 colours].
#include <stdlib.h>

#include "common.cuh"
#include "sort.cuh"

namespace mwe {

// iteration that owns transition idx: largest it with offsets[it] <= idx
__device__ __forceinline__ int64_t find_iter(const int64_t* __restrict__ offsets, int64_t n_iters, int64_t idx) {
    int64_t lo = 0, hi = n_iters;  // answer in [lo, hi)
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
    flux_keys_kernel(const int64_t* __restrict__ start, const int64_t* __restrict__ end,
                     const uint8_t* __restrict__ flag0, const uint8_t* __restrict__ flag1,
                     const uint8_t* __restrict__ col0, const uint8_t* __restrict__ col1, int64_t N, int64_t n_clusters,
                     int C, uint64_t sentinel, const int64_t* __restrict__ iter_offsets, int64_t n_iters, int iter_shift,
                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* __restrict__ err_count,
                     unsigned long long* __restrict__ tile_state, int64_t n_tile_state) {
    pdl_wait();
    pdl_launch_dependents();
    // look-back state + ticket of the mark/scan pass that follows the sort (zeroed here: no memset between kernels)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tile_state; i += (int64_t)gridDim.x * blockDim.x)
        tile_state[i] = 0ull;
    const int64_t M = n_clusters + 2;
    const uint64_t CM = (uint64_t)C * (uint64_t)M;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        int64_t s = start[i], e = end[i];
        const uint8_t f0 = flag0 ? flag0[i] : (uint8_t)0;
        const uint8_t f1 = flag1 ? flag1[i] : (uint8_t)0;
        // override order of _fluxmatrix.py:135-137 (a child in both regions ends up in the basis)
        if (f1 & MWE_FLAG_TARGET) e = n_clusters + 1;
        if (f0 & MWE_FLAG_BASIS) s = n_clusters;
        if (f1 & MWE_FLAG_BASIS) e = n_clusters;
        uint64_t key = sentinel;
        if (s < 0 || s >= M || e < 0 || e >= M) {
            atomicAdd(&err_count[MWE_ERR_LABEL_RANGE], 1);
        } else {
            const uint64_t c0 = col0 ? (uint64_t)col0[i] : 0ull;
            const uint64_t c1 = col1 ? (uint64_t)col1[i] : 0ull;
            if (c0 >= (uint64_t)C || c1 >= (uint64_t)C) atomicAdd(&err_count[MWE_ERR_LABEL_RANGE], 1);
            else key = ((uint64_t)C * (uint64_t)s + c0) * CM + ((uint64_t)C * (uint64_t)e + c1);
        }
        // The WE iteration of the transition rides in the key bits ABOVE the cell (the sort only looks at the
        // cell bits, the rest is carried along), so the passes after the sort find (cell, iteration) group
        // boundaries by comparing neighbouring keys instead of searching iter_offsets per element.
        if (iter_shift && key != sentinel) {
            // iterations of a WE run have similar sizes: start from the proportional guess, walk a few steps, and
            // only fall back to the binary search (a chain of log2(n_iters) dependent loads) when that fails
            int64_t it = (int64_t)(((unsigned long long)i * (unsigned long long)n_iters) / (unsigned long long)N);
            int steps = 0;
            while (it > 0 && iter_offsets[it] > i && steps < 4) { --it; ++steps; }
            while (it + 1 < n_iters && iter_offsets[it + 1] <= i && steps < 4) { ++it; ++steps; }
            if (iter_offsets[it] > i || (it + 1 < n_iters && iter_offsets[it + 1] <= i)) it = find_iter(iter_offsets, n_iters, i);
            key |= (uint64_t)it << iter_shift;
        }
        keys[i] = key;
        vals[i] = (uint32_t)i;
    }
}

// Pass 1 + group numbering in ONE launch: the flags of pass 1 and their exclusive prefix sum (= the index of
// the group every element belongs to), single pass with decoupled look-back across tiles (each tile publishes
// its aggregate, then its inclusive prefix, in one 64-bit word: top 2 bits = status, rest = value).  Tiles take
// their number from an atomic ticket, so a tile only ever waits for tiles that are already running.
static constexpr int FM_THREADS = 256;
static constexpr int FM_TILE_MIN = FM_THREADS * 2;   // smallest tile the launcher uses (sizes the look-back state)

// FM_ITEMS consecutive elements per thread: 2 while the pass is latency-bound (more tiles in flight), 8 for large N
template <int FM_ITEMS>
__global__ void __launch_bounds__(FM_THREADS)
    flux_mark_scan_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t N,
                          uint64_t sentinel, uint64_t cell_mask, bool iter_in_key, const double* __restrict__ w,
                          const int64_t* __restrict__ iter_offsets, int64_t n_iters, double* __restrict__ wv,
                          int32_t* __restrict__ sub_head,
                          int32_t* __restrict__ cell_head, int32_t* __restrict__ sub_pos,
                          unsigned long long* __restrict__ tile_state, unsigned int* __restrict__ ticket,
                          int64_t* __restrict__ total_out, int32_t* __restrict__ err_count) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int FM_TILE = FM_THREADS * FM_ITEMS;
    __shared__ int scratch[9];
    __shared__ unsigned int s_tile;
    __shared__ long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * FM_TILE + (int64_t)threadIdx.x * FM_ITEMS;   // blocked: 8 consecutive items
    int sh[FM_ITEMS], ch[FM_ITEMS];
    int tsum = 0;
#pragma unroll
    for (int j = 0; j < FM_ITEMS; ++j) {
        const int64_t q = base + j;
        sh[j] = ch[j] = 0;
        if (q < N) {
            const uint64_t k = keys[q];
            if ((k & cell_mask) != sentinel) {
                const uint32_t idx = vals[q];
                if (w) wv[q] = w[idx];
                const uint64_t kp = q ? keys[q - 1] : ~k;
                ch[j] = ((kp ^ k) & cell_mask) ? 1 : 0;
                sh[j] = ch[j];
                if (iter_in_key) {
                    sh[j] = (kp != k) ? 1 : 0;          // cell or iteration differs
                } else if (!ch[j] && iter_offsets) {
                    // same cell as the previous element (which has a smaller transition index): a new group
                    // starts when the previous element belongs to an earlier iteration
                    const int64_t it = find_iter(iter_offsets, n_iters, (int64_t)idx);
                    sh[j] = ((int64_t)vals[q - 1] < iter_offsets[it]) ? 1 : 0;
                }
            }
            sub_head[q] = sh[j];
            cell_head[q] = ch[j];
            tsum += sh[j];
        }
    }
    int blk_total;
    const int excl = block_excl_scan_256(tsum, scratch, &blk_total);
    // ---- decoupled look-back (warp 0: 32 predecessors per round) ----
    if (threadIdx.x < 32) {
        const unsigned long long AGG = 1ull << 62, PRE = 2ull << 62, MASK = (1ull << 62) - 1;
        const int lane = threadIdx.x;
        long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(tile_state + 0, PRE | (unsigned long long)blk_total);
        } else {
            if (lane == 0) atomicExch(tile_state + tile, AGG | (unsigned long long)blk_total);
            long long hi = (long long)tile - 1;   // nearest predecessor not yet accounted for
            bool done = false;
            int spins = 0;
            while (!done) {
                const long long t = hi - lane;
                unsigned long long st = PRE;      // lanes before tile 0 behave like a finished prefix of 0
                if (t >= 0) {
                    while (((st = *reinterpret_cast<volatile unsigned long long*>(tile_state + t)) >> 62) == 0) {
                        if (++spins > (1 << 22)) {   // never hang the GPU on a protocol bug
                            atomicAdd(&err_count[MWE_ERR_INTERNAL], 1);
                            st = PRE;
                            break;
                        }
                    }
                }
                const uint32_t pre_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                const int stop = pre_mask ? (__ffs(pre_mask) - 1) : 32;   // first lane (nearest tile) with an inclusive prefix
                long long contrib = (lane <= stop && t >= 0) ? (long long)(st & MASK) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                done = pre_mask != 0;
                hi -= 32;
            }
            if (lane == 0) atomicExch(tile_state + tile, PRE | (unsigned long long)(prefix + blk_total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if ((int64_t)(tile + 1) * FM_TILE >= N && total_out) *total_out = prefix + blk_total;
        }
    }
    __syncthreads();
    long long run = s_prefix + excl;
#pragma unroll
    for (int j = 0; j < FM_ITEMS; ++j) {
        const int64_t q = base + j;
        if (q < N) sub_pos[q] = (int32_t)run;
        run += sh[j];
    }
}

// Pass 2 (one thread per sub head): sum the weights of one (cell, iteration) group in segment order --
// what scipy's coo_matrix -> dense does for that iteration's matrix.  Unit weights: the count, exactly.
__global__ void __launch_bounds__(256)
    flux_group_sum_kernel(const uint64_t* __restrict__ keys, int64_t N, uint64_t sentinel, uint64_t cell_mask,
                          const double* __restrict__ wv,
                          bool have_w, const int32_t* __restrict__ sub_head, const int32_t* __restrict__ sub_pos,
                          const int32_t* __restrict__ cell_head, double* __restrict__ group_sum,
                          uint8_t* __restrict__ group_is_cell_head, uint64_t* __restrict__ group_key) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += stride) {
        if (!sub_head[q]) continue;
        const int32_t gidx = sub_pos[q];
        double part = 0.0;
        int64_t e = q;
        if (have_w) {
            part = wv[q];
            e = q + 1;
            // 4 loads in flight; the adds stay strictly sequential
            while (e + 4 <= N && !(sub_head[e] | sub_head[e + 1] | sub_head[e + 2] | sub_head[e + 3]) &&
                   (keys[e + 3] & cell_mask) != sentinel) {
                const double a = wv[e], b = wv[e + 1], c = wv[e + 2], d = wv[e + 3];
                part = __dadd_rn(part, a);
                part = __dadd_rn(part, b);
                part = __dadd_rn(part, c);
                part = __dadd_rn(part, d);
                e += 4;
            }
            while (e < N && !sub_head[e] && (keys[e] & cell_mask) != sentinel) {
                part = __dadd_rn(part, wv[e]);
                ++e;
            }
        } else {
            e = q + 1;
            while (e < N && !sub_head[e] && (keys[e] & cell_mask) != sentinel) ++e;
            part = (double)(e - q);
        }
        group_sum[gidx] = part;
        group_is_cell_head[gidx] = (uint8_t)cell_head[q];
        group_key[gidx] = keys[q] & cell_mask;
    }
}

// Pass 3: add every cell's per-iteration groups in iteration order onto the running value of the dense matrix
// (the reference's `fluxMatrix = fluxMatrix + fluxMatrixI`), single writer per cell.
// A warp owns a window of 32 consecutive groups (coalesced loads).  Cells that end inside the window are summed
// lane-parallel: round u adds the value u places to the right to every head lane whose run is longer than u, so
// each cell still sees its addends strictly in order.  Only the last cell of a window can continue past it; the
// warp then walks the following windows together (one coalesced load per 32 groups, next window prefetched)
// while the owner lane keeps adding in order.  Hot cells (basis -> basis, one group per WE iteration) therefore
// cost one memory latency per 32 groups instead of one per group.
__global__ void __launch_bounds__(256)
    flux_cell_sum_kernel(const double* __restrict__ group_sum, const uint8_t* __restrict__ group_is_cell_head,
                         const uint64_t* __restrict__ group_key, const int64_t* __restrict__ n_groups_p, uint64_t CM,
                         double* __restrict__ dense, const int32_t* __restrict__ cell_pos, int64_t* __restrict__ coo_row,
                         int64_t* __restrict__ coo_col, double* __restrict__ coo_val) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t n_groups = *n_groups_p;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp_global * 32; base < n_groups; base += n_warps * 32) {
        const int64_t gi = base + lane;
        const bool in = gi < n_groups;
        const double v = in ? group_sum[gi] : 0.0;
        const bool head = in && group_is_cell_head[gi];
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t valid = __ballot_sync(0xffffffffu, in);
        // run length inside the window: distance to the next head (or to the end of the window / data)
        const uint32_t after = heads & ~((2u << lane) - 1u);          // heads strictly to the right of this lane
        const int nvalid = __popc(valid);
        const int next = after ? (__ffs(after) - 1) : nvalid;
        const int len = head ? next - lane : 0;
        const bool open_end = head && !after && base + 32 < n_groups;  // the cell may continue in the next window
        uint64_t k = 0;
        double total = 0.0, fresh = 0.0;
        if (head) {
            k = group_key[gi];
            total = dense ? dense[k] : 0.0;
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, len);
        for (int u = 0; u < maxlen; ++u) {
            const double vu = __shfl_down_sync(0xffffffffu, v, u);
            if (u < len) {
                total = __dadd_rn(total, vu);
                fresh = (u == 0) ? vu : __dadd_rn(fresh, vu);
            }
        }
        // continuation of the window's last cell
        const uint32_t open_mask = __ballot_sync(0xffffffffu, open_end);
        if (open_mask) {
            const int owner = __ffs(open_mask) - 1;
            int64_t wb = base + 32;
            bool more = true;
            double nv = (wb + lane < n_groups) ? group_sum[wb + lane] : 0.0;
            bool nh = (wb + lane < n_groups) ? group_is_cell_head[wb + lane] != 0 : true;   // past the end acts as a head
            while (more) {
                const double cv = nv;
                const uint32_t ch = __ballot_sync(0xffffffffu, nh);
                const int take = ch ? (__ffs(ch) - 1) : 32;
                more = take == 32 && wb + 32 < n_groups;
                if (more) {   // prefetch the next window before the ordered adds of this one
                    const int64_t q = wb + 32 + lane;
                    nv = (q < n_groups) ? group_sum[q] : 0.0;
                    nh = (q < n_groups) ? group_is_cell_head[q] != 0 : true;
                }
                for (int u = 0; u < take; ++u) {
                    const double vu = __shfl_sync(0xffffffffu, cv, u);
                    if (lane == owner) {
                        total = __dadd_rn(total, vu);
                        fresh = __dadd_rn(fresh, vu);
                    }
                }
                wb += 32;
            }
        }
        if (head) {
            if (dense) dense[k] = total;
            if (coo_val) {
                const int32_t pos = cell_pos[gi];
                const uint64_t r = k / CM;
                coo_row[pos] = (int64_t)r;
                coo_col[pos] = (int64_t)(k - r * CM);
                coo_val[pos] = fresh;
            }
        }
    }
}

// group -> 1 if it starts a cell (input of the COO position scan)
__global__ void __launch_bounds__(256)
    flux_cellflag_kernel(const uint8_t* __restrict__ group_is_cell_head, const int64_t* __restrict__ n_groups_p,
                         int64_t cap, int32_t* __restrict__ out) {
    const int64_t n_groups = *n_groups_p;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < cap; gi += stride)
        out[gi] = (gi < n_groups && group_is_cell_head[gi]) ? 1 : 0;
}

static size_t flux_ws_bytes(int64_t N) {
    if (N < 1) N = 1;
    size_t b = 0;
    b += align_up((size_t)N * sizeof(uint64_t), 256);  // keys
    b += align_up((size_t)N * sizeof(uint32_t), 256);  // vals
    b += 3 * align_up((size_t)N * sizeof(int32_t), 256);   // sub heads / cell heads / positions
    b += 2 * align_up((size_t)N * sizeof(double), 256);    // gathered weights, group sums
    b += align_up((size_t)N * sizeof(uint64_t), 256);      // group keys
    b += align_up((size_t)N, 256) + 256;                   // group flags, group counter
    b += align_up((size_t)(N / FM_TILE_MIN + 4) * sizeof(unsigned long long), 256);   // look-back tile states + ticket
    b += sort_workspace_bytes(N);
    b += scan_workspace_bytes(N);
    return b + 1024;
}

}  // namespace mwe

extern "C" size_t mwe_flux_workspace_bytes(int64_t N) { return mwe::flux_ws_bytes(N); }

extern "C" int mwe_flux_accumulate_f64(const int64_t* start, const int64_t* end, const uint8_t* flag0,
                                       const uint8_t* flag1, const uint8_t* col0, const uint8_t* col1, const double* w,
                                       int64_t N, int64_t n_clusters, int C, const int64_t* iter_offsets,
                                       int64_t n_iters, double* dense_inout, int64_t* coo_row, int64_t* coo_col,
                                       double* coo_val, int64_t* nnz_out, void* workspace, size_t workspace_bytes,
                                       int32_t* err_count, void* stream) {
    using namespace mwe;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MWE_REQUIRE(N >= 0 && N < ((int64_t)1 << 32), "flux: N must be < 2^32 per call");
    MWE_REQUIRE(n_clusters >= 0 && C >= 1 && C <= 255, "flux: bad n_clusters / C");
    MWE_REQUIRE(start && end && err_count, "flux: null pointer");
    MWE_REQUIRE((coo_val == nullptr) == (coo_row == nullptr) && (coo_val == nullptr) == (coo_col == nullptr),
                "flux: coo_row/coo_col/coo_val must be given together");
    MWE_REQUIRE(coo_val == nullptr || nnz_out != nullptr, "flux: COO output needs nnz_out");
    MWE_REQUIRE(iter_offsets == nullptr || n_iters >= 1, "flux: iter_offsets needs n_iters >= 1");
    const uint64_t CM = (uint64_t)C * (uint64_t)(n_clusters + 2);
    MWE_REQUIRE(CM < ((uint64_t)1 << 31), "flux: matrix side too large");
    if (N == 0) {
        if (nnz_out) MWE_CHECK_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), s));
        return MWE_OK;
    }
    if (workspace_bytes < flux_ws_bytes(N)) {
        set_last_error("flux: workspace too small (%zu < %zu)", workspace_bytes, flux_ws_bytes(N));
        return MWE_E_WORKSPACE;
    }
    Carver cv(workspace, workspace_bytes);
    uint64_t* keys = cv.take<uint64_t>((size_t)N);
    uint32_t* vals = cv.take<uint32_t>((size_t)N);
    int32_t* sub_head = cv.take<int32_t>((size_t)N);
    int32_t* cell_head = cv.take<int32_t>((size_t)N);
    int32_t* pos = cv.take<int32_t>((size_t)N);
    double* wv = cv.take<double>((size_t)N);
    double* group_sum = cv.take<double>((size_t)N);
    uint64_t* group_key = cv.take<uint64_t>((size_t)N);
    uint8_t* group_flag = cv.take<uint8_t>((size_t)N);
    int64_t* n_groups = cv.take<int64_t>(1);
    unsigned long long* tile_state = cv.take<unsigned long long>((size_t)((N + FM_TILE_MIN - 1) / FM_TILE_MIN) + 2);
    const size_t sort_bytes = sort_workspace_bytes(N);
    void* sort_ws = cv.take<char>(sort_bytes);
    const size_t scan_bytes = scan_workspace_bytes(N);
    void* scan_ws = cv.take<char>(scan_bytes);

    const uint64_t sentinel = CM * CM;  // one past the last cell
    const int key_bits = ceil_log2_u64(sentinel + 1);
    int64_t blocks = (N + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    // iteration id above the cell bits when both fit in 64 bits (always, for any realistic matrix)
    // (above the last radix digit, not just above the cell: the sort must not see iteration bits)
    int iter_shift = 0;
    const int sorted_bits = ((key_bits + 7) / 8) * 8;
    if (iter_offsets && sorted_bits + ceil_log2_u64((uint64_t)n_iters + 1) <= 63) iter_shift = sorted_bits;
    const uint64_t cell_mask = iter_shift ? (((uint64_t)1 << iter_shift) - 1) : ~(uint64_t)0;
    MWE_CHECK_CUDA(launch_pdl(flux_keys_kernel, dim3((unsigned)blocks), dim3(256), 0, s, start, end, flag0, flag1, col0, col1, N, n_clusters,
                              C, sentinel, iter_offsets, n_iters, iter_shift, keys, vals, err_count, tile_state,
                              (N + FM_TILE_MIN - 1) / FM_TILE_MIN + 1));
    uint64_t* ks;
    uint32_t* vs;
    int rc = sort_pairs(keys, vals, N, key_bits, sort_ws, sort_bytes, s, &ks, &vs);
    if (rc != MWE_OK) return rc;
    {
        // 2 consecutive elements per thread: a warp's accesses stay within two sectors per lane pair (8 per thread measured
        // 8 % slower at 3.4e7 transitions: 64-byte lane stride)
        int items = 2;
        if (const char* e = getenv("MWE_FLUX_ITEMS")) items = atoi(e) == 8 ? 8 : 2;   // tuning knob
        const int64_t ntiles = (N + FM_THREADS * items - 1) / (FM_THREADS * items);
        // the ticket lives right behind the state words of the smallest tiling (flux_keys_kernel zeroed all of them)
        unsigned int* ticket = reinterpret_cast<unsigned int*>(tile_state + (N + FM_TILE_MIN - 1) / FM_TILE_MIN);
        if (items == 2)
            MWE_CHECK_CUDA(launch_pdl(flux_mark_scan_kernel<2>, dim3((unsigned)ntiles), dim3(FM_THREADS), 0, s, ks, vs, N, sentinel,
                                      cell_mask, iter_shift != 0, w, iter_offsets, n_iters, wv, sub_head, cell_head, pos, tile_state,
                                      ticket, n_groups, err_count));
        else
            MWE_CHECK_CUDA(launch_pdl(flux_mark_scan_kernel<8>, dim3((unsigned)ntiles), dim3(FM_THREADS), 0, s, ks, vs, N, sentinel,
                                      cell_mask, iter_shift != 0, w, iter_offsets, n_iters, wv, sub_head, cell_head, pos, tile_state,
                                      ticket, n_groups, err_count));
    }
    MWE_CHECK_CUDA(launch_pdl(flux_group_sum_kernel, dim3((unsigned)blocks), dim3(256), 0, s, ks, N, sentinel, cell_mask, wv, w != nullptr,
                              sub_head, pos, cell_head, group_sum, group_flag, group_key));
    if (coo_val) {
        // COO slot of every cell = rank of its head group among the cell heads
        flux_cellflag_kernel<<<(unsigned)blocks, 256, 0, s>>>(group_flag, n_groups, N, pos);
        MWE_CHECK_LAUNCH();
        rc = exclusive_scan_i32(pos, pos, N, nnz_out, scan_ws, scan_bytes, s);
        if (rc != MWE_OK) return rc;
    }
    MWE_CHECK_CUDA(launch_pdl(flux_cell_sum_kernel, dim3((unsigned)blocks), dim3(256), 0, s, group_sum, group_flag, group_key, n_groups, CM,
                              dense_inout, pos, coo_row, coo_col, coo_val));
    return MWE_OK;
}
