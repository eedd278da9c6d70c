// K2: centroid accumulation / update (minibatch running mean and Lloyd mean), order-deterministic.
//
// reference: update_center_dense (sklearn/cluster/_k_means_minibatch.pyx:59-111), reached from
// MiniBatchKMeans.partial_fit at msm_we/_hamsm/_clustering.py:909; and the M step of
// lloyd_iter_chunked_dense (sklearn/cluster/_k_means_lloyd.pyx:23-165, _clustering.py:289,491).
// Both add x*w over a cluster's members in sample order with separately rounded product and sum;
// this kernel keeps that order and rounding (no FMA contraction), so with identical labels the
// centroids are reproduced to the last bit on one GPU.
//
//   keys   : key = cluster label (points outside [0, sumK) get a sentinel and are ignored)
//   sort   : stable radix sort by label (radix_sort.cu) -> each cluster's members contiguous, in
//            input order; segment starts are read off the sorted keys
//   sum    : one CTA per cluster; member rows staged 32 at a time through a cp.async shared-memory ring,
//            threads own feature columns and add the staged rows in member order
// HBM-bound: D*8 + 8 + 8 bytes per point per pass, partial sums written once per cluster.
#include "common.cuh"
#include "sort.cuh"

namespace mwe {

__global__ void __launch_bounds__(256)
    centroid_keys_kernel(const int64_t* __restrict__ label, int64_t N, int64_t sumK, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const int64_t l = label[i];
        keys[i] = (l >= 0 && l < sumK) ? (uint64_t)l : (uint64_t)sumK;
        vals[i] = (uint32_t)i;
    }
}

// Segment starts straight from the sorted keys (no per-label atomics, no scan): wherever the key changes at
// position q, every label in (previous key, key] starts at q; labels above the last key start at N.
// seg_start has sumK + 2 entries; [sumK] is where the ignored points (sentinel key) begin.
__global__ void __launch_bounds__(256)
    centroid_bounds_kernel(const uint64_t* __restrict__ keys, int64_t N, int64_t sumK, int32_t* __restrict__ seg_start) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += stride) {
        const int64_t k = (int64_t)keys[q];
        const int64_t kp = q ? (int64_t)keys[q - 1] : -1;
        for (int64_t l = kp + 1; l <= k; ++l) seg_start[l] = (int32_t)q;
        if (q == N - 1)
            for (int64_t l = k + 1; l <= sumK + 1; ++l) seg_start[l] = (int32_t)N;
    }
}

// Launch order of the per-cluster CTAs: largest clusters first.  One CTA adds one cluster's member rows in sample order,
// so a cluster with 10^4 members keeps its CTA busy for most of the kernel; in cluster-index order the last such
// cluster to start IS the tail (config-5 shape, Lloyd on WE data: a few clusters per bin hold a quarter of the bin's
// points; 2.2-2.8 ms per 4e6 points against 1.4 ms with evenly sized clusters).  Counting sort by size class (quarter
// octaves, descending) in one CTA; the position inside a class is arbitrary, which changes the schedule, never a sum.
static constexpr int CS_CLASSES = 128;
static constexpr int CS_BIG = 2048, CS_SPLIT = 4;   // see centroid_sum_kernel
__device__ __forceinline__ int cs_size_class(int n) {
    if (n <= 0) return CS_CLASSES;
    const int lg = 31 - __clz(n);
    const int sub = lg >= 2 ? ((n >> (lg - 2)) & 3) : ((n << (2 - lg)) & 3);
    return CS_CLASSES - 1 - (lg * 4 + sub);
}
__global__ void __launch_bounds__(1024)
    centroid_order_kernel(const int32_t* __restrict__ seg_start, int32_t sumK, int32_t* __restrict__ order) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ int hist[CS_CLASSES + 1];
    for (int i = threadIdx.x; i <= CS_CLASSES; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < sumK; k += blockDim.x) atomicAdd(&hist[cs_size_class(seg_start[k + 1] - seg_start[k])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i <= CS_CLASSES; ++i) {
            const int c = hist[i];
            hist[i] = run;
            run += c;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) order[sumK] = hist[cs_size_class(CS_BIG) + 1];   // classes of >= CS_BIG members come first
    __syncthreads();
    for (int k = threadIdx.x; k < sumK; k += blockDim.x)
        order[atomicAdd(&hist[cs_size_class(seg_start[k + 1] - seg_start[k])], 1)] = k;
}

enum CentroidMode { CM_ACCUMULATE = 0, CM_MINIBATCH = 1 };

static constexpr int CS_ST = 3;       // ring depth: two chunks in flight per CTA while one is being added
static constexpr int CS_SUPER = 8;    // chunks per metadata super-chunk
// Two geometries.  FT = feature columns handled per pass over a cluster's member list (thread f < FT owns column
// ft*FT + f).  Narrow rows (D <= 64... cfg2) use 64 columns x 24 staged rows; wide rows (cfg5: 256, cfg3: 3000) use 256
// columns x 6 rows, so a cluster's member list is walked ONCE per 256 features instead of once per 64 -- with a few
// hundred members per cluster the per-pass prologue (two dependent metadata loads + pipeline fill) is what costs.
template <int FT> struct CsGeom;
template <> struct CsGeom<64> { static constexpr int THREADS = 128, ROWS = 24; };
template <> struct CsGeom<256> { static constexpr int THREADS = 288, ROWS = 6; };   // 256 owners + one warp (weight sum)

// One CTA per cluster.  The add chain of a (cluster, feature) pair is sequential by definition (sample order, product
// and sum rounded separately), so the parallelism is clusters x features; what must not be serial is the MEMORY
// side.  Member rows are therefore staged in chunks of 32 through a double-buffered shared-memory ring with cp.async
// (all 128 threads copy, coalesced row segments), the member indices and weights of the chunk after next are
// prefetched into registers, and the 64 feature threads only ever read shared memory.  One extra thread adds the
// weights in member order while the others work (the first version made every thread walk the whole weight list
// through two dependent global loads per member before it started: 416 us on the cfg2 shape, now HBM-bound).
// Clusters with at least CS_BIG members are SPLIT over CS_SPLIT CTAs by feature columns (wide geometry, accumulate
// mode): a (cluster, feature) add chain stays sequential, but four CTAs pull the cluster's rows instead of one.  For
// D <= 768 the four work on interleaved 64-column blocks with 24 staged rows (the same shared-memory footprint
// reinterpreted), above that on interleaved 256-column blocks.  The order kernel puts the big clusters first and
// leaves their number in order[sumK].
template <int MODE, int VEC, int CS_FT>
__global__ void __launch_bounds__(CsGeom<CS_FT>::THREADS, CS_FT == 256 ? 4 : 5)
    centroid_sum_kernel(const double* __restrict__ X, int64_t ldx, int D, const double* __restrict__ w,
                        const uint32_t* __restrict__ members, const int32_t* __restrict__ seg_start,
                        const int32_t* __restrict__ order, int32_t sumK, double* __restrict__ out_wx,
                        double* __restrict__ out_w) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int CS_THREADS = CsGeom<CS_FT>::THREADS;
    constexpr int CS_ROWS = CsGeom<CS_FT>::ROWS;
    constexpr bool CAN_SPLIT = (MODE == CM_ACCUMULATE) && CS_FT == 256;
    constexpr int SROWS_MAX = CS_SUPER * CS_ROWS * (CAN_SPLIT ? CS_FT / 64 : 1);   // members whose index + weight are staged together
    static_assert(SROWS_MAX <= 2 * CS_THREADS, "two metadata entries per thread");
    __shared__ __align__(16) double s_x[CS_ST][CS_ROWS * CS_FT];
    __shared__ double s_w[2][SROWS_MAX];
    __shared__ uint32_t s_idx[2][SROWS_MAX];
    __shared__ double s_wsum;
    // work item -> (cluster, first column block j, stride S)
    int j = 0, S = 1;
    int64_t k;
    {
        const int n_big = CAN_SPLIT ? order[sumK] : 0;
        const int item = blockIdx.x;
        if (item < n_big * CS_SPLIT) {
            k = order[item / CS_SPLIT];
            j = item % CS_SPLIT;
            S = CS_SPLIT;
        } else {
            const int idx = item - n_big * (CS_SPLIT - 1);
            if (idx >= sumK) return;
            k = order[idx];
        }
    }
    const bool narrow_blocks = CAN_SPLIT && S > 1 && D <= 3 * CS_FT;
    const int FTe = narrow_blocks ? 64 : CS_FT;               // columns per block of this item
    const int ROWSe = narrow_blocks ? CS_ROWS * (CS_FT / 64) : CS_ROWS;
    const int SROWSe = CS_SUPER * ROWSe;
    const int seg_shift = (narrow_blocks ? 6 : (CS_FT == 256 ? 8 : 6)) - (VEC == 2 ? 1 : 0);   // log2(copy segments per row)
    const int32_t s = seg_start[k], e = seg_start[k + 1];
    const int n = e - s;
    const int nchunks = (n + ROWSe - 1) / ROWSe;
    const int tid = threadIdx.x;
    const double count0 = (MODE == CM_MINIBATCH) ? out_w[k] : 0.0;
    const int n_ft = (D + FTe - 1) / FTe;
    double wsum = 0.0;                                // thread CS_THREADS-1: sum of the weights in member order
    bool first = true;

    for (int ft = j; ft < n_ft; ft += S) {
        const int col0 = ft * FTe;
        const int ncols = min(FTe, D - col0);
        // Member indices and weights (two dependent global loads per member) are fetched a whole super-chunk
        // (CS_SUPER chunks) ahead into registers and published to shared memory when the copies reach it, so that
        // chain is paid once per super-chunk, behind 8 chunks of work, instead of once per chunk.
        uint32_t idx_r[2];
        double w_r[2];
        auto fetch_super = [&](int sc) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = sc * SROWSe + h * CS_THREADS + tid;
                idx_r[h] = 0u;
                w_r[h] = 0.0;
                if (h * CS_THREADS + tid < SROWSe && m < n) {
                    idx_r[h] = members[s + m];
                    w_r[h] = w ? w[idx_r[h]] : 1.0;
                }
            }
        };
        auto publish_super = [&](int sc) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (h * CS_THREADS + tid < SROWSe) {
                    s_idx[sc & 1][h * CS_THREADS + tid] = idx_r[h];
                    s_w[sc & 1][h * CS_THREADS + tid] = w_r[h];
                }
            __syncthreads();
        };
        auto copy_chunk = [&](int c) {
            const int sc = c / CS_SUPER, r0 = (c % CS_SUPER) * ROWSe;
            const int rows = min(ROWSe, n - c * ROWSe);
            for (int u = tid; u < (rows << seg_shift); u += CS_THREADS) {
                const int r = u >> seg_shift, sg = u - (r << seg_shift);
                const int c_in = sg * VEC;
                if (c_in < ncols) {
                    const double* src = X + (int64_t)s_idx[sc & 1][r0 + r] * ldx + col0 + c_in;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_x[c % CS_ST][r * FTe + c_in]);
                    if (VEC == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                }
            }
        };
        double acc = 0.0;
        const bool owner = tid < ncols;
        if (MODE == CM_MINIBATCH && owner) acc = __dmul_rn(out_wx[k * D + col0 + tid], count0);
        // copies run CS_ST-1 chunks ahead of the adds; one commit group per trip (possibly empty) keeps the group
        // count in step with the chunk count
        fetch_super(0);
        publish_super(0);
        fetch_super(1);
        for (int c = 0; c < CS_ST - 1; ++c) {
            if (c < nchunks) copy_chunk(c);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int c = 0; c < nchunks; ++c) {
            const int cn = c + CS_ST - 1;                     // chunk whose copies are issued this trip
            if (cn < nchunks) {
                if (cn % CS_SUPER == 0) {                     // first chunk of a new super-chunk: its metadata is in registers
                    publish_super(cn / CS_SUPER);
                    fetch_super(cn / CS_SUPER + 1);
                }
                copy_chunk(cn);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(CS_ST - 1) : "memory");
            __syncthreads();
            const int rows = min(ROWSe, n - c * ROWSe);
            const double* wc = &s_w[(c / CS_SUPER) & 1][(c % CS_SUPER) * ROWSe];
            if (owner) {
                const double* xs = &s_x[c % CS_ST][tid];
                for (int r = 0; r < rows; ++r) acc = __dadd_rn(acc, __dmul_rn(xs[r * FTe], wc[r]));
            } else if (tid == CS_THREADS - 1 && first) {
                for (int r = 0; r < rows; ++r) wsum = __dadd_rn(wsum, wc[r]);
            }
            __syncthreads();     // the chunk is consumed before its buffers are refilled
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (first) {
            if (tid == CS_THREADS - 1) s_wsum = wsum;
            __syncthreads();
            first = false;
        }
        const double ws = s_wsum;
        if (MODE == CM_MINIBATCH) {
            if (!(ws > 0.0)) continue;   // centre untouched (_k_means_minibatch.pyx:108-111)
            if (owner) out_wx[k * D + col0 + tid] = __dmul_rn(acc, 1.0 / __dadd_rn(count0, ws));
        } else if (owner) {
            out_wx[k * D + col0 + tid] = acc;
        }
    }
    if (tid == 0 && j == 0) {
        const double ws = s_wsum;
        if (MODE == CM_MINIBATCH) { if (ws > 0.0) out_w[k] = __dadd_rn(count0, ws); }
        else out_w[k] = ws;
    }
}

__global__ void __launch_bounds__(256)
    lloyd_finalize_kernel(const double* __restrict__ sum_wx, const double* __restrict__ sum_w, int64_t sumK, int D,
                          double* __restrict__ centers) {
    const int64_t total = sumK * D;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t k = i / D;
        const double ws = sum_w[k];
        // _average_centers: alpha = 1/weight; centre *= alpha; empty clusters keep their centre
        if (ws > 0.0) centers[i] = __dmul_rn(sum_wx[i], 1.0 / ws);
    }
}

__global__ void __launch_bounds__(256)
    minibatch_finalize_kernel(const double* __restrict__ sum_wx, const double* __restrict__ sum_w, int64_t sumK, int D,
                              double* __restrict__ centers, const double* __restrict__ counts) {
    const int64_t total = sumK * D;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t k = i / D;
        const double ws = sum_w[k];
        if (ws > 0.0) {
            const double c0 = counts[k];
            const double acc = __dadd_rn(__dmul_rn(centers[i], c0), sum_wx[i]);
            centers[i] = __dmul_rn(acc, 1.0 / __dadd_rn(c0, ws));
        }
    }
}

__global__ void __launch_bounds__(256)
    counts_add_kernel(double* __restrict__ counts, const double* __restrict__ sum_w, int64_t sumK) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < sumK && sum_w[i] > 0.0) counts[i] = __dadd_rn(counts[i], sum_w[i]);
}

static size_t centroid_ws_bytes(int64_t N, int64_t sumK) {
    if (N < 1) N = 1;
    size_t b = 0;
    b += align_up((size_t)N * sizeof(uint64_t), 256);
    b += align_up((size_t)N * sizeof(uint32_t), 256);
    b += 2 * align_up((size_t)(sumK + 2) * sizeof(int32_t), 256);
    b += sort_workspace_bytes(N);
    b += scan_workspace_bytes(sumK + 2);
    return b + 1024;
}

template <int MODE>
static int centroid_run(const double* X, int64_t N, int D, int64_t ldx, const double* w, const int64_t* label,
                        int64_t sumK, double* out_wx, double* out_w, void* workspace, size_t workspace_bytes,
                        cudaStream_t s) {
    MWE_REQUIRE(N >= 0 && N < ((int64_t)1 << 31), "centroid: N must be < 2^31 per call");
    MWE_REQUIRE(D >= 1 && ldx >= D && sumK >= 1 && sumK < ((int64_t)1 << 31), "centroid: bad shape");
    MWE_REQUIRE(X && label && out_wx && out_w, "centroid: null pointer");
    if (workspace_bytes < centroid_ws_bytes(N, sumK)) {
        set_last_error("centroid: workspace too small (%zu < %zu)", workspace_bytes, centroid_ws_bytes(N, sumK));
        return MWE_E_WORKSPACE;
    }
    Carver cv(workspace, workspace_bytes);
    uint64_t* keys = cv.take<uint64_t>((size_t)(N > 0 ? N : 1));
    uint32_t* vals = cv.take<uint32_t>((size_t)(N > 0 ? N : 1));
    int32_t* order = cv.take<int32_t>((size_t)sumK + 2);
    int32_t* seg_start = cv.take<int32_t>((size_t)sumK + 2);
    const size_t sort_bytes = sort_workspace_bytes(N);
    void* sort_ws = cv.take<char>(sort_bytes);

    uint64_t* ks = keys;
    uint32_t* vs = vals;
    if (N > 0) {
        int64_t blocks = (N + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        MWE_CHECK_CUDA(launch_pdl(centroid_keys_kernel, dim3((unsigned)blocks), dim3(256), 0, s, label, N, sumK, keys, vals));
        int rc = sort_pairs(keys, vals, N, ceil_log2_u64((uint64_t)sumK + 1), sort_ws, sort_bytes, s, &ks, &vs);
        if (rc != MWE_OK) return rc;
        MWE_CHECK_CUDA(launch_pdl(centroid_bounds_kernel, dim3((unsigned)blocks), dim3(256), 0, s, ks, N, sumK, seg_start));
    } else {
        MWE_CHECK_CUDA(cudaMemsetAsync(seg_start, 0, (size_t)(sumK + 2) * sizeof(int32_t), s));
    }
    MWE_CHECK_CUDA(launch_pdl(centroid_order_kernel, dim3(1), dim3(1024), 0, s, seg_start, (int32_t)sumK, order));
    const bool vec2 = (D % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const bool wide = D > 96;
    // wide accumulate launches carry the extra items of the split clusters (at most N / CS_BIG of them)
    const int64_t max_big = (wide && MODE == CM_ACCUMULATE) ? std::min<int64_t>(sumK, N / CS_BIG) : 0;
    const dim3 g((unsigned)(sumK + max_big * (CS_SPLIT - 1)));
    if (wide) {
        const dim3 b(CsGeom<256>::THREADS);
        if (vec2) MWE_CHECK_CUDA(launch_pdl(centroid_sum_kernel<MODE, 2, 256>, g, b, 0, s, X, ldx, D, w, vs, seg_start, order, (int32_t)sumK, out_wx, out_w));
        else MWE_CHECK_CUDA(launch_pdl(centroid_sum_kernel<MODE, 1, 256>, g, b, 0, s, X, ldx, D, w, vs, seg_start, order, (int32_t)sumK, out_wx, out_w));
    } else {
        const dim3 b(CsGeom<64>::THREADS);
        if (vec2) MWE_CHECK_CUDA(launch_pdl(centroid_sum_kernel<MODE, 2, 64>, g, b, 0, s, X, ldx, D, w, vs, seg_start, order, (int32_t)sumK, out_wx, out_w));
        else MWE_CHECK_CUDA(launch_pdl(centroid_sum_kernel<MODE, 1, 64>, g, b, 0, s, X, ldx, D, w, vs, seg_start, order, (int32_t)sumK, out_wx, out_w));
    }
    return MWE_OK;
}


// ---- label-keyed grouping and per-label statistics ---------------------------------
// reference: ClusteringMixin.get_cluster_centers (msm_we/_hamsm/_clustering.py:1528-1599) walks all dtrajs once
// PER CLUSTER (np.where over every iteration) and takes nanmean / nanmin / nanmax of the members' end pcoords;
// update_cluster_structures (:1398-1526) appends every segment to a per-cluster Python list.  Both are a
// group-by-label: the stable sort of K2 yields each label's members contiguous and in input order.

// One warp per label.  Lane l takes members s+l, s+l+32, ... in order; the 32 partials are combined with a fixed
// butterfly, so the result does not depend on scheduling.  NaN values are skipped (numpy nan* semantics).
__global__ void __launch_bounds__(256)
    label_stats_kernel(const double* __restrict__ v, int64_t ldv, const uint32_t* __restrict__ members,
                       const int32_t* __restrict__ seg_start, int64_t n_labels, int64_t* __restrict__ count,
                       double* __restrict__ sum, double* __restrict__ mn, double* __restrict__ mx) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); k < n_labels; k += warps) {
        const int32_t s = seg_start[k], e = seg_start[k + 1];
        double acc = 0.0, lo = __longlong_as_double(0x7ff0000000000000LL), hi = __longlong_as_double(0xfff0000000000000LL);
        long long c = 0;
        for (int32_t m = s + lane; m < e; m += 32) {
            const double x = v[(int64_t)members[m] * ldv];
            if (x == x) {
                acc += x;
                lo = fmin(lo, x);
                hi = fmax(hi, x);
                ++c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            count[k] = c;
            sum[k] = acc;
            mn[k] = lo;
            mx[k] = hi;
        }
    }
}

// Few labels with many members each (the per-WE-bin maximum of the relocation step: 100 labels x 10^4 members): one
// CTA per label instead of one warp.  Thread t takes members s+t, s+t+256, ...; fixed butterfly inside each warp, then
// the eight warp partials in warp order: as deterministic as the kernel above, with a different (fixed) association.
__global__ void __launch_bounds__(256)
    label_stats_block_kernel(const double* __restrict__ v, int64_t ldv, const uint32_t* __restrict__ members,
                             const int32_t* __restrict__ seg_start, int64_t n_labels, int64_t* __restrict__ count,
                             double* __restrict__ sum, double* __restrict__ mn, double* __restrict__ mx) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ double s_acc[8], s_lo[8], s_hi[8];
    __shared__ long long s_c[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t k = blockIdx.x; k < n_labels; k += gridDim.x) {
        const int32_t s = seg_start[k], e = seg_start[k + 1];
        double acc = 0.0, lo = __longlong_as_double(0x7ff0000000000000LL), hi = __longlong_as_double(0xfff0000000000000LL);
        long long c = 0;
        for (int32_t m = s + (int32_t)threadIdx.x; m < e; m += 256) {
            const double x = v[(int64_t)members[m] * ldv];
            if (x == x) {
                acc += x;
                lo = fmin(lo, x);
                hi = fmax(hi, x);
                ++c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            s_acc[warp] = acc;
            s_lo[warp] = lo;
            s_hi[warp] = hi;
            s_c[warp] = c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) {
                acc += s_acc[w];
                lo = fmin(lo, s_lo[w]);
                hi = fmax(hi, s_hi[w]);
                c += s_c[w];
            }
            count[k] = c;
            sum[k] = acc;
            mn[k] = lo;
            mx[k] = hi;
        }
        __syncthreads();
    }
}

// Squared distance of listed points to their own centre (empty-cluster relocation of the Lloyd M step: sklearn moves an
// empty cluster onto the point farthest from its centre, _k_means_common.pyx _relocate_empty_clusters_dense).  One warp
// per listed point, lanes stride the features, fixed butterfly.
template <int VEC>
__global__ void __launch_bounds__(256, 4)
    point_center_dist2_kernel(const double* __restrict__ X, int64_t ldx, int D, const int32_t* __restrict__ list, int64_t n,
                              const int64_t* __restrict__ label, const double* __restrict__ centers, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < n; e += warps) {
        const int64_t pt = list[e];
        const double* x = X + pt * ldx;
        const double* c = centers + label[pt] * D;
        double acc = 0.0;
        if (VEC == 2) {
            // 16-byte loads, four row pieces in flight per lane before the first add (the order of the adds is fixed)
            double acc1 = 0.0;
            int k = 2 * lane;
#pragma unroll 1
            for (; k + 192 < D; k += 256) {
                const double2 x0 = *reinterpret_cast<const double2*>(x + k), x1 = *reinterpret_cast<const double2*>(x + k + 64),
                              x2 = *reinterpret_cast<const double2*>(x + k + 128), x3 = *reinterpret_cast<const double2*>(x + k + 192);
                const double2 c0 = *reinterpret_cast<const double2*>(c + k), c1 = *reinterpret_cast<const double2*>(c + k + 64),
                              c2 = *reinterpret_cast<const double2*>(c + k + 128), c3 = *reinterpret_cast<const double2*>(c + k + 192);
                double d;
                d = x0.x - c0.x; acc = fma(d, d, acc);  d = x0.y - c0.y; acc1 = fma(d, d, acc1);
                d = x1.x - c1.x; acc = fma(d, d, acc);  d = x1.y - c1.y; acc1 = fma(d, d, acc1);
                d = x2.x - c2.x; acc = fma(d, d, acc);  d = x2.y - c2.y; acc1 = fma(d, d, acc1);
                d = x3.x - c3.x; acc = fma(d, d, acc);  d = x3.y - c3.y; acc1 = fma(d, d, acc1);
            }
#pragma unroll 1
            for (; k < D; k += 64) {
                const double2 xv = *reinterpret_cast<const double2*>(x + k), cv = *reinterpret_cast<const double2*>(c + k);
                double d = xv.x - cv.x;
                acc = fma(d, d, acc);
                d = xv.y - cv.y;
                acc1 = fma(d, d, acc1);
            }
            acc += acc1;
        } else {
            for (int k = lane; k < D; k += 32) {
                const double d = x[k] - c[k];
                acc = fma(d, d, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[e] = acc;
    }
}

// The k largest values of each listed segment (k <= ST_KMAX), largest first, equal values in member order -- what a
// stable descending sort of the segment would put first (the farthest points of a WE bin that lost several clusters:
// ranking every point of every such bin took two device-wide sorts).  One CTA per listed segment: every thread keeps
// the best k of its strided share in registers, then k rounds of a block-wide argmax over the threads' current heads.
static constexpr int ST_KMAX = 8;
__device__ __forceinline__ bool st_better(double va, int pa, double vb, int pb) { return va > vb || (va == vb && pa < pb); }
__global__ void __launch_bounds__(256)
    segment_topk_kernel(const double* __restrict__ values, const uint32_t* __restrict__ members, const int32_t* __restrict__ seg_start,
                        const int32_t* __restrict__ seg_ids, int k, int32_t* __restrict__ out_pos, double* __restrict__ out_val) {
    __shared__ double s_v[256 * ST_KMAX];
    __shared__ int s_p[256 * ST_KMAX];
    __shared__ double s_wv[8];
    __shared__ int s_wp[8], s_wt[8];
    const int seg = seg_ids[blockIdx.x];
    const int32_t s = seg_start[seg], e = seg_start[seg + 1];
    const double ninf = __longlong_as_double(0xfff0000000000000LL);
    double tv[ST_KMAX];
    int tp[ST_KMAX];
#pragma unroll
    for (int j = 0; j < ST_KMAX; ++j) {
        tv[j] = ninf;
        tp[j] = 0x7fffffff;
    }
    for (int32_t m = s + (int32_t)threadIdx.x; m < e; m += 256) {
        int cp = (int)members[m];
        double cvv = values[cp];
        if (cvv == cvv && st_better(cvv, cp, tv[ST_KMAX - 1], tp[ST_KMAX - 1])) {
#pragma unroll
            for (int j = 0; j < ST_KMAX; ++j)
                if (st_better(cvv, cp, tv[j], tp[j])) {
                    const double a = tv[j];
                    const int b = tp[j];
                    tv[j] = cvv;
                    tp[j] = cp;
                    cvv = a;
                    cp = b;
                }
        }
    }
#pragma unroll
    for (int j = 0; j < ST_KMAX; ++j) {
        s_v[threadIdx.x * ST_KMAX + j] = tv[j];
        s_p[threadIdx.x * ST_KMAX + j] = tp[j];
    }
    __syncthreads();
    int cursor = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < k; ++r) {
        double v = cursor < ST_KMAX ? s_v[threadIdx.x * ST_KMAX + cursor] : ninf;
        int p = cursor < ST_KMAX ? s_p[threadIdx.x * ST_KMAX + cursor] : 0x7fffffff;
        int t = (int)threadIdx.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int op = __shfl_xor_sync(0xffffffffu, p, o), ot = __shfl_xor_sync(0xffffffffu, t, o);
            if (st_better(ov, op, v, p)) {
                v = ov;
                p = op;
                t = ot;
            }
        }
        if (lane == 0) {
            s_wv[warp] = v;
            s_wp[warp] = p;
            s_wt[warp] = t;
        }
        __syncthreads();
        v = s_wv[0];
        p = s_wp[0];
        t = s_wt[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            if (st_better(s_wv[w], s_wp[w], v, p)) {
                v = s_wv[w];
                p = s_wp[w];
                t = s_wt[w];
            }
        if ((int)threadIdx.x == t && p != 0x7fffffff) ++cursor;
        if (threadIdx.x == 0) {
            out_pos[(int64_t)blockIdx.x * k + r] = (p == 0x7fffffff) ? -1 : p;     // fewer than k members: -1
            out_val[(int64_t)blockIdx.x * k + r] = v;
        }
        __syncthreads();
    }
}

static int group_by_label(const int64_t* label, int64_t N, int64_t n_labels, uint32_t* members_out, int32_t* seg_start_out,
                          void* workspace, size_t workspace_bytes, cudaStream_t s) {
    MWE_REQUIRE(N >= 0 && N < ((int64_t)1 << 31), "group_by_label: N must be < 2^31 per call");
    MWE_REQUIRE(n_labels >= 1 && n_labels < ((int64_t)1 << 31), "group_by_label: bad label count");
    MWE_REQUIRE(members_out && seg_start_out && (label || N == 0), "group_by_label: null pointer");
    if (workspace_bytes < centroid_ws_bytes(N, n_labels)) {
        set_last_error("group_by_label: workspace too small (%zu < %zu)", workspace_bytes, centroid_ws_bytes(N, n_labels));
        return MWE_E_WORKSPACE;
    }
    if (N == 0) {
        MWE_CHECK_CUDA(cudaMemsetAsync(seg_start_out, 0, (size_t)(n_labels + 2) * sizeof(int32_t), s));
        return MWE_OK;
    }
    Carver cv(workspace, workspace_bytes);
    uint64_t* keys = cv.take<uint64_t>((size_t)N);
    uint32_t* vals = cv.take<uint32_t>((size_t)N);
    (void)cv.take<int32_t>((size_t)n_labels + 2);
    (void)cv.take<int32_t>((size_t)n_labels + 2);
    const size_t sort_bytes = sort_workspace_bytes(N);
    void* sort_ws = cv.take<char>(sort_bytes);
    uint64_t* ks = keys;
    uint32_t* vs = vals;
    int64_t blocks = (N + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    MWE_CHECK_CUDA(launch_pdl(centroid_keys_kernel, dim3((unsigned)blocks), dim3(256), 0, s, label, N, n_labels, keys, vals));
    int rc = sort_pairs(keys, vals, N, ceil_log2_u64((uint64_t)n_labels + 1), sort_ws, sort_bytes, s, &ks, &vs);
    if (rc != MWE_OK) return rc;
    MWE_CHECK_CUDA(launch_pdl(centroid_bounds_kernel, dim3((unsigned)blocks), dim3(256), 0, s, ks, N, n_labels, seg_start_out));
    MWE_CHECK_CUDA(cudaMemcpyAsync(members_out, vs, (size_t)N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    return MWE_OK;
}

}  // namespace mwe

extern "C" size_t mwe_centroid_workspace_bytes(int64_t N, int64_t sumK) { return mwe::centroid_ws_bytes(N, sumK); }

extern "C" int mwe_centroid_accumulate_f64(const double* X, int64_t N, int D, int64_t ldx, const double* w,
                                           const int64_t* label, int64_t sumK, double* sum_wx, double* sum_w,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    return mwe::centroid_run<mwe::CM_ACCUMULATE>(X, N, D, ldx, w, label, sumK, sum_wx, sum_w, workspace,
                                                 workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int mwe_minibatch_update_f64(const double* X, int64_t N, int D, int64_t ldx, const double* w,
                                        const int64_t* label, int64_t sumK, double* centers, double* counts,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    return mwe::centroid_run<mwe::CM_MINIBATCH>(X, N, D, ldx, w, label, sumK, centers, counts, workspace,
                                                workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int mwe_lloyd_finalize_f64(const double* sum_wx, const double* sum_w, int64_t sumK, int D, double* centers,
                                      void* stream) {
    using namespace mwe;
    MWE_REQUIRE(sumK >= 0 && D >= 1, "lloyd_finalize: bad shape");
    if (sumK == 0) return MWE_OK;
    int64_t blocks = (sumK * D + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    lloyd_finalize_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(sum_wx, sum_w, sumK, D, centers);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

extern "C" int mwe_minibatch_finalize_f64(const double* sum_wx, const double* sum_w, int64_t sumK, int D,
                                          double* centers, double* counts, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(sumK >= 0 && D >= 1, "minibatch_finalize: bad shape");
    if (sumK == 0) return MWE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int64_t blocks = (sumK * D + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    minibatch_finalize_kernel<<<(unsigned)blocks, 256, 0, s>>>(sum_wx, sum_w, sumK, D, centers, counts);
    counts_add_kernel<<<(unsigned)((sumK + 255) / 256), 256, 0, s>>>(counts, sum_w, sumK);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

extern "C" int mwe_group_by_label(const int64_t* label, int64_t N, int64_t n_labels, uint32_t* members_out,
                                  int32_t* seg_start_out, void* workspace, size_t workspace_bytes, void* stream) {
    return mwe::group_by_label(label, N, n_labels, members_out, seg_start_out, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int mwe_label_stats_f64(const double* values, int64_t ldv, const uint32_t* members, const int32_t* seg_start,
                                   int64_t n_labels, int64_t* count, double* sum, double* vmin, double* vmax, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(n_labels >= 0 && ldv >= 1, "label_stats: bad shape");
    if (n_labels == 0) return MWE_OK;
    MWE_REQUIRE(values && members && seg_start && count && sum && vmin && vmax, "label_stats: null pointer");
    if (n_labels <= 512) {   // few labels: a CTA each
        MWE_CHECK_CUDA(launch_pdl(label_stats_block_kernel, dim3((unsigned)n_labels), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                  values, ldv, members, seg_start, n_labels, count, sum, vmin, vmax));
        return MWE_OK;
    }
    int64_t blocks = (n_labels + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    MWE_CHECK_CUDA(launch_pdl(label_stats_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                              values, ldv, members, seg_start, n_labels, count, sum, vmin, vmax));
    return MWE_OK;
}

extern "C" int mwe_point_center_dist2_f64(const double* X, int64_t ldx, int D, const int32_t* list, int64_t n,
                                          const int64_t* label, const double* centers, double* out, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(n >= 0 && D >= 1 && ldx >= D, "point_center_dist2: bad shape");
    if (n == 0) return MWE_OK;
    MWE_REQUIRE(X && list && label && centers && out, "point_center_dist2: null pointer");
    int64_t blocks = (n + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const bool vec2 = (D % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(centers) & 15) == 0);
    if (vec2) point_center_dist2_kernel<2><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, D, list, n, label, centers, out);
    else point_center_dist2_kernel<1><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, ldx, D, list, n, label, centers, out);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

extern "C" int mwe_segment_topk_f64(const double* values, const uint32_t* members, const int32_t* seg_start, const int32_t* seg_ids,
                                    int32_t n_sel, int k, int32_t* out_pos, double* out_val, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(n_sel >= 0 && k >= 1 && k <= ST_KMAX, "segment_topk: k must be in [1, 8]");
    if (n_sel == 0) return MWE_OK;
    MWE_REQUIRE(values && members && seg_start && seg_ids && out_pos && out_val, "segment_topk: null pointer");
    segment_topk_kernel<<<(unsigned)n_sel, 256, 0, static_cast<cudaStream_t>(stream)>>>(values, members, seg_start, seg_ids, k, out_pos, out_val);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}
