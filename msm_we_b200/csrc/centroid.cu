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
//            input order; a histogram + scan gives every cluster's segment start
//   sum    : one CTA per cluster; threads own feature columns (coalesced row reads), members are
//            visited in order, 4 rows in flight per thread
// HBM-bound: D*8 + 8 + 8 bytes per point per pass, partial sums written once per cluster.
#include "common.cuh"
#include "sort.cuh"

namespace mwe {

__global__ void __launch_bounds__(256)
    centroid_keys_kernel(const int64_t* __restrict__ label, int64_t N, int64_t sumK, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ vals, int32_t* __restrict__ count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const int64_t l = label[i];
        const bool ok = l >= 0 && l < sumK;
        keys[i] = ok ? (uint64_t)l : (uint64_t)sumK;
        vals[i] = (uint32_t)i;
        if (ok) atomicAdd(&count[l], 1);
    }
}

enum CentroidMode { CM_ACCUMULATE = 0, CM_MINIBATCH = 1 };

static constexpr int CS_THREADS = 128;

template <int MODE>
__global__ void __launch_bounds__(CS_THREADS)
    centroid_sum_kernel(const double* __restrict__ X, int64_t ldx, int D, const double* __restrict__ w,
                        const uint32_t* __restrict__ members, const int32_t* __restrict__ seg_start,
                        double* __restrict__ out_wx, double* __restrict__ out_w) {
    const int64_t k = blockIdx.x;
    const int32_t s = seg_start[k], e = seg_start[k + 1];
    // weight sum in member order (every thread computes the same value)
    double wsum = 0.0;
    for (int32_t q = s; q < e; ++q) wsum = __dadd_rn(wsum, w ? w[members[q]] : 1.0);
    double count0 = 0.0;
    if (MODE == CM_MINIBATCH) {
        count0 = out_w[k];
        __syncthreads();  // everyone has read the old count before thread 0 overwrites it
        if (!(wsum > 0.0)) return;  // centre untouched (_k_means_minibatch.pyx:108-111)
    }
    const double new_count = __dadd_rn(count0, wsum);
    const double alpha = 1.0 / new_count;
    for (int f = threadIdx.x; f < D; f += CS_THREADS) {
        double acc = 0.0;
        if (MODE == CM_MINIBATCH) acc = __dmul_rn(out_wx[k * D + f], count0);
        int32_t q = s;
        for (; q + 4 <= e; q += 4) {
            const uint32_t i0 = members[q], i1 = members[q + 1], i2 = members[q + 2], i3 = members[q + 3];
            const double x0 = X[(int64_t)i0 * ldx + f], x1 = X[(int64_t)i1 * ldx + f];
            const double x2 = X[(int64_t)i2 * ldx + f], x3 = X[(int64_t)i3 * ldx + f];
            const double w0 = w ? w[i0] : 1.0, w1 = w ? w[i1] : 1.0, w2 = w ? w[i2] : 1.0, w3 = w ? w[i3] : 1.0;
            acc = __dadd_rn(acc, __dmul_rn(x0, w0));
            acc = __dadd_rn(acc, __dmul_rn(x1, w1));
            acc = __dadd_rn(acc, __dmul_rn(x2, w2));
            acc = __dadd_rn(acc, __dmul_rn(x3, w3));
        }
        for (; q < e; ++q) {
            const uint32_t i0 = members[q];
            acc = __dadd_rn(acc, __dmul_rn(X[(int64_t)i0 * ldx + f], w ? w[i0] : 1.0));
        }
        if (MODE == CM_MINIBATCH) out_wx[k * D + f] = __dmul_rn(acc, alpha);
        else out_wx[k * D + f] = acc;
    }
    if (threadIdx.x == 0) out_w[k] = (MODE == CM_MINIBATCH) ? new_count : wsum;
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
    int32_t* count = cv.take<int32_t>((size_t)sumK + 2);
    int32_t* seg_start = cv.take<int32_t>((size_t)sumK + 2);
    const size_t sort_bytes = sort_workspace_bytes(N);
    void* sort_ws = cv.take<char>(sort_bytes);
    const size_t scan_bytes = scan_workspace_bytes(sumK + 2);
    void* scan_ws = cv.take<char>(scan_bytes);

    MWE_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)(sumK + 2) * sizeof(int32_t), s));
    uint64_t* ks = keys;
    uint32_t* vs = vals;
    if (N > 0) {
        int64_t blocks = (N + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        centroid_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(label, N, sumK, keys, vals, count);
        MWE_CHECK_LAUNCH();
        int rc = sort_pairs(keys, vals, N, ceil_log2_u64((uint64_t)sumK + 1), sort_ws, sort_bytes, s, &ks, &vs);
        if (rc != MWE_OK) return rc;
    }
    int rc = exclusive_scan_i32(count, seg_start, sumK + 1, nullptr, scan_ws, scan_bytes, s);
    if (rc != MWE_OK) return rc;
    centroid_sum_kernel<MODE><<<(unsigned)sumK, CS_THREADS, 0, s>>>(X, ldx, D, w, vs, seg_start, out_wx, out_w);
    MWE_CHECK_LAUNCH();
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
