/*
 * msm_we_b200 -- C ABI of the B200-native discretization + flux hot path of jdrusso/msm_we.
 *
 * The reference is pure Python; its "FFI" for this path is the call into scikit-learn's Cython
 * kernels and scipy.sparse.  Each entry point below names the reference call site it replaces
 * (paths relative to the reference root, or sklearn/ for the third-party kernels).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / Python types.
 *   - every data pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless stated; calls are re-entrant per stream (workspaces are caller-owned);
 *   - return value: 0 on success, negative MWE_E_* otherwise; mwe_last_error() gives the text
 *     (thread-local);
 *   - data-dependent errors the reference raises as Python exceptions (coordinate outside the bin
 *     space, label out of range, WE bin without centres) cannot be raised from inside a kernel:
 *     they are counted into a caller-supplied device int32 `err_count[MWE_ERR_SLOTS]` which the
 *     host layer reads at its next synchronisation point and converts into the reference's
 *     exception type.
 */
#ifndef MSM_WE_B200_H
#define MSM_WE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MWE_ABI_VERSION 1
#define MWE_API __attribute__((visibility("default")))

#define MWE_OK 0
#define MWE_E_INVALID (-1)   /* bad argument                                   */
#define MWE_E_CUDA (-2)      /* CUDA runtime error, see mwe_last_error()        */
#define MWE_E_WORKSPACE (-3) /* workspace too small                             */
#define MWE_E_UNSUPPORTED (-4)

/* slots of the device error counter array */
#define MWE_ERR_SLOTS 4
#define MWE_ERR_OUT_OF_BINSPACE 0 /* westpa ValueError("coordinate outside of bin space")          */
#define MWE_ERR_NO_CENTERS 1      /* AssertionError, msm_we/stratified_clustering.py:187-189      */
#define MWE_ERR_LABEL_RANGE 2     /* ValueError from coo_matrix, msm_we/_hamsm/_fluxmatrix.py:147 */
#define MWE_ERR_INTERNAL 3

/* flag bits written by mwe_bin_flags_f64 */
#define MWE_FLAG_BASIS 1u
#define MWE_FLAG_TARGET 2u

/* mapper kinds */
#define MWE_MAPPER_RECTILINEAR 0
#define MWE_MAPPER_VORONOI 1
#define MWE_MAPPER_PRECOMPUTED 2

/* precision paths of the assignment kernel */
#define MWE_ASSIGN_FP64 0      /* fp64 tensor (DMMA) distances, the parity path                    */
#define MWE_ASSIGN_TF32X3 1    /* tcgen05 split-TF32 candidate pass + fp64 re-check of near-ties   */
#define MWE_ASSIGN_AUTO 2      /* the faster of the two for the shape (labels are identical)        */
/* OR-ed into precision_path of mwe_assign_stratified_f64: the workspace, label_out and local_out still hold what the
 * previous call with the SAME N, bin, flag, nbins and bin_offset left there (nothing else wrote to them since), so the
 * bucketing of the points by WE bin is not recomputed -- Lloyd iterations re-label the same points against new centres
 * (sklearn lloyd_iter_chunked_dense called max_iter times on one X, _kmeans.py:700-760). */
#define MWE_ASSIGN_REUSE_BUCKETS 0x100

MWE_API int mwe_abi_version(void);
MWE_API const char* mwe_last_error(void);
/* number of SMs of the current device (grid sizing is done inside the library) */
MWE_API int mwe_device_sm_count(void);

/* Page-lock / unlock a host range the caller owns (cudaHostRegister / cudaHostUnregister), so that the
 * per-iteration coordinate arrays a model already holds (msm_we/_hamsm/_data.py:557-618 hands them out as numpy
 * arrays) can be copied to the device asynchronously at link speed without a staging copy.  Returns MWE_OK, or
 * MWE_E_CUDA when the range cannot be registered (already registered, overlapping, over the lock limit). */
MWE_API int mwe_host_register(void* ptr, size_t bytes);
MWE_API int mwe_host_unregister(void* ptr);

/* Measurement hook: CUDA events (cudaEvent_t as void*, nullable) that the following calls on this
 * host thread record immediately before / after their dominant kernel (K1: the DMMA assignment
 * kernel), on the stream the kernel is launched on.  Pass NULL, NULL to switch it off. */
MWE_API int mwe_set_timing_events(void* start, void* stop);

/* ---- K0: WE-bin lookup + basis/target flags -------------------------------------------------
 * Replaces bin_mapper.assign(pcoords) + we_remap (msm_we/stratified_clustering.py:134-135,
 * msm_we/_hamsm/_clustering.py:877) and modelWE.is_WE_basis / is_WE_target
 * (msm_we/msm_we.py:462-527; call sites stratified_clustering.py:137-138, _fluxmatrix.py:33-53).
 *   pcoord      [N, P] f64 row-major
 *   mapper_kind MWE_MAPPER_RECTILINEAR: mapper_data = all boundaries concatenated (float32, as
 *               westpa stores them), mapper_lens_host[P] (HOST) = boundary count per dimension;
 *               MWE_MAPPER_VORONOI: mapper_data = centres [nbins, P] float32 (Euclidean dfunc);
 *               MWE_MAPPER_PRECOMPUTED: bin_out already holds the raw bin of every point (a host
 *               mapper produced it); only we_remap and the flags are applied.
 *   basis_lohi_host / target_lohi_host [P, 2] f64 (HOST, passed to the kernel by value), strict
 *               lo < p < hi on every dimension.
 *   we_remap    [nbins] int32 (nullable = identity)
 *   bin_out     [N] int32 remapped WE bin; flag_out [N] uint8 (MWE_FLAG_* bits)
 *   bin_count   nullable [nbins] int32, += number of points with flag == 0 in each bin
 */
MWE_API int mwe_bin_flags_f64(const double* pcoord, int64_t N, int P, int mapper_kind, const float* mapper_data,
                      const int32_t* mapper_lens_host, int32_t nbins, const double* basis_lohi_host,
                      const double* target_lohi_host, const int32_t* we_remap, int32_t* bin_out, uint8_t* flag_out,
                      int32_t* bin_count, int32_t* err_count, void* stream);

/* Rows of X [N, D] (row stride ldx) that contain a NaN -> out [N] uint8 (1 / 0).
 * Replaces the NaN scan of get_transition_data_lag0 (msm_we/_hamsm/_data.py:302-313: such segments get transition
 * weight 0), which the reference repeats in the flux pass by re-reading every structure from disk. */
MWE_API int mwe_rows_with_nan_f64(const double* X, int64_t N, int D, int64_t ldx, uint8_t* out, void* stream);

/* ---- K1: stratified nearest-centre assignment -----------------------------------------------
 * Replaces the per-segment MiniBatchKMeans.predict([coord]) loop of StratifiedClusters.predict
 * (msm_we/stratified_clustering.py:152-203) and the E step of partial_fit / KMeans.fit
 * (sklearn/cluster/_k_means_lloyd.pyx:168-218: ||c||^2 - 2 x.c, strict-< argmin, lowest index).
 *   X [N, D] f64, row stride ldx (elements); bin/flag from K0 (flag nullable = all free);
 *   centers [sumK, D] f64 all bins concatenated; csq [sumK] = row squared norms
 *   (mwe_centers_sqnorm_f64); bin_offset [nbins+1] int64 prefix of per-bin centre counts.
 *   label_out [N] int64: target -> T+1, basis -> T (target tested first), else
 *   bin_offset[bin] + argmin, T = bin_offset[nbins]   (stratified_clustering.py:143-196).
 *   bin_count_in nullable [nbins] int32: per-bin count of unflagged points as written by
 *   mwe_bin_flags_f64 (saves one pass over bin/flag).
 *   local_out  nullable [N] int32: argmin local to the bin (what partial_fit's E step needs).
 *   Near-ties: scores that differ by less than tol = 4 (D+8) 2^-53 cmax (2||x|| + cmax) -- the rounding
 *   noise of their own evaluation -- count as tied and the lowest index wins (exact duplicates of a
 *   centre therefore resolve as in the reference; see DESIGN.md "near-tie re-check").
 */
MWE_API size_t mwe_assign_workspace_bytes(int64_t N, int32_t nbins);            /* MWE_ASSIGN_FP64 */
MWE_API size_t mwe_assign_workspace_bytes_ex(int64_t N, int32_t nbins, int D, int32_t max_k, int precision_path);
MWE_API int mwe_centers_sqnorm_f64(const double* centers, int64_t sumK, int D, double* csq, void* stream);
MWE_API int mwe_assign_stratified_f64(const double* X, int64_t N, int D, int64_t ldx, const int32_t* bin,
                              const uint8_t* flag, const double* centers, const double* csq,
                              const int64_t* bin_offset, int32_t nbins, int32_t max_k, int precision_path,
                              const int32_t* bin_count_in, int64_t* label_out, int32_t* local_out, void* workspace,
                              size_t workspace_bytes, int32_t* err_count, void* stream);

/* ---- K2: centroid accumulation / update -----------------------------------------------------
 * Replaces update_center_dense (sklearn/cluster/_k_means_minibatch.pyx:59-111, reached from
 * msm_we/_hamsm/_clustering.py:909) and the M step of lloyd_iter_chunked_dense
 * (sklearn/cluster/_k_means_lloyd.pyx:23-165, reached from _clustering.py:289,491).
 * Deterministic: points are grouped by label with a stable sort and every cluster's rows are
 * summed in original index order (the order the single-threaded reference uses).
 *   label [N] int64 global cluster index in [0, sumK); entries outside are skipped
 *   (basis/target points).  w nullable (= 1).
 *   mwe_centroid_accumulate_f64: sum_wx [sumK, D], sum_w [sumK] are OVERWRITTEN with the sums
 *   (these are the buffers a multi-GPU run all-reduces).
 *   mwe_minibatch_update_f64: centers/counts updated in place with the running-mean rule,
 *   starting each sum from centers*counts exactly as the reference does.
 *   mwe_lloyd_finalize_f64 / mwe_minibatch_finalize_f64: apply the rule to reduced partial sums.
 */
MWE_API size_t mwe_centroid_workspace_bytes(int64_t N, int64_t sumK);
MWE_API int mwe_centroid_accumulate_f64(const double* X, int64_t N, int D, int64_t ldx, const double* w,
                                const int64_t* label, int64_t sumK, double* sum_wx, double* sum_w,
                                void* workspace, size_t workspace_bytes, void* stream);
MWE_API int mwe_minibatch_update_f64(const double* X, int64_t N, int D, int64_t ldx, const double* w,
                             const int64_t* label, int64_t sumK, double* centers, double* counts,
                             void* workspace, size_t workspace_bytes, void* stream);
MWE_API int mwe_lloyd_finalize_f64(const double* sum_wx, const double* sum_w, int64_t sumK, int D, double* centers,
                           void* stream);
MWE_API int mwe_minibatch_finalize_f64(const double* sum_wx, const double* sum_w, int64_t sumK, int D, double* centers,
                               double* counts, void* stream);

/* ||x_i - centre[label_i]||^2 (fp64) of the n listed points (list [n] int32 indices into X / label).  Used by the
 * empty-cluster relocation of the Lloyd M step (sklearn/cluster/_k_means_common.pyx _relocate_empty_clusters_dense,
 * reached through KMeans.fit from msm_we/_hamsm/_clustering.py:289,491): only the points of the WE bins that own an
 * empty cluster are listed, and only the handful of farthest ones ever leave the device. */
MWE_API int mwe_point_center_dist2_f64(const double* X, int64_t ldx, int D, const int32_t* list, int64_t n,
                                       const int64_t* label, const double* centers, double* out, void* stream);

/* ---- group-by-label + per-label statistics ----------------------------
 * Replaces the O(n_clusters x iterations) np.where loops of ClusteringMixin.get_cluster_centers
 * (msm_we/_hamsm/_clustering.py:1528-1599: nanmean / nanmin / nanmax of the end pcoord of every cluster's
 * members) and the per-segment list appends of update_cluster_structures (:1398-1526).
 *   mwe_group_by_label: stable sort of the N labels; members_out [N] uint32 = indices grouped by label, each
 *     group in input order; seg_start_out [n_labels + 2] int32 = first position of every label's group
 *     ([n_labels] = where the out-of-range labels begin, [n_labels+1] = N).  Workspace:
 *     mwe_centroid_workspace_bytes(N, n_labels).
 *   mwe_label_stats_f64: NaN-skipping count / sum / min / max of values[member * ldv] per label (an empty
 *     label gets count 0, sum 0, min +inf, max -inf); one warp per label, fixed reduction order. */
MWE_API int mwe_group_by_label(const int64_t* label, int64_t N, int64_t n_labels, uint32_t* members_out,
                               int32_t* seg_start_out, void* workspace, size_t workspace_bytes, void* stream);
MWE_API int mwe_label_stats_f64(const double* values, int64_t ldv, const uint32_t* members, const int32_t* seg_start,
                                int64_t n_labels, int64_t* count, double* sum, double* vmin, double* vmax, void* stream);
/*   mwe_segment_topk_f64: for each of the n_sel listed groups (seg_ids, indices into seg_start) the k <= 8 largest
 *     values[member] with their member indices, largest first, equal values in member order (= the head of a stable
 *     descending sort of the group); groups with fewer than k (non-NaN) members are padded with index -1.  Picks the
 *     farthest points of a WE-bin model that lost several clusters in one Lloyd iteration (sklearn
 *     _relocate_empty_clusters_dense, _k_means_common.pyx; reached from msm_we/_hamsm/_clustering.py:289,491). */
MWE_API int mwe_segment_topk_f64(const double* values, const uint32_t* members, const int32_t* seg_start,
                                 const int32_t* seg_ids, int32_t n_sel, int k, int32_t* out_pos, double* out_val, void* stream);


/* ---- K3: weighted transition scatter into the flux matrix -----------------------------------
 * Replaces FluxMatrixMixin.build_flux_matrix + .todense() + the per-iteration accumulation of
 * get_fluxMatrix (msm_we/_hamsm/_fluxmatrix.py:97-164, 61-72, 232-260, 311-342) and the coloured
 * count scatter of NonMarkovModel.fit (msm_we/nmm.py:132-158).
 *   start/end [N] int64 cluster labels of parent and child; flag0/flag1 [N] uint8 from K0 on
 *   pcoord0 / pcoord1 (nullable): end[target]=n+1, start[basis]=n, end[basis]=n in that order
 *   (_fluxmatrix.py:135-137).  col0/col1 nullable uint8 history colours (C=2):
 *   row = C*start+col0, col = C*end+col1 (nmm.py:147-154).  w nullable (= 1.0).
 *   Matrix side: CM = C*(n_clusters+2).  Labels outside [0, n_clusters+2) count an
 *   MWE_ERR_LABEL_RANGE error and are dropped.
 *   The transitions are stably sorted by (row, col); every cell is summed in input order --
 *   per iteration first and then across iterations when iter_offsets [n_iters+1] is given,
 *   which is the association of the reference's serial path -- so the result does not depend on
 *   thread scheduling.
 *   dense_inout nullable [CM, CM] f64: cell += sum.  coo_* nullable: sorted unique (row, col, sum),
 *   *nnz_out (device int64) = count; coo capacity must be >= N.
 */
MWE_API size_t mwe_flux_workspace_bytes(int64_t N);
MWE_API int mwe_flux_accumulate_f64(const int64_t* start, const int64_t* end, const uint8_t* flag0, const uint8_t* flag1,
                            const uint8_t* col0, const uint8_t* col1, const double* w, int64_t N,
                            int64_t n_clusters, int C, const int64_t* iter_offsets, int64_t n_iters,
                            double* dense_inout, int64_t* coo_row, int64_t* coo_col, double* coo_val,
                            int64_t* nnz_out, void* workspace, size_t workspace_bytes, int32_t* err_count,
                            void* stream);
/* buf[i] = buf[i] / divisor      (fluxMatrix / nI, _fluxmatrix.py:342) */
/* ---- history colours along WE lineages ----------------------------------------
 * Replaces the per-trajectory Python walk of NonMarkovModel.fit (msm_we/nmm.py:117-167) for trajectories that are the
 * lineages of a weighted-ensemble run (traced through seg_index['parent_id'], msm_we/_hamsm/_data.py:807-932):
 *   mwe_lineage_colour : one iteration forward.  colour_now[s] = 0 if state_class[label_now[s]] == 1 (state in A),
 *                        1 if == 2 (state in B), else colour_prev[parent[s]] (-1 = undefined; colour_prev NULL for the
 *                        first coloured iteration, parent < 0 = no parent);
 *   mwe_lineage_leaves : one iteration backward.  leaves_prev[parent[s]] += leaves_now[s] (caller zeroes leaves_prev
 *                        and seeds the last iteration with 1): how many traced trajectories pass through a segment;
 *   mwe_lineage_records: the coloured transition records of one iteration for mwe_flux_accumulate_f64 with C = 2:
 *                        (label_prev[parent], label_now, colour_prev[parent], colour_now, weight = leaves_now), weight 0
 *                        where a colour is undefined or the segment has no parent. */
MWE_API int mwe_lineage_colour(const int64_t* label_now, const int64_t* parent, int64_t S_now, const int8_t* colour_prev,
                               int64_t S_prev, const uint8_t* state_class, int64_t n_states, int8_t* colour_now, void* stream);
MWE_API int mwe_lineage_leaves(const int64_t* parent, const uint64_t* leaves_now, int64_t S_now, uint64_t* leaves_prev,
                               int64_t S_prev, void* stream);
MWE_API int mwe_lineage_records(const int64_t* label_prev, const int8_t* colour_prev, int64_t S_prev, const int64_t* label_now,
                                const int8_t* colour_now, const int64_t* parent, const uint64_t* leaves_now, int64_t S_now,
                                int64_t* start, int64_t* end, uint8_t* col0, uint8_t* col1, double* w, void* stream);

MWE_API int mwe_divide_f64(double* buf, int64_t count, double divisor, void* stream);

/* ---- one-call hot path ------------------------------------------------------------------------
 * K0 -> K1 -> K3 (-> / divisor) for a batch of WE iterations, enqueued without returning to the host
 * language between kernels.  Replaces the per-iteration body of do_stratified_ray_discretization
 * (msm_we/_hamsm/_clustering.py:1278-1316) plus get_fluxMatrix's accumulation
 * (msm_we/_hamsm/_fluxmatrix.py:232-260, 342) over that batch.
 *   X2 [2*n_frames, D] f64: parent features of every frame, then child features (row stride ldx);
 *   pcoord2 [2*n_frames, P] f64: pcoord0 of every frame, then pcoord1; w [n_frames] (nullable = 1);
 *   iter_offsets [n_iters+1] nullable; mapper_* / bounds / we_remap as mwe_bin_flags_f64;
 *   centers / csq / bin_offset / max_k / precision_path as mwe_assign_stratified_f64.
 *   labels2_out [2*n_frames] int64 (parents then children, predict convention);
 *   dense_inout nullable [(n_clusters+2)^2]: += this batch's transitions, then /= divisor when
 *   divisor is neither 0 nor 1 (pass the total iteration count on the last batch). */
MWE_API size_t mwe_hotpath_workspace_bytes(int64_t n_frames, int32_t nbins);      /* MWE_ASSIGN_FP64 */
MWE_API size_t mwe_hotpath_workspace_bytes_ex(int64_t n_frames, int32_t nbins, int D, int32_t max_k, int precision_path);
MWE_API int mwe_hotpath_step_f64(const double* X2, int64_t ldx, int D, const double* pcoord2, int P, const double* w,
                         int64_t n_frames, const int64_t* iter_offsets, int64_t n_iters, int mapper_kind,
                         const float* mapper_data, const int32_t* mapper_lens_host, int32_t nbins,
                         const double* basis_lohi_host, const double* target_lohi_host, const int32_t* we_remap,
                         const double* centers, const double* csq, const int64_t* bin_offset, int32_t max_k,
                         int precision_path, int64_t n_clusters, double divisor, int64_t* labels2_out,
                         double* dense_inout, void* workspace, size_t workspace_bytes, int32_t* err_count,
                         void* stream);

/* ---- test hooks for the tcgen05 assignment path ------------------------------------------------
 * mwe_debug_set_tc_scores: device buffer [N][mwe_debug_tc_columns(max_k)] fp32 (or NULL) that the next
 * MWE_ASSIGN_TF32X3 calls fill with the scores the tensor cores produced (error-bound validation). */
MWE_API int mwe_debug_set_tc_scores(float* buf);
MWE_API int mwe_debug_tc_columns(int32_t max_k);
/* tuning hook: device uint64[20] (or NULL) accumulating per-role wait cycles of the tcgen05 kernel */
MWE_API int mwe_debug_set_tc_profile(unsigned long long* buf);
/* tuning hook: device uint64[8] (or NULL) accumulating per-phase clock cycles of the resident-centre fp64 kernel
 * (assign_res.cu: consumers [0] wait for data, [1] metadata, [2] k loop, [3] fold + epilogue, [4] groups;
 * producers [5] wait for a buffer, [6] claim + copy issue) */
MWE_API int mwe_debug_set_k1_profile(unsigned long long* buf);

/* ---- shared primitive, exported for tests ---------------------------------------------------
 * Stable LSD radix sort of (key, value) pairs on the low `key_bits` bits of the key.
 * keys/vals are sorted in place (a ping-pong copy lives in the workspace). */
MWE_API size_t mwe_sort_workspace_bytes(int64_t N);
MWE_API int mwe_sort_pairs_u64_u32(uint64_t* keys, uint32_t* vals, int64_t N, int key_bits, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- projection in front of K1 ---------------------------------------
 * Replaces `self.coordinates.transform(...)` = IncrementalPCA's `(X - mean_) @ components_.T`
 * (msm_we/_hamsm/_dimensionality.py:243; call sites _clustering.py:1291-1296, :894) for frames that are already on
 * the device.
 *   X          [N, D_in] f64, row stride ldx      components [d_out, D_in] f64 row-major (sklearn's components_)
 *   mean       [D_in] f64 or NULL                 Y          [N, d_out] f64, row stride ldy */
MWE_API int mwe_project_f64(const double* X, int64_t N, int D_in, int64_t ldx, const double* components, const double* mean,
                            int d_out, double* Y, int64_t ldy, void* stream);

/* ---- multi-GPU exchange step of the flux path, over NVLink peer memory -------------------------
 * Replaces the driver-side sum of the per-worker iteration matrices and the final "/ nI"
 * (msm_we/_hamsm/_fluxmatrix.py:311-327, :342) when WE iterations are sharded over one process per GPU.
 * Setup (once): every rank allocates its partial-sum buffer, result buffer ([count] f64 each) and flag words
 * ([2*world] u32, zero) with mwe_device_malloc, exports them (mwe_ipc_export -> 64-byte handle), sends the handles
 * to the other ranks by any transport, and opens theirs (mwe_ipc_open).  Per call: all ranks call
 * mwe_flux_peer_allreduce_f64 with the same `epoch` (1, 2, 3, ...), host arrays of the `world` device pointers
 * (own buffers at index `rank`), and a zeroed u32 `cta_counter`.  On return (stream order) outs[rank] holds
 * (partial_0 + partial_1 + ...) / divisor, summed in rank order, on every rank. */
MWE_API int mwe_device_malloc(size_t bytes, void** out);
MWE_API int mwe_device_free(void* ptr);
MWE_API int mwe_ipc_export(void* device_ptr, unsigned char* handle64);
MWE_API int mwe_ipc_open(const unsigned char* handle64, void** out);
MWE_API int mwe_ipc_close(void* ptr);
MWE_API int mwe_flux_peer_allreduce_f64(const void* const* partials, void* const* outs, void* const* flags, int rank,
                                        int world, int64_t count, double divisor, uint32_t epoch, void* cta_counter,
                                        int32_t* err_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSM_WE_B200_H */
