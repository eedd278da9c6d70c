// One-call hot path: K0 (bins + basis/target flags of the stacked parent/child pcoords) -> K1
// (stratified assignment of the stacked parent/child features) -> K3 (flux accumulation, / nI).
//
// reference: the body of do_stratified_ray_discretization (msm_we/_hamsm/_clustering.py:1278-1316)
// followed by get_iter_fluxMatrix / get_fluxMatrix (msm_we/_hamsm/_fluxmatrix.py:21-72, 232-260, 342)
// for a whole range of iterations at once.  Exists so that a host can enqueue the ~20 kernels of a
// step without returning to the interpreter between them (the small-configuration step is
// launch-bound: BASELINE cfg2 is 32 us of HBM time).
#include "common.cuh"

extern "C" size_t mwe_hotpath_workspace_bytes_ex(int64_t n_frames, int32_t nbins, int D, int32_t max_k, int precision_path) {
    return mwe_hotpath_workspace_bytes(n_frames, nbins) +
           (mwe_assign_workspace_bytes_ex(2 * n_frames, nbins, D, max_k, precision_path) - mwe_assign_workspace_bytes(2 * n_frames, nbins));
}

extern "C" size_t mwe_hotpath_workspace_bytes(int64_t n_frames, int32_t nbins) {
    const int64_t N2 = 2 * n_frames;
    size_t b = 0;
    b += mwe::align_up((size_t)(N2 > 0 ? N2 : 1) * sizeof(int32_t), 256);  // bins
    b += mwe::align_up((size_t)(N2 > 0 ? N2 : 1), 256);                    // flags
    b += mwe::align_up((size_t)(nbins + 1) * sizeof(int32_t), 256);        // per-bin counts
    b += mwe_assign_workspace_bytes(N2, nbins);
    b += mwe_flux_workspace_bytes(n_frames);
    return b + 1024;
}

extern "C" int mwe_hotpath_step_f64(const double* X2, int64_t ldx, int D, const double* pcoord2, int P, const double* w,
                                    int64_t n_frames, const int64_t* iter_offsets, int64_t n_iters, int mapper_kind,
                                    const float* mapper_data, const int32_t* mapper_lens_host, int32_t nbins,
                                    const double* basis_lohi_host, const double* target_lohi_host,
                                    const int32_t* we_remap, const double* centers, const double* csq,
                                    const int64_t* bin_offset, int32_t max_k, int precision_path, int64_t n_clusters,
                                    double divisor, int64_t* labels2_out, double* dense_inout, void* workspace,
                                    size_t workspace_bytes, int32_t* err_count, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(n_frames >= 0, "hotpath: negative frame count");
    if (workspace_bytes < mwe_hotpath_workspace_bytes_ex(n_frames, nbins, D, max_k, precision_path)) {
        set_last_error("hotpath: workspace too small");
        return MWE_E_WORKSPACE;
    }
    const int64_t N2 = 2 * n_frames;
    Carver cv(workspace, workspace_bytes);
    int32_t* bins = cv.take<int32_t>((size_t)(N2 > 0 ? N2 : 1));
    uint8_t* flags = cv.take<uint8_t>((size_t)(N2 > 0 ? N2 : 1));
    int32_t* bin_count = cv.take<int32_t>((size_t)nbins + 1);
    const size_t aws = mwe_assign_workspace_bytes_ex(N2, nbins, D, max_k, precision_path);
    void* assign_ws = cv.take<char>(aws);
    const size_t fws = mwe_flux_workspace_bytes(n_frames);
    void* flux_ws = cv.take<char>(fws);
    MWE_CHECK_CUDA(cudaMemsetAsync(bin_count, 0, (size_t)(nbins + 1) * sizeof(int32_t), static_cast<cudaStream_t>(stream)));
    int rc = mwe_bin_flags_f64(pcoord2, N2, P, mapper_kind, mapper_data, mapper_lens_host, nbins, basis_lohi_host,
                               target_lohi_host, we_remap, bins, flags, bin_count, err_count, stream);
    if (rc != MWE_OK) return rc;
    rc = mwe_assign_stratified_f64(X2, N2, D, ldx, bins, flags, centers, csq, bin_offset, nbins, max_k, precision_path,
                                   bin_count, labels2_out, nullptr, assign_ws, aws, err_count, stream);
    if (rc != MWE_OK) return rc;
    if (dense_inout) {
        rc = mwe_flux_accumulate_f64(labels2_out, labels2_out + n_frames, flags, flags + n_frames, nullptr, nullptr, w,
                                     n_frames, n_clusters, 1, iter_offsets, n_iters, dense_inout, nullptr, nullptr,
                                     nullptr, nullptr, flux_ws, fws, err_count, stream);
        if (rc != MWE_OK) return rc;
        if (divisor != 0.0 && divisor != 1.0) {
            const int64_t M = n_clusters + 2;
            rc = mwe_divide_f64(dense_inout, M * M, divisor, stream);
        }
    }
    return rc;
}
