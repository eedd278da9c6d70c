// K0: WE-bin lookup + basis/target flags for every point.
//
// reference: bin_mapper.assign + we_remap (msm_we/stratified_clustering.py:134-135,
// msm_we/_hamsm/_clustering.py:877); modelWE.is_WE_basis / is_WE_target (msm_we/msm_we.py:462-527).
// westpa's rectilinear mapper compares float32 coordinates against float32 boundaries
// (lower <= x < upper); the basis/target tests are strict fp64 comparisons on the fp64 pcoord.
// HBM-bound and tiny next to the feature stream: N*(8P) bytes in, 5 bytes out per point.
#include "common.cuh"

namespace mwe {

static constexpr int BF_THREADS = 256;
static constexpr int BF_MAX_P = 8;
static constexpr int BF_SMEM_BINS = 2048;
static constexpr int BF_SMEM_BOUNDS = 2048;

struct BinFlagParams {
    const double* pcoord;
    int64_t N;
    int P;
    int kind;
    const float* mapper_data;
    int32_t nbins;
    int32_t n_bounds;
    const int32_t* we_remap;
    int32_t* bin_out;
    uint8_t* flag_out;
    int32_t* bin_count;
    int32_t* err_count;
    int32_t lens[BF_MAX_P];   // rectilinear: boundaries per dimension
    int32_t starts[BF_MAX_P]; // rectilinear: offset of each dimension's boundaries
    double basis[BF_MAX_P][2];
    double target[BF_MAX_P][2];
};

// PP: compile-time pcoord_ndim for the common 1-D / 2-D progress coordinates (0 = read p.P at run time); the
// generic body carries 8-way predicated loops that cost ~5x the instructions of the 1-D case.
template <int PP>
__global__ void __launch_bounds__(BF_THREADS) bin_flags_kernel(const BinFlagParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const int P = PP ? PP : p.P;
    constexpr int MAXP = PP ? PP : BF_MAX_P;
    __shared__ int32_t s_count[BF_SMEM_BINS];
    __shared__ float s_bounds[BF_SMEM_BOUNDS];
    const bool use_smem = p.bin_count != nullptr && p.nbins <= BF_SMEM_BINS;
    if (use_smem)
        for (int b = threadIdx.x; b < p.nbins; b += BF_THREADS) s_count[b] = 0;
    // rectilinear boundaries are searched once per point and dimension: keep them in shared memory
    const bool smem_bounds = p.kind == MWE_MAPPER_RECTILINEAR && p.n_bounds <= BF_SMEM_BOUNDS;
    if (smem_bounds)
        for (int b = threadIdx.x; b < p.n_bounds; b += BF_THREADS) s_bounds[b] = p.mapper_data[b];
    __syncthreads();
    const float* bounds = smem_bounds ? s_bounds : p.mapper_data;
    const int64_t stride = (int64_t)gridDim.x * BF_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * BF_THREADS + threadIdx.x; i < p.N; i += stride) {
        double pc[MAXP];
#pragma unroll
        for (int d = 0; d < MAXP; ++d)
            if (d < P) pc[d] = p.pcoord[i * P + d];
        bool in_basis = true, in_target = true;
#pragma unroll
        for (int d = 0; d < MAXP; ++d)
            if (d < P) {
                in_basis = in_basis && (pc[d] > p.basis[d][0]) && (pc[d] < p.basis[d][1]);
                in_target = in_target && (pc[d] > p.target[d][0]) && (pc[d] < p.target[d][1]);
            }
        int32_t bin = -1;
        if (p.kind == MWE_MAPPER_RECTILINEAR) {
            int32_t index = 0;
            bool ok = true;
#pragma unroll
            for (int d = 0; d < MAXP; ++d)
                if (d < P) {
                    const float x = (float)pc[d];  // westpa casts coordinates to float32
                    const float* b = bounds + p.starts[d];
                    const int nb = p.lens[d];
                    // number of boundaries <= x (upper_bound), minus one
                    int lo = 0, hi = nb;
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if (b[mid] <= x) lo = mid + 1; else hi = mid;
                    }
                    const int pos = lo - 1;
                    if (pos < 0 || pos >= nb - 1 || x != x) ok = false;
                    index = index * (nb - 1) + (ok ? pos : 0);
                }
            bin = ok ? index : -1;
        } else if (p.kind == MWE_MAPPER_VORONOI) {
            float best = 0.f;
            int32_t besti = 0;
            for (int32_t c = 0; c < p.nbins; ++c) {
                float d2 = 0.f;
#pragma unroll
                for (int d = 0; d < MAXP; ++d)
                    if (d < P) {
                        const float diff = __fsub_rn((float)pc[d], p.mapper_data[(size_t)c * P + d]);
                        d2 = __fadd_rn(d2, __fmul_rn(diff, diff));
                    }
                if (c == 0 || d2 < best) { best = d2; besti = c; }
            }
            bin = besti;
        } else {
            bin = p.bin_out[i];
            if (bin < 0 || bin >= p.nbins) bin = -1;
        }
        if (bin < 0) {
            atomicAdd(&p.err_count[MWE_ERR_OUT_OF_BINSPACE], 1);
        } else if (p.we_remap) {
            bin = p.we_remap[bin];
        }
        const uint8_t flag = (in_basis ? MWE_FLAG_BASIS : 0u) | (in_target ? MWE_FLAG_TARGET : 0u);
        p.bin_out[i] = bin;
        p.flag_out[i] = flag;
        if (p.bin_count && flag == 0 && bin >= 0) {
            if (use_smem) atomicAdd(&s_count[bin], 1);
            else atomicAdd(&p.bin_count[bin], 1);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < p.nbins; b += BF_THREADS)
            if (s_count[b]) atomicAdd(&p.bin_count[b], s_count[b]);
    }
}


// Rows holding a NaN.  reference: get_transition_data_lag0 zeroes the weight of every segment whose start or end
// structure contains a NaN (msm_we/_hamsm/_data.py:302-313) and re-reads all structures in the flux pass only to
// find them; the discretization pass has the rows on the device anyway.  One warp per row, coalesced.
__global__ void __launch_bounds__(256)
    rows_with_nan_kernel(const double* __restrict__ X, int64_t N, int D, int64_t ldx, uint8_t* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < N; r += warps) {
        const double* row = X + r * ldx;
        bool bad = false;
        for (int c = lane; c < D; c += 32) {
            const double x = row[c];
            bad |= (x != x);
        }
        const unsigned any = __ballot_sync(0xffffffffu, bad);
        if (lane == 0) out[r] = any ? 1 : 0;
    }
}

}  // namespace mwe

extern "C" int mwe_bin_flags_f64(const double* pcoord, int64_t N, int P, int mapper_kind, const float* mapper_data,
                                 const int32_t* mapper_lens_host, int32_t nbins, const double* basis_lohi_host,
                                 const double* target_lohi_host, const int32_t* we_remap, int32_t* bin_out,
                                 uint8_t* flag_out, int32_t* bin_count, int32_t* err_count, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(N >= 0, "bin_flags: negative N");
    MWE_REQUIRE(P >= 1 && P <= BF_MAX_P, "bin_flags: pcoord_ndim must be in [1, 8]");
    MWE_REQUIRE(mapper_kind >= 0 && mapper_kind <= 2, "bin_flags: unknown mapper kind");
    MWE_REQUIRE(bin_out && flag_out && err_count, "bin_flags: null output");
    MWE_REQUIRE(basis_lohi_host && target_lohi_host, "bin_flags: null bounds");
    if (N == 0) return MWE_OK;
    BinFlagParams p;
    p.pcoord = pcoord; p.N = N; p.P = P; p.kind = mapper_kind; p.mapper_data = mapper_data; p.nbins = nbins;
    p.we_remap = we_remap; p.bin_out = bin_out; p.flag_out = flag_out; p.bin_count = bin_count; p.err_count = err_count;
    int32_t start = 0;
    int64_t prod = 1;
    p.n_bounds = 0;
    for (int d = 0; d < BF_MAX_P; ++d) {
        p.lens[d] = 0; p.starts[d] = 0;
        p.basis[d][0] = p.basis[d][1] = p.target[d][0] = p.target[d][1] = 0.0;
        if (d < P) {
            p.basis[d][0] = basis_lohi_host[2 * d]; p.basis[d][1] = basis_lohi_host[2 * d + 1];
            p.target[d][0] = target_lohi_host[2 * d]; p.target[d][1] = target_lohi_host[2 * d + 1];
            if (mapper_kind == MWE_MAPPER_RECTILINEAR) {
                MWE_REQUIRE(mapper_lens_host && mapper_lens_host[d] >= 2, "bin_flags: each dimension needs >= 2 boundaries");
                p.lens[d] = mapper_lens_host[d];
                p.starts[d] = start;
                start += mapper_lens_host[d];
                p.n_bounds = start;
                prod *= (mapper_lens_host[d] - 1);
            }
        }
    }
    if (mapper_kind == MWE_MAPPER_RECTILINEAR) MWE_REQUIRE(prod == nbins, "bin_flags: nbins != product of per-dimension bins");
    if (mapper_kind != MWE_MAPPER_PRECOMPUTED) MWE_REQUIRE(mapper_data, "bin_flags: null mapper data");
    int64_t blocks = (N + BF_THREADS - 1) / BF_THREADS;
    // latency-bound per point (one search per dimension): one point per thread until the grid is many waves deep
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 g((unsigned)blocks), b(BF_THREADS);
    if (P == 1) MWE_CHECK_CUDA(launch_pdl(bin_flags_kernel<1>, g, b, 0, st, p));
    else if (P == 2) MWE_CHECK_CUDA(launch_pdl(bin_flags_kernel<2>, g, b, 0, st, p));
    else MWE_CHECK_CUDA(launch_pdl(bin_flags_kernel<0>, g, b, 0, st, p));
    return MWE_OK;
}

extern "C" int mwe_rows_with_nan_f64(const double* X, int64_t N, int D, int64_t ldx, uint8_t* out, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(N >= 0 && D >= 1 && ldx >= D, "rows_with_nan: bad shape");
    if (N == 0) return MWE_OK;
    MWE_REQUIRE(X && out, "rows_with_nan: null pointer");
    int64_t blocks = (N + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    MWE_CHECK_CUDA(launch_pdl(rows_with_nan_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                              X, N, D, ldx, out));
    return MWE_OK;
}
