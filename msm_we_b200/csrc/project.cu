// Affine projection in front of K1: Y = (X - mean) @ components^T, fp64.
//
// reference: `self.coordinates.transform(self.processCoordinates(coords))` immediately before every predict /
// partial_fit (msm_we/_hamsm/_clustering.py:1291-1296, :894), where `coordinates` is the fitted IncrementalPCA of
// msm_we/_hamsm/_dimensionality.py:243, i.e. sklearn's `(X - mean_) @ components_.T`.  SURVEY section 8(f) rank 1:
// it runs on every frame, right before assignment, and as a host numpy matmul it is the first thing left on the
// CPU once K1 is fast.
//
// The centred coordinate is formed exactly as numpy forms it (one rounded subtraction per element); the products
// are accumulated in k order on the fp64 tensor pipe (mma.sync.m8n8k4.f64), which differs from a blocked BLAS
// dgemm only in summation order (tests: 1e-12 relative).  HBM-bound: D_in*8 bytes in, d_out*8 bytes out per row;
// a tile is 64 rows, k-chunks of 16 columns of X and of the components go through a double-buffered cp.async ring.
#include "assign_common.cuh"

namespace mwe {

static constexpr int PJ_THREADS = 128;
static constexpr int PJ_ROWS = 64;
static constexpr int PJ_KC = 16;
static constexpr int PJ_LD = PJ_KC + 4;   // padded rows: conflict-free LDS.64 fragments

template <int NT, int VEC>
__global__ void __launch_bounds__(PJ_THREADS)
    project_kernel(const double* __restrict__ X, int64_t N, int D_in, int64_t ldx, const double* __restrict__ W,
                   const double* __restrict__ mean, int d_out, int col_base, double* __restrict__ Y, int64_t ldy) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ __align__(16) double sX[2][PJ_ROWS][PJ_LD];
    __shared__ __align__(16) double sW[2][NT * 8][PJ_LD];
    __shared__ double sM[2][PJ_KC];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t row0 = (int64_t)blockIdx.x * PJ_ROWS;
    const int nch = (D_in + PJ_KC - 1) / PJ_KC;
    constexpr int SEGS = PJ_KC / VEC;

    auto stage = [&](int kc, int buf) {
        const int k0 = kc * PJ_KC;
        for (int u = tid; u < PJ_ROWS * SEGS; u += PJ_THREADS) {
            const int r = u / SEGS, sg = u - r * SEGS;
            const int k = k0 + sg * VEC;
            int bytes = (D_in - k) * 8;
            bytes = bytes < 0 ? 0 : (bytes > VEC * 8 ? VEC * 8 : bytes);
            const int64_t row = row0 + r;
            if (row < N) cp_async_zfill<VEC>(&sX[buf][r][sg * VEC], X + row * ldx + (bytes ? k : 0), bytes);
        }
        for (int u = tid; u < NT * 8 * SEGS; u += PJ_THREADS) {
            const int r = u / SEGS, sg = u - r * SEGS;
            const int k = k0 + sg * VEC;
            int bytes = (D_in - k) * 8;
            bytes = bytes < 0 ? 0 : (bytes > VEC * 8 ? VEC * 8 : bytes);
            const int c = col_base + r;
            if (c >= d_out) bytes = 0;          // padded output columns: zero rows
            cp_async_zfill<VEC>(&sW[buf][r][sg * VEC], W + (int64_t)(bytes ? c : 0) * D_in + (bytes ? k : 0), bytes);
        }
        if (tid < PJ_KC) sM[buf][tid] = (mean && k0 + tid < D_in) ? mean[k0 + tid] : 0.0;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[2][NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    stage(0, 0);
    for (int kc = 0; kc < nch; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nch) {
            stage(kc + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* xa0 = &sX[buf][warp * 16 + g][t];
        const double* xa1 = xa0 + 8 * PJ_LD;
        const double* wb = &sW[buf][g][t];
#pragma unroll
        for (int ks = 0; ks < PJ_KC / 4; ++ks) {
            const double m = sM[buf][ks * 4 + t];
            const double a0 = __dsub_rn(xa0[ks * 4], m);      // the centred coordinate, rounded as numpy rounds it
            const double a1 = __dsub_rn(xa1[ks * 4], m);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double bv = wb[nt * 8 * PJ_LD + ks * 4];
                dmma8x8x4(acc[0][nt][0], acc[0][nt][1], a0, bv);
                dmma8x8x4(acc[1][nt][0], acc[1][nt][1], a1, bv);
            }
        }
        __syncthreads();      // the chunk is consumed before the next stage() overwrites this buffer
    }
    // accumulator fragment: rows g / g+8 of the warp's 16, columns 2t, 2t+1 of every 8-column tile
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int64_t row = row0 + warp * 16 + mt * 8 + g;
        if (row >= N) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = col_base + nt * 8 + 2 * t + j;
                if (c < d_out) Y[row * ldy + c] = acc[mt][nt][j];
            }
    }
}

template <int VEC>
static int launch_project(int nt, dim3 grid, cudaStream_t s, const double* X, int64_t N, int D_in, int64_t ldx, const double* W,
                          const double* mean, int d_out, int col_base, double* Y, int64_t ldy) {
    const dim3 b(PJ_THREADS);
    switch (nt) {
        case 1: MWE_CHECK_CUDA(launch_pdl(project_kernel<1, VEC>, grid, b, 0, s, X, N, D_in, ldx, W, mean, d_out, col_base, Y, ldy)); break;
        case 2: MWE_CHECK_CUDA(launch_pdl(project_kernel<2, VEC>, grid, b, 0, s, X, N, D_in, ldx, W, mean, d_out, col_base, Y, ldy)); break;
        case 4: MWE_CHECK_CUDA(launch_pdl(project_kernel<4, VEC>, grid, b, 0, s, X, N, D_in, ldx, W, mean, d_out, col_base, Y, ldy)); break;
        default: MWE_CHECK_CUDA(launch_pdl(project_kernel<8, VEC>, grid, b, 0, s, X, N, D_in, ldx, W, mean, d_out, col_base, Y, ldy)); break;
    }
    return MWE_OK;
}

}  // namespace mwe

extern "C" int mwe_project_f64(const double* X, int64_t N, int D_in, int64_t ldx, const double* components,
                               const double* mean, int d_out, double* Y, int64_t ldy, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(N >= 0 && D_in >= 1 && d_out >= 1 && ldx >= D_in && ldy >= d_out, "project: bad shape");
    MWE_REQUIRE(X && components && Y, "project: null pointer");
    if (N == 0) return MWE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool vec2 = (D_in % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(components) & 15) == 0);
    const dim3 grid((unsigned)((N + PJ_ROWS - 1) / PJ_ROWS));
    // output columns in blocks of up to 64 (8 accumulator tiles); the last block takes the smallest width that fits
    for (int col_base = 0; col_base < d_out; col_base += 64) {
        const int left = d_out - col_base;
        const int nt = left <= 8 ? 1 : left <= 16 ? 2 : left <= 32 ? 4 : 8;
        const int rc = vec2 ? launch_project<2>(nt, grid, s, X, N, D_in, ldx, components, mean, d_out, col_base, Y, ldy)
                            : launch_project<1>(nt, grid, s, X, N, D_in, ldx, components, mean, d_out, col_base, Y, ldy);
        if (rc != MWE_OK) return rc;
    }
    return MWE_OK;
}
