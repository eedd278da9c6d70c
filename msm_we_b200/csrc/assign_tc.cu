// K1, fast path: tcgen05 (5th-gen tensor core) candidate pass + exact fp64 re-check of close calls.
//
// reference arithmetic: sklearn/cluster/_k_means_lloyd.pyx:168-218 (score_j = ||c_j||^2 - 2 x.c_j, first
// minimum wins), reached from msm_we/stratified_clustering.py:152-203.
//
// Why: the fp64 tensor pipe (DMMA, 37 TFLOP/s measured) cannot keep up with HBM once K_b*D grows
// (BASELINE cfg3/cfg5 need 4x the fp64 FLOPs HBM time allows; even cfg2 sits on the ridge).  The labels only
// need the ORDER of the scores, and an order decided with a margin larger than the evaluation error is the
// fp64 order.  So: evaluate x.c on the tensor cores with split TF32 operands (x = hi + lo, each TF32;
// hi.hi + hi.lo + lo.hi accumulated in fp32 in TMEM, ~2^-20 relative), take the argmin and the runner-up,
// and hand every point whose margin does not clear a rigorous error bound to the same fp64 re-check
// kernel the DMMA path uses (assign.cu, assign_recheck_kernel).  Result: identical labels, HBM-bound kernel.
//
// Data flow of one CTA (persistent, one per SM, 128 points per tile, all of one WE bin):
//   loader warps (8)  : gather the tile's fp64 rows straight from HBM into registers (2 x LDG.128 per
//                       thread and row chunk), subtract the bin's mean centre (translation invariance ->
//                       smaller norms -> tighter error bound), split into TF32 hi / lo and store them in the
//                       UMMA canonical K-major layout (8-row x 16-byte core matrices, no swizzle, padded
//                       leading offset so the stores are bank-conflict free); also accumulate ||x||^2
//   centre warp (1)   : one cp.async.bulk (TMA) per k-chunk brings the bin's PRE-SPLIT centre block
//                       (prepared once per call, already in the canonical layout) into the same stage
//   MMA warp (1 lane) : per k-chunk 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=padded K_b, K=8) into a
//                       TMEM accumulator; tcgen05.commit releases the smem stage / publishes the accumulator
//   epilogue warps (4): tcgen05.ld their 32 TMEM lanes (one point per thread), scores = ||c'||^2 - 2 dot,
//                       running best / runner-up, label store, near-tie list for the re-check
// Two TMEM accumulators alternate so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdlib.h>

#include "assign_common.cuh"

namespace mwe {

static constexpr int TC_TP = 128;                          // points per tile (UMMA M)
static constexpr int TC_KC = 32;                           // TF32 elements per k-chunk
static constexpr int TC_LBO = 144;                         // bytes between K-adjacent core matrices (128 + 16 pad)
static constexpr int TC_SBO = 8 * TC_LBO;                  // bytes between 8-row groups (8 core matrices / chunk)
static constexpr int TC_A_BYTES = (TC_TP / 8) * TC_SBO;    // one hi or lo point tile
static constexpr int TC_EPI_WARP0 = 0;                     // epilogue warps 0..3 (warp % 4 == TMEM lane quarter)
static constexpr int TC_MMA_WARP = 4;
static constexpr int TC_CENTRE_WARP = 5;
static constexpr int TC_CONV_WARP0 = 6;                    // staging (cp.async) + fp64 -> TF32 conversion warps
static constexpr int TC_CONV_WARPS = 8;
static constexpr int TC_THREADS = (TC_CONV_WARP0 + TC_CONV_WARPS) * 32;
static constexpr int TC_RAW_LD = TC_KC + 2;                // padded fp64 row (272 B): conflict-free 16-byte reads
static constexpr int TC_RAW_BYTES = (TC_TP + 1) * TC_RAW_LD * 8;   // 128 rows + the bin-mean chunk
static constexpr int TC_MAX_STAGES = 6;
static constexpr size_t TC_SMEM_BUDGET = 222 * 1024;

struct TcParams {
    AssignParams a;
    const unsigned char* bprep;   // [bin][cb][kc] -> hi block, lo block (canonical layout, TC_SBO per 8 rows)
    const double* mean;           // [bin][d_pad]   bin mean centre (zero padded)
    const float* csqf;            // [bin][ncb * n_pad]  centred ||c'||^2 (+inf for padding columns)
    const float* cmaxf;           // [bin][3]  upper bounds: max ||c'||^2 (centred), max ||c||^2 (raw), ||mean||
    int n_pad;                    // UMMA N: centres per block, multiple of 16, <= 256
    int d_pad;                    // nch * TC_KC
    int nstages;                  // TF32 operand ring depth
    int nstages_raw;              // fp64 staging ring depth
    uint32_t tmem_cols;           // power of two >= 2 * n_pad
    float err_coef;               // bound of |score error| / (cmax' (2 ||x'|| + cmax'))
    unsigned long long* dbg_prof; // tuning only: cycles spent waiting per role/barrier (16 slots) or nullptr
    float* dbg_scores;            // tests only: [N][ncb * n_pad] fp32 scores as the tensor cores produced them
};

// ---------------------------------------------------------------------------------------------------------
// preparation: bin means, pre-split centres in the canonical layout, centred norms
// ---------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128)
    tc_mean_kernel(const double* __restrict__ centers, const int64_t* __restrict__ bin_offset, int D, int d_pad,
                   double* __restrict__ mean) {
    const int b = blockIdx.x;
    const int k = blockIdx.y * 128 + threadIdx.x;
    if (k >= d_pad) return;
    const int64_t c0 = bin_offset[b], c1 = bin_offset[b + 1];
    double s = 0.0;
    if (k < D) {
        // eight loads in flight, the adds stay in centre order
        int64_t j = c0;
        for (; j + 8 <= c1; j += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = centers[(j + u) * D + k];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; j < c1; ++j) s += centers[j * D + k];
    }
    mean[(size_t)b * d_pad + k] = (c1 > c0 && k < D) ? s / (double)(c1 - c0) : 0.0;
}

// round an fp32 value to TF32 (10 explicit mantissa bits), nearest with ties away from zero: add half a TF32
// ulp to the magnitude bits and clear the 13 low bits.  Two integer ops (cvt.rna.tf32.f32 is emulated with
// a longer sequence on sm_100a); the carry into the exponent is the correct rounding up to the next binade.
__device__ __forceinline__ float tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// one thread per (bin, centre block, row of the block, 4-element k group)
__global__ void __launch_bounds__(256)
    tc_split_centers_kernel(const double* __restrict__ centers, const int64_t* __restrict__ bin_offset, int32_t nbins,
                            int D, int d_pad, int n_pad, int ncb, const double* __restrict__ mean,
                            unsigned char* __restrict__ bprep) {
    const int k4n = d_pad / 4;
    const int64_t total = (int64_t)nbins * ncb * n_pad * k4n;
    const int nch = d_pad / TC_KC;
    const int ng = n_pad / 8;
    const size_t block_bytes = (size_t)2 * ng * TC_SBO;   // hi + lo of one (bin, cb, kc)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k4 = (int)(i % k4n);
        int64_t r = i / k4n;
        const int n = (int)(r % n_pad);
        r /= n_pad;
        const int cb = (int)(r % ncb);
        const int b = (int)(r / ncb);
        const int64_t c0 = bin_offset[b];
        const int kb = (int)(bin_offset[b + 1] - c0);
        const int c = cb * n_pad + n;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k4 * 4 + e;
            double v = 0.0;
            if (c < kb && k < D) v = centers[(c0 + c) * D + k] - mean[(size_t)b * d_pad + k];
            const float vf = (float)v;
            hi[e] = tf32_rna(vf);
            lo[e] = vf - hi[e];
        }
        const int kc = k4 / 8, kc8 = k4 % 8;
        unsigned char* base = bprep + ((size_t)(b * ncb + cb) * nch + kc) * block_bytes;
        const size_t off = (size_t)(n / 8) * TC_SBO + (size_t)kc8 * TC_LBO + (size_t)(n % 8) * 16;
        *reinterpret_cast<float4*>(base + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(base + (size_t)ng * TC_SBO + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// one warp per (bin, padded column): centred squared norm (fp64 sum, rounded to fp32) and per-bin maxima
__global__ void __launch_bounds__(256)
    tc_csq_kernel(const double* __restrict__ centers, const double* __restrict__ csq_raw,
                  const int64_t* __restrict__ bin_offset, int32_t nbins, int D, int d_pad, int ncols,
                  const double* __restrict__ mean, float* __restrict__ csqf, float* __restrict__ cmaxf) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)nbins * ncols) return;
    const int b = (int)(w / ncols), c = (int)(w % ncols);
    const int64_t c0 = bin_offset[b];
    const int kb = (int)(bin_offset[b + 1] - c0);
    const int lane = threadIdx.x & 31;
    float out = __int_as_float(0x7f800000);  // +inf: padding columns can never win
    if (c < kb) {
        double s = 0.0;
        for (int k = lane; k < D; k += 32) {
            const double v = centers[(c0 + c) * D + k] - mean[(size_t)b * d_pad + k];
            s = fma(v, v, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        out = (float)s;
        if (lane == 0) {
            // positive floats order like their bit patterns
            atomicMax(reinterpret_cast<int*>(cmaxf + 3 * b), __float_as_int(__double2float_ru(s)));
            atomicMax(reinterpret_cast<int*>(cmaxf + 3 * b + 1), __float_as_int(__double2float_ru(csq_raw[c0 + c])));
        }
    }
    if (lane == 0) csqf[(size_t)b * ncols + c] = out;
    if (c == 0) {
        // ||mean|| of the bin (upper bound), used for ||x|| <= ||x'|| + ||mean||
        double s = 0.0;
        for (int k = lane; k < D; k += 32) {
            const double v = mean[(size_t)b * d_pad + k];
            s = fma(v, v, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) cmaxf[3 * b + 2] = __double2float_ru(sqrt(s)) * 1.000001f;
    }
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 helpers
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    // K-major, no swizzle: start address, leading (K) byte offset, stride (8-row group) byte offset, version 1
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(TC_LBO >> 4) << 16;
    d |= (uint64_t)(TC_SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, long long& acc, bool on) {
    if (on) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity);
    }
}

__device__ __forceinline__ void timed_wait(uint32_t bar32, uint32_t parity, long long& acc, bool on) {
    if (on) {
        const long long t0 = clock64();
        mbar_wait(bar32, parity);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar32, parity);
    }
}

template <int VEC>
__global__ void __launch_bounds__(TC_THREADS, 1) assign_tc_kernel(const TcParams q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const AssignParams& p = q.a;
    __shared__ uint64_t raw_full[TC_MAX_STAGES], raw_empty[TC_MAX_STAGES];   // fp64 staging ring (cp.async)
    __shared__ uint64_t tf_full[TC_MAX_STAGES], tf_empty[TC_MAX_STAGES];     // TF32 operand ring (UMMA)
    __shared__ uint64_t tmem_full[2], tmem_empty[2], xn_full[2], xn_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_xn[2][TC_TP];                     // [buffer][row]: centred ||x'||^2 (upper bound)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_raw = q.nstages_raw, n_tf = q.nstages;
    const int ng = q.n_pad / 8;
    const uint32_t b_bytes = (uint32_t)(2 * ng * TC_SBO);
    const uint32_t tf_bytes = 2u * TC_A_BYTES + b_bytes;
    unsigned char* tf_base = smem_raw;
    unsigned char* raw_base = smem_raw + (size_t)n_tf * tf_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_raw; ++s) {
            mbar_init(&raw_full[s], TC_CONV_WARPS * 32);      // every staging thread: cp.async ... arrive.noinc
            mbar_init(&raw_empty[s], TC_CONV_WARPS);          // one lane per converter warp
        }
        for (int s = 0; s < n_tf; ++s) {
            mbar_init(&tf_full[s], TC_CONV_WARPS + 1);        // converter warps + the centre warp's expect_tx arrive
            mbar_init(&tf_empty[s], 1);                       // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);    // tcgen05.commit
            mbar_init(&tmem_empty[i], 4);   // one lane per epilogue warp
            mbar_init(&xn_full[i], TC_CONV_WARPS);
            mbar_init(&xn_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // bin tables -> shared memory
    TileTables tt{p.tile_prefix, p.bin_start, p.bin_offset, p.nbins};
    if (p.nbins <= AS_TABLE_BINS) {
        int32_t* s_tp = reinterpret_cast<int32_t*>(raw_base + (size_t)n_raw * TC_RAW_BYTES);
        int32_t* s_bs = s_tp + (p.nbins + 1);
        int64_t* s_bo = reinterpret_cast<int64_t*>(s_bs + (p.nbins + 1));
        for (int b = threadIdx.x; b <= p.nbins; b += TC_THREADS) {
            s_tp[b] = p.tile_prefix[b];
            s_bs[b] = p.bin_start[b];
            s_bo[b] = p.bin_offset[b];
        }
        tt = TileTables{s_tp, s_bs, s_bo, p.nbins};
    }
    if (warp == TC_MMA_WARP) {
        // TMEM allocation (whole warp), address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(q.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int32_t n_tiles = tt.tile_prefix[p.nbins];
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int ncb = p.ncb, nch = p.nch;
    const bool prof = q.dbg_prof != nullptr;
    long long w0 = 0, w1 = 0, w2 = 0;   // cycles waited on up to three barriers of this role
    const long long t_begin = clock64();

    if (warp == TC_MMA_WARP) {
        // =========================== MMA issuer (one lane) ===========================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(q.n_pad >> 3) << 17) | ((uint32_t)(TC_TP >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t tph0 = 0, tph1 = 0;
            int unit = 0;
            for (int ti = 0; ti < my_tiles; ++ti)
                for (int cb = 0; cb < ncb; ++cb, ++unit) {
                    const int ab = unit & 1;
                    timed_wait(&tmem_empty[ab], (ab ? tph1 : tph0) ^ 1u, w0, prof);   // epilogue has drained this accumulator
                    if (ab) tph1 ^= 1u; else tph0 ^= 1u;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ab * q.n_pad);
                    for (int kc = 0; kc < nch; ++kc) {
                        timed_wait(&tf_full[stage], phase, w1, prof);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tf_base + (size_t)stage * tf_bytes);
                        const uint32_t a_hi = sa, a_lo = sa + TC_A_BYTES;
                        const uint32_t b_hi = sa + 2 * TC_A_BYTES, b_lo = b_hi + (uint32_t)(ng * TC_SBO);
#pragma unroll
                        for (int j = 0; j < TC_KC / 8; ++j) {
                            const uint32_t ko = (uint32_t)(j * 2 * TC_LBO);
                            const uint64_t dah = make_smem_desc(a_hi + ko), dal = make_smem_desc(a_lo + ko);
                            const uint64_t dbh = make_smem_desc(b_hi + ko), dbl = make_smem_desc(b_lo + ko);
                            umma_tf32(d_tmem, dah, dbh, idesc, (kc | j) != 0);   // hi.hi (first MMA overwrites)
                            umma_tf32(d_tmem, dah, dbl, idesc, 1);              // hi.lo
                            umma_tf32(d_tmem, dal, dbh, idesc, 1);              // lo.hi
                        }
                        umma_commit(&tf_empty[stage]);   // stage reusable once these MMAs have read it
                        if (kc == nch - 1) umma_commit(&tmem_full[ab]);
                        if (++stage == n_tf) { stage = 0; phase ^= 1u; }
                    }
                }
        }
    } else if (warp == TC_CENTRE_WARP) {
        // =========================== centre-block producer (one lane) ===========================
        if (lane == 0) {
            TileWalk<TC_TP> w{0, 0, 0, 0, 0, 0, 0, 0};
            w.load(tt, my_tiles);
            int stage = 0;
            uint32_t phase = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                for (int cb = 0; cb < ncb; ++cb)
                    for (int kc = 0; kc < nch; ++kc) {
                        timed_wait(&tf_empty[stage], phase ^ 1u, w0, prof);
                        unsigned char* dst = tf_base + (size_t)stage * tf_bytes + 2 * TC_A_BYTES;
                        const unsigned char* src = q.bprep + ((size_t)(w.bin * ncb + cb) * nch + kc) * b_bytes;
                        mbar_expect_tx(&tf_full[stage], b_bytes);
                        bulk_copy_g2s(dst, src, b_bytes, &tf_full[stage]);
                        if (++stage == n_tf) { stage = 0; phase ^= 1u; }
                    }
                w.next_tile(tt, my_tiles);
            }
        }
    } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + 4) {
        // =========================== epilogue: one point per thread ===========================
        const int row = (warp - TC_EPI_WARP0) * 32 + lane;
        const uint32_t lane_addr = (uint32_t)((warp - TC_EPI_WARP0) * 32) << 16;
        const float finf = __int_as_float(0x7f800000);
        TileWalk<TC_TP> w{0, 0, 0, 0, 0, 0, 0, 0};
        w.load(tt, my_tiles);
        uint32_t tph0 = 0, tph1 = 0, xph0 = 0, xph1 = 0;
        int unit = 0;
        const int ncols = ncb * q.n_pad;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int32_t pt = (row < w.pcount) ? p.perm[w.pstart + row] : -1;   // label destination, fetched early
            const float* csqf = q.csqf + (size_t)w.bin * ncols;
            float m1 = finf, m2 = finf;
            int32_t bi = 0;
            for (int cb = 0; cb < ncb; ++cb, ++unit) {
                const int ab = unit & 1;
                timed_wait(&tmem_full[ab], ab ? tph1 : tph0, w0, prof);
                if (ab) tph1 ^= 1u; else tph0 ^= 1u;
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(ab * q.n_pad);
                for (int c0 = 0; c0 < q.n_pad; c0 += 16) {
                    if (cb * q.n_pad + c0 >= w.kb) break;   // only padding beyond here (warp-uniform)
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = cb * q.n_pad + c0 + j;
                        const float sf = fmaf(-2.0f, v[j], __ldg(csqf + c));   // +inf on padding columns
                        if (q.dbg_scores && pt >= 0) q.dbg_scores[(size_t)pt * ncols + c] = sf;
                        const bool lt = sf < m1;
                        m2 = lt ? m1 : fminf(m2, sf);
                        bi = lt ? c : bi;
                        m1 = lt ? sf : m1;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[ab]);
            }
            // norms of this tile's rows from the converter warps
            const int xb = ti & 1;
            timed_wait(&xn_full[xb], xb ? xph1 : xph0, w1, prof);
            if (xb) xph1 ^= 1u; else xph0 ^= 1u;
            const float xn2c = s_xn[xb][row];
            __syncwarp();
            if (lane == 0) mbar_arrive(&xn_empty[xb]);
            if (pt >= 0) {
                p.label_out[pt] = w.coff + bi;
                if (p.local_out) p.local_out[pt] = bi;
                const float cmc = sqrtf(q.cmaxf[3 * w.bin]) * 1.000001f;       // centred max ||c'||
                const float cmr = sqrtf(q.cmaxf[3 * w.bin + 1]) * 1.000001f;   // raw max ||c||
                const float xnc = sqrtf(xn2c) * 1.000001f;
                const float xnr = xnc + q.cmaxf[3 * w.bin + 2];                  // ||x|| <= ||x'|| + ||mean||
                const float err = q.err_coef * cmc * (2.0f * xnc + cmc);           // tensor-core evaluation error
                const float tol = (float)p.tie_scale * cmr * (2.0f * xnr + cmr);   // fp64 tie band
                // both scores may be off by err: the order is certain only beyond 2 err (+ the tie band)
                if (!(m2 - m1 > 2.0f * err + 2.0f * tol)) p.recheck_list[atomicAdd(p.recheck_count, 1)] = pt;
            }
            w.next_tile(tt, my_tiles);
        }
    } else if (warp >= TC_CONV_WARP0) {
        // =========================== staging + conversion warps ===========================
        // Each of the 16 warps (a) issues the cp.async (LDGSTS, zero-filling) copies of 8 point rows of the
        // chunk n_raw-1 steps ahead into the fp64 staging ring -- fire-and-forget, completion lands on raw_full;
        // spreading the issue over all warps matters: one warp sustains only ~16 copies in flight -- and
        // (b) converts the current chunk: fp64 (shared) -> centred TF32 hi/lo tiles in the UMMA layout.
        // Conversion mapping: 8 lanes cover one row chunk (4 elements each), 16 warps cover 64 rows per pass, a
        // thread owns rows rsub and rsub + 64.  Rows past the tile end hold stale data: garbage rows of the
        // product are never read.
        const int cwp = warp - TC_CONV_WARP0;              // 0..15
        const int lt = threadIdx.x - TC_CONV_WARP0 * 32;   // 0..511
        const int slot = lt & 7;
        const int rsub = lt >> 3;                          // 0..63
        constexpr int RPP = TC_CONV_WARPS * 4;             // rows per conversion pass (64)
        constexpr int NP = TC_TP / RPP;
        constexpr int CR = TC_TP / TC_CONV_WARPS;          // rows each warp copies (8)
        // copy geometry: SEGS lanes per row chunk, RPI rows per warp instruction, XQ instructions for 16 rows
        constexpr int SEGS = TC_KC / VEC;
        constexpr int RPI = 32 / SEGS;
        constexpr int XQ = CR / RPI;
        const int seg = lane % SEGS, crs = lane / SEGS;
        const int kcol0 = seg * VEC;
        TileWalk<TC_TP> cw{0, 0, 0, 0, 0, 0, 0, 0};       // step being converted
        cw.load(tt, my_tiles);
        TileWalk<TC_TP> iw = cw;                           // step whose copies are being issued
        TileWalk<TC_TP> nw = cw;                           // tile whose point indices are being prefetched
        const double* xsrc[XQ];
        int32_t pidx_next[XQ];
        auto fetch = [&](const TileWalk<TC_TP>& t) {
#pragma unroll
            for (int qq = 0; qq < XQ; ++qq) {
                const int r = cwp * CR + qq * RPI + crs;
                pidx_next[qq] = (r < t.pcount) ? p.perm[t.pstart + r] : -1;
            }
        };
        auto set_xsrc = [&]() {
#pragma unroll
            for (int qq = 0; qq < XQ; ++qq)
                xsrc[qq] = (pidx_next[qq] >= 0) ? p.X + (int64_t)pidx_next[qq] * p.ldx + kcol0 : nullptr;
        };
        fetch(iw);
        set_xsrc();
        nw.next_tile(tt, my_tiles);
        fetch(nw);
        const int64_t total_steps = (int64_t)my_tiles * ncb * nch;
        int is = 0;
        uint32_t iphase = 0;
        int64_t issued = 0;
        auto issue_one = [&]() {
            timed_wait(&raw_empty[is], iphase ^ 1u, w2, prof);
            double* st = reinterpret_cast<double*>(raw_base + (size_t)is * TC_RAW_BYTES);
            const int k0 = iw.kc * TC_KC;
            int vbytes = (p.D - k0 - kcol0) * 8;
            vbytes = vbytes < 0 ? 0 : (vbytes > VEC * 8 ? VEC * 8 : vbytes);
            double* dst = st + (cwp * CR + crs) * TC_RAW_LD + kcol0;
#pragma unroll
            for (int qq = 0; qq < XQ; ++qq)   // (a zero-size copy still gets an in-range source address)
                if (xsrc[qq]) cp_async_zfill<VEC>(dst + qq * RPI * TC_RAW_LD, vbytes ? xsrc[qq] + k0 : p.X, vbytes);
            if (cwp == 0 && lane < 16)   // the bin-mean chunk rides along (zero padded past D: always 16 x 16 B)
                cp_async_zfill<2>(st + TC_TP * TC_RAW_LD + 2 * lane, q.mean + (size_t)iw.bin * q.d_pad + k0 + 2 * lane, 16);
            cp_async_arrive_noinc(&raw_full[is]);
            if (++is == n_raw) { is = 0; iphase ^= 1u; }
            ++issued;
            if (iw.advance(tt, ncb, nch, my_tiles)) {
                set_xsrc();
                nw.next_tile(tt, my_tiles);
                fetch(nw);
            }
        };
        while (issued < total_steps && issued < n_raw - 1) issue_one();

        int rs = 0, ts = 0;
        uint32_t rphase = 0, tphase = 0, xph0 = 0, xph1 = 0;
        const uint32_t soff = (uint32_t)(rsub >> 3) * TC_SBO + (uint32_t)slot * TC_LBO + (uint32_t)(rsub & 7) * 16;
        float xc[NP];
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) xc[ps] = 0.f;
        for (int64_t step = 0; step < total_steps; ++step) {
            if (issued < total_steps) issue_one();
            timed_wait(&raw_full[rs], rphase, w0, prof);
            const double* st = reinterpret_cast<const double*>(raw_base + (size_t)rs * TC_RAW_BYTES);
            const double2 mu0 = *reinterpret_cast<const double2*>(st + TC_TP * TC_RAW_LD + 4 * slot);
            const double2 mu1 = *reinterpret_cast<const double2*>(st + TC_TP * TC_RAW_LD + 4 * slot + 2);
            double2 xv[NP][2];
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
                const double* src = st + (rsub + RPP * ps) * TC_RAW_LD + 4 * slot;
                xv[ps][0] = *reinterpret_cast<const double2*>(src);
                xv[ps][1] = *reinterpret_cast<const double2*>(src + 2);
            }
            timed_wait(&tf_empty[ts], tphase ^ 1u, w1, prof);
            unsigned char* sa = tf_base + (size_t)ts * tf_bytes + soff;
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
                const float x0 = (float)(xv[ps][0].x - mu0.x), x1 = (float)(xv[ps][0].y - mu0.y);
                const float x2 = (float)(xv[ps][1].x - mu1.x), x3 = (float)(xv[ps][1].y - mu1.y);
                const float h0 = tf32_rna(x0), h1 = tf32_rna(x1), h2 = tf32_rna(x2), h3 = tf32_rna(x3);
                if (cw.cb == 0) xc[ps] = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, xc[ps]))));
                *reinterpret_cast<float4*>(sa + ps * ((RPP / 8) * TC_SBO)) = make_float4(h0, h1, h2, h3);
                *reinterpret_cast<float4*>(sa + ps * ((RPP / 8) * TC_SBO) + TC_A_BYTES) = make_float4(x0 - h0, x1 - h1, x2 - h2, x3 - h3);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tf_full[ts]);
                mbar_arrive(&raw_empty[rs]);
            }
            if (++rs == n_raw) { rs = 0; rphase ^= 1u; }
            if (++ts == n_tf) { ts = 0; tphase ^= 1u; }
            if (cw.kc == nch - 1 && cw.cb == ncb - 1) {
                // centred ||x'||^2 of this tile's rows -> epilogue (fp32 sums of fp32 roundings, inflated a little)
                const int xbuf = cw.ti & 1;
                mbar_wait(&xn_empty[xbuf], (xbuf ? xph1 : xph0) ^ 1u);
                if (xbuf) xph1 ^= 1u; else xph0 ^= 1u;
#pragma unroll
                for (int ps = 0; ps < NP; ++ps) {
                    float a = xc[ps];
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 4);
                    if (slot == 0) s_xn[xbuf][rsub + RPP * ps] = a * 1.0001f;
                    xc[ps] = 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xn_full[xbuf]);
            }
            cw.advance(tt, ncb, nch, my_tiles);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }

    if (prof && lane == 0 && (warp == TC_MMA_WARP || warp == TC_CENTRE_WARP || warp == TC_EPI_WARP0 || warp == TC_CONV_WARP0)) {
        // slots: role * 4 + {total cycles, wait 0, wait 1, wait 2}; roles: 0 MMA, 1 centre, 2 epilogue, 3 raw, 4 converter
        const int role = warp == TC_MMA_WARP ? 0 : warp == TC_CENTRE_WARP ? 1 : warp == TC_EPI_WARP0 ? 2 : 4;
        atomicAdd(q.dbg_prof + role * 4 + 0, (unsigned long long)(clock64() - t_begin));
        atomicAdd(q.dbg_prof + role * 4 + 1, (unsigned long long)w0);
        atomicAdd(q.dbg_prof + role * 4 + 2, (unsigned long long)w1);
        atomicAdd(q.dbg_prof + role * 4 + 3, (unsigned long long)w2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(q.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// main kernel, second generation: the split A operand lives in TENSOR MEMORY
// ---------------------------------------------------------------------------------------------------------
// The kernel above is bound by shared-memory bandwidth, not by HBM or the tensor pipe: per 32-element k-chunk of a
// 128-point tile (32 KB of HBM traffic) the shared-memory port moves 32 KB of cp.async writes, 32 KB of converter
// reads, 32 KB of hi/lo operand stores, 48 KB of UMMA A-operand reads (hi twice, lo once), 43 KB of UMMA B reads and
// 28 KB of centre-block TMA writes = 215 KB, i.e. 1 680 port cycles against the 1 460 cycles the chunk's HBM share
// takes -- and the two 69 KB operand stages leave room for only two fp64 staging buffers (~32 KB of loads in flight
// per SM, short of what the HBM latency needs).  Here the converter threads own one point ROW each (the TMEM lane of
// that row), read its k-slice from the fp64 staging ring and write the TF32 hi / lo values straight into tensor
// memory with tcgen05.st; the MMAs take A from TMEM ([d], [a], b-desc form).  That removes the operand stores and the
// A reads from shared memory (135 KB per chunk instead of 215 KB) and frees 74 KB for the staging ring (4 stages).
//
// TMEM columns: [0, 2 n_pad) two accumulators; then n_a A stages of 64 columns (32 hi + 32 lo, K = 32 per chunk).
// Warp roles as above; a converter warp may only touch the TMEM lane quarter (warp id % 4), so warp w converts rows
// 32 (w % 4) .. +31 and the k-half ((w - 6) / 4) of the chunk: 16 elements per thread and chunk.

static constexpr int TC2_A_STAGE_COLS = 64;
static constexpr int TC2_MAX_A_STAGES = 4;
static constexpr int TC2_MAX_B_STAGES = 4;
static constexpr int TC2_CONV_WARPS = 8;                   // 2 per TMEM lane quarter: each converts 16 of the chunk's 32 elements
                                                           // (16 warps x 8 elements measured slower: 0.999 vs 0.903 ms, cfg5 slice)
static constexpr int TC2_KSUB = TC_KC / (TC2_CONV_WARPS / 4);   // elements per thread and chunk
static constexpr int TC2_THREADS = (TC_CONV_WARP0 + TC2_CONV_WARPS) * 32;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

struct Tc2Params {
    TcParams t;
    int ablate;       // tuning only: bit 0 = issue no MMAs, bit 1 = skip the fp64 -> TF32 conversion arithmetic, bit 2 = no centre-block loads
    int n_a;          // A stages in tensor memory
    int n_b;          // centre-block stages in shared memory
    uint32_t a_col0;  // first TMEM column of the A stages
};

// FULLK: D is a multiple of the k-chunk, so no copy needs a zero-filled tail; PROF: per-role wait-cycle counters.
template <int VEC, bool FULLK, bool PROF>
__global__ void __launch_bounds__(TC2_THREADS, 1) assign_tc2_kernel(const Tc2Params qq) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TcParams& q = qq.t;
    const AssignParams& p = q.a;
    __shared__ uint64_t raw_full[TC_MAX_STAGES], raw_empty[TC_MAX_STAGES];   // fp64 staging ring (cp.async)
    __shared__ uint64_t a_full[TC2_MAX_A_STAGES], a_empty[TC2_MAX_A_STAGES]; // A stages in TMEM
    __shared__ uint64_t b_full[TC2_MAX_B_STAGES], b_empty[TC2_MAX_B_STAGES]; // centre blocks in shared memory
    __shared__ uint64_t tmem_full[2], tmem_empty[2], xn_full[2], xn_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_xn[2][TC2_CONV_WARPS / 4][TC_TP];    // [buffer][k-slice][row]: partial centred ||x'||^2

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_raw = q.nstages_raw, n_a = qq.n_a, n_b = qq.n_b;
    const int ablate = PROF ? qq.ablate : 0;          // measurement switches exist only in the profiled instantiation
    const int ng = q.n_pad / 8;
    const uint32_t b_bytes = (uint32_t)(2 * ng * TC_SBO);
    unsigned char* b_base = smem_raw;
    unsigned char* raw_base = smem_raw + (size_t)n_b * b_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_raw; ++s) {
            mbar_init(&raw_full[s], TC2_CONV_WARPS * 32);     // every staging thread: cp.async ... arrive.noinc
            mbar_init(&raw_empty[s], TC2_CONV_WARPS);         // one lane per converter warp
        }
        for (int s = 0; s < n_a; ++s) {
            mbar_init(&a_full[s], TC2_CONV_WARPS);            // one lane per converter warp, after its tcgen05.st
            mbar_init(&a_empty[s], 1);                        // tcgen05.commit
        }
        for (int s = 0; s < n_b; ++s) {
            mbar_init(&b_full[s], 1);                         // the centre warp's expect_tx arrive
            mbar_init(&b_empty[s], 1);                        // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);    // tcgen05.commit
            mbar_init(&tmem_empty[i], 4);   // one lane per epilogue warp
            mbar_init(&xn_full[i], TC2_CONV_WARPS);
            mbar_init(&xn_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    TileTables tt{p.tile_prefix, p.bin_start, p.bin_offset, p.nbins};
    if (p.nbins <= AS_TABLE_BINS) {
        int32_t* s_tp = reinterpret_cast<int32_t*>(raw_base + (size_t)n_raw * TC_RAW_BYTES);
        int32_t* s_bs = s_tp + (p.nbins + 1);
        int64_t* s_bo = reinterpret_cast<int64_t*>(s_bs + (p.nbins + 1));
        for (int b = threadIdx.x; b <= p.nbins; b += TC2_THREADS) {
            s_tp[b] = p.tile_prefix[b];
            s_bs[b] = p.bin_start[b];
            s_bo[b] = p.bin_offset[b];
        }
        tt = TileTables{s_tp, s_bs, s_bo, p.nbins};
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(q.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int32_t n_tiles = tt.tile_prefix[p.nbins];
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int ncb = p.ncb, nch = p.nch;
    constexpr bool prof = PROF;
    long long w0 = 0, w1 = 0, w2 = 0;
    const long long t_begin = clock64();

    if (warp == TC_MMA_WARP) {
        // =========================== MMA issuer (one lane) ===========================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(q.n_pad >> 3) << 17) | ((uint32_t)(TC_TP >> 4) << 24);
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            uint32_t tph0 = 0, tph1 = 0;
            int unit = 0;
            for (int ti = 0; ti < my_tiles; ++ti)
                for (int cb = 0; cb < ncb; ++cb, ++unit) {
                    const int ab = unit & 1;
                    timed_wait(&tmem_empty[ab], (ab ? tph1 : tph0) ^ 1u, w0, prof);   // epilogue has drained this accumulator
                    if (ab) tph1 ^= 1u; else tph0 ^= 1u;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ab * q.n_pad);
                    for (int kc = 0; kc < nch; ++kc) {
                        timed_wait(&b_full[bs], bph, w2, prof);
                        timed_wait(&a_full[as], aph, w1, prof);
                        tc_fence_after();
                        const uint32_t a_hi = tmem_base + qq.a_col0 + (uint32_t)(as * TC2_A_STAGE_COLS), a_lo = a_hi + TC_KC;
                        const uint32_t b_hi = smem_u32(b_base + (size_t)bs * b_bytes), b_lo = b_hi + (uint32_t)(ng * TC_SBO);
                        if (!(ablate & 1))
#pragma unroll
                        for (int j = 0; j < TC_KC / 8; ++j) {
                            const uint32_t ko = (uint32_t)(j * 2 * TC_LBO);
                            const uint64_t dbh = make_smem_desc(b_hi + ko), dbl = make_smem_desc(b_lo + ko);
                            umma_tf32_ts(d_tmem, a_hi + 8u * j, dbh, idesc, (kc | j) != 0);   // hi.hi (first MMA overwrites)
                            umma_tf32_ts(d_tmem, a_hi + 8u * j, dbl, idesc, 1);              // hi.lo
                            umma_tf32_ts(d_tmem, a_lo + 8u * j, dbh, idesc, 1);              // lo.hi
                        }
                        umma_commit(&a_empty[as]);       // stages reusable once these MMAs have read them
                        umma_commit(&b_empty[bs]);
                        if (kc == nch - 1) umma_commit(&tmem_full[ab]);
                        if (++as == n_a) { as = 0; aph ^= 1u; }
                        if (++bs == n_b) { bs = 0; bph ^= 1u; }
                    }
                }
        }
    } else if (warp == TC_CENTRE_WARP) {
        // =========================== centre-block producer (one lane) ===========================
        if (lane == 0) {
            TileWalk<TC_TP> w{0, 0, 0, 0, 0, 0, 0, 0};
            w.load(tt, my_tiles);
            int bs = 0;
            uint32_t bph = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                for (int cb = 0; cb < ncb; ++cb)
                    for (int kc = 0; kc < nch; ++kc) {
                        timed_wait(&b_empty[bs], bph ^ 1u, w0, prof);
                        unsigned char* dst = b_base + (size_t)bs * b_bytes;
                        const unsigned char* src = q.bprep + ((size_t)(w.bin * ncb + cb) * nch + kc) * b_bytes;
                        if (ablate & 4) {
                            mbar_arrive(&b_full[bs]);          // measurement only: no centre-block traffic
                        } else {
                            mbar_expect_tx(&b_full[bs], b_bytes);
                            bulk_copy_g2s(dst, src, b_bytes, &b_full[bs]);
                        }
                        if (++bs == n_b) { bs = 0; bph ^= 1u; }
                    }
                w.next_tile(tt, my_tiles);
            }
        }
    } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + 4) {
        // =========================== epilogue: one point per thread ===========================
        const int row = (warp - TC_EPI_WARP0) * 32 + lane;
        const uint32_t lane_addr = (uint32_t)((warp - TC_EPI_WARP0) * 32) << 16;
        const float finf = __int_as_float(0x7f800000);
        TileWalk<TC_TP> w{0, 0, 0, 0, 0, 0, 0, 0};
        w.load(tt, my_tiles);
        uint32_t tph0 = 0, tph1 = 0, xph0 = 0, xph1 = 0;
        int unit = 0;
        const int ncols = ncb * q.n_pad;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int32_t pt = (row < w.pcount) ? p.perm[w.pstart + row] : -1;   // label destination, fetched early
            const float* csqf = q.csqf + (size_t)w.bin * ncols;
            float m1 = finf, m2 = finf;
            int32_t bi = 0;
            for (int cb = 0; cb < ncb; ++cb, ++unit) {
                const int ab = unit & 1;
                timed_wait(&tmem_full[ab], ab ? tph1 : tph0, w0, prof);
                if (ab) tph1 ^= 1u; else tph0 ^= 1u;
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(ab * q.n_pad);
                for (int c0 = 0; c0 < q.n_pad; c0 += 16) {
                    if (cb * q.n_pad + c0 >= w.kb) break;   // only padding beyond here (warp-uniform)
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = cb * q.n_pad + c0 + j;
                        const float sf = fmaf(-2.0f, v[j], __ldg(csqf + c));   // +inf on padding columns
                        if (q.dbg_scores && pt >= 0) q.dbg_scores[(size_t)pt * ncols + c] = sf;
                        const bool lt = sf < m1;
                        m2 = lt ? m1 : fminf(m2, sf);
                        bi = lt ? c : bi;
                        m1 = lt ? sf : m1;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[ab]);
            }
            const int xb = ti & 1;
            timed_wait(&xn_full[xb], xb ? xph1 : xph0, w1, prof);
            if (xb) xph1 ^= 1u; else xph0 ^= 1u;
            float xn2c = 0.f;
#pragma unroll
            for (int h = 0; h < TC2_CONV_WARPS / 4; ++h) xn2c += s_xn[xb][h][row];
            __syncwarp();
            if (lane == 0) mbar_arrive(&xn_empty[xb]);
            if (pt >= 0) {
                p.label_out[pt] = w.coff + bi;
                if (p.local_out) p.local_out[pt] = bi;
                const float cmc = sqrtf(q.cmaxf[3 * w.bin]) * 1.000001f;       // centred max ||c'||
                const float cmr = sqrtf(q.cmaxf[3 * w.bin + 1]) * 1.000001f;   // raw max ||c||
                const float xnc = sqrtf(xn2c) * 1.000001f;
                const float xnr = xnc + q.cmaxf[3 * w.bin + 2];                  // ||x|| <= ||x'|| + ||mean||
                const float err = q.err_coef * cmc * (2.0f * xnc + cmc);           // tensor-core evaluation error
                const float tol = (float)p.tie_scale * cmr * (2.0f * xnr + cmr);   // fp64 tie band
                if (!(m2 - m1 > 2.0f * err + 2.0f * tol)) p.recheck_list[atomicAdd(p.recheck_count, 1)] = pt;
            }
            w.next_tile(tt, my_tiles);
        }
    } else if (warp >= TC_CONV_WARP0) {
        // =========================== staging + conversion warps ===========================
        // Each of the 8 warps (a) issues the cp.async (LDGSTS) copies of 16 point rows of the chunk n_raw-1 steps ahead
        // into the fp64 staging ring -- fire and forget, completion lands on raw_full -- and (b) converts ITS rows of
        // the current chunk: thread = one point row (the TMEM lane it may write), 16 consecutive elements.
        // These warps are ISSUE-bound (ncu: nothing else on the SM is above 45 %; the first version of this loop spent
        // ~550 instructions per step and warp, 170 of them on predicates / selects / 64-bit address pairs around the 8
        // copies), so the loop is kept lean: row pointers are formed once per tile, rows past the end of a tile read row
        // 0 instead of being predicated off (their products are never looked at), full chunks use the plain copy, and
        // the walk over (tile, centre block, k-chunk) is three counters.
        const int cwp = warp - TC_CONV_WARP0;              // 0..TC2_CONV_WARPS-1
        const int quarter = warp & 3;                      // TMEM lane quarter this warp can access
        const int khalf = cwp >> 2;                        // which TC2_KSUB elements of the chunk
        const int row = quarter * 32 + lane;
        constexpr int CR = TC_TP / TC2_CONV_WARPS;         // rows each warp copies
        constexpr int SEGS = TC_KC / VEC;
        constexpr int RPI = 32 / SEGS;
        constexpr int XQ = CR / RPI;
        const int seg = lane % SEGS, crs = lane / SEGS;
        const int kcol0 = seg * VEC;
        TileWalk<TC_TP> iw{0, 0, 0, 0, 0, 0, 0, 0};       // tile whose copies are being issued
        iw.load(tt, my_tiles);
        TileWalk<TC_TP> nw = iw;                           // tile whose point indices are being prefetched
        uint32_t xrow[XQ];                                 // point indices of this thread's rows of the tile being issued
        int32_t pidx_next[XQ];                             // (an index, not a pointer: one IMAD.WIDE per copy, no register-pair moves)
        const char* const Xb = reinterpret_cast<const char*>(p.X + kcol0);
        const uint32_t ldx8 = (uint32_t)p.ldx * 8u;
        auto fetch = [&](const TileWalk<TC_TP>& t) {
#pragma unroll
            for (int k = 0; k < XQ; ++k) {
                const int r = cwp * CR + k * RPI + crs;
                pidx_next[k] = (r < t.pcount) ? p.perm[t.pstart + r] : 0;
            }
        };
        auto set_rows = [&]() {
#pragma unroll
            for (int k = 0; k < XQ; ++k) xrow[k] = (uint32_t)pidx_next[k];
        };
        fetch(iw);
        set_rows();
        nw.next_tile(tt, my_tiles);
        fetch(nw);
        const int64_t total_steps = (int64_t)my_tiles * ncb * nch;
        const uint32_t raw32 = smem_u32(raw_base) + (uint32_t)(((cwp * CR + crs) * TC_RAW_LD + kcol0) * 8);
        const uint32_t mean32 = smem_u32(raw_base) + (uint32_t)((TC_TP * TC_RAW_LD + 2 * lane) * 8);
        const double* mean_src = q.mean + (size_t)iw.bin * q.d_pad + 2 * lane;
        // one opaque base register + constant offsets (left to itself the compiler re-derives every barrier address from
        // the shared-window base at each use: 4 instructions per wait / arrive)
        uint32_t bar0;
        asm volatile("mov.u32 %0, %1;" : "=r"(bar0) : "r"(smem_u32(raw_full)));
        const uint32_t raw_full32 = bar0, raw_empty32 = bar0 + (smem_u32(raw_empty) - smem_u32(raw_full)),
                       a_full32 = bar0 + (smem_u32(a_full) - smem_u32(raw_full)),
                       a_empty32 = bar0 + (smem_u32(a_empty) - smem_u32(raw_full)),
                       xn_full32 = bar0 + (smem_u32(xn_full) - smem_u32(raw_full)),
                       xn_empty32 = bar0 + (smem_u32(xn_empty) - smem_u32(raw_full));
        int is = 0, ikc = 0, icb = 0;
        uint32_t iphase = 0;
        int64_t issued = 0;
        auto issue_one = [&]() {
            timed_wait(raw_empty32 + 8u * (uint32_t)is, iphase ^ 1u, w2, prof);
            const uint32_t dst = raw32 + (uint32_t)is * (uint32_t)TC_RAW_BYTES;
            const int k0 = ikc * TC_KC;
            if (FULLK) {
                const char* const Xk = Xb + (size_t)k0 * 8;
#pragma unroll
                for (int k = 0; k < XQ; ++k) {
                    const char* src = Xk + (size_t)xrow[k] * ldx8;
                    if (VEC == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(k * RPI * TC_RAW_LD * 8)), "l"(src) : "memory");
                    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (uint32_t)(k * RPI * TC_RAW_LD * 8)), "l"(src) : "memory");
                }
            } else {
                int vbytes = (p.D - k0 - kcol0) * 8;
                vbytes = vbytes < 0 ? 0 : (vbytes > VEC * 8 ? VEC * 8 : vbytes);
                const char* const Xk = vbytes ? Xb + (size_t)k0 * 8 : reinterpret_cast<const char*>(p.X);   // (a zero-size copy still gets an in-range source address)
                const uint32_t stride8 = vbytes ? ldx8 : 0u;
#pragma unroll
                for (int k = 0; k < XQ; ++k) {
                    const char* src = Xk + (size_t)xrow[k] * stride8;
                    if (VEC == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(k * RPI * TC_RAW_LD * 8)), "l"(src), "r"(vbytes) : "memory");
                    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + (uint32_t)(k * RPI * TC_RAW_LD * 8)), "l"(src), "r"(vbytes) : "memory");
                }
            }
            if (cwp == 0 && lane < 16)   // the bin-mean chunk rides along (zero padded past D: always 16 x 16 B)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mean32 + (uint32_t)is * (uint32_t)TC_RAW_BYTES), "l"(mean_src + k0) : "memory");
            cp_async_arrive_noinc(raw_full32 + 8u * (uint32_t)is);
            if (++is == n_raw) { is = 0; iphase ^= 1u; }
            ++issued;
            if (++ikc == nch) {
                ikc = 0;
                if (++icb == ncb) {                        // next tile: its indices were prefetched a tile ago
                    icb = 0;
                    iw.next_tile(tt, my_tiles);
                    set_rows();
                    mean_src = q.mean + (size_t)iw.bin * q.d_pad + 2 * lane;
                    nw.next_tile(tt, my_tiles);
                    fetch(nw);
                }
            }
        };
        while (issued < total_steps && issued < n_raw - 1) issue_one();

        int rs = 0, as = 0, ckc = 0, ccb = 0, cti = 0;
        uint32_t rphase = 0, aphase = 0, xph0 = 0, xph1 = 0;
        const uint32_t ta0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + qq.a_col0 + (uint32_t)(TC2_KSUB * khalf);
        const uint32_t src32 = smem_u32(raw_base) + (uint32_t)((row * TC_RAW_LD + TC2_KSUB * khalf) * 8);
        const uint32_t mu32 = smem_u32(raw_base) + (uint32_t)((TC_TP * TC_RAW_LD + TC2_KSUB * khalf) * 8);
        float xc = 0.f;
        for (int64_t step = 0; step < total_steps; ++step) {
            if (issued < total_steps) issue_one();
            timed_wait(raw_full32 + 8u * (uint32_t)rs, rphase, w0, prof);
            const uint32_t so = (uint32_t)rs * (uint32_t)TC_RAW_BYTES;
            float hi[TC2_KSUB], lo[TC2_KSUB];
            if (ablate & 2) {
#pragma unroll
                for (int e = 0; e < TC2_KSUB; ++e) hi[e] = lo[e] = 0.f;
            } else {
                double2 xv[TC2_KSUB / 2];
                float xs = xc;
#pragma unroll
                for (int e = 0; e < TC2_KSUB / 2; ++e)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xv[e].x), "=d"(xv[e].y) : "r"(src32 + so + 16u * e));
#pragma unroll
                for (int e = 0; e < TC2_KSUB / 2; ++e) {
                    double2 mv;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(mv.x), "=d"(mv.y) : "r"(mu32 + so + 16u * e));
                    const float x0 = (float)(xv[e].x - mv.x), x1 = (float)(xv[e].y - mv.y);
                    hi[2 * e] = tf32_rna(x0);
                    hi[2 * e + 1] = tf32_rna(x1);
                    lo[2 * e] = x0 - hi[2 * e];
                    lo[2 * e + 1] = x1 - hi[2 * e + 1];
                    xs = fmaf(x0, x0, fmaf(x1, x1, xs));
                }
                if (ccb == 0) xc = xs;                     // (one select, not one per pair)
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(raw_empty32 + 8u * (uint32_t)rs);   // the staged fp64 chunk is in registers
            timed_wait(a_empty32 + 8u * (uint32_t)as, aphase ^ 1u, w1, prof);   // the MMAs that read this TMEM stage have completed
            tc_fence_after();
            const uint32_t ta = ta0 + (uint32_t)(as * TC2_A_STAGE_COLS);
            if (TC2_KSUB == 8) {
                tmem_st8(ta, hi);
                tmem_st8(ta + TC_KC, lo);
            } else {
                tmem_st16(ta, hi);
                tmem_st16(ta + TC_KC, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full32 + 8u * (uint32_t)as);
            if (++rs == n_raw) { rs = 0; rphase ^= 1u; }
            if (++as == n_a) { as = 0; aphase ^= 1u; }
            if (++ckc == nch) {
                ckc = 0;
                if (++ccb == ncb) {
                    // partial centred ||x'||^2 of this thread's row -> epilogue (fp32 sums of fp32 roundings, inflated a little)
                    ccb = 0;
                    const int xbuf = cti & 1;
                    mbar_wait(xn_empty32 + 8u * (uint32_t)xbuf, (xbuf ? xph1 : xph0) ^ 1u);
                    if (xbuf) xph1 ^= 1u; else xph0 ^= 1u;
                    s_xn[xbuf][khalf][row] = xc * 1.0001f;
                    xc = 0.f;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(xn_full32 + 8u * (uint32_t)xbuf);
                    ++cti;
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }

    if (prof && lane == 0 && (warp == TC_MMA_WARP || warp == TC_CENTRE_WARP || warp == TC_EPI_WARP0 || warp == TC_CONV_WARP0)) {
        const int role = warp == TC_MMA_WARP ? 0 : warp == TC_CENTRE_WARP ? 1 : warp == TC_EPI_WARP0 ? 2 : 4;
        atomicAdd(q.dbg_prof + role * 4 + 0, (unsigned long long)(clock64() - t_begin));
        atomicAdd(q.dbg_prof + role * 4 + 1, (unsigned long long)w0);
        atomicAdd(q.dbg_prof + role * 4 + 2, (unsigned long long)w1);
        atomicAdd(q.dbg_prof + role * 4 + 3, (unsigned long long)w2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(q.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------

static int tc_n_pad(int32_t max_k) {
    int n = (max_k + 15) / 16 * 16;
    return n > 256 ? 256 : (n < 16 ? 16 : n);
}

struct TcLayout {
    int n_pad, ncb, nch, d_pad;
    size_t bprep_bytes, mean_bytes, csqf_bytes, cmax_bytes;
};
static TcLayout tc_layout(int32_t nbins, int D, int32_t max_k) {
    TcLayout L;
    L.n_pad = tc_n_pad(max_k);
    L.ncb = (max_k + L.n_pad - 1) / L.n_pad;
    L.nch = (D + TC_KC - 1) / TC_KC;
    L.d_pad = L.nch * TC_KC;
    L.bprep_bytes = (size_t)nbins * L.ncb * L.nch * 2 * (L.n_pad / 8) * TC_SBO;
    L.mean_bytes = (size_t)nbins * L.d_pad * sizeof(double);
    L.csqf_bytes = (size_t)nbins * L.ncb * L.n_pad * sizeof(float);
    L.cmax_bytes = (size_t)nbins * 3 * sizeof(float);
    return L;
}

size_t assign_tc_prep_bytes(int32_t nbins, int D, int32_t max_k) {
    const TcLayout L = tc_layout(nbins, D, max_k);
    return align_up(L.bprep_bytes, 256) + align_up(L.mean_bytes, 256) + align_up(L.csqf_bytes, 256) +
           align_up(L.cmax_bytes, 256) + 1024;
}

static unsigned long long* g_dbg_prof = nullptr;
static float* g_dbg_scores = nullptr;   // set through mwe_debug_set_tc_scores (tests only)

int launch_assign_tc(const AssignParams& p_in, int32_t max_k, int64_t N, void* prep, size_t prep_bytes, cudaStream_t stream) {
    const TcLayout L = tc_layout(p_in.nbins, p_in.D, max_k);
    if (prep_bytes < assign_tc_prep_bytes(p_in.nbins, p_in.D, max_k)) {
        set_last_error("assign(tc): preparation workspace too small");
        return MWE_E_WORKSPACE;
    }
    Carver cv(prep, prep_bytes);
    unsigned char* bprep = cv.take<unsigned char>(L.bprep_bytes);
    double* mean = cv.take<double>((size_t)p_in.nbins * L.d_pad);
    float* csqf = cv.take<float>((size_t)p_in.nbins * L.ncb * L.n_pad);
    float* cmaxf = cv.take<float>((size_t)p_in.nbins * 3);

    // ---- preparation (depends on the centres only) ----
    MWE_CHECK_CUDA(cudaMemsetAsync(cmaxf, 0, L.cmax_bytes, stream));
    tc_mean_kernel<<<dim3((unsigned)p_in.nbins, (unsigned)((L.d_pad + 127) / 128)), 128, 0, stream>>>(p_in.centers, p_in.bin_offset,
                                                                                                  p_in.D, L.d_pad, mean);
    {
        const int64_t total = (int64_t)p_in.nbins * L.ncb * L.n_pad * (L.d_pad / 4);
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        tc_split_centers_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p_in.centers, p_in.bin_offset, p_in.nbins, p_in.D, L.d_pad,
                                                                     L.n_pad, L.ncb, mean, bprep);
        const int64_t warps = (int64_t)p_in.nbins * L.ncb * L.n_pad;
        tc_csq_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(p_in.centers, p_in.csq, p_in.bin_offset, p_in.nbins, p_in.D,
                                                                      L.d_pad, L.ncb * L.n_pad, mean, csqf, cmaxf);
    }
    MWE_CHECK_LAUNCH();

    // ---- main kernel ----
    TcParams q;
    q.a = p_in;
    q.a.ncb = L.ncb;
    q.a.nch = L.nch;
    q.bprep = bprep; q.mean = mean; q.csqf = csqf; q.cmaxf = cmaxf;
    q.n_pad = L.n_pad; q.d_pad = L.d_pad;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * L.n_pad)) cols <<= 1;
    q.tmem_cols = cols;
    // |score error| <= err_coef * cmax' (2 ||x'|| + cmax'):
    //   operands: fp32 rounding (2^-24) + TF32 split residual (<= 2^-21 each)      -> 2^-20 on a product
    //   dropped lo.lo term                                                       -> 2^-22
    //   fp32 accumulation in the tensor core: <= 2 ulp per MMA, 3 D/8 MMAs         -> 3 ceil(D/8) 2^-22
    //   fp32 score arithmetic (csq rounding, the fma)                            -> 2^-21
    q.err_coef = (float)(1.0 / 1048576.0 + (3.0 * ((p_in.D + 7) / 8) + 10.0) / 4194304.0);
    q.dbg_scores = g_dbg_scores;
    q.dbg_prof = g_dbg_prof;
    const size_t table_bytes = (p_in.nbins <= AS_TABLE_BINS) ? (size_t)(p_in.nbins + 2) * 16 + 16 : 0;
    const bool vec2 = (p_in.D % 2 == 0) && (p_in.ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(p_in.X) & 15) == 0);
    const int64_t max_tiles = (N + TC_TP - 1) / TC_TP + p_in.nbins;
    int64_t grid = sm_count();
    if (grid > max_tiles) grid = max_tiles;
    if (grid < 1) grid = 1;
    cudaEvent_t ev0, ev1;
    timing_events(&ev0, &ev1);

    // generation 2 (A operand in tensor memory) whenever two accumulators and >= 2 A stages fit the 512 TMEM columns
    int gen = (2 * L.n_pad + 2 * TC2_A_STAGE_COLS <= 512) ? 2 : 1;
    if (const char* e = getenv("MWE_TC_KERNEL")) gen = (atoi(e) == 1 || gen == 1) ? 1 : 2;
    if (gen == 2) {
        Tc2Params q2;
        const size_t b_bytes = (size_t)2 * (L.n_pad / 8) * TC_SBO;
        int n_a = (512 - 2 * L.n_pad) / TC2_A_STAGE_COLS;
        if (n_a > TC2_MAX_A_STAGES) n_a = TC2_MAX_A_STAGES;
        int n_b = 2;
        if (const char* e = getenv("MWE_TC_B_STAGES")) n_b = atoi(e);
        if (n_b < 1) n_b = 1;
        if (n_b > TC2_MAX_B_STAGES) n_b = TC2_MAX_B_STAGES;
        while (n_b > 1 && (size_t)n_b * b_bytes + 2 * (size_t)TC_RAW_BYTES + table_bytes > TC_SMEM_BUDGET) --n_b;
        if ((size_t)n_b * b_bytes + 2 * (size_t)TC_RAW_BYTES + table_bytes > TC_SMEM_BUDGET) {
            set_last_error("assign(tc): pipeline stages do not fit in shared memory");
            return MWE_E_UNSUPPORTED;
        }
        int n_raw = (int)((TC_SMEM_BUDGET - table_bytes - (size_t)n_b * b_bytes) / TC_RAW_BYTES);
        if (n_raw > TC_MAX_STAGES) n_raw = TC_MAX_STAGES;
        if (const char* e = getenv("MWE_TC_RAW_STAGES")) n_raw = atoi(e) < n_raw ? (atoi(e) < 2 ? 2 : atoi(e)) : n_raw;
        q.nstages = n_b;
        q.nstages_raw = n_raw;
        uint32_t cols2 = 32;
        while (cols2 < (uint32_t)(2 * L.n_pad + n_a * TC2_A_STAGE_COLS)) cols2 <<= 1;
        q.tmem_cols = cols2;
        q2.t = q;
        q2.n_a = n_a;
        q2.ablate = 0;
        if (const char* e = getenv("MWE_TC_ABLATE")) q2.ablate = atoi(e);
        q2.n_b = n_b;
        q2.a_col0 = (uint32_t)(2 * L.n_pad);
        const size_t smem = b_bytes * n_b + (size_t)TC_RAW_BYTES * n_raw + table_bytes;
        const bool fullk = (p_in.D % TC_KC) == 0;
        const bool profiled = q.dbg_prof != nullptr;
        void (*kern)(const Tc2Params) = nullptr;
        const int variant = (vec2 ? 4 : 0) | (fullk ? 2 : 0) | (profiled ? 1 : 0);
        switch (variant) {
            case 7: kern = assign_tc2_kernel<2, true, true>; break;
            case 6: kern = assign_tc2_kernel<2, true, false>; break;
            case 5: kern = assign_tc2_kernel<2, false, true>; break;
            case 4: kern = assign_tc2_kernel<2, false, false>; break;
            case 3: kern = assign_tc2_kernel<1, true, true>; break;
            case 2: kern = assign_tc2_kernel<1, true, false>; break;
            case 1: kern = assign_tc2_kernel<1, false, true>; break;
            default: kern = assign_tc2_kernel<1, false, false>; break;
        }
        static size_t configured_dev[MWE_MAX_DEVICES][8] = {};   // the attribute is per device (and per instantiation)
        size_t& configured = configured_dev[device_slot()][variant];
        if (configured < smem) {
            MWE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        if (ev0) MWE_CHECK_CUDA(cudaEventRecord(ev0, stream));
        kern<<<(unsigned)grid, TC2_THREADS, smem, stream>>>(q2);
        MWE_CHECK_LAUNCH();
        if (ev1) MWE_CHECK_CUDA(cudaEventRecord(ev1, stream));
        return MWE_OK;
    }

    const size_t tf_bytes = 2 * (size_t)TC_A_BYTES + (size_t)2 * (L.n_pad / 8) * TC_SBO;
    // two TF32 stages when they fit next to two fp64 stages, the rest of the budget goes to the fp64 ring
    // (that ring is what keeps HBM requests in flight)
    int n_tf = (2 * tf_bytes + 2 * (size_t)TC_RAW_BYTES + table_bytes <= TC_SMEM_BUDGET) ? 2 : 1;
    if (const char* e = getenv("MWE_TC_TF_STAGES")) n_tf = atoi(e);
    if ((size_t)n_tf * tf_bytes + 2 * (size_t)TC_RAW_BYTES + table_bytes > TC_SMEM_BUDGET) {
        set_last_error("assign(tc): pipeline stages do not fit in shared memory");
        return MWE_E_UNSUPPORTED;
    }
    int n_raw = (int)((TC_SMEM_BUDGET - table_bytes - (size_t)n_tf * tf_bytes) / TC_RAW_BYTES);
    if (n_raw > TC_MAX_STAGES) n_raw = TC_MAX_STAGES;
    q.nstages = n_tf;
    q.nstages_raw = n_raw;
    const size_t smem = tf_bytes * n_tf + (size_t)TC_RAW_BYTES * n_raw + table_bytes;
    if (vec2) {
        static size_t configured_dev[MWE_MAX_DEVICES] = {};   // the attribute is per device, not per process
        size_t& configured = configured_dev[device_slot()];
        if (configured < smem) {
            MWE_CHECK_CUDA(cudaFuncSetAttribute(assign_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        if (ev0) MWE_CHECK_CUDA(cudaEventRecord(ev0, stream));
        assign_tc_kernel<2><<<(unsigned)grid, TC_THREADS, smem, stream>>>(q);
    } else {
        static size_t configured_dev[MWE_MAX_DEVICES] = {};   // the attribute is per device, not per process
        size_t& configured = configured_dev[device_slot()];
        if (configured < smem) {
            MWE_CHECK_CUDA(cudaFuncSetAttribute(assign_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        if (ev0) MWE_CHECK_CUDA(cudaEventRecord(ev0, stream));
        assign_tc_kernel<1><<<(unsigned)grid, TC_THREADS, smem, stream>>>(q);
    }
    MWE_CHECK_LAUNCH();
    if (ev1) MWE_CHECK_CUDA(cudaEventRecord(ev1, stream));
    return MWE_OK;
}

}  // namespace mwe

// tests only: device buffer [N][ncb * n_pad] that receives the fp32 scores of the next tcgen05 assignment calls
extern "C" int mwe_debug_set_tc_scores(float* buf) {
    mwe::g_dbg_scores = buf;
    return MWE_OK;
}
extern "C" int mwe_debug_set_tc_profile(unsigned long long* buf) {
    mwe::g_dbg_prof = buf;
    return MWE_OK;
}
extern "C" int mwe_debug_tc_columns(int32_t max_k) {
    const int n = mwe::tc_n_pad(max_k);
    return ((max_k + n - 1) / n) * n;
}
