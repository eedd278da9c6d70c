// K1: stratified nearest-centre assignment -- bucketing, dispatch, the streaming fp64 kernel and the exact re-check.
//
// reference: StratifiedClusters.predict (msm_we/stratified_clustering.py:101-212) whose inner
// call is MiniBatchKMeans.predict([coord]) -> sklearn/cluster/_k_means_lloyd.pyx:168-218:
//     score[j] = ||c_j||^2 - 2 x.c_j ;  label = first j with the smallest score (strict <).
//
// Design.  Points of one WE bin only ever meet that bin's K_b centres, so the points are first
// bucketed by bin (count -> scan -> scatter of indices, order inside a bucket is irrelevant
// because every point's label is computed independently); then one of three main kernels labels the
// buckets (mwe_assign_stratified_f64 picks):
//   * assign_res.cu  fp64, all centres of the bin resident in shared memory, producer/consumer warps --
//                    K <= 64 and 16-byte aligned rows (BASELINE cfg2);
//   * this file      fp64, centres streamed with the points: every shape (assign_dmma_kernel below);
//   * assign_tc.cu   tcgen05 split-TF32 candidate pass (precision path 1), large K*D.
// All three keep fp32 candidates (best, its column, runner-up) and hand every point whose gap does not clear
// the tie tolerance to assign_recheck_kernel, which evaluates the scores in the reference's order.
//
// assign_dmma_kernel: a tile is 64 (or 128) points of ONE bin; the x.c products of a tile are a
// [TP x K_b x D] GEMM on the fp64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA): warp w owns points
// [16w, 16w+16) x all centres of the block, accumulators stay in registers, the argmin epilogue is fused
// (quad shuffles), and no distance matrix ever reaches HBM.  Every warp copies its own point rows and a share of
// the centre block into an mbarrier-guarded shared-memory ring with cp.async (zero-filling tails), nstages-1
// steps ahead of the step it computes.  The grid is persistent (multiple of the SM count) and the ring runs
// across tile boundaries, so short-D tiles (D=64 is two chunks) do not drain the pipeline.
//
// Algorithmic traffic per point: D*8 bytes of features (+4 B bucket index, +8 B label); centres
// are re-read from L2.  FLOPs per point: 2*K_b*D.
#include <stdlib.h>

#include "assign_common.cuh"
#include "sort.cuh"

namespace mwe {

static constexpr int AS_DC = 32;                // doubles per k-chunk
static constexpr int AS_LD = AS_DC + 4;         // padded row: 72 words == 8 (mod 32): conflict-free LDS.64 fragments
static constexpr int AS_MAX_STAGES = 8;
static constexpr size_t AS_SMEM_BUDGET = 200 * 1024;        // one 8-warp CTA per SM
static constexpr size_t AS_SMEM_BUDGET_SMALL = 54 * 1024;   // four 4-warp CTAs per SM

// ---- bucketing -----------------------------------------------------------------------------

static constexpr int AS_SMEM_BINS = 2048;   // bins whose counters fit the shared-memory histogram
static constexpr int AS_BK_ITEMS = 2;       // points per thread in the bucketing kernels (latency-bound: keep the grid wide)

__global__ void __launch_bounds__(256)
    assign_count_kernel(const int32_t* __restrict__ bin, const uint8_t* __restrict__ flag, int64_t N, int32_t nbins,
                        int32_t* __restrict__ bin_count) {
    __shared__ int32_t s_cnt[AS_SMEM_BINS];
    const bool use_smem = nbins <= AS_SMEM_BINS;
    if (use_smem) {
        for (int b = threadIdx.x; b < nbins; b += 256) s_cnt[b] = 0;
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * (256 * AS_BK_ITEMS);
#pragma unroll
    for (int j = 0; j < AS_BK_ITEMS; ++j) {
        const int64_t i = base + j * 256 + threadIdx.x;
        if (i < N) {
            const uint8_t f = flag ? flag[i] : (uint8_t)0;
            const int32_t b = bin[i];
            if (!f && b >= 0 && b < nbins) {
                if (use_smem) atomicAdd(&s_cnt[b], 1);
                else atomicAdd(&bin_count[b], 1);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += 256)
            if (s_cnt[b]) atomicAdd(&bin_count[b], s_cnt[b]);
    }
}

// one CTA: bucket starts, per-bin tile prefix, cursors; flags bins that hold points but no centres
__global__ void __launch_bounds__(256)
    assign_scan_kernel(const int32_t* __restrict__ bin_count, const int64_t* __restrict__ bin_offset, int32_t nbins,
                       int32_t* __restrict__ bin_start, int32_t* __restrict__ bin_cursor,
                       int32_t* __restrict__ tile_prefix, int32_t* __restrict__ err_count, int tile_points,
                       int32_t* __restrict__ recheck_count) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ int scratch[9];
    __shared__ int s_carry[2];
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; *recheck_count = 0; }   // (saves a memset node between kernels)
    __syncthreads();
    for (int base = 0; base < nbins; base += 256) {
        const int b = base + threadIdx.x;
        int cnt = 0, tiles = 0;
        if (b < nbins) {
            cnt = bin_count[b];
            const int64_t kb = bin_offset[b + 1] - bin_offset[b];
            if (cnt > 0 && kb <= 0) {
                atomicAdd(&err_count[MWE_ERR_NO_CENTERS], cnt);
                tiles = 0;
            } else {
                tiles = (cnt + tile_points - 1) / tile_points;
            }
        }
        int tot_c, tot_t;
        const int ex_c = block_excl_scan_256(cnt, scratch, &tot_c);
        const int ex_t = block_excl_scan_256(tiles, scratch, &tot_t);
        const int c0 = s_carry[0], c1 = s_carry[1];
        if (b < nbins) {
            bin_start[b] = c0 + ex_c;
            bin_cursor[b] = c0 + ex_c;
            tile_prefix[b] = c1 + ex_t;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] = c0 + tot_c; s_carry[1] = c1 + tot_t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bin_start[nbins] = s_carry[0];
        tile_prefix[nbins] = s_carry[1];
    }
}

// Writes the bucket permutation and the labels of flagged / out-of-space points.  Ranks inside a CTA come
// from shared-memory atomics, one global atomic per (CTA, bin) reserves the CTA's slice of the bucket.
__global__ void __launch_bounds__(256)
    assign_scatter_kernel(const int32_t* __restrict__ bin, const uint8_t* __restrict__ flag, int64_t N, int32_t nbins,
                          const int64_t* __restrict__ bin_offset, int32_t* __restrict__ bin_cursor,
                          int32_t* __restrict__ perm, int64_t* __restrict__ label_out, int32_t* __restrict__ local_out,
                          const int32_t* __restrict__ bin_start, const int32_t* __restrict__ tile_prefix, int tile_points,
                          int4* __restrict__ tile_desc) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ int32_t s_cnt[AS_SMEM_BINS];
    const bool use_smem = nbins <= AS_SMEM_BINS;
    const int64_t T = bin_offset[nbins];
    if (use_smem) {
        for (int b = threadIdx.x; b < nbins; b += 256) s_cnt[b] = 0;
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * (256 * AS_BK_ITEMS);
    int32_t mybin[AS_BK_ITEMS];
    int32_t rank[AS_BK_ITEMS];
#pragma unroll
    for (int j = 0; j < AS_BK_ITEMS; ++j) {
        const int64_t i = base + j * 256 + threadIdx.x;
        mybin[j] = -1;
        rank[j] = 0;
        if (i < N) {
            const uint8_t f = flag ? flag[i] : (uint8_t)0;
            const int32_t b = bin[i];
            if (f) {
                // target is tested before basis (stratified_clustering.py:159-169)
                label_out[i] = (f & MWE_FLAG_TARGET) ? T + 1 : T;
                if (local_out) local_out[i] = -1;
            } else if (b < 0 || b >= nbins) {
                label_out[i] = -1;
                if (local_out) local_out[i] = -1;
            } else {
                mybin[j] = b;
                rank[j] = use_smem ? atomicAdd(&s_cnt[b], 1) : atomicAdd(&bin_cursor[b], 1);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += 256) {
            const int32_t c = s_cnt[b];
            s_cnt[b] = c ? atomicAdd(&bin_cursor[b], c) : 0;   // now the CTA's base inside the bucket
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < AS_BK_ITEMS; ++j) {
        if (mybin[j] >= 0) {
            const int64_t i = base + j * 256 + threadIdx.x;
            perm[(use_smem ? s_cnt[mybin[j]] : 0) + rank[j]] = (int32_t)i;
        }
    }
    if (tile_desc) {
        // tile records for the resident-centre kernel: the bin of a tile by binary search over the tile prefix
        const int32_t n_tiles = tile_prefix[nbins];
        for (int32_t tile = blockIdx.x * 256 + threadIdx.x; tile < n_tiles; tile += gridDim.x * 256) {
            int lo = 0, hi = nbins - 1;          // last bin with tile_prefix[bin] <= tile
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (tile_prefix[mid] <= tile) lo = mid; else hi = mid - 1;
            }
            const int32_t in_bin = (tile - tile_prefix[lo]) * tile_points;
            const int32_t left = (bin_start[lo + 1] - bin_start[lo]) - in_bin;
            const int64_t coff = bin_offset[lo];
            tile_desc[tile] = make_int4(bin_start[lo] + in_bin, left < tile_points ? left : tile_points, (int32_t)coff,
                                        (int32_t)(bin_offset[lo + 1] - coff));
        }
    }
}

// K1 main kernel.
//   NT  : 8-column centre sub-tiles per centre block (block = NT*8 centres)
//   VEC : 2 = 16-byte cp.async (LDGSTS; needs 16-byte aligned rows: even D and row stride), 1 = 8-byte
//         cp.async for odd D / unaligned views.  Both zero-fill past the end of a row, so the k-tail and
//         short tiles need no special casing in the math.
//   CW  : warps per CTA; a tile is CW*16 points.  Small centre blocks run 4-warp CTAs, several per SM, so
//         one CTA's latency stalls are covered by another; large blocks run one 8-warp CTA per SM.
// Every warp computes; each also copies its own 16 point rows and 16 of the centre rows, nstages-1 steps
// ahead of the step it computes (one step = one 32-column k-chunk of one centre block of one tile).
// (A cp.async.bulk / UBLKCP row copy was tried first: one instruction per 256-byte row, issued lane by
// lane through an ELECT loop, cost more issue slots than the vectorised LDGSTS below -- profiles/.)
template <int NT, int VEC, int CW>
__global__ void __launch_bounds__(CW * 32, (CW == 4 ? (NT <= 4 ? 4 : 2) : 1)) assign_dmma_kernel(const AssignParams p) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TP = CW * 16;
    constexpr int THREADS = CW * 32;
    constexpr int CROWS = NT * 8;
    constexpr int STAGE_DOUBLES = (TP + CROWS) * AS_LD + CROWS;   // rows + the block's ||c||^2
    constexpr int SEGS = AS_DC / VEC;          // copies per row chunk (16 or 32)
    constexpr int RPI = 32 / SEGS;             // rows covered by one warp-wide copy instruction (2 or 1)
    constexpr int XQ = 16 / RPI;               // copy instructions for the warp's 16 point rows
    constexpr int CPW = (CROWS + CW - 1) / CW; // centre rows copied per warp
    constexpr int CQ = (CPW + RPI - 1) / RPI;
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    __shared__ uint64_t full_bar[AS_MAX_STAGES];
    __shared__ uint64_t empty_bar[AS_MAX_STAGES];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nstages = p.nstages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full_bar[s], THREADS);   // every thread: cp.async.mbarrier.arrive.noinc
            mbar_init(&empty_bar[s], CW);       // one elected lane per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // bin tables -> shared memory (they are walked once per tile by every warp)
    TileTables tt{p.tile_prefix, p.bin_start, p.bin_offset, p.nbins};
    if (p.nbins <= AS_TABLE_BINS) {
        int32_t* s_tp = reinterpret_cast<int32_t*>(stage_base + (size_t)nstages * STAGE_DOUBLES);
        int32_t* s_bs = s_tp + (p.nbins + 1);
        int64_t* s_bo = reinterpret_cast<int64_t*>(s_bs + (p.nbins + 1) + ((2 * (p.nbins + 1)) & 1));
        for (int b = threadIdx.x; b <= p.nbins; b += THREADS) {
            s_tp[b] = p.tile_prefix[b];
            s_bs[b] = p.bin_start[b];
            s_bo[b] = p.bin_offset[b];
        }
        tt = TileTables{s_tp, s_bs, s_bo, p.nbins};
    }
    __syncthreads();

    const int32_t n_tiles = tt.tile_prefix[p.nbins];
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int64_t total_steps = (int64_t)my_tiles * p.ncb * p.nch;
    const int g = lane >> 2;  // fragment row / column group
    const int t = lane & 3;   // position inside the k4 step
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const float finf = __int_as_float(0x7f800000);

    TileWalk<TP> iw{0, 0, 0, 0, 0, 0, 0, 0};   // copies being issued
    iw.load(tt, my_tiles);
    TileWalk<TP> cw = iw;                     // tile being computed
    TileWalk<TP> pw = iw;                     // tile whose point indices are being prefetched (one ahead of iw)
    // copy geometry of this lane: column segment `seg`, row sub-index `rsub` inside each instruction
    const int seg = lane % SEGS;
    const int rsub = lane / SEGS;
    const int kcol0 = seg * VEC;
    const double* xsrc[XQ];   // source of the rows this lane copies for the tile of iw (nullptr = no such row)
    int32_t pidx_next[XQ];    // point indices for the following tile (in flight while iw's tile is being issued)
    auto fetch_pidx = [&](const TileWalk<TP>& w, int32_t* out) {
#pragma unroll
        for (int q = 0; q < XQ; ++q) {
            const int r = warp * 16 + q * RPI + rsub;
            out[q] = (r < w.pcount) ? p.perm[w.pstart + r] : -1;
        }
    };
    auto set_xsrc = [&]() {
#pragma unroll
        for (int q = 0; q < XQ; ++q)
            xsrc[q] = (pidx_next[q] >= 0) ? p.X + (int64_t)pidx_next[q] * p.ldx + kcol0 : nullptr;
    };
    fetch_pidx(iw, pidx_next);
    set_xsrc();
    pw.next_tile(tt, my_tiles);
    fetch_pidx(pw, pidx_next);
    int istage = 0;
    uint32_t iphase = 0;
    int64_t issued = 0;
    // rows outside the tile / centre block are simply not copied: rows and columns of the product are
    // independent, the stale shared memory they leave only reaches accumulators that are never read.
    // The k-tail of a row, however, is zero-filled (src-size < copy size), so the math needs no masks.
    auto issue_one = [&]() {
        mbar_wait(&empty_bar[istage], iphase ^ 1u);
        double* st = stage_base + (size_t)istage * STAGE_DOUBLES;
        double* sX = st + (warp * 16 + rsub) * AS_LD + kcol0;
        const int k0 = iw.kc * AS_DC;
        int vbytes = (p.D - k0 - kcol0) * 8;   // bytes of this lane's segment that exist in the row
        vbytes = vbytes < 0 ? 0 : (vbytes > VEC * 8 ? VEC * 8 : vbytes);
        const int crows = iw.kb - iw.cb * CROWS;   // centre rows of this block (<= 0 for a ragged trailing block)
#pragma unroll
        for (int q = 0; q < XQ; ++q)
            if (xsrc[q]) cp_async_zfill<VEC>(sX + q * RPI * AS_LD, xsrc[q] + k0, vbytes);
        {
            double* sC = st + (TP + warp * CPW + rsub) * AS_LD + kcol0;
            const double* cbase = p.centers + (iw.coff + iw.cb * CROWS + warp * CPW + rsub) * p.D + k0 + kcol0;
#pragma unroll
            for (int q = 0; q < CQ; ++q) {
                const int rr = q * RPI + rsub;          // row inside this warp's share
                if (rr < CPW && warp * CPW + rr < crows)
                    cp_async_zfill<VEC>(sC + q * RPI * AS_LD, cbase + (int64_t)q * RPI * p.D, vbytes);
            }
        }
        if (iw.kc == p.nch - 1) {
            // the block's ||c||^2 rides in the stage of the last k-chunk
            double* sQ = st + (TP + CROWS) * AS_LD;
            for (int c = threadIdx.x; c < crows && c < CROWS; c += THREADS)
                cp_async_zfill<1>(sQ + c, p.csq + iw.coff + iw.cb * CROWS + c, 8);
        }
        cp_async_arrive_noinc(&full_bar[istage]);
        if (++istage == nstages) { istage = 0; iphase ^= 1u; }
        ++issued;
        if (iw.advance(tt, p.ncb, p.nch, my_tiles)) {
            // entered the next tile: its indices were fetched a tile ago; start fetching the one after
            set_xsrc();
            pw.next_tile(tt, my_tiles);
            fetch_pidx(pw, pidx_next);
        }
    };

    // prologue: fill nstages-1 slots
    while (issued < total_steps && issued < nstages - 1) issue_one();

    int stage = 0;
    uint32_t phase = 0;
    // Candidate tracking is done on fp32 roundings (toward -inf) of the fp64 scores: smallest, its column,
    // and second smallest.  A point whose fp32 gap does not clear the tie tolerance plus one fp32 ulp is
    // handed to the exact fp64 re-check kernel, so the fp32 filter never decides a close call.
    float m1f[2] = {finf, finf};
    float m2f[2] = {finf, finf};
    int32_t besti[2] = {0, 0};
    uint32_t xhi[2] = {0u, 0u};       // largest |x_k| high word of rows g / g+8 (this thread's k positions)
    float cmaxf = 0.f;                // upper bound of the largest ||c||^2 among this thread's columns
    int32_t out_pt[2] = {-1, -1};     // point index of rows g / g+8 of the tile being computed (lanes t == 0)
    double acc[2][NT][2];

    for (int64_t step = 0; step < total_steps; ++step) {
        if (issued < total_steps) issue_one();
        if (cw.kc == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
            if (cw.cb == 0 && t == 0) {
                // label destinations, fetched now so the load latency hides under the tile's math
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int r = warp * 16 + mt * 8 + g;
                    out_pt[mt] = (r < cw.pcount) ? p.perm[cw.pstart + r] : -1;
                }
            }
        }
        mbar_wait(&full_bar[stage], phase);
        const double* st = stage_base + (size_t)stage * STAGE_DOUBLES;
        const double* xa0 = st + (warp * 16 + g) * AS_LD + t;
        const double* xa1 = xa0 + 8 * AS_LD;
        const double* cb0 = st + (TP + g) * AS_LD + t;
        if (cw.cb == 0) {
#pragma unroll
            for (int ks = 0; ks < AS_DC / 4; ++ks) {
                const double a0 = xa0[ks * 4];
                const double a1 = xa1[ks * 4];
                xhi[0] = max(xhi[0], (uint32_t)__double2hiint(a0) & 0x7fffffffu);   // bound of |x_k| (feeds the tie tolerance)
                xhi[1] = max(xhi[1], (uint32_t)__double2hiint(a1) & 0x7fffffffu);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double bv = cb0[nt * 8 * AS_LD + ks * 4];
                    dmma8x8x4(acc[0][nt][0], acc[0][nt][1], a0, bv);
                    dmma8x8x4(acc[1][nt][0], acc[1][nt][1], a1, bv);
                }
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < AS_DC / 4; ++ks) {
                const double a0 = xa0[ks * 4];
                const double a1 = xa1[ks * 4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double bv = cb0[nt * 8 * AS_LD + ks * 4];
                    dmma8x8x4(acc[0][nt][0], acc[0][nt][1], a0, bv);
                    dmma8x8x4(acc[1][nt][0], acc[1][nt][1], a1, bv);
                }
            }
        }
        if (cw.kc == p.nch - 1) {
            // fold this centre block into the running argmin: score = ||c||^2 - 2 x.c
            const double* sQ = st + (TP + CROWS) * AS_LD;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double2 cs2 = *reinterpret_cast<const double2*>(sQ + nt * 8 + 2 * t);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = cw.cb * CROWS + nt * 8 + 2 * t + j;
                    if (c < cw.kb) {
                        const double cs = j ? cs2.y : cs2.x;
                        cmaxf = fmaxf(cmaxf, __double2float_ru(cs));
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const float sf = __double2float_rd(fma(-2.0, acc[mt][nt][j], cs));
                            const bool lt = sf < m1f[mt];          // strict: the first of equal roundings stays
                            m2f[mt] = lt ? m1f[mt] : fminf(m2f[mt], sf);
                            besti[mt] = lt ? c : besti[mt];
                            m1f[mt] = lt ? sf : m1f[mt];
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == nstages) { stage = 0; phase ^= 1u; }

        if (cw.kc == p.nch - 1 && cw.cb == p.ncb - 1) {
            // every quad sees all columns of the tile: reduce over its 4 lanes
            cmaxf = fmaxf(cmaxf, __shfl_xor_sync(0xffffffffu, cmaxf, 1));
            cmaxf = fmaxf(cmaxf, __shfl_xor_sync(0xffffffffu, cmaxf, 2));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float bs = m1f[mt], ru = m2f[mt];
                int32_t bi = besti[mt];
                uint32_t xh = xhi[mt];
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
                    const float o2 = __shfl_xor_sync(0xffffffffu, ru, o);
                    const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    xh = max(xh, __shfl_xor_sync(0xffffffffu, xh, o));
                    // runner-up of the union = min(both runner-ups, the larger of the two bests)
                    ru = fminf(fminf(ru, o2), fmaxf(bs, os));
                    if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
                }
                if (t == 0 && out_pt[mt] >= 0) {
                    const int32_t pt = out_pt[mt];
                    p.label_out[pt] = cw.coff + bi;
                    if (p.local_out) p.local_out[pt] = bi;
                    // true gap > (ru - bs) - ulp32(bs); flag unless that clears the (over-estimated) tolerance
                    const float cmax = sqrtf(cmaxf) * 1.000001f;
                    // ||x|| <= sqrt(D) max|x_k|, max|x_k| < the double whose high word is xh + 1 (inf/NaN -> inf)
                    const float xnorm = __double2float_ru(__hiloint2double((int)(min(xh, 0x7ff00000u) + 1u), 0)) * p.sqrt_d;
                    const float tolf = 2.0f * (float)p.tie_scale * cmax * (2.0f * xnorm + cmax);
                    // lower bound of the true gap, in fp32 with directed rounding (fp64 arithmetic here would queue
                    // behind the other warps' DMMAs on the fp64 pipe)
                    const float gap_lb = __fsub_rd(__fsub_rd(ru, bs), __fmul_ru(1.2e-7f, fabsf(bs)));
                    if (!(gap_lb > tolf)) p.recheck_list[atomicAdd(p.recheck_count, 1)] = pt;
                }
                m1f[mt] = m2f[mt] = finf;
                besti[mt] = 0;
                xhi[mt] = 0u;
            }
            cmaxf = 0.f;
        }
        cw.advance(tt, p.ncb, p.nch, my_tiles);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// Near-tie re-check.  One warp per flagged point: fp64 scores of ALL centres of the point's bin (32 interleaved FMA
// chains over k, combined by a fixed butterfly: a deterministic order; any fp64 order is well inside the tie band), then
//     label = lowest j with score_j <= min_j score_j + tol,
//     tol   = TIE_C (D+8) 2^-53 cmax (2 ||x|| + cmax),  cmax = max_j ||c_j||,
// i.e. scores that differ by less than the rounding noise of their own evaluation count as tied and the
// reference's "first minimum wins" rule applies to the tie set.  Exact duplicate centres (common after
// MiniBatchKMeans' random reassignment) and ulp-level near-duplicates therefore resolve to the lowest index.
__global__ void __launch_bounds__(128)
    assign_recheck_kernel(const double* __restrict__ X, int64_t ldx, int D, const int32_t* __restrict__ bin,
                          const double* __restrict__ centers, const double* __restrict__ csq,
                          const int64_t* __restrict__ bin_offset, const int32_t* __restrict__ list,
                          const int32_t* __restrict__ count, double tie_scale, int64_t* __restrict__ label_out,
                          int32_t* __restrict__ local_out) {
    pdl_wait();
    pdl_launch_dependents();
    // One warp per listed point, 8 centres per pass.  Lanes stride the D dimension (coalesced reads of the centre rows;
    // the point row stays in L1) and keep one partial fp64 FMA chain per centre of the pass; the 32 x 8 partials are
    // folded by a transposing butterfly over lane bits 0-2 (4 + 2 + 1 exchanges) and a plain one over bits 3-4, after
    // which lane l holds the score of centre (l & 7) of the pass: 9 exchanges per 8 centres instead of 5 per centre.
    // Fixed order, independent of scheduling.  Four passes fill a group of 32 centres; lane j % 32 keeps s_j of group
    // j / 32 in sc[], so the tolerance pass needs no second evaluation for K_b <= 32 * RC_ROUNDS.
    constexpr int RC_ROUNDS = 8;
    const int n = *count;
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += warps_total) {
        const int32_t pt = list[e];
        const int32_t b = bin[pt];
        const int64_t coff = bin_offset[b];
        const int kb = (int)(bin_offset[b + 1] - coff);
        const double* x = X + (int64_t)pt * ldx;
        double xx = 0.0;
        for (int k = lane; k < D; k += 32) xx = fma(x[k], x[k], xx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, o);
        // scores of centres [j0, j0 + 8): lane l gets centre j0 + (l & 7); +inf past K_b
        auto pass_scores = [&](int j0) {
            double acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.0;
            const double* cbase = centers + (coff + j0) * D;
            const int nc = kb - j0;                         // >= 1; rows past K_b are not read
            for (int k = lane; k < D; k += 32) {
                const double xv = x[k];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < nc) acc[c] = fma(xv, cbase[(int64_t)c * D + k], acc[c]);
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {               // transposing fold over lane bits 2, 1, 0
                const bool upper = (lane & o) != 0;
#pragma unroll
                for (int c = 0; c < o; ++c) {
                    const double keep = upper ? acc[c + o] : acc[c];
                    const double give = upper ? acc[c] : acc[c + o];
                    acc[c] = keep + __shfl_xor_sync(0xffffffffu, give, o);
                }
            }
            double tot = acc[0];
            tot += __shfl_xor_sync(0xffffffffu, tot, 8);
            tot += __shfl_xor_sync(0xffffffffu, tot, 16);
            const int j = j0 + (lane & 7);
            return (j < kb) ? fma(-2.0, tot, csq[coff + j]) : inf;
        };
        // scores of centres [g0, g0 + 32) on lane (j - g0)
        auto group_scores = [&](int g0) {
            double s = inf;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (g0 + 8 * q < kb) {
                    const double v = pass_scores(g0 + 8 * q);
                    if ((lane >> 3) == q) s = v;
                }
            }
            return s;
        };
        double sc[RC_ROUNDS];
#pragma unroll
        for (int r = 0; r < RC_ROUNDS; ++r) sc[r] = inf;
        double smin = inf, cm = 0.0;
#pragma unroll
        for (int r = 0; r < RC_ROUNDS; ++r) {
            if (r * 32 < kb) {
                sc[r] = group_scores(r * 32);
                smin = fmin(smin, sc[r]);
                if (r * 32 + lane < kb) cm = fmax(cm, csq[coff + r * 32 + lane]);
            }
        }
        for (int g0 = 32 * RC_ROUNDS; g0 < kb; g0 += 32) {                          // very wide bins
            smin = fmin(smin, group_scores(g0));
            if (g0 + lane < kb) cm = fmax(cm, csq[coff + g0 + lane]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            smin = fmin(smin, __shfl_xor_sync(0xffffffffu, smin, o));
            cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, o));
        }
        const double cmax = sqrt(cm);
        const double tol = tie_scale * cmax * (2.0 * sqrt(xx) + cmax);
        int best = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < RC_ROUNDS; ++r) {
            if (best != 0x7fffffff || r * 32 >= kb) break;
            const unsigned m = __ballot_sync(0xffffffffu, sc[r] <= smin + tol);      // (+inf on lanes past K_b: never set)
            if (m) best = r * 32 + (__ffs(m) - 1);
        }
        for (int g0 = 32 * RC_ROUNDS; g0 < kb && best == 0x7fffffff; g0 += 32) {     // very wide bins: evaluate again
            const unsigned m = __ballot_sync(0xffffffffu, group_scores(g0) <= smin + tol);
            if (m) best = g0 + (__ffs(m) - 1);
        }
        if (best == 0x7fffffff) best = 0;  // all scores NaN: the reference's scan keeps index 0
        if (lane == 0) {
            label_out[pt] = coff + best;
            if (local_out) local_out[pt] = best;
        }
    }
}

__global__ void __launch_bounds__(256) centers_sqnorm_kernel(const double* __restrict__ centers, int64_t sumK, int D,
                                                            double* __restrict__ csq) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= sumK) return;
    const double* c = centers + row * D;
    double s = 0.0;
    for (int k = lane_id(); k < D; k += 32) s = fma(c[k], c[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane_id() == 0) csq[row] = s;
}


struct AssignWs {
    int4* tile_desc;
    int32_t* recheck_list;
    int32_t* recheck_count;
    int32_t* perm;
    int32_t* bin_count;
    int32_t* bin_cursor;
    int32_t* bin_start;
    int32_t* tile_prefix;
};

static size_t assign_ws_bytes(int64_t N, int32_t nbins) {
    size_t b = 0;
    b += 2 * align_up((size_t)(N > 0 ? N : 1) * sizeof(int32_t), 256);
    b += 5 * align_up((size_t)(nbins + 1) * sizeof(int32_t), 256);
    b += align_up(((size_t)(N > 0 ? N : 1) / 128 + (size_t)nbins + 2) * sizeof(int4), 256);   // tile records
    return b + 1024;
}

template <int NT, int VEC, int CW>
static int launch_assign(AssignParams p, int64_t max_tiles, cudaStream_t stream) {
    constexpr int TP = CW * 16;
    constexpr size_t stage_bytes = ((size_t)(TP + NT * 8) * AS_LD + NT * 8) * sizeof(double);
    const size_t table_bytes = (p.nbins <= AS_TABLE_BINS) ? (size_t)(p.nbins + 2) * 16 + 16 : 0;
    const size_t budget = (CW == 4 ? AS_SMEM_BUDGET_SMALL : AS_SMEM_BUDGET) - table_bytes;
    int nstages = (int)(budget / stage_bytes);
    if (const char* e = getenv("MWE_ASSIGN_STAGES")) nstages = atoi(e);   // tuning knob
    if (nstages > AS_MAX_STAGES) nstages = AS_MAX_STAGES;
    if (nstages < 2) nstages = 2;
    const size_t smem = stage_bytes * nstages + table_bytes;
    static size_t configured_dev[MWE_MAX_DEVICES] = {};   // the attribute is per device, not per process
        size_t& configured = configured_dev[device_slot()];
    if (configured < smem) {
        MWE_CHECK_CUDA(cudaFuncSetAttribute(assign_dmma_kernel<NT, VEC, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    p.nstages = nstages;
    int occ = 1;
    MWE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, assign_dmma_kernel<NT, VEC, CW>, CW * 32, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sm_count() * occ;   // persistent: every resident CTA walks its share of the tiles
    if (grid > max_tiles) grid = max_tiles;
    if (grid < 1) grid = 1;
    cudaEvent_t ev0, ev1;
    timing_events(&ev0, &ev1);
    if (ev0) MWE_CHECK_CUDA(cudaEventRecord(ev0, stream));
    MWE_CHECK_CUDA(launch_pdl(assign_dmma_kernel<NT, VEC, CW>, dim3((unsigned)grid), dim3(CW * 32), smem, stream, p));
    if (ev1) MWE_CHECK_CUDA(cudaEventRecord(ev1, stream));
    return MWE_OK;
}

// centre-block widths that are instantiated, and the CTA shape used with each
static const int kNtList[] = {2, 3, 4, 7, 8, 13, 16};
static int pick_nt(int max_k) {
    for (int i = 0; i < 7; ++i)
        if (kNtList[i] * 8 >= max_k) return kNtList[i];
    return 16;
}
static bool small_cta(int nt) {
    if (const char* e = getenv("MWE_ASSIGN_CW")) return atoi(e) == 4;   // tuning knob
    return nt <= 4;
}
static int tile_points_for(int nt) { return small_cta(nt) ? 64 : 128; }

template <int VEC>
static int dispatch_nt(int nt, const AssignParams& p, int64_t max_tiles, cudaStream_t stream) {
    if (small_cta(nt)) {
        switch (nt) {
            case 2: return launch_assign<2, VEC, 4>(p, max_tiles, stream);
            case 3: return launch_assign<3, VEC, 4>(p, max_tiles, stream);
            case 4: return launch_assign<4, VEC, 4>(p, max_tiles, stream);
            case 7: return launch_assign<7, VEC, 4>(p, max_tiles, stream);
            default: break;
        }
    }
    switch (nt) {
        case 2: return launch_assign<2, VEC, 8>(p, max_tiles, stream);
        case 3: return launch_assign<3, VEC, 8>(p, max_tiles, stream);
        case 4: return launch_assign<4, VEC, 8>(p, max_tiles, stream);
        case 7: return launch_assign<7, VEC, 8>(p, max_tiles, stream);
        case 8: return launch_assign<8, VEC, 8>(p, max_tiles, stream);
        case 13: return launch_assign<13, VEC, 8>(p, max_tiles, stream);
        default: return launch_assign<16, VEC, 8>(p, max_tiles, stream);
    }
}

// MWE_ASSIGN_AUTO: the fp64 DMMA kernel needs 2 K D fp64 FLOPs per point (37 TFLOP/s measured), the tcgen05
// kernel streams at ~2.9 TB/s whatever K is (measured, tools/assign_bench.py): cross-over near K*D ~ 2000.
int resolve_assign_path(int precision_path, int D, int32_t max_k) {
    if (precision_path != MWE_ASSIGN_AUTO) return precision_path;
    return ((int64_t)max_k * D >= 2048) ? MWE_ASSIGN_TF32X3 : MWE_ASSIGN_FP64;
}

}  // namespace mwe

extern "C" size_t mwe_assign_workspace_bytes(int64_t N, int32_t nbins) { return mwe::assign_ws_bytes(N, nbins); }

extern "C" size_t mwe_assign_workspace_bytes_ex(int64_t N, int32_t nbins, int D, int32_t max_k, int precision_path) {
    precision_path = mwe::resolve_assign_path(precision_path, D, max_k);
    size_t b = mwe::assign_ws_bytes(N, nbins);
    if (precision_path == MWE_ASSIGN_TF32X3) b += mwe::assign_tc_prep_bytes(nbins, D, max_k) + 256;
    return b;
}

extern "C" int mwe_centers_sqnorm_f64(const double* centers, int64_t sumK, int D, double* csq, void* stream) {
    MWE_REQUIRE(sumK >= 0 && D >= 1, "centers_sqnorm: bad shape");
    if (sumK == 0) return MWE_OK;
    const int64_t blocks = (sumK + 7) / 8;
    mwe::centers_sqnorm_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(centers, sumK, D, csq);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

extern "C" int mwe_assign_stratified_f64(const double* X, int64_t N, int D, int64_t ldx, const int32_t* bin,
                                         const uint8_t* flag, const double* centers, const double* csq,
                                         const int64_t* bin_offset, int32_t nbins, int32_t max_k, int precision_path,
                                         const int32_t* bin_count_in, int64_t* label_out, int32_t* local_out,
                                         void* workspace, size_t workspace_bytes, int32_t* err_count, void* stream) {
    using namespace mwe;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MWE_REQUIRE(N >= 0 && N < ((int64_t)1 << 31), "assign: N must be < 2^31 per call");
    MWE_REQUIRE(D >= 1 && ldx >= D, "assign: bad D / ldx");
    MWE_REQUIRE(nbins >= 1 && max_k >= 1, "assign: bad nbins / max_k");
    MWE_REQUIRE(bin && centers && csq && bin_offset && label_out && err_count, "assign: null pointer");
    const bool reuse_buckets = (precision_path & MWE_ASSIGN_REUSE_BUCKETS) != 0;
    precision_path = mwe::resolve_assign_path(precision_path & ~MWE_ASSIGN_REUSE_BUCKETS, D, max_k);
    if (precision_path != MWE_ASSIGN_FP64 && precision_path != MWE_ASSIGN_TF32X3) {
        set_last_error("assign: unknown precision path %d", precision_path);
        return MWE_E_UNSUPPORTED;
    }
    if (N == 0) return MWE_OK;
    const bool use_tc = precision_path == MWE_ASSIGN_TF32X3;
    const size_t need = assign_ws_bytes(N, nbins) + (use_tc ? assign_tc_prep_bytes(nbins, D, max_k) : 0);
    if (workspace_bytes < need) {
        set_last_error("assign: workspace too small (%zu < %zu)", workspace_bytes, need);
        return MWE_E_WORKSPACE;
    }
    Carver cv(workspace, workspace_bytes);
    AssignWs ws;
    ws.perm = cv.take<int32_t>((size_t)N);
    ws.recheck_list = cv.take<int32_t>((size_t)N);
    ws.bin_count = cv.take<int32_t>((size_t)nbins + 2);   // [nbins + 1] doubles as the re-check counter
    ws.recheck_count = ws.bin_count + nbins + 1;
    ws.bin_cursor = cv.take<int32_t>((size_t)nbins + 1);
    ws.bin_start = cv.take<int32_t>((size_t)nbins + 1);
    ws.tile_prefix = cv.take<int32_t>((size_t)nbins + 1);
    ws.tile_desc = cv.take<int4>((size_t)N / 128 + (size_t)nbins + 2);

    const int nt = pick_nt(max_k);
    const bool vec2 = (D % 2 == 0) && (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(centers) & 15) == 0);
    const int res_points = (use_tc || (int64_t)nbins * max_k >= ((int64_t)1 << 31) || ldx >= ((int64_t)1 << 28)) ? 0 : assign_resident_tile_points(D, max_k, vec2);
    const bool use_res = res_points > 0;
    const int tile_points = use_tc ? 128 : use_res ? res_points : tile_points_for(nt);
    const int64_t blocks = (N + 256 * AS_BK_ITEMS - 1) / (256 * AS_BK_ITEMS);
    if (reuse_buckets) {
        // perm / bin_start / tile tables and the labels of basis, target and unfitted points are still in place
        MWE_CHECK_CUDA(cudaMemsetAsync(ws.recheck_count, 0, sizeof(int32_t), s));
    } else {
        if (!bin_count_in) {
            MWE_CHECK_CUDA(cudaMemsetAsync(ws.bin_count, 0, (size_t)(nbins + 1) * sizeof(int32_t), s));
            assign_count_kernel<<<(unsigned)blocks, 256, 0, s>>>(bin, flag, N, nbins, ws.bin_count);
        }
        MWE_CHECK_CUDA(launch_pdl(assign_scan_kernel, dim3(1), dim3(256), 0, s, bin_count_in ? bin_count_in : ws.bin_count, bin_offset,
                                  nbins, ws.bin_start, ws.bin_cursor, ws.tile_prefix, err_count, tile_points, ws.recheck_count));
        MWE_CHECK_CUDA(launch_pdl(assign_scatter_kernel, dim3((unsigned)blocks), dim3(256), 0, s, bin, flag, N, nbins, bin_offset,
                                  ws.bin_cursor, ws.perm, label_out, local_out, ws.bin_start, ws.tile_prefix, tile_points,
                                  use_res ? ws.tile_desc : nullptr));
    }

    AssignParams p;
    p.X = X; p.ldx = ldx; p.D = D; p.centers = centers; p.csq = csq; p.bin_offset = bin_offset; p.nbins = nbins;
    p.perm = ws.perm; p.bin_start = ws.bin_start; p.tile_prefix = ws.tile_prefix; p.tile_desc = ws.tile_desc;
    p.label_out = label_out; p.local_out = local_out;
    p.recheck_list = ws.recheck_list; p.recheck_count = ws.recheck_count;
    p.tie_scale = AS_TIE_C * (double)(D + 8) * 1.1102230246251565e-16;
    p.sqrt_d = (float)(sqrt((double)D) * 1.000001);
    p.ncb = (max_k + nt * 8 - 1) / (nt * 8);
    p.nch = (D + AS_DC - 1) / AS_DC;
    const int64_t max_tiles = (N + tile_points - 1) / tile_points + nbins;
    int rc;
    if (use_tc) {
        const size_t prep_bytes = assign_tc_prep_bytes(nbins, D, max_k);
        void* prep = cv.take<char>(prep_bytes);
        rc = launch_assign_tc(p, max_k, N, prep, prep_bytes, s);
    } else if (use_res) {
        rc = launch_assign_resident(p, max_k, vec2, max_tiles, s);
    } else {
        rc = vec2 ? dispatch_nt<2>(nt, p, max_tiles, s) : dispatch_nt<1>(nt, p, max_tiles, s);
    }
    if (rc != MWE_OK) return rc;
    MWE_CHECK_CUDA(launch_pdl(assign_recheck_kernel, dim3(sm_count() * 8), dim3(128), 0, s, X, ldx, D, bin, centers, csq, bin_offset,
                              ws.recheck_list, ws.recheck_count, p.tie_scale, label_out, local_out));
    return MWE_OK;
}
