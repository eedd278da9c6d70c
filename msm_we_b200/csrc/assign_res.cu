// K1, fp64 path, resident-centre variant: the kernel the small-K*D regime (BASELINE cfg2: K=20, D=64) runs.
//
// reference: StratifiedClusters.predict (msm_we/stratified_clustering.py:101-212) -> sklearn
// _k_means_lloyd.pyx:168-218.  Same arithmetic as assign.cu (score_j = ||c_j||^2 - 2 x.c_j, x.c from a
// k-sequential DMMA chain, fp32 round-down candidate filter + exact re-check), so labels are identical.
//
// Structure: one PRODUCER warp and NW CONSUMER warps per CTA, one CTA per SM, CTAs own contiguous tile ranges.
//   * The producer turns the CTA's tiles into GROUPS of 16 points of one WE bin.  For each group it waits for a
//     free row buffer, writes the group's metadata (point indices, bin's centre offset / count) next to it and
//     gathers the 16 point rows with cp.async.bulk -- one instruction copies a whole D*8-byte row, 32 lanes issue
//     32 rows at once, completion is byte-counted on the buffer's mbarrier.  No per-segment address arithmetic,
//     no shuffles: the copy path costs the SM ~2 instructions per row.
//   * Consumers take groups from a shared ticket counter (any warp, any group: no static imbalance), wait on the
//     group's mbarrier, run the [16 x K_pad x D] product on the DMMA pipe straight out of shared memory, release
//     the buffer, and finish the argmin in registers.  Nothing but LDS + DMMA + the fold is on their path, which
//     matters because one warp can issue a DMMA only every ~32 cycles while the pipe takes one per ~15.5 cycles
//     per scheduler (tools/dmma_chain_bench.cu): the pipe stays fed only if >= 2 warps per scheduler are inside
//     their DMMA phase at any time.
//   * All centres of the current WE bin live in shared memory (double-buffered across bin changes, loaded by
//     the producer with bulk copies), so centre rows are read from L2 once per (CTA, bin).
// Measured background (tools/ldgsts_gather_bw.cu): random 512-byte row gathers reach 6.3-6.9 TB/s once >= 64 KB
// per SM are in flight; the ring here holds up to ~200 KB.
#include <stdlib.h>

#include "assign_common.cuh"

namespace mwe {

static constexpr size_t AR_SMEM_MAX = 226 * 1024;
static constexpr int AR_TP = 256;              // points per tile record (16 groups)
static constexpr int AR_GROUP = 16;            // points per group = one consumer warp's m16 slice
static constexpr int AR_MAX_BUFS = 32;
static constexpr int AR_NP = 4;                // producer warps (groups are dealt round-robin to them)
static constexpr int AR_META_INTS = 21;        // idx[16], coff, kb, cbuf, nrows, tag (group number being filled)

// tuning hook (mwe_debug_set_k1_profile): device uint64[8] accumulating clock cycles per phase, summed over warps
//   consumers: [0] ticket -> data ready, [1] metadata + accumulator seed, [2] k loop, [3] fold + epilogue, [4] groups
//   producers: [5] waiting for a free buffer, [6] claiming + copy issue, [7] bin changes
static unsigned long long* g_k1_profile = nullptr;

struct ResShared {
    int32_t pidx[AR_NP][AR_TP];    // per producer warp: point indices of the tile being issued (16-byte aligned rows)
    uint64_t full[AR_MAX_BUFS];
    uint64_t empty[AR_MAX_BUFS];
    uint64_t cbar[2];
    int32_t meta[AR_MAX_BUFS][AR_META_INTS];
    int32_t ticket;
    int32_t total_groups;
};

//   NT : 8-column centre sub-tiles (K_pad = NT*8 >= every bin's centre count)
//   NW : consumer warps
// PROF: compile-time switch of the per-phase cycle counters (the product path runs the PROF = false instantiation)
template <int NT, int NW, bool PROF>
__global__ void __launch_bounds__((NW + AR_NP) * 32, 1) assign_dmma_resident_kernel(const AssignParams p, int nbufs, int xld, int nks, unsigned long long* prof_buf) {
    unsigned long long* const prof = PROF ? prof_buf : nullptr;
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int KP = NT * 8;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)p.D * 8u;
    double* sC = reinterpret_cast<double*>(smem_raw);            // [2][KP][xld]
    double* sQ = sC + (size_t)2 * KP * xld;                       // [2][KP]
    double* bufs = sQ + 2 * KP;                                   // [nbufs][16][xld]
    ResShared* sh = reinterpret_cast<ResShared*>(bufs + (size_t)nbufs * AR_GROUP * xld);

    const int32_t n_tiles = p.tile_prefix[p.nbins];
    const int first = (int)(((int64_t)n_tiles * blockIdx.x) / gridDim.x);
    const int count = (int)(((int64_t)n_tiles * (blockIdx.x + 1)) / gridDim.x) - first;
    const int4* __restrict__ desc = p.tile_desc + first;         // {pstart, pcount, coff, kb}

    // Columns [D, 8*ceil(D/8)) of every point and centre row are read by the k loop but never written by a copy:
    // zero them once.  Nothing else needs initialising: rows of a short group / centre rows past the bin's count
    // only feed accumulator rows / columns that are never read (rows and columns of the product are independent).
    if (p.D != xld - 8) {
        const int npad = xld - 8 - p.D;
        const int nrows_all = 2 * KP + nbufs * AR_GROUP;          // sC rows, then (after sQ) the buffer rows
        for (int e = threadIdx.x; e < nrows_all * npad; e += blockDim.x) {
            const int r = e / npad, c = p.D + (e - r * npad);
            double* row = (r < 2 * KP) ? sC + (size_t)r * xld : bufs + (size_t)(r - 2 * KP) * xld;
            row[c] = 0.0;
        }
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < nbufs; ++b) { mbar_init(&sh->full[b], 33); mbar_init(&sh->empty[b], 1); }
        mbar_init(&sh->cbar[0], 1);
        mbar_init(&sh->cbar[1], 1);
        sh->ticket = 0;
        for (int b = 0; b < nbufs; ++b) sh->meta[b][20] = -1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NW) {
        int tot = 0;
        for (int ti = lane; ti < count; ti += 32) tot += (desc[ti].y + AR_GROUP - 1) / AR_GROUP;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) sh->total_groups = tot;
    }
    __syncthreads();

    if (warp >= NW) {
        // ================================ producers ================================
        const int pw = warp - NW;
        const int segs = p.D >> 1;                       // 16-byte segments per row
        const int row_step = 32 / segs, seg_step = 32 % segs;
        const int row0 = lane / segs, seg0 = lane % segs;
        const uint32_t ldxb = (uint32_t)(p.ldx * 8);     // row stride in bytes (< 2^32, checked by the launcher)
        const char* xb = reinterpret_cast<const char*>(p.X);
        uint64_t xl64 = reinterpret_cast<uint64_t>(xb + lane * 16);   // this lane's segment when a row is exactly 32 segments
        uint32_t xld8 = (uint32_t)xld * 8u;
        asm volatile("" : "+l"(xl64), "+r"(xld8));       // keep both in registers (else re-derived from constants per row)
        // point indices of the tile being issued live in shared memory (group gi = entries [16 gi, 16 gi + 16)),
        // those of the next tile are in flight in registers
        int32_t* s_idx = sh->pidx[pw];
        auto fetch_idx = [&](const int4& d, int32_t* out) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j * 32 + lane;
                out[j] = (r < d.y) ? p.perm[d.x + r] : -1;
            }
        };
        int4 d_cur = count > 0 ? desc[0] : make_int4(0, 0, 0, 0);
        int4 d_next = count > 1 ? desc[1] : make_int4(0, 0, 0, 0);
        int32_t idx_next[8];
        fetch_idx(d_cur, idx_next);
#pragma unroll
        for (int j = 0; j < 8; ++j) s_idx[j * 32 + lane] = idx_next[j];
        __syncwarp();
        fetch_idx(d_next, idx_next);
        int32_t q = 0;                 // next group number
        int32_t cur_coff = -1, cur_cbuf = 1, bins_seen = 0, groups_in_bin = 0;   // (the last two: centre-buffer reuse rule below)
        uint32_t cuse[2] = {0u, 0u};   // fills of each centre buffer so far
        for (int ti = 0; ti < count; ++ti) {
            if (d_cur.z != cur_coff) {
                // ---- bin change: producer 0 loads the bin's centres into the other centre buffer ----
                const int nb = 1 - cur_cbuf;
                asm volatile("bar.sync 1, %0;" ::"n"(AR_NP * 32) : "memory");   // both producers have published every group < q
                if (pw == 0) {
                    if (bins_seen >= 2 && groups_in_bin < nbufs) {
                        // groups of the bin before the previous one may still be in flight and read buffer nb: drain
                        if (lane < nbufs) {
                            const int32_t last = q - 1 - ((q - 1 - lane) % nbufs + nbufs) % nbufs;   // last group that used buffer `lane`
                            if (last >= 0) mbar_wait(&sh->empty[lane], (uint32_t)(last / nbufs) & 1u);
                        }
                        __syncwarp();
                    }
                    const int kb = d_cur.w < KP ? d_cur.w : KP;
                    double* cdst = sC + (size_t)nb * KP * xld;
                    if (lane == 0) mbar_expect_tx(&sh->cbar[nb], (uint32_t)kb * row_bytes);
                    __syncwarp();
                    for (int r = lane; r < KP; r += 32) {
                        if (r < kb) bulk_copy_g2s(cdst + (size_t)r * xld, p.centers + ((int64_t)d_cur.z + r) * p.D, row_bytes, &sh->cbar[nb]);
                        sQ[nb * KP + r] = (r < kb) ? -0.5 * p.csq[d_cur.z + r] : 0.0;     // accumulator seed: acc = x.c - ||c||^2/2
                    }
                    mbar_wait(&sh->cbar[nb], cuse[nb] & 1u);
                }
                ++cuse[nb];
                asm volatile("bar.sync 1, %0;" ::"n"(AR_NP * 32) : "memory");   // centres (and ||c||^2) are in place
                cur_cbuf = nb;
                cur_coff = d_cur.z;
                ++bins_seen;
                groups_in_bin = 0;
            }
            const int ngroups = (d_cur.y + AR_GROUP - 1) / AR_GROUP;
            // buffer and use count of group q + pw, then advanced by AR_NP groups per step (no divisions in the loop)
            // (group q + gi goes to producer (q + gi) % AR_NP: with nbufs a multiple of AR_NP a buffer is always
            // refilled by the same warp, in order)
            const int gi0 = ((pw - q) % AR_NP + AR_NP) % AR_NP;
            int b = (q + gi0) % nbufs;
            uint32_t use = (uint32_t)((q + gi0) / nbufs);
            for (int gi = gi0; gi < ngroups; gi += AR_NP) {
                const int32_t qq = q + gi;
                const long long t_p0 = prof ? clock64() : 0;
                mbar_wait(&sh->empty[b], (use & 1u) ^ 1u);
                const long long t_p1 = prof ? clock64() : 0;
                const int nrows = min(AR_GROUP, d_cur.y - gi * AR_GROUP);
                const int32_t* gidx = s_idx + gi * AR_GROUP;
                if (lane < AR_GROUP) sh->meta[b][lane] = gidx[lane];
                if (lane == 0) {
                    sh->meta[b][16] = d_cur.z;
                    sh->meta[b][17] = d_cur.w;
                    sh->meta[b][18] = cur_cbuf;
                    sh->meta[b][19] = nrows;
                    // the buffer now belongs to group qq: its consumer may start waiting on full[b] (see there)
                    *reinterpret_cast<volatile int32_t*>(&sh->meta[b][20]) = qq;
                }
                const uint32_t dst0 = smem_u32(bufs + (size_t)b * AR_GROUP * xld);
                if (segs == 32) {
                    // a row is one warp-wide copy: 32 lanes x 16 bytes.  Indices first (4 x LDS.128), then 16 copies
                    // whose addresses are one IMAD.WIDE each.
                    int32_t pis[AR_GROUP];
#pragma unroll
                    for (int v = 0; v < AR_GROUP / 4; ++v) {
                        const int4 w = reinterpret_cast<const int4*>(gidx)[v];
                        pis[4 * v] = w.x; pis[4 * v + 1] = w.y; pis[4 * v + 2] = w.z; pis[4 * v + 3] = w.w;
                    }
                    const uint32_t dst = dst0 + lane * 16;
                    if (nrows == AR_GROUP) {
#pragma unroll
                        for (int r = 0; r < AR_GROUP; ++r)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)r * xld8),
                                         "l"(xl64 + (uint64_t)(uint32_t)pis[r] * ldxb));
                    } else {
#pragma unroll
                        for (int r = 0; r < AR_GROUP; ++r)
                            if (r < nrows)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)r * xld8),
                                             "l"(xl64 + (uint64_t)(uint32_t)pis[r] * ldxb));
                    }
                } else {
                    // lane copies the 16-byte segments e = lane, lane+32, ... of the group's [nrows x D] block
                    int row = row0, sg = seg0;
                    while (row < nrows) {
                        const int32_t pi = gidx[row];
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)((row * xld + sg * 2) * 8)),
                                     "l"(xb + (uint64_t)(uint32_t)pi * ldxb + sg * 16) : "memory");
                        row += row_step;
                        sg += seg_step;
                        if (sg >= segs) { sg -= segs; ++row; }
                    }
                }
                cp_async_arrive_noinc(&sh->full[b]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh->full[b]);   // releases the metadata written above
                if (prof && lane == 0) {
                    atomicAdd(prof + 5, (unsigned long long)(t_p1 - t_p0));
                    atomicAdd(prof + 6, (unsigned long long)(clock64() - t_p1));
                }
                b += AR_NP;
                if (b >= nbufs) { b -= nbufs; ++use; }
            }
            q += ngroups;
            groups_in_bin += ngroups;
            d_cur = d_next;
            __syncwarp();              // every lane is done reading this tile's indices
#pragma unroll
            for (int j = 0; j < 8; ++j) s_idx[j * 32 + lane] = idx_next[j];
            __syncwarp();
            d_next = (ti + 2 < count) ? desc[ti + 2] : make_int4(0, 0, 0, 0);
            fetch_idx(d_next, idx_next);
        }
    } else {
        // ================================ consumers ================================
        const int g = lane >> 2, t = lane & 3;
        const float finf = __int_as_float(0x7f800000);
        const int32_t total = sh->total_groups;
        const uint32_t nbufs_magic = (uint32_t)((((uint64_t)1 << 32) + (uint32_t)nbufs - 1) / (uint32_t)nbufs);
        int32_t q_next = 0;
        long long c_wait = 0, c_meta = 0, c_mma = 0, c_fold = 0, c_groups = 0, t_a = prof ? clock64() : 0;
        while (true) {
            // the ticket is taken only now, when the warp is free: a ticket reserved earlier would pin this warp to a
            // group whose data may land long after groups that idle warps could have taken
            if (lane == 0) q_next = atomicAdd(&sh->ticket, 1);
            const int32_t q = __shfl_sync(0xffffffffu, q_next, 0);
            if (q >= total) break;
            const uint32_t use = __umulhi((uint32_t)q, nbufs_magic);     // q / nbufs (exact for q < 2^32 / nbufs)
            const int b = q - (int32_t)use * nbufs;
            // Tickets can run ahead of the fills by more than nbufs groups (other consumers keep finishing groups
            // while one waits), and an mbarrier parity cannot tell phase u from phase u-2.  So first wait until
            // the producer has claimed the buffer for THIS group (then full[b] is in phase u or u+1), then wait
            // for the data.
            {
                const volatile int32_t* tag = &sh->meta[b][20];
                int spins = 0;
                while (*tag != q) {
                    __nanosleep(20);       // do not take issue slots from the producer warps while waiting
                    if (++spins > AS_SPIN_LIMIT) __trap();
                }
            }
            mbar_wait(&sh->full[b], use & 1u);
            if (prof) { const long long t_b = clock64(); c_wait += t_b - t_a; t_a = t_b; }
            const int32_t* meta = sh->meta[b];
            const int32_t coff = meta[16], kb = meta[17], cbuf = meta[18], nrows = meta[19];
            int32_t out_pt[2];
            out_pt[0] = (g < nrows) ? meta[g] : -1;
            out_pt[1] = (g + 8 < nrows) ? meta[g + 8] : -1;
            // Fragment addressing: lane (g, t) reads the 16-byte pair of columns {8p + 2t, 8p + 2t + 1} of its rows
            // with ONE LDS.128 and feeds the first to k-step 2p and the second to k-step 2p + 1 -- the dot product
            // is summed in a permuted column order (the same permutation for points and centres), which the
            // candidate pass may do (see the note on evaluation order below); row stride = 8 mod 16 doubles keeps
            // the quarter-warp phases of LDS.128 conflict-free.
            const double* xa0 = bufs + ((size_t)b * AR_GROUP + g) * xld + 2 * t;
            const double* xa1 = xa0 + 8 * xld;
            const double* cb0 = sC + ((size_t)cbuf * KP + g) * xld + 2 * t;
            // Accumulators start at -||c_j||^2/2, so after the k loop acc = x.c_j - ||c_j||^2/2 = -score_j/2 and the
            // fold needs no fp64 arithmetic.  This is NOT the reference's evaluation order (dot product from 0 in
            // column order, then one fma with ||c||^2; here also columns 0,2,4,6 before 1,3,5,7 inside every block
            // of 8): any two orders differ by < D u cmax (cmax + 2||x||), a quarter of the tie band, and the
            // filter below only trusts gaps of two tie bands, so everything it accepts has the reference's argmin;
            // everything else is re-evaluated in the reference's order by assign_recheck_kernel.
            double acc[2][NT][2];
            float cmaxf = 0.f;
            {
                const double* sq = sQ + cbuf * KP;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double2 h = *reinterpret_cast<const double2*>(sq + nt * 8 + 2 * t);
                    acc[0][nt][0] = acc[1][nt][0] = h.x;
                    acc[0][nt][1] = acc[1][nt][1] = h.y;
                    cmaxf = fmaxf(cmaxf, fmaxf(__double2float_rd(h.x) * -2.0f, __double2float_rd(h.y) * -2.0f));   // >= ||c||^2
                }
            }
            uint32_t xhi[2] = {0u, 0u};
            if (prof) { const long long t_b = clock64(); c_meta += t_b - t_a; t_a = t_b; }
#pragma unroll 2
            for (int kp = 0; kp < (nks >> 1); ++kp) {
                const double2 a0 = *reinterpret_cast<const double2*>(xa0 + kp * 8);
                const double2 a1 = *reinterpret_cast<const double2*>(xa1 + kp * 8);
                // bound of |x_k| (feeds the tie tolerance)
                xhi[0] = max(xhi[0], max((uint32_t)__double2hiint(a0.x) & 0x7fffffffu, (uint32_t)__double2hiint(a0.y) & 0x7fffffffu));
                xhi[1] = max(xhi[1], max((uint32_t)__double2hiint(a1.x) & 0x7fffffffu, (uint32_t)__double2hiint(a1.y) & 0x7fffffffu));
                double2 bv[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bv[nt] = *reinterpret_cast<const double2*>(cb0 + (size_t)nt * 8 * xld + kp * 8);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    dmma8x8x4(acc[0][nt][0], acc[0][nt][1], a0.x, bv[nt].x);
                    dmma8x8x4(acc[1][nt][0], acc[1][nt][1], a1.x, bv[nt].x);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    dmma8x8x4(acc[0][nt][0], acc[0][nt][1], a0.y, bv[nt].y);
                    dmma8x8x4(acc[1][nt][0], acc[1][nt][1], a1.y, bv[nt].y);
                }
            }

            // rows and metadata are in registers and the centre buffers were read for the last time: the buffer goes
            // back ("groups in flight <= buffers" therefore also bounds who can still be reading a centre buffer,
            // which the producer relies on when it reuses one)
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->empty[b]);
            if (prof) { const long long t_b = clock64(); c_mma += t_b - t_a; t_a = t_b; }

            // fp32 candidates: sf = round-down(score) = -2 * round-up(acc); smallest, its column, second smallest
            float m1f[2] = {finf, finf}, m2f[2] = {finf, finf};
            int32_t besti[2] = {0, 0};
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = nt * 8 + 2 * t + j;
                    if (c < kb) {
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const float sf = -2.0f * __double2float_ru(acc[mt][nt][j]);
                            const bool lt = sf < m1f[mt];          // strict: the first of equal roundings stays
                            m2f[mt] = lt ? m1f[mt] : fminf(m2f[mt], sf);
                            besti[mt] = lt ? c : besti[mt];
                            m1f[mt] = lt ? sf : m1f[mt];
                        }
                    }
                }
            }
            cmaxf = fmaxf(cmaxf, __shfl_xor_sync(0xffffffffu, cmaxf, 1));
            cmaxf = fmaxf(cmaxf, __shfl_xor_sync(0xffffffffu, cmaxf, 2));
            const float cmax = sqrtf(cmaxf) * 1.000001f;
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
                    p.label_out[pt] = (int64_t)coff + bi;
                    if (p.local_out) p.local_out[pt] = bi;
                    // ||x|| <= sqrt(D) max|x_k|, max|x_k| < the double whose high word is xh + 1 (inf/NaN -> NaN -> re-check)
                    const float xnorm = __double2float_ru(__hiloint2double((int)(min(xh, 0x7ff00000u) + 1u), 0)) * p.sqrt_d;
                    const float tolf = 2.0f * (float)p.tie_scale * cmax * (2.0f * xnorm + cmax);
                    // lower bound of the true gap, in fp32 with directed rounding (fp64 arithmetic here would queue
                    // behind the other warps' DMMAs on the fp64 pipe)
                    const float gap_lb = __fsub_rd(__fsub_rd(ru, bs), __fmul_ru(1.2e-7f, fabsf(bs)));
                    if (!(gap_lb > tolf)) p.recheck_list[atomicAdd(p.recheck_count, 1)] = pt;
                }
            }
            if (prof) { const long long t_b = clock64(); c_fold += t_b - t_a; t_a = t_b; ++c_groups; }
        }
        if (prof && lane == 0) {
            atomicAdd(prof + 0, (unsigned long long)c_wait);
            atomicAdd(prof + 1, (unsigned long long)c_meta);
            atomicAdd(prof + 2, (unsigned long long)c_mma);
            atomicAdd(prof + 3, (unsigned long long)c_fold);
            atomicAdd(prof + 4, (unsigned long long)c_groups);
        }
    }
}

struct ResidentPlan {
    int nt, nw, xld, nbufs;
    size_t smem;
};

static bool resident_plan(int D, int32_t max_k, bool vec2, ResidentPlan* pl) {
    if (const char* e = getenv("MWE_ASSIGN_RESIDENT"))
        if (atoi(e) == 0) return false;   // tuning knob: force the streaming kernel
    if (!vec2 || max_k > 64) return false;   // 16-byte copies need aligned rows; K_pad <= 64 accumulator columns
    pl->nt = max_k <= 16 ? 2 : max_k <= 24 ? 3 : max_k <= 32 ? 4 : max_k <= 48 ? 6 : 8;
    pl->nw = pl->nt <= 3 ? 16 : pl->nt == 4 ? 14 : 12;
    if (const char* e = getenv("MWE_ASSIGN_NW")) { if (pl->nt == 3 && (atoi(e) == 18 || atoi(e) == 20)) pl->nw = atoi(e); }   // tuning knob
    pl->xld = ((D + 7) / 8) * 8 + 8;      // 8 mod 16 doubles: conflict-free LDS.128 fragment loads
    const size_t kp = (size_t)pl->nt * 8;
    const size_t fixed = (2 * kp * pl->xld + 2 * kp) * sizeof(double) + sizeof(ResShared) + 128;
    const size_t buf_bytes = (size_t)AR_GROUP * pl->xld * sizeof(double);
    if (fixed + (size_t)pl->nw * buf_bytes > AR_SMEM_MAX) return false;
    int nbufs = (int)((AR_SMEM_MAX - fixed) / buf_bytes);
    if (const char* e = getenv("MWE_ASSIGN_STAGES")) nbufs = atoi(e) < nbufs ? atoi(e) : nbufs;   // tuning knob
    if (nbufs > AR_MAX_BUFS) nbufs = AR_MAX_BUFS;
    // a consumer may hold a ticket for a group that is up to NW groups ahead of the oldest unfilled one: with
    // fewer buffers than consumers it could not tell its buffer's phase from the one two fills earlier
    if (nbufs < pl->nw) return false;
    nbufs -= nbufs % AR_NP;              // a buffer is always filled by the same producer warp
    pl->nbufs = nbufs;
    pl->smem = fixed + (size_t)nbufs * buf_bytes;
    return true;
}

template <int NT, int NW>
static int launch_resident(const AssignParams& p, const ResidentPlan& pl, int64_t max_tiles, cudaStream_t stream) {
    static size_t configured_dev[MWE_MAX_DEVICES] = {};   // the attribute is per device, not per process
        size_t& configured = configured_dev[device_slot()];
    if (configured < pl.smem) {
        MWE_CHECK_CUDA(cudaFuncSetAttribute(assign_dmma_resident_kernel<NT, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        MWE_CHECK_CUDA(cudaFuncSetAttribute(assign_dmma_resident_kernel<NT, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        configured = pl.smem;
    }
    int64_t grid = sm_count();
    if (grid > max_tiles) grid = max_tiles;
    if (grid < 1) grid = 1;
    cudaEvent_t ev0, ev1;
    timing_events(&ev0, &ev1);
    if (ev0) MWE_CHECK_CUDA(cudaEventRecord(ev0, stream));
    if (g_k1_profile)
        MWE_CHECK_CUDA(launch_pdl(assign_dmma_resident_kernel<NT, NW, true>, dim3((unsigned)grid), dim3((NW + AR_NP) * 32), pl.smem, stream, p,
                                  pl.nbufs, pl.xld, (pl.xld - 8) / 4, g_k1_profile));
    else
        MWE_CHECK_CUDA(launch_pdl(assign_dmma_resident_kernel<NT, NW, false>, dim3((unsigned)grid), dim3((NW + AR_NP) * 32), pl.smem, stream, p,
                                  pl.nbufs, pl.xld, (pl.xld - 8) / 4, g_k1_profile));
    if (ev1) MWE_CHECK_CUDA(cudaEventRecord(ev1, stream));
    return MWE_OK;
}

// points per tile when the shape is inside this kernel's envelope, 0 otherwise (caller uses the streaming kernel)
int assign_resident_tile_points(int D, int32_t max_k, bool vec2) {
    ResidentPlan pl;
    return resident_plan(D, max_k, vec2, &pl) ? AR_TP : 0;
}

}  // namespace mwe

extern "C" int mwe_debug_set_k1_profile(unsigned long long* buf) {
    mwe::g_k1_profile = buf;
    return MWE_OK;
}

namespace mwe {

int launch_assign_resident(AssignParams p, int32_t max_k, bool vec2, int64_t max_tiles, cudaStream_t stream) {
    ResidentPlan pl;
    if (!resident_plan(p.D, max_k, vec2, &pl)) return MWE_E_UNSUPPORTED;
    switch (pl.nt) {
        case 2: return launch_resident<2, 16>(p, pl, max_tiles, stream);
        case 3: return pl.nw == 20 ? launch_resident<3, 20>(p, pl, max_tiles, stream)
                     : pl.nw == 18 ? launch_resident<3, 18>(p, pl, max_tiles, stream)
                                   : launch_resident<3, 16>(p, pl, max_tiles, stream);
        case 4: return launch_resident<4, 14>(p, pl, max_tiles, stream);
        case 6: return launch_resident<6, 12>(p, pl, max_tiles, stream);
        default: return launch_resident<8, 12>(p, pl, max_tiles, stream);
    }
}

}  // namespace mwe
