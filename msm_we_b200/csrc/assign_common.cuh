// Shared pieces of the K1 assignment kernels (assign.cu: fp64 DMMA path, assign_tc.cu: tcgen05 path).
#pragma once
#include "common.cuh"

namespace mwe {

static constexpr int AS_TABLE_BINS = 1024;     // bins whose tile tables are cached in shared memory
static constexpr int AS_SPIN_LIMIT = 1 << 22;     // a protocol bug traps after a few seconds instead of hanging the GPU
static constexpr double AS_TIE_C = 4.0;        // tie band = AS_TIE_C (D+8) 2^-53 cmax (2||x|| + cmax)

// ---- PTX helpers -----------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > AS_SPIN_LIMIT) __trap();  // never hang the GPU on a protocol bug
    }
}
// the same three with a precomputed 32-bit shared-memory address: the generic -> shared conversion of a pointer costs
// 4-5 instructions per call, which shows in loops that are bound by instruction issue (assign_tc.cu, conversion warps)
__device__ __forceinline__ void mbar_arrive(uint32_t bar32) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar32) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar32, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar32), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > AS_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar32) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar32) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int VEC>
__device__ __forceinline__ void cp_async_zfill(void* dst, const void* src, int src_bytes) {
    if (VEC == 2) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA bulk copy global -> shared, completion (in bytes) signalled on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct AssignParams {
    const double* X;
    int64_t ldx;
    int D;
    const double* centers;
    const double* csq;
    const int64_t* bin_offset;
    int32_t nbins;
    const int32_t* perm;
    const int32_t* bin_start;
    const int32_t* tile_prefix;
    const int4* tile_desc;   // per tile {first bucket slot, points, first centre row, centres} (resident kernel)
    int64_t* label_out;
    int32_t* local_out;
    int32_t* recheck_list;   // points whose best two scores are within rounding noise
    int32_t* recheck_count;
    double tie_scale;        // TIE_C * (D + 8) * 2^-53
    float sqrt_d;            // sqrt(D), rounded up
    int ncb;      // centre blocks of NT*8 per tile
    int nch;      // k-chunks of AS_DC per centre block
    int nstages;  // depth of the shared-memory ring
};

// Tile bookkeeping tables (shared memory copies when the bin count allows, else the global arrays).
struct TileTables {
    const int32_t* tile_prefix;   // [nbins+1]
    const int32_t* bin_start;     // [nbins+1]
    const int64_t* bin_offset;    // [nbins+1]
    int32_t nbins;
};

// Per-warp view of the tile sequence of this CTA.  Three walkers run at different distances: the tile
// whose point indices are being prefetched, the tile whose copies are being issued, the tile being
// computed.  Tiles handled by one CTA are increasing, so the bin is found by scanning forward.
template <int TP>
struct TileWalk {
    int ti, cb, kc;      // tile ordinal inside the CTA, centre block, k-chunk
    int32_t bin;
    int32_t pstart, pcount, kb;
    int64_t coff;
    __device__ __forceinline__ void load(const TileTables& tt, int my_tiles) {
        if (ti >= my_tiles) { pcount = 0; return; }
        const int32_t tile = (int32_t)blockIdx.x + ti * (int32_t)gridDim.x;
        while (bin + 1 < tt.nbins && tile >= tt.tile_prefix[bin + 1]) ++bin;
        const int32_t in_bin = (tile - tt.tile_prefix[bin]) * TP;
        pstart = tt.bin_start[bin] + in_bin;
        pcount = min(TP, (tt.bin_start[bin + 1] - tt.bin_start[bin]) - in_bin);
        coff = tt.bin_offset[bin];
        kb = (int32_t)(tt.bin_offset[bin + 1] - coff);
    }
    __device__ __forceinline__ void next_tile(const TileTables& tt, int my_tiles) {
        ++ti;
        load(tt, my_tiles);
    }
    // returns true when the walk entered a new tile
    __device__ __forceinline__ bool advance(const TileTables& tt, int ncb, int nch, int my_tiles) {
        if (++kc < nch) return false;
        kc = 0;
        if (++cb < ncb) return false;
        cb = 0;
        next_tile(tt, my_tiles);
        return true;
    }
};


// fp64 path with the bin's centres resident in shared memory (assign_res.cu)
int assign_resident_tile_points(int D, int32_t max_k, bool vec2);   // 0 = shape outside the kernel's envelope
int launch_assign_resident(AssignParams p, int32_t max_k, bool vec2, int64_t max_tiles, cudaStream_t stream);

// tcgen05 path (assign_tc.cu).  prep_bytes: workspace for the pre-split centres.
int resolve_assign_path(int precision_path, int D, int32_t max_k);
size_t assign_tc_prep_bytes(int32_t nbins, int D, int32_t max_k);
int launch_assign_tc(const AssignParams& p, int32_t max_k, int64_t N, void* prep, size_t prep_bytes, cudaStream_t stream);

}  // namespace mwe
