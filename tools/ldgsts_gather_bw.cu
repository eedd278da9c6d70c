// What bandwidth can the K1 access pattern reach at all?  Rows picked by a random permutation are copied into a
// shared-memory ring with cp.async (LDGSTS) + mbarrier, exactly like the assignment kernels, but nothing is
// computed.  Sweeps the contiguous piece per row (bytes), CTAs per SM and ring depth.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
    uint32_t d = 0;
    while (!d) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(d) : "r"(s32(b)), "r"(ph) : "memory");
}
// THREADS threads; tile = ROWS rows x PIECE bytes; every thread copies 16-byte segments
template <int THREADS>
__global__ void gather_ring(const char* __restrict__ x, const int* __restrict__ perm, long long n_tiles, int rows, int piece,
                            long long row_stride, int pieces_per_row, int nstages, unsigned long long* sink, long long nrows_total) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[8], empty[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) { mb_init(&full[s], THREADS); mb_init(&empty[s], THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int segs = piece / 16;                 // 16-byte segments per row piece
    const int per_tile = rows * segs;
    const int seg_shift = 31 - __clz(segs);
    const size_t stage_bytes = (size_t)rows * (piece + 16);
    long long issued = 0, total = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) total += pieces_per_row;
    int is = 0, cs = 0; uint32_t iph = 0, cph = 0;
    long long it_tile = blockIdx.x; int it_piece = 0;
    auto issue = [&]() {
        mb_wait(&empty[is], iph ^ 1u);
        unsigned char* st = smem + (size_t)is * stage_bytes;
        for (int e = threadIdx.x; e < per_tile; e += THREADS) {
            const int r = e >> seg_shift, sg = e & (segs - 1);
            const long long row = (long long)(((uint32_t)(it_tile * rows + r) * 2654435761u) & (uint32_t)(nrows_total - 1));   // nrows_total is a power of two
            const char* src = x + row * row_stride + (long long)it_piece * piece + sg * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(st + (size_t)r * (piece + 16) + sg * 16)), "l"(src) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(&full[is])) : "memory");
        if (++is == nstages) { is = 0; iph ^= 1u; }
        ++issued;
        if (++it_piece == pieces_per_row) { it_piece = 0; it_tile += gridDim.x; }
    };
    while (issued < total && issued < nstages - 1) issue();
    unsigned long long acc = 0;
    for (long long step = 0; step < total; ++step) {
        if (issued < total) issue();
        mb_wait(&full[cs], cph);
        acc += smem[(size_t)cs * stage_bytes + threadIdx.x];
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mb_arrive(&empty[cs]);
        if (++cs == nstages) { cs = 0; cph ^= 1u; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (acc == 0x123456789ull) *sink = acc;
}
// same ring, rows moved with cp.async.bulk (TMA bulk copy): one instruction per row piece, byte-counted completion
template <int THREADS>
__global__ void gather_ring_bulk(const char* __restrict__ x, const int* __restrict__ perm, long long n_tiles, int rows, int piece,
                                 long long row_stride, int pieces_per_row, int nstages, unsigned long long* sink, long long nrows_total) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[8], empty[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t stage_bytes = (size_t)rows * (piece + 16);
    long long issued = 0, total = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) total += pieces_per_row;
    int is = 0, cs = 0; uint32_t iph = 0, cph = 0;
    long long it_tile = blockIdx.x; int it_piece = 0;
    auto issue = [&]() {
        mb_wait(&empty[is], iph ^ 1u);
        unsigned char* st = smem + (size_t)is * stage_bytes;
        if (threadIdx.x == 0)
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s32(&full[is])), "r"((uint32_t)(rows * piece)) : "memory");
        __syncthreads();
        for (int r = threadIdx.x; r < rows; r += THREADS) {
            const long long row = (long long)(((uint32_t)(it_tile * rows + r) * 2654435761u) & (uint32_t)(nrows_total - 1));   // nrows_total is a power of two
            const char* src = x + row * row_stride + (long long)it_piece * piece;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(st + (size_t)r * (piece + 16))),
                         "l"(src), "r"((uint32_t)piece), "r"(s32(&full[is])) : "memory");
        }
        if (++is == nstages) { is = 0; iph ^= 1u; }
        ++issued;
        if (++it_piece == pieces_per_row) { it_piece = 0; it_tile += gridDim.x; }
    };
    while (issued < total && issued < nstages - 1) issue();
    unsigned long long acc = 0;
    for (long long step = 0; step < total; ++step) {
        if (issued < total) issue();
        mb_wait(&full[cs], cph);
        acc += smem[(size_t)cs * stage_bytes + threadIdx.x];
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mb_arrive(&empty[cs]);
        if (++cs == nstages) { cs = 0; cph ^= 1u; }
    }
    if (acc == 0x123456789ull) *sink = acc;
}
int main() {
    const long long row_bytes_list[] = {512, 2048};
    unsigned long long* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (long long row_bytes : row_bytes_list) {
        const size_t total_bytes = (size_t)2 << 30;
        const long long nrows = total_bytes / row_bytes;
        char* x; cudaMalloc(&x, total_bytes); cudaMemset(x, 1, total_bytes);
        std::vector<int> perm(nrows); for (long long i = 0; i < nrows; ++i) perm[i] = (int)i;
        std::mt19937 rng(3); std::shuffle(perm.begin(), perm.end(), rng);
        int* dperm; cudaMalloc(&dperm, nrows * 4); cudaMemcpy(dperm, perm.data(), nrows * 4, cudaMemcpyHostToDevice);
        for (int piece : {256, 512, 1024, 2048}) {
            if (piece > row_bytes) continue;
            for (int ctas : {1, 2, 4})
            for (int rows : {64, 128}) {
                for (int nst : {2, 3, 4}) {
                    const size_t smem = (size_t)nst * rows * (piece + 16);
                    if (smem * ctas > 200 * 1024) continue;
                    cudaFuncSetAttribute(gather_ring<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    const long long n_tiles = nrows / rows;
                    float best = 1e9;
                    for (int rep = 0; rep < 3; ++rep) {
                        cudaEventRecord(e0);
                        gather_ring<128><<<148 * ctas, 128, smem>>>(x, dperm, n_tiles, rows, piece, row_bytes, (int)(row_bytes / piece), nst, sink, nrows);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
                    }
                    cudaFuncSetAttribute(gather_ring_bulk<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    float bestb = 1e9;
                    for (int rep = 0; rep < 3; ++rep) {
                        cudaEventRecord(e0);
                        gather_ring_bulk<128><<<148 * ctas, 128, smem>>>(x, dperm, n_tiles, rows, piece, row_bytes, (int)(row_bytes / piece), nst, sink, nrows);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float ms; cudaEventElapsedTime(&ms, e0, e1); bestb = std::min(bestb, ms);
                    }
                    printf("row %5lld B piece %5d B rows/stage %3d CTAs/SM %d stages %d (%3zu KB/SM in flight): LDGSTS %7.1f GB/s   bulk %7.1f GB/s\n", row_bytes, piece, rows, ctas, nst,
                           (size_t)(nst - 1) * rows * piece * ctas / 1024, total_bytes / best * 1e-6, total_bytes / bestb * 1e-6);
                }
            }
        }
        cudaFree(x); cudaFree(dperm);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
