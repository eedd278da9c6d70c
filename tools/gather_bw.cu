// HBM read bandwidth of a random row gather as a function of the contiguous granule (bytes per row piece).
// Answers: what can ANY kernel reach when it gathers 256-byte / 512-byte / 2-KB pieces of rows picked by a
// permutation (the access pattern of the bucketed K1 assignment)?   nvcc -arch=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
__global__ void gather(const double2* __restrict__ x, const int* __restrict__ perm, long long nrows, int vec_per_row,
                       long long stride_vec, double* out) {
    // one warp per row piece; lanes read consecutive 16-byte vectors
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    double acc = 0;
    for (long long r = w; r < nrows; r += nw) {
        const double2* src = x + (long long)perm[r] * stride_vec;
        for (int v = lane; v < vec_per_row; v += 32) { double2 t = src[v]; acc += t.x + t.y; }
    }
    if (acc == 123.456) out[0] = acc;
}
int main() {
    const size_t total_bytes = (size_t)3 << 30;  // 3 GiB working set
    double2* x; cudaMalloc(&x, total_bytes); cudaMemset(x, 0, total_bytes);
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int gran : {128, 256, 512, 1024, 2048, 4096, 24576}) {
        const long long nrows = total_bytes / gran;
        std::vector<int> perm(nrows);
        for (long long i = 0; i < nrows; ++i) perm[i] = (int)i;
        std::mt19937 rng(1); std::shuffle(perm.begin(), perm.end(), rng);
        int* dperm; cudaMalloc(&dperm, nrows * 4); cudaMemcpy(dperm, perm.data(), nrows * 4, cudaMemcpyHostToDevice);
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            gather<<<148 * 16, 256>>>(x, dperm, nrows, gran / 16, gran / 16, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("granule %6d B: %7.1f GB/s\n", gran, total_bytes / best * 1e-6);
        cudaFree(dperm);
    }
    // sequential reference
    {
        const long long nrows = total_bytes / 4096;
        std::vector<int> perm(nrows); for (long long i = 0; i < nrows; ++i) perm[i] = (int)i;
        int* dperm; cudaMalloc(&dperm, nrows * 4); cudaMemcpy(dperm, perm.data(), nrows * 4, cudaMemcpyHostToDevice);
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); gather<<<148 * 16, 256>>>(x, dperm, nrows, 256, 256, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("sequential      : %7.1f GB/s\n", total_bytes / best * 1e-6);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
