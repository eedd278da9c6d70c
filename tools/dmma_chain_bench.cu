// DMMA issue/latency facts for the K1 inner loop: throughput as a function of the number of independent
// accumulator chains per warp (dependent DMMAs are `chains` instructions apart) and of where operands come from.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH, bool SMEM>
__global__ void k(double* out, int iters) {
    __shared__ double sa[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sa[i] = i * 1e-4;
    __syncthreads();
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (SMEM) { a = sa[g * 36 + ks * 4 + t]; b = sa[(8 + g) * 36 + ks * 4 + t + (it & 1)]; }
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH, bool SMEM>
void run(int warps_per_sm, int sms, double* out) {
    const int iters = 4000 / CH * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<CH, SMEM><<<sms, warps_per_sm * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double flops = (double)sms * warps_per_sm * iters * 8.0 * CH * 512.0;
    printf("chains %2d  %s operands  warps/SM %2d : %6.2f TFLOP/s\n", CH, SMEM ? "smem" : "reg ", warps_per_sm, flops / best * 1e-9);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int w : {8, 16, 32}) {
        run<1, false>(w, sms, out); run<2, false>(w, sms, out); run<4, false>(w, sms, out); run<6, false>(w, sms, out); run<12, false>(w, sms, out);
        run<2, true>(w, sms, out); run<6, true>(w, sms, out); run<12, true>(w, sms, out);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
