// Measures the fp64 ceilings of the B200 this runs on: DMMA (mma.sync m8n8k4 f64) and DFMA issue
// rates with register operands only, so the assignment kernel's tensor roofline has a measured
// denominator (MEASURED_PEAKS.json has no fp64 figure).   nvcc -arch=sm_100a -O3 -o fp64_microbench
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dmma_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
        const int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 256 * 8 * (double)iters * (blocks * threads / 32);
            if (rep) printf("DMMA  warps/SM=%2d  %.2f TFLOP/s  (%.3f ms)\n", warps_per_sm, flops / ms * 1e-9, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 16 * (double)iters * (blocks * threads);
            if (rep) printf("DFMA  warps/SM=%2d  %.2f TFLOP/s  (%.3f ms)\n", warps_per_sm, flops / ms * 1e-9, ms);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
