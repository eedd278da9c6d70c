// Does an occasional DFMA (or F2F.F32.F64) disturb a DMMA stream?  G DMMAs then F extra fp64-pipe instructions.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int F, int KIND>
__global__ void k(double* out, int iters) {
    double c[6][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i][0] = c[i][1] = 0.0;
    double x[4] = {1.0, 2.0, 3.0, 4.0};
    float y[4] = {0, 0, 0, 0};
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 6; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[f % 4]) : "d"(a), "d"(b));
            else asm volatile("cvt.rm.f32.f64 %0, %1;" : "=f"(y[f % 4]) : "d"(x[f % 4]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + x[0] + x[1] + x[2] + x[3] + y[0] + y[1] + y[2] + y[3];
}
template <int F, int KIND>
void run(int warps_per_sm, int sms, double* out) {
    const int iters = 1000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<F, KIND><<<sms, warps_per_sm * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double flops = (double)sms * warps_per_sm * iters * 24.0 * 512.0;
    printf("%s per 24 DMMA: %2d  warps/SM %2d : %6.2f DMMA-TFLOP/s\n", KIND ? "F2F " : "DFMA", F, warps_per_sm, flops / best * 1e-9);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int w : {8, 16}) {
        run<0, 0>(w, sms, out); run<1, 0>(w, sms, out); run<2, 0>(w, sms, out); run<4, 0>(w, sms, out); run<8, 0>(w, sms, out);
        run<2, 1>(w, sms, out); run<8, 1>(w, sms, out);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
