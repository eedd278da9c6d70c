// Does a DMMA hold the scheduler's issue port?  DMMA stream with F independent integer IMADs per DMMA:
// if time = max(DMMA, IMAD) the two overlap; if time = sum, non-DMMA instructions cost DMMA throughput 1:1.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int F, bool DM>
__global__ void k(double* out, int iters, int mul) {
    double c[6][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i][0] = c[i][1] = 0.0;
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (DM) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int f = 0; f < F; ++f) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[f % 8]) : "r"(mul), "r"(f));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1];
    int xs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) xs += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + xs;
}
template <int F, bool DM>
void run(int warps_per_sm, int sms, double* out) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<F, DM><<<sms, warps_per_sm * 32>>>(out, iters, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double n = (double)warps_per_sm / 4 * iters * 6;   // DMMA slots per SMSP
    printf("%s IMAD/slot %2d warps/SM %2d : %7.3f ms  = %5.1f ns per slot per SMSP\n", DM ? "DMMA+" : "     ", F, warps_per_sm, best, best * 1e6 / n);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int w : {8, 16}) {
        run<0, true>(w, sms, out); run<4, true>(w, sms, out); run<8, true>(w, sms, out); run<12, true>(w, sms, out); run<16, true>(w, sms, out);
        run<4, false>(w, sms, out); run<8, false>(w, sms, out); run<16, false>(w, sms, out);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
