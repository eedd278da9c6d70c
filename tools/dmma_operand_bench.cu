// Does DMMA throughput depend on operand reuse?  The K1 inner loop: per k-step 2 A fragments and 3 B fragments from
// shared memory feed 6 DMMAs (each A used 3 times, each B twice) -- versus the same operands for every DMMA.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>   // 0: one (a, b) pair per k-step for all 6 DMMAs; 1: the kernel's 2 x 3 pattern; 2: 6 distinct pairs
__global__ void k(double* out, int iters) {
    __shared__ double sa[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sa[i] = i * 1e-4;
    __syncthreads();
    double c[6][2];
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i][0] = c[i][1] = 0.0;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int o = ks * 4 + t + (it & 1);
            if (MODE == 0) {
                const double a = sa[g * 36 + o], b = sa[(8 + g) * 36 + o];
#pragma unroll
                for (int i = 0; i < 6; ++i) dmma(c[i][0], c[i][1], a, b);
            } else if (MODE == 1) {
                const double a0 = sa[g * 36 + o], a1 = sa[(8 + g) * 36 + o];
                const double b0 = sa[(16 + g) * 36 + o], b1 = sa[(24 + g) * 36 + o], b2 = sa[(32 + g) * 36 + o];
                dmma(c[0][0], c[0][1], a0, b0); dmma(c[1][0], c[1][1], a1, b0);
                dmma(c[2][0], c[2][1], a0, b1); dmma(c[3][0], c[3][1], a1, b1);
                dmma(c[4][0], c[4][1], a0, b2); dmma(c[5][0], c[5][1], a1, b2);
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) dmma(c[i][0], c[i][1], sa[(i * 8 + g) * 36 + o], sa[((i + 2) * 8 % 56 + g) * 36 + o]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(int warps_per_sm, int sms, double* out) {
    const int iters = 1500;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<sms, warps_per_sm * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double flops = (double)sms * warps_per_sm * iters * 8.0 * 6 * 512.0;
    const char* names[] = {"same operands for all 6", "K1 pattern (2 A x 3 B)", "6 distinct operand pairs"};
    printf("%-26s warps/SM %2d : %6.2f TFLOP/s\n", names[MODE], warps_per_sm, flops / best * 1e-9);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int w : {8, 16}) { run<0>(w, sms, out); run<1>(w, sms, out); run<2>(w, sms, out); }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
