#include <cuda_runtime.h>
__global__ void k16n8k4(double* out, const double* a, const double* b) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    double a0 = a[threadIdx.x], a1 = a[threadIdx.x + 32], b0 = b[threadIdx.x];
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(b0));
    out[threadIdx.x] = c0 + c1 + c2 + c3;
}
__global__ void k16n8k8(double* out, const double* a, const double* b) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    double a0 = a[threadIdx.x], a1 = a[threadIdx.x + 32], a2 = a[threadIdx.x + 64], a3 = a[threadIdx.x + 96];
    double b0 = b[threadIdx.x], b1 = b[threadIdx.x + 32];
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    out[threadIdx.x] = c0 + c1 + c2 + c3;
}
__global__ void k16n8k16(double* out, const double* a, const double* b) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    double av[8], bv[4];
    for (int i = 0; i < 8; ++i) av[i] = a[threadIdx.x + 32 * i];
    for (int i = 0; i < 4; ++i) bv[i] = b[threadIdx.x + 32 * i];
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
                 : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(av[4]), "d"(av[5]), "d"(av[6]), "d"(av[7]), "d"(bv[0]), "d"(bv[1]), "d"(bv[2]), "d"(bv[3]));
    out[threadIdx.x] = c0 + c1 + c2 + c3;
}
__global__ void k8n8k4(double* out, const double* a, const double* b) {
    double c0 = 0, c1 = 0;
    double a0 = a[threadIdx.x], b0 = b[threadIdx.x];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a0), "d"(b0));
    out[threadIdx.x] = c0 + c1;
}
