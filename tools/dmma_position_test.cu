// Does DMMA (mma.sync m8n8k4 f64) round identically for identical operand columns/rows placed at
// different positions of the tile?  Prints the number of bitwise mismatches.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void k(const double* a, const double* bcol, const double* c_in, double* out, int iters) {
    // A: 8x4 from a (row-major), B: every column equals bcol[0..3]
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double c0 = c_in[0], c1 = c_in[0];
    for (int it = 0; it < iters; ++it) {
        double av = a[it * 32 + g * 4 + t];
        double bv = bcol[it * 4 + t];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
    }
    out[g * 8 + 2 * t] = c0; out[g * 8 + 2 * t + 1] = c1;
}
int main() {
    const int iters = 16;
    double ha[iters * 32], hb[iters * 4], hc[1] = {0.3}, ho[64];
    srand(7);
    int mism_cols = 0, mism_rows = 0, trials = 2000;
    double *da, *db, *dc, *dout;
    cudaMalloc(&da, sizeof(ha)); cudaMalloc(&db, sizeof(hb)); cudaMalloc(&dc, 8); cudaMalloc(&dout, 64 * 8);
    for (int tr = 0; tr < trials; ++tr) {
        for (int i = 0; i < iters * 4; ++i) hb[i] = (rand() / (double)RAND_MAX - 0.5) * 3.0;
        // all 8 rows of A identical too
        for (int it = 0; it < iters; ++it) for (int r = 0; r < 8; ++r) for (int q = 0; q < 4; ++q) ha[it * 32 + r * 4 + q] = (r == 0) ? (rand() / (double)RAND_MAX - 0.5) * 3.0 : ha[it * 32 + q];
        cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice); cudaMemcpy(dc, hc, 8, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(da, db, dc, dout, iters);
        cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
        bool bc = false, br = false;
        for (int r = 0; r < 8; ++r) for (int c = 1; c < 8; ++c) if (ho[r * 8 + c] != ho[r * 8]) bc = true;
        for (int r = 1; r < 8; ++r) if (ho[r * 8] != ho[0]) br = true;
        mism_cols += bc; mism_rows += br;
        if (tr == 0) { // compare with sequential fma chain
            double s = hc[0];
            for (int it = 0; it < iters; ++it) for (int q = 0; q < 4; ++q) s = fma(ha[it * 32 + q], hb[it * 4 + q], s);
            printf("dmma %.17g  seq-fma %.17g  diff %.3g\n", ho[0], s, ho[0] - s);
        }
    }
    printf("trials %d: column-position mismatches %d, row-position mismatches %d\n", trials, mism_cols, mism_rows);
    return 0;
}
