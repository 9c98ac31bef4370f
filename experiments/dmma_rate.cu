// Microbenchmark: DMMA.8x8x4 (mma.sync m8n8k4 f64) against DFMA on B200 — issue rate per SM with 12 single-warp CTAs per
// SM, and the rounding of a k = 4 accumulation against an FMA chain.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_rate experiments/dmma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>

template <int CH>
__global__ void __launch_bounds__(32) k_dmma(const double* a, const double* b, double* c, int iters) {
    const int lane = threadIdx.x;
    double acc[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i][0] = c[lane] * i, acc[i][1] = c[lane] + i;
    const double av = a[lane], bv = b[lane];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1];
    c[blockIdx.x * 32 + lane] = s;
}

template <int CH>
__global__ void __launch_bounds__(32) k_dfma(const double* a, const double* b, double* c, int iters) {
    const int lane = threadIdx.x;
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = c[lane] * i;
    const double av = a[lane], bv = b[lane];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(av, bv, acc[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    c[blockIdx.x * 32 + lane] = s;
}

// both pipes at once: does DMMA share the FP64 FMA pipe?
template <int CH>
__global__ void __launch_bounds__(32) k_both(const double* a, const double* b, double* c, int iters) {
    const int lane = threadIdx.x;
    double acc[CH][2], f[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i][0] = c[lane] * i, acc[i][1] = c[lane] + i, f[i] = c[lane] - i;
    const double av = a[lane], bv = b[lane];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(av), "d"(bv));
            f[i] = fma(av, bv, f[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1] + f[i];
    c[blockIdx.x * 32 + lane] = s;
}

__global__ void k_round(const double* a, const double* b, const double* c, double* d) {  // one 8x8x4 product
    const int lane = threadIdx.x;
    double c0 = c[2 * lane], c1 = c[2 * lane + 1];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a[lane]), "d"(b[lane]));
    d[2 * lane] = c0, d[2 * lane + 1] = c1;
}

int main() {
    double *a, *b, *c;
    cudaMalloc(&a, 256);
    cudaMalloc(&b, 256);
    cudaMalloc(&c, 8 * 32 * 148 * 16);
    cudaMemset(a, 0, 256), cudaMemset(b, 0, 256), cudaMemset(c, 0, 8 * 32 * 148 * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int iters = 20000;
    for (int per_sm : {4, 8, 12, 16}) {
        const int grid = 148 * per_sm;
        float ms[3];
        for (int v = 0; v < 3; ++v) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (v == 0) k_dmma<15><<<grid, 32>>>(a, b, c, iters);
                if (v == 1) k_dfma<15><<<grid, 32>>>(a, b, c, iters);
                if (v == 2) k_both<15><<<grid, 32>>>(a, b, c, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[v], e0, e1);
            }
        }
        const double n = (double)iters * 15 * per_sm;  // warp instructions per SM (per pipe)
        const double clk = 1.965e6;                    // cycles per ms at 1965 MHz (nominal)
        printf("warps/SM %2d: DMMA %.3f ms = %.2f cyc/inst/SM | DFMA %.3f ms = %.2f cyc/inst/SM | both %.3f ms (sum %.3f)\n",
               per_sm, ms[0], ms[0] * clk / n, ms[1], ms[1] * clk / n, ms[2], ms[0] + ms[1]);
    }
    // rounding: d = c + sum_k a[g][k] b[k][n] against fma chains in both orders
    double ha[32], hb[32], hc[64], hd[64];
    srand(1);
    for (int i = 0; i < 32; ++i) ha[i] = rand() / (double)RAND_MAX - 0.5, hb[i] = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 64; ++i) hc[i] = rand() / (double)RAND_MAX - 0.5;
    double *da, *db, *dc, *dd;
    cudaMalloc(&da, 256), cudaMalloc(&db, 256), cudaMalloc(&dc, 512), cudaMalloc(&dd, 512);
    cudaMemcpy(da, ha, 256, cudaMemcpyHostToDevice), cudaMemcpy(db, hb, 256, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, hc, 512, cudaMemcpyHostToDevice);
    k_round<<<1, 32>>>(da, db, dc, dd);
    cudaMemcpy(hd, dd, 512, cudaMemcpyDeviceToHost);
    int same_fwd = 0, same_rev = 0;
    double worst = 0;
    for (int lane = 0; lane < 32; ++lane)
        for (int e = 0; e < 2; ++e) {
            const int g = lane >> 2, n = 2 * (lane & 3) + e;
            double f = hc[2 * lane + e], r = hc[2 * lane + e];
            for (int k = 0; k < 4; ++k) f = fma(ha[4 * g + k], hb[4 * n + k], f);   // A[g][k] at lane 4g+k, B[k][n] at lane 4n+k
            for (int k = 3; k >= 0; --k) r = fma(ha[4 * g + k], hb[4 * n + k], r);
            same_fwd += f == hd[2 * lane + e];
            same_rev += r == hd[2 * lane + e];
            worst = fmax(worst, fabs(f - hd[2 * lane + e]));
        }
    printf("rounding: %d / 64 equal to the k = 0..3 FMA chain, %d / 64 to the reversed chain, worst |diff| %.3e\n", same_fwd,
           same_rev, worst);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
