// Does SHFL compete with shared-memory loads for the same LSU data pipe on sm_100a?
// 16 warps per SM (one warp per CTA, 16 CTAs per SM), each runs `iters` rounds of 8 LDS.64 and / or 8 SHFL.32.
#include <cstdio>
#include <cuda_runtime.h>
template <int LDS, int SHF>
__global__ void __launch_bounds__(32, 16) k(double* out, int iters) {
    __shared__ double buf[32 * 9];
    const int lane = threadIdx.x;
    for (int i = lane; i < 32 * 9; i += 32) buf[i] = i * 0.5;
    __syncwarp();
    double acc = 0.0;
    unsigned v = lane * 2654435761u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (LDS) acc += buf[((it + u) & 7) * 32 + lane];
            if (SHF) v += __shfl_sync(0xffffffffu, v, (lane + u + it) & 31);
        }
    }
    out[blockIdx.x * 32 + lane] = acc + v;
}
template <int LDS, int SHF>
float run(double* d, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<LDS, SHF><<<148 * 16, 32>>>(d, iters);
    cudaEventRecord(a);
    k<LDS, SHF><<<148 * 16, 32>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
int main() {
    double* d;
    cudaMalloc(&d, 148 * 16 * 32 * 8);
    const int iters = 20000;
    const float l = run<1, 0>(d, iters), s = run<0, 1>(d, iters), b = run<1, 1>(d, iters);
    // per SM per cycle: 16 warps * iters * 8 ops / (ms * 1.965e6 cycles)
    const double ops = 16.0 * iters * 8;
    printf("LDS.64 only : %.3f ms  (%.2f warp-instr/clk/SM)\n", l, ops / (l * 1.965e6));
    printf("SHFL only   : %.3f ms  (%.2f warp-instr/clk/SM)\n", s, ops / (s * 1.965e6));
    printf("both        : %.3f ms  (sum of the two alone: %.3f ms)\n", b, l + s);
    return 0;
}
