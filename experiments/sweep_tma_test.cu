// Tuning / validation aid (not part of the product): TMA-staged sweep vs the register-path sweep, bitwise + timing.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../ekf-slam-ml_b200/csrc/ekf_large_tma.cuh"
using namespace ekf;

template <int P>
void run(double* a, double* b, const double* init, long long ld, int N, double2* Kp, double2* Wp, int sm) {
    const size_t bytes = sizeof(double) * ld * N;
    cudaMemcpy(a, init, bytes, cudaMemcpyDeviceToDevice);
    cudaMemcpy(b, init, bytes, cudaMemcpyDeviceToDevice);
    launch_sweep_p(P, a, ld, N, Kp, Wp, 0, nullptr, 0, nullptr, sm, 0);
    cudaFuncSetAttribute(k_large_sweep_tma<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
    const long long units = ((ld + kTmaCols - 1) / kTmaCols) * ((N + kUnitRows - 1) / kUnitRows);
    const unsigned grid = (unsigned)std::min<long long>(units, sm);
    k_large_sweep_tma<P><<<grid, kTmaThreads, kTmaSmemBytes>>>(b, ld, N, Kp, Wp, 0, nullptr, 0, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("P=%d error %s\n", P, cudaGetErrorString(e)); exit(1); }
    std::vector<double> ha((size_t)ld * 64), hb((size_t)ld * 64);
    // compare the whole matrix on the device-side cheaply: copy back in slabs
    size_t bad = 0;
    for (int r0 = 0; r0 < N; r0 += 64) {
        const int nr = std::min(64, N - r0);
        cudaMemcpy(ha.data(), a + (size_t)r0 * ld, sizeof(double) * ld * nr, cudaMemcpyDeviceToHost);
        cudaMemcpy(hb.data(), b + (size_t)r0 * ld, sizeof(double) * ld * nr, cudaMemcpyDeviceToHost);
        if (memcmp(ha.data(), hb.data(), sizeof(double) * ld * nr) != 0)
            for (size_t k = 0; k < (size_t)ld * nr; ++k) bad += (ha[k] != hb[k]);
    }
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int reps = 5;
    float ms_reg = 0, ms_tma = 0;
    cudaEventRecord(t0);
    for (int i = 0; i < reps; ++i) launch_sweep_p(P, a, ld, N, Kp, Wp, 0, nullptr, 0, nullptr, sm, 0);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    cudaEventElapsedTime(&ms_reg, t0, t1);
    cudaEventRecord(t0);
    for (int i = 0; i < reps; ++i) k_large_sweep_tma<P><<<grid, kTmaThreads, kTmaSmemBytes>>>(b, ld, N, Kp, Wp, 0, nullptr, 0, nullptr);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    cudaEventElapsedTime(&ms_tma, t0, t1);
    const double gb = 16.0 * N * (double)N / 1e9;
    printf("P=%d mismatches=%zu  reg %.3f ms (%.0f GB/s)   tma %.3f ms (%.0f GB/s)\n", P, bad, ms_reg / reps,
           gb / (ms_reg / reps * 1e-3), ms_tma / reps, gb / (ms_tma / reps * 1e-3));
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 16387;
    const long long ld = (N + 15) & ~15LL;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    double *a, *b, *init;
    double2 *Kp, *Wp;
    cudaMalloc(&a, sizeof(double) * ld * N);
    cudaMalloc(&b, sizeof(double) * ld * N);
    cudaMalloc(&init, sizeof(double) * ld * N);
    cudaMalloc(&Kp, sizeof(double2) * (ld * kMaxPending + 16));
    cudaMalloc(&Wp, sizeof(double2) * (ld * kMaxPending + 16));
    std::vector<double> h((size_t)ld * N);
    for (auto& v : h) v = rand() / (double)RAND_MAX;
    cudaMemcpy(init, h.data(), sizeof(double) * ld * N, cudaMemcpyHostToDevice);
    std::vector<double2> f(ld * kMaxPending);
    for (auto& v : f) v = make_double2(1e-3 * (rand() / (double)RAND_MAX), 1e-3 * (rand() / (double)RAND_MAX));
    cudaMemcpy(Kp, f.data(), sizeof(double2) * f.size(), cudaMemcpyHostToDevice);
    for (auto& v : f) v = make_double2(1e-3 * (rand() / (double)RAND_MAX), 1e-3 * (rand() / (double)RAND_MAX));
    cudaMemcpy(Wp, f.data(), sizeof(double2) * f.size(), cudaMemcpyHostToDevice);
    run<1>(a, b, init, ld, N, Kp, Wp, sm);
    run<4>(a, b, init, ld, N, Kp, Wp, sm);
    run<6>(a, b, init, ld, N, Kp, Wp, sm);
    run<8>(a, b, init, ld, N, Kp, Wp, sm);
    return 0;
}
