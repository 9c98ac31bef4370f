// Tuning aid (not part of the product): register-path sweep variants.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ekf-slam-ml_b200/csrc/ekf_large_delayed.cuh"
using namespace ekf;
template <int P, int COLS, int U, int ROWS, int TH>
float time_one(double* sig, long long ld, int N, double2* Kp, double2* Wp, int reps) {
    const long long chunk = (long long)TH * COLS;
    const long long tiles = ((ld + chunk - 1) / chunk) * ((N + ROWS - 1) / ROWS);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_large_sweep_p<P, COLS, U, ROWS, TH><<<(unsigned)tiles, TH>>>(sig, ld, N, Kp, Wp, 0, nullptr, 0, nullptr);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) k_large_sweep_p<P, COLS, U, ROWS, TH><<<(unsigned)tiles, TH>>>(sig, ld, N, Kp, Wp, 0, nullptr, 0, nullptr);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return ms / reps;
}
int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 16387;
    const long long ld = (N + 15) & ~15LL;
    double* sig; double2 *Kp, *Wp;
    cudaMalloc(&sig, sizeof(double) * ld * N); cudaMalloc(&Kp, sizeof(double2) * ld * kMaxPending); cudaMalloc(&Wp, sizeof(double2) * ld * kMaxPending);
    cudaMemset(sig, 0, sizeof(double) * ld * N);
    std::vector<double2> h(ld * kMaxPending);
    for (auto& v : h) v = make_double2(1e-3 * (rand() / (double)RAND_MAX), 1e-3 * (rand() / (double)RAND_MAX));
    cudaMemcpy(Kp, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(Wp, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice);
    const double gb = 16.0 * N * (double)N / 1e9;
#define T(P, C, U, R, TH) { float ms = time_one<P, C, U, R, TH>(sig, ld, N, Kp, Wp, 5); printf("P=%d COLS=%d U=%d ROWS=%d TH=%d  %8.3f ms  %7.1f GB/s\n", P, C, U, R, TH, ms, gb / (ms * 1e-3)); }
    T(4, 2, 4, 64, 128) T(4, 2, 4, 64, 256) T(4, 2, 8, 64, 128) T(4, 2, 2, 64, 256)
    T(6, 2, 4, 64, 128) T(6, 2, 4, 64, 256) T(6, 2, 8, 64, 128) T(6, 2, 2, 64, 256)
    T(8, 2, 4, 64, 128) T(8, 2, 4, 64, 256) T(8, 2, 8, 64, 128) T(8, 2, 2, 64, 256) T(8, 2, 4, 128, 256)
    return 0;
}
