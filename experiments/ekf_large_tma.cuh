// TMA-staged variant of the multi-factor covariance sweep (streamed engine).
//
// Same arithmetic as k_large_sweep_p (ekf_large_delayed.cuh) — Sigma[r][c] <- Sigma[r][c] - sum_{j<P} K_j[r] W_j[c],
// factors applied in order with the same two FMAs per factor, so the result is bit-identical — but Sigma moves
// HBM -> shared memory -> HBM with the bulk copy engine in a 4-stage mbarrier pipeline instead of through
// registers: one producer warp keeps kStages x 32 KB of loads in flight per SM no matter how long the consumer
// warps spend on the 2*P FMAs per element, so the sweep stays on the HBM roofline for every P <= 8 (the
// register-path kernel loses 20-35 % for P >= 3 because one CTA per SM cannot overlap its FMAs with its loads).
//
// CTA = 8 consumer warps + 1 producer warp.  A work unit is kTmaCols (512) columns x kUnitRows rows; the unit's W
// pairs live in consumer registers, its rows stream through the stages kStageRows (8) at a time.
#pragma once
#include "../ekf-slam-ml_b200/csrc/bulk_copy.cuh"
#include "../ekf-slam-ml_b200/csrc/ekf_large_delayed.cuh"

namespace ekf {

constexpr int kTmaCols = 512;                       // columns per work unit (4 KB per row segment)
constexpr int kStageRows = 8;                       // rows per pipeline stage (32 KB)
constexpr int kStages = 4;
constexpr int kUnitRows = 128;                      // rows per work unit
constexpr int kTmaConsumers = 256;                  // consumer threads (each owns 2 adjacent columns)
constexpr int kTmaThreads = kTmaConsumers + 32;
constexpr int kTmaSmemBytes = kStages * kStageRows * kTmaCols * 8 + 2 * kStages * 8 + 64;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int P>
__global__ void __launch_bounds__(kTmaThreads, 1)
    k_large_sweep_tma(double* __restrict__ sig, long long ld, int n_rows, const double2* __restrict__ Kp,
                      const double2* __restrict__ Wp, long long row0, unsigned long long* __restrict__ n_updates,
                      int n_counted, const UpdateCmd* __restrict__ cmd) {
    if (cmd && !cmd->do_update) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_updates) *n_updates += (unsigned long long)n_counted;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);  // [kStages][kStageRows][kTmaCols]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kStageRows * kTmaCols * 8);
    uint64_t* empty = full + kStages;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int chunks = (int)((ld + kTmaCols - 1) / kTmaCols);
    const int row_units = (n_rows + kUnitRows - 1) / kUnitRows;
    const long long units = (long long)chunks * row_units;
    const bool producer = tid >= kTmaConsumers;
    int stage = 0;
    uint32_t phase = 0;  // parity of the current pass over the stage ring

    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int cu = (int)(u % chunks), ru = (int)(u / chunks);  // neighbouring CTAs share rows -> K stays in L1/L2
        const long long c0 = (long long)cu * kTmaCols;
        const int width = (int)(ld - c0 < kTmaCols ? ld - c0 : kTmaCols);  // multiple of 16 doubles
        const int r_begin = ru * kUnitRows;
        const int r_end = r_begin + kUnitRows < n_rows ? r_begin + kUnitRows : n_rows;
        const int groups = (r_end - r_begin + kStageRows - 1) / kStageRows;
        if (producer) {
            if (tid == kTmaConsumers) {
                for (int g = 0; g < groups; ++g) {
                    mbar_wait(empty + stage, phase ^ 1u);  // stage free (passes immediately on the first lap)
                    const int r = r_begin + g * kStageRows;
                    const int nr = r_end - r < kStageRows ? r_end - r : kStageRows;
                    mbar_arrive_expect_tx(full + stage, (uint32_t)(nr * width * 8));
                    for (int k = 0; k < nr; ++k)
                        bulk_g2s(tiles + ((size_t)stage * kStageRows + k) * kTmaCols, sig + (long long)(r + k) * ld + c0,
                                 (uint32_t)(width * 8), full + stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        } else {
            // consumer: columns {2 tid, 2 tid + 1} -> one conflict-free LDS.128 / STS.128 per row
            const int ca = 2 * tid;
            const bool has_a = ca < width;
            double2 w[P][2];
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const double2* wj = Wp + (long long)j * ld + c0;
                w[j][0] = has_a ? wj[ca] : make_double2(0.0, 0.0);
                w[j][1] = has_a ? wj[ca + 1] : make_double2(0.0, 0.0);
            }
            for (int g = 0; g < groups; ++g) {
                const int r = r_begin + g * kStageRows;
                const int nr = r_end - r < kStageRows ? r_end - r : kStageRows;
                mbar_wait(full + stage, phase);
                double* tile = tiles + (size_t)stage * kStageRows * kTmaCols;
                if (has_a) {
                    double2 v[kStageRows];
#pragma unroll
                    for (int k = 0; k < kStageRows; ++k)
                        if (k < nr) v[k] = *reinterpret_cast<const double2*>(tile + k * kTmaCols + ca);
#pragma unroll
                    for (int k = 0; k < kStageRows; ++k) {
                        if (k < nr) {
#pragma unroll
                            for (int j = 0; j < P; ++j) {
                                const double2 kk = Kp[(long long)j * ld + row0 + r + k];
                                v[k].x = apply_factor(v[k].x, kk, w[j][0]);
                                v[k].y = apply_factor(v[k].y, kk, w[j][1]);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kStageRows; ++k)
                        if (k < nr) *reinterpret_cast<double2*>(tile + k * kTmaCols + ca) = v[k];
                }
                fence_proxy_async_smem();             // generic-proxy writes -> visible to the bulk store
                named_bar_sync(1, kTmaConsumers);     // all four consumer warps are done with this stage
                if (tid == 0) {
                    for (int k = 0; k < nr; ++k)
                        bulk_s2g(sig + (long long)(r + k) * ld + c0, tile + k * kTmaCols, (uint32_t)(width * 8));
                    bulk_commit();
                    // the PREVIOUS stage's store has finished reading shared memory: hand that stage back
                    bulk_wait_read_1();
                    if (g > 0 || u != (long long)blockIdx.x) mbar_arrive(empty + (stage == 0 ? kStages - 1 : stage - 1));
                }
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    }
    if (!producer && tid == 0) bulk_wait_all();  // drain the last stores before the CTA's shared memory goes away
}

}  // namespace ekf
