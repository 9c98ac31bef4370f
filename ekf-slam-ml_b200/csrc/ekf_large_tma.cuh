// TMA-staged variant of the multi-factor covariance sweep (streamed engine).
//
// Same arithmetic as k_large_sweep_p (ekf_large_delayed.cuh) — Sigma[r][c] <- Sigma[r][c] - sum_{j<P} K_j[r] W_j[c],
// factors applied in order with the same two FMAs per factor, so the result is bit-identical — but Sigma moves
// HBM -> shared memory -> HBM with the bulk copy engine in a kStages-deep mbarrier pipeline instead of through
// registers, so the loads of the next tiles are in flight no matter how long the 2*P FMAs per element take.
// Measured on B200 at N = 16,387 (warm, back to back): P=4 0.677 ms (6.35 TB/s), P=6 0.687 ms (6.25 TB/s = 0.95 of the
// measured copy peak), P=8 0.80 ms, against 0.79 / 0.82 / 0.98 ms for the register path; for P <= 2 the register path
// (0.65 ms) wins, so launch_sweep() below uses it there.
//
// CTA = 1 producer warp + 8 consumer warps + 1 store warp (320 threads), one CTA per SM, persistent over work units of
// kTmaCols (512) columns x kUnitRows rows.  Per stage (kStageRows = 8 rows x 512 columns = 32 KB):
//   producer : waits empty[s], bulk-loads the 8 row segments and the stage's K rows (P x 8 pairs), full[s] counts bytes;
//   consumers: wait full[s]; thread (w, lane) owns columns {2t, 2t+1, 256+2t, 256+2t+1}, t = 32 (w % 4) + lane, and the
//              rows 4 (w / 4) .. +3 of the stage; W pairs of its columns live in registers for the whole unit, K pairs
//              come from shared memory (warp-uniform); arrive on done[s] (one arrive per warp);
//   storer   : waits done[s], bulk-stores the 8 row segments, and hands stage s-1 back (empty[s-1]) once the store
//              of s-1 has finished reading shared memory.
#pragma once
#include "bulk_copy.cuh"
#include "ekf_large_delayed.cuh"

#ifndef EKF_TMA_NARROW_FROM
#define EKF_TMA_NARROW_FROM 5  // pending-factor count from which the consumers use the 2-column mapping
#endif

namespace ekf {

constexpr int kTmaCols = 512;
constexpr int kStageRows = 8;
constexpr int kStages = 4;
constexpr int kUnitRows = 128;
constexpr int kTmaConsumerWarps = 8;
constexpr int kTmaThreads = 32 * (kTmaConsumerWarps + 2);
constexpr int kTileBytes = kStageRows * kTmaCols * 8;
constexpr int kKBytes = kMaxPending * kStageRows * 16;
constexpr int kTmaSmemBytes = kStages * (kTileBytes + kKBytes) + 3 * kStages * 8 + 64;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// NARROW = false: a consumer thread owns 4 columns x 4 rows of a stage (16 P registers of W pairs);
// NARROW = true : 2 columns x all 8 rows (8 P registers), which keeps P = 7, 8 out of local memory.
template <int P, bool NARROW = (P >= EKF_TMA_NARROW_FROM)>
__global__ void __launch_bounds__(kTmaThreads, 1)
    k_large_sweep_tma(double* __restrict__ sig, long long ld, int n_rows, const double2* __restrict__ Kp,
                      const double2* __restrict__ Wp, long long row0, unsigned long long* __restrict__ n_updates,
                      int n_counted, const UpdateCmd* __restrict__ cmd) {
    if (cmd && !cmd->do_update) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_updates) *n_updates += (unsigned long long)n_counted;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);                                   // [kStages][8][512]
    double2* ksm = reinterpret_cast<double2*>(smem_raw + (size_t)kStages * kTileBytes);    // [kStages][kMaxPending][8]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * (kTileBytes + kKBytes));
    uint64_t* done = full + kStages;
    uint64_t* empty = done + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(done + s, kTmaConsumerWarps);
            mbar_init(empty + s, 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int chunks = (int)((ld + kTmaCols - 1) / kTmaCols);
    const int row_units = (n_rows + kUnitRows - 1) / kUnitRows;
    const long long units = (long long)chunks * row_units;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;

    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int cu = (int)(u % chunks), ru = (int)(u / chunks);
        const long long c0 = (long long)cu * kTmaCols;
        const int width = (int)(ld - c0 < kTmaCols ? ld - c0 : kTmaCols);  // multiple of 16 doubles
        const int r_begin = ru * kUnitRows;
        const int r_end = r_begin + kUnitRows < n_rows ? r_begin + kUnitRows : n_rows;
        const int groups = (r_end - r_begin + kStageRows - 1) / kStageRows;
        if (warp == 0) {
            // ---------------------------------------------------------------- producer
            if (lane == 0) {
                for (int g = 0; g < groups; ++g) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    const int r = r_begin + g * kStageRows;
                    const int nr = r_end - r < kStageRows ? r_end - r : kStageRows;
                    mbar_arrive_expect_tx(full + stage, (uint32_t)(nr * width * 8 + P * kStageRows * 16));
                    double* tile = tiles + (size_t)stage * kStageRows * kTmaCols;
                    for (int k = 0; k < nr; ++k)
                        bulk_g2s(tile + k * kTmaCols, sig + (long long)(r + k) * ld + c0, (uint32_t)(width * 8), full + stage);
                    // K rows of this stage (8 pairs = 128 B per factor; Kp is padded to ld >= n_rows + 15)
                    for (int j = 0; j < P; ++j)
                        bulk_g2s(ksm + ((size_t)stage * kMaxPending + j) * kStageRows, Kp + (long long)j * ld + row0 + r,
                                 kStageRows * 16, full + stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        } else if (warp <= kTmaConsumerWarps) {
            // ---------------------------------------------------------------- consumers
            const int w = warp - 1;
            if constexpr (!NARROW) {
                const int t = 32 * (w & 3) + lane;
                const int ca = 2 * t, cb = kTmaCols / 2 + 2 * t;
                const bool has_a = ca < width, has_b = cb < width;
                const int kr0 = 4 * (w >> 2);  // first row of this warp's half of the stage
                double2 wv[P][4];
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const double2* wj = Wp + (long long)j * ld + c0;
                    wv[j][0] = has_a ? wj[ca] : make_double2(0.0, 0.0);
                    wv[j][1] = has_a ? wj[ca + 1] : make_double2(0.0, 0.0);
                    wv[j][2] = has_b ? wj[cb] : make_double2(0.0, 0.0);
                    wv[j][3] = has_b ? wj[cb + 1] : make_double2(0.0, 0.0);
                }
                for (int g = 0; g < groups; ++g) {
                    const int r = r_begin + g * kStageRows;
                    const int nr = r_end - r < kStageRows ? r_end - r : kStageRows;
                    mbar_wait(full + stage, phase);
                    double* tile = tiles + (size_t)stage * kStageRows * kTmaCols;
                    const double2* kst = ksm + (size_t)stage * kMaxPending * kStageRows;
                    double2 va[4], vb[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        va[k] = make_double2(0.0, 0.0);
                        vb[k] = make_double2(0.0, 0.0);
                        if (kr0 + k < nr) {
                            if (has_a) va[k] = *reinterpret_cast<const double2*>(tile + (kr0 + k) * kTmaCols + ca);
                            if (has_b) vb[k] = *reinterpret_cast<const double2*>(tile + (kr0 + k) * kTmaCols + cb);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < P; ++j) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const double2 kk = kst[j * kStageRows + kr0 + k];
                            va[k].x = apply_factor(va[k].x, kk, wv[j][0]);
                            va[k].y = apply_factor(va[k].y, kk, wv[j][1]);
                            vb[k].x = apply_factor(vb[k].x, kk, wv[j][2]);
                            vb[k].y = apply_factor(vb[k].y, kk, wv[j][3]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (kr0 + k < nr) {
                            if (has_a) *reinterpret_cast<double2*>(tile + (kr0 + k) * kTmaCols + ca) = va[k];
                            if (has_b) *reinterpret_cast<double2*>(tile + (kr0 + k) * kTmaCols + cb) = vb[k];
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(done + stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            } else {
                const int ca = 2 * (32 * w + lane);  // 256 threads x 2 columns = the 512-column tile
                const bool has = ca < width;
                double2 wv[P][2];
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const double2* wj = Wp + (long long)j * ld + c0;
                    wv[j][0] = has ? wj[ca] : make_double2(0.0, 0.0);
                    wv[j][1] = has ? wj[ca + 1] : make_double2(0.0, 0.0);
                }
                for (int g = 0; g < groups; ++g) {
                    const int r = r_begin + g * kStageRows;
                    const int nr = r_end - r < kStageRows ? r_end - r : kStageRows;
                    mbar_wait(full + stage, phase);
                    double* tile = tiles + (size_t)stage * kStageRows * kTmaCols;
                    const double2* kst = ksm + (size_t)stage * kMaxPending * kStageRows;
                    // the stage's rows go through in batches of RBN so that v[] + the 8 P registers of W pairs fit
                    constexpr int RBN = (P > 12) ? 4 : kStageRows;
#pragma unroll
                    for (int k0 = 0; k0 < kStageRows; k0 += RBN) {
                        double2 v[RBN];
#pragma unroll
                        for (int k = 0; k < RBN; ++k) {
                            v[k] = make_double2(0.0, 0.0);
                            if (k0 + k < nr && has) v[k] = *reinterpret_cast<const double2*>(tile + (k0 + k) * kTmaCols + ca);
                        }
#pragma unroll
                        for (int j = 0; j < P; ++j) {
#pragma unroll
                            for (int k = 0; k < RBN; ++k) {
                                const double2 kk = kst[j * kStageRows + k0 + k];
                                v[k].x = apply_factor(v[k].x, kk, wv[j][0]);
                                v[k].y = apply_factor(v[k].y, kk, wv[j][1]);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < RBN; ++k)
                            if (k0 + k < nr && has) *reinterpret_cast<double2*>(tile + (k0 + k) * kTmaCols + ca) = v[k];
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(done + stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        } else {
            // ---------------------------------------------------------------- store warp
            if (lane == 0) {
                for (int g = 0; g < groups; ++g) {
                    const int r = r_begin + g * kStageRows;
                    const int nr = r_end - r < kStageRows ? r_end - r : kStageRows;
                    mbar_wait(done + stage, phase);
                    double* tile = tiles + (size_t)stage * kStageRows * kTmaCols;
                    for (int k = 0; k < nr; ++k)
                        bulk_s2g(sig + (long long)(r + k) * ld + c0, tile + k * kTmaCols, (uint32_t)(width * 8));
                    bulk_commit();
                    bulk_wait_read_1();  // every store but the newest has finished reading shared memory
                    if (!first) mbar_arrive(empty + (stage == 0 ? kStages - 1 : stage - 1));
                    first = false;
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    }
    if (warp == kTmaConsumerWarps + 1 && lane == 0) bulk_wait_all();  // drain before the CTA's shared memory goes away
}

// Dispatch: register path for 1-2 factors, TMA pipeline for 3-8.  Kp must have at least kStageRows pairs of slack
// after its last used row (the stage's K rows are fetched 8 at a time).
inline cudaError_t launch_sweep(int pending, double* sig, long long ld, int n_rows, const double2* Kp, const double2* Wp,
                                long long row0, unsigned long long* n_updates, int n_counted, const UpdateCmd* cmd,
                                int sm_count, cudaStream_t stream) {
    if (pending <= 2) return launch_sweep_p(pending, sig, ld, n_rows, Kp, Wp, row0, n_updates, n_counted, cmd, sm_count, stream);
    const long long units = ((ld + kTmaCols - 1) / kTmaCols) * ((n_rows + kUnitRows - 1) / kUnitRows);
    const unsigned grid = (unsigned)(units < sm_count ? (units < 1 ? 1 : units) : sm_count);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
#define EKF_TMA_CASE(PP)                                                                                              \
    case PP: {                                                                                                        \
        /* the opt-in is per device: one flag per device, indexed by the device the launch goes to */                 \
        static bool attr_set[64] = {false};                                                                           \
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {                                                                 \
            cudaError_t e = cudaFuncSetAttribute(k_large_sweep_tma<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                                 kTmaSmemBytes);                                                      \
            if (e != cudaSuccess) return e;                                                                           \
            if (dev >= 0 && dev < 64) attr_set[dev] = true;                                                           \
        }                                                                                                             \
        k_large_sweep_tma<PP><<<grid, kTmaThreads, kTmaSmemBytes, stream>>>(sig, ld, n_rows, Kp, Wp, row0, n_updates, \
                                                                             n_counted, cmd);                        \
        break;                                                                                                        \
    }
    switch (pending) {
        EKF_TMA_CASE(3)
        EKF_TMA_CASE(4)
        EKF_TMA_CASE(5)
        EKF_TMA_CASE(6)
        EKF_TMA_CASE(7)
        EKF_TMA_CASE(8)
#if EKF_MAX_PENDING > 8
        EKF_TMA_CASE(9)
        EKF_TMA_CASE(10)
        EKF_TMA_CASE(11)
        EKF_TMA_CASE(12)
#endif
#if EKF_MAX_PENDING > 12
        EKF_TMA_CASE(13)
        EKF_TMA_CASE(14)
        EKF_TMA_CASE(15)
        EKF_TMA_CASE(16)
#endif
        default:
            return cudaErrorInvalidValue;
    }
#undef EKF_TMA_CASE
    return cudaGetLastError();
}

}  // namespace ekf
