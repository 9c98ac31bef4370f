// Engine 2: one filter whose covariance lives in HBM (maps too large for shared memory; cfg4 has
// n = 8,192 landmarks, Sigma = 16,387^2 fp64 = 2.1 GB).  Sigma is row-major with the row stride `ld`
// padded to a multiple of 16 doubles (128 B) so that every row starts on a cache line and 256-bit
// loads/stores are aligned.  Per landmark correction the work is
//   gain  : O(N)  — 5 rows + 5 columns of Sigma -> W = Hj Sigma (2 x N), K = Sigma Hj^T S^-1 (N x 2), state
//   sweep : O(N^2)— Sigma -= K W, streamed once through the SMs: one HBM read + one HBM write of Sigma.
// Restates rigid2d/src/ekf_slam.cpp:55-106, :108-197, :200-214, :217-276, :278-402.
#pragma once
#include "ekf_math.cuh"

namespace ekf {

// What the association step decided for one measurement; consumed by gain + sweep without a host round trip.
struct UpdateCmd {
    int do_update;  // 1: correct landmark `lm` with (sx, sy)
    int lm;
    int created;    // a new landmark was initialised for this measurement
    int pad;
    double sx, sy;
};

// New values of the five state entries that H_j depends on.  The gain kernel must not overwrite them while
// other CTAs still read the old ones, so they are parked here and committed by the sweep kernel.
struct Special5 {
    double v[5];
    int idx[5];
    int valid;
};

struct AssocPartial {
    double best, second;
    int best_i;
    int pad;
};

// ---------------------------------------------------------------- prediction (ekf_slam.cpp:55-106)
// One thread: motion model, state[0..2], the 3x3 robot block of At*Sigma*At^T + Q, and (a1, a2) for the strips.
__global__ void k_large_motion(double* __restrict__ state, double* __restrict__ sig, long long ld, double dtheta,
                               double dx, double* __restrict__ motion_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Motion m = motion_model(state[0], dtheta, dx);
    double s[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) s[r][c] = sig[r * ld + c];
    for (int c = 0; c < 3; ++c) {
        s[1][c] = fma(m.a1, s[0][c], s[1][c]);
        s[2][c] = fma(m.a2, s[0][c], s[2][c]);
    }
    for (int r = 0; r < 3; ++r) {
        s[r][1] = fma(s[r][0], m.a1, s[r][1]);
        s[r][2] = fma(s[r][0], m.a2, s[r][2]);
    }
    s[0][0] += kQ;
    s[1][1] += kQ;
    s[2][2] += kQ;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) sig[r * ld + c] = s[r][c];
    state[0] = state[0] + m.u0;
    state[1] = state[1] + m.u1;
    state[2] = state[2] + m.u2;
    motion_out[0] = m.a1;
    motion_out[1] = m.a2;
}

// Robot-landmark strips: rows 1,2 += a*row0 and cols 1,2 += a*col0 for indices k >= 3.  6N doubles touched.
__global__ void k_large_predict_strips(double* __restrict__ sig, long long ld, int N,
                                       const double* __restrict__ motion) {
    const int k = 3 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double a1 = motion[0], a2 = motion[1];
    const double r0 = sig[k];
    sig[ld + k] = fma(a1, r0, sig[ld + k]);
    sig[2 * ld + k] = fma(a2, r0, sig[2 * ld + k]);
    double* row = sig + (long long)k * ld;
    const double c0 = row[0];
    row[1] = fma(c0, a1, row[1]);
    row[2] = fma(c0, a2, row[2]);
}

// First measurement() call: every slot initialised from its reading (ekf_slam.cpp:113-128).  Also snapshots
// the entry-time pose, which measurement() keeps using for every landmark of the call (:109-111).
__global__ void k_large_init_landmarks(double* __restrict__ state, const double* __restrict__ xy, int n,
                                       int32_t* __restrict__ init_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *init_flag = 1;
    if (i >= n) return;
    double mx, my;
    landmark_from_reading(xy[2 * i], xy[2 * i + 1], state[0], state[1], state[2], mx, my);
    state[3 + 2 * i] = mx;
    state[4 + 2 * i] = my;
}

// ---------------------------------------------------------------- association (ekf_slam.cpp:217-276, 291-330)
// Mahalanobis distance of measurement j to every known landmark in parallel, block argmin with lowest-index
// tie-break, then the last block to finish reduces the per-block partials and takes the decision
// (new landmark / update / drop) on the device.
__device__ __forceinline__ void assoc_merge(double& best, double& second, int& best_i, double ob, double os, int oi) {
    if (better(ob, oi, best, best_i)) {
        second = fmin(best, os);
        best = ob;
        best_i = oi;
    } else {
        second = fmin(second, fmin(ob, os));
    }
}

__global__ void __launch_bounds__(256)
    k_large_assoc(const double* __restrict__ sig, long long ld, double* __restrict__ state, int n,
                  const double* __restrict__ meas, int j, int* __restrict__ known_count_p,
                  AssocPartial* __restrict__ partials, unsigned int* __restrict__ done_counter,
                  UpdateCmd* __restrict__ cmd, int32_t* __restrict__ assoc_out, double* __restrict__ dmin_out,
                  double* __restrict__ second_out, uint8_t* __restrict__ created_out) {
    __shared__ AssocPartial sh[8];
    __shared__ bool is_last;
    const int known_count = *known_count_p;
    const double sx = meas[2 * j], sy = meas[2 * j + 1];
    double zr, zphi;
    range_bearing(sx, sy, zr, zphi);
    const double theta = state[0], x = state[1], y = state[2];
    double best = INFINITY, second = INFINITY;
    int best_i = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < known_count; i += gridDim.x * blockDim.x) {
        double d = maha_distance(sig, ld, i, state[3 + 2 * i], state[4 + 2 * i], zr, zphi, theta, x, y);
        if (!(d == d)) d = INFINITY;
        if (d < best) {
            second = best;
            best = d;
            best_i = i;
        } else if (d < second) {
            second = d;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const double os = __shfl_xor_sync(0xffffffffu, second, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        assoc_merge(best, second, best_i, ob, os, oi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = AssocPartial{best, second, best_i, 0};
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) assoc_merge(best, second, best_i, sh[w].best, sh[w].second, sh[w].best_i);
        partials[blockIdx.x] = AssocPartial{best, second, best_i, 0};
        __threadfence();
        const unsigned int t = atomicAdd(done_counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence();
    best = INFINITY;
    second = INFINITY;
    best_i = 0x7fffffff;
    for (unsigned int b = 0; b < gridDim.x; ++b) {
        const volatile AssocPartial* pp = partials + b;
        assoc_merge(best, second, best_i, pp->best, pp->second, pp->best_i);
    }
    *done_counter = 0;  // re-arm for the next measurement (stream-ordered)
    double min_d = kGateNew;
    int min_idx = known_count;
    if (best < kGateNew) {
        min_d = best;
        min_idx = best_i;
        second = fmin(second, kGateNew);
    } else {
        second = best;
    }
    if (dmin_out) dmin_out[j] = min_d;
    if (second_out) second_out[j] = second;
    int created = 0;
    if (min_idx == known_count && min_idx < n) {  // ekf_slam.cpp:318-327
        double mx, my;
        landmark_from_reading(sx, sy, theta, x, y, mx, my);
        state[3 + 2 * min_idx] = mx;
        state[4 + 2 * min_idx] = my;
        *known_count_p = known_count + 1;
        min_d = 0.0;
        created = 1;
    }
    const int upd = min_d < kGateUpdate;  // :330
    cmd->do_update = upd;
    cmd->lm = min_idx;
    cmd->created = created;
    cmd->sx = sx;
    cmd->sy = sy;
    if (assoc_out) assoc_out[j] = upd ? min_idx : -1;
    if (created_out) created_out[j] = (uint8_t)created;
}

// ---------------------------------------------------------------- gain (ekf_slam.cpp:138-187 / 331-385)
// grid covers k in [0, ld): thread k produces W[:,k] (from the 5 rows), K[k,:] (from the 5 columns of row k)
// and the new state[k].  Every CTA first recomputes the O(1) part (Hj, S^-1, innovation) from the 5x5 block.
struct GainShared {
    Hj h;
    Sym2 si;
    double nu0, nu1;
    int i3;
    int active;
};

__global__ void __launch_bounds__(256)
    k_large_gain(const double* __restrict__ sig, long long ld, int N, double* __restrict__ state,
                 const double* __restrict__ pose_src, const UpdateCmd* __restrict__ cmd, int lm_arg, double sx_arg,
                 double sy_arg, double2* __restrict__ K2, double2* __restrict__ W2, Special5* __restrict__ sp) {
    __shared__ GainShared g;
    if (threadIdx.x == 0) {
        int lm = lm_arg;
        double sx = sx_arg, sy = sy_arg;
        int active = 1;
        if (cmd) {
            active = cmd->do_update;
            lm = cmd->lm;
            sx = cmd->sx;
            sy = cmd->sy;
        }
        g.active = active;
        if (active) {
            const int i3 = 3 + 2 * lm;
            const double theta = pose_src[0], x = pose_src[1], y = pose_src[2];
            const Hj h = make_hj(state[i3], state[i3 + 1], theta, x, y);
            const long long id[5] = {0, 1, 2, i3, i3 + 1};
            double w0[5], w1[5];
            for (int l = 0; l < 5; ++l) {
                const double s0 = sig[id[0] * ld + id[l]], s1 = sig[id[1] * ld + id[l]], s2 = sig[id[2] * ld + id[l]];
                const double s3 = sig[id[3] * ld + id[l]], s4 = sig[id[4] * ld + id[l]];
                w0[l] = h_row0(h, s1, s2, s3, s4);
                w1[l] = h_row1(h, s0, s1, s2, s3, s4);
            }
            const double s00 = h_row0(h, w0[1], w0[2], w0[3], w0[4]) + kR;
            const double s01 = h_row1(h, w0[0], w0[1], w0[2], w0[3], w0[4]);
            const double s10 = h_row0(h, w1[1], w1[2], w1[3], w1[4]);
            const double s11 = h_row1(h, w1[0], w1[1], w1[2], w1[3], w1[4]) + kR;
            g.h = h;
            g.si = inv2x2(s00, s01, s10, s11);
            double zr, zphi;
            range_bearing(sx, sy, zr, zphi);
            g.nu0 = __dsub_rn(zr, h.zr);
            g.nu1 = normalize_angle(__dsub_rn(zphi, h.zphi));
            g.i3 = i3;
        }
    }
    __syncthreads();
    if (!g.active) return;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ld) return;
    if (k >= N) {  // padding columns/rows stay inert in the sweep
        K2[k] = make_double2(0.0, 0.0);
        W2[k] = make_double2(0.0, 0.0);
        return;
    }
    const Hj h = g.h;
    const int i3 = g.i3, i4 = i3 + 1;
    {
        const double s0 = sig[k], s1 = sig[ld + k], s2 = sig[2 * ld + k];
        const double s3 = sig[i3 * ld + k], s4 = sig[i4 * ld + k];
        W2[k] = make_double2(h_row0(h, s1, s2, s3, s4), h_row1(h, s0, s1, s2, s3, s4));
    }
    const double* row = sig + k * ld;
    const double r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[i3], r4 = row[i4];
    const double p0 = h_row0(h, r1, r2, r3, r4), p1 = h_row1(h, r0, r1, r2, r3, r4);
    const double k0 = fma(p1, g.si.i10, p0 * g.si.i00);
    const double k1 = fma(p1, g.si.i11, p0 * g.si.i01);
    K2[k] = make_double2(k0, k1);
    double ns = state[k] + fma(k1, g.nu1, k0 * g.nu0);
    int slot = -1;
    if (k < 3)
        slot = (int)k;
    else if (k == i3)
        slot = 3;
    else if (k == i4)
        slot = 4;
    if (slot < 0) {
        state[k] = ns;
    } else {
        if (k == 0) ns = normalize_angle(ns);  // ekf_slam.cpp:187
        sp->v[slot] = ns;
        sp->idx[slot] = (int)k;
        if (k == 0) sp->valid = 1;
    }
}

// ---------------------------------------------------------------- sweep (ekf_slam.cpp:191-192 / 389-390)
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

constexpr int kSweepThreads = 256;
constexpr int kSweepColsPerThread = 4;
constexpr int kSweepChunk = kSweepThreads * kSweepColsPerThread;  // 1024 columns = 8 KB of a row
constexpr int kSweepRows = 32;                                    // rows per tile

// Sigma[r][c] -= K[r][0] W[0][c] + K[r][1] W[1][c], streamed.  A tile is kSweepRows x kSweepChunk doubles
// (256 KB); each thread keeps its four columns of W in registers for the whole tile and moves 32 B per row
// with one 256-bit load and one 256-bit store.  Tiles are walked by a grid-stride loop so the grid can be
// sized to a multiple of the SM count.
__global__ void __launch_bounds__(kSweepThreads)
    k_large_sweep(double* __restrict__ sig, long long ld, int N, const double2* __restrict__ K2,
                  const double2* __restrict__ W2, const UpdateCmd* __restrict__ cmd, Special5* __restrict__ sp,
                  double* __restrict__ state, unsigned long long* __restrict__ n_updates) {
    if (cmd && !cmd->do_update) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (sp && sp->valid) {
            for (int s = 0; s < 5; ++s) state[sp->idx[s]] = sp->v[s];
            sp->valid = 0;
        }
        if (n_updates) *n_updates += 1;
    }
    const int chunks = (int)((ld + kSweepChunk - 1) / kSweepChunk);
    const int row_blocks = (N + kSweepRows - 1) / kSweepRows;
    const long long tiles = (long long)chunks * row_blocks;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int rb = (int)(t / chunks), cc = (int)(t - (long long)rb * chunks);
        const long long c = (long long)cc * kSweepChunk + threadIdx.x * kSweepColsPerThread;
        if (c >= ld) continue;
        const double2 w0 = W2[c], w1 = W2[c + 1], w2 = W2[c + 2], w3 = W2[c + 3];
        const int r_begin = rb * kSweepRows;
        const int r_end = min(N, r_begin + kSweepRows);
        double* p = sig + (long long)r_begin * ld + c;
        int r = r_begin;
        for (; r + 4 <= r_end; r += 4) {
            double v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ld256(p + u * ld, v[u][0], v[u][1], v[u][2], v[u][3]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 k = K2[r + u];
                v[u][0] = fma(-k.y, w0.y, fma(-k.x, w0.x, v[u][0]));
                v[u][1] = fma(-k.y, w1.y, fma(-k.x, w1.x, v[u][1]));
                v[u][2] = fma(-k.y, w2.y, fma(-k.x, w2.x, v[u][2]));
                v[u][3] = fma(-k.y, w3.y, fma(-k.x, w3.x, v[u][3]));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) st256(p + u * ld, v[u][0], v[u][1], v[u][2], v[u][3]);
            p += 4 * ld;
        }
        for (; r < r_end; ++r) {
            double a, b, cdd, d;
            ld256(p, a, b, cdd, d);
            const double2 k = K2[r];
            a = fma(-k.y, w0.y, fma(-k.x, w0.x, a));
            b = fma(-k.y, w1.y, fma(-k.x, w1.x, b));
            cdd = fma(-k.y, w2.y, fma(-k.x, w2.x, cdd));
            d = fma(-k.y, w3.y, fma(-k.x, w3.x, d));
            st256(p, a, b, cdd, d);
            p += ld;
        }
    }
}

// Sigma0 = blockdiag(0_3, 100 I) (ekf_slam.cpp:29-36) on an already-zeroed buffer.
__global__ void k_large_init_sigma(double* __restrict__ sig, long long ld, int N) {
    const long long k = 3 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) sig[k * ld + k] = kSigma0;
}

}  // namespace ekf
