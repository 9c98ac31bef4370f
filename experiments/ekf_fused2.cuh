// Engine 1b: the known-association SLAM step (prediction + measurement()) of one filter per CTA with TWO warps in
// different roles.  The single-warp kernel (ekf_fused.cuh) is latency-bound: per correction a ~500-instruction scalar
// fp64 chain (sqrt, 8 divisions, atan2, angle wraps) and a ~500-instruction shared-memory rank-2 update run one after
// the other in one in-order instruction stream, with only 12 warps per SM (Sigma in shared memory caps the resident
// filters).  Here the same 12 filters per SM get 24 warps:
//   warp S (scalar): H_j / z_hat of the next landmark, S = H Sigma H^T + R, its inverse, the innovation;
//   warp M (matrix): prediction strips, the row / column gathers (W = H Sigma, P = Sigma H^T), K, the state update,
//                    and the rank-2 update of Sigma;
// so the long scalar chain of correction i+1 really overlaps the rank-2 update of correction i.  The two warps
// hand over through shared memory and four named barriers per correction.  Same arithmetic, same order of
// operations per element as ekf_fused.cuh (results are bit-identical to it).
// Restates rigid2d/src/ekf_slam.cpp:55-106 (prediction) and :108-197 (measurement).
#pragma once
#include "../ekf-slam-ml_b200/csrc/ekf_fused.cuh"

namespace ekf {

struct Mailbox {  // scalar warp -> matrix warp
    Hj h;
    Sym2 si;
    double nu0, nu1;
};

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

struct Fused2Smem {
    int off_sig, off_st, off_k2, off_w2, off_z, off_mail, off_bar, total;
    __host__ __device__ explicit Fused2Smem(int n) {
        const int N = 3 + 2 * n;
        int o = 0;
        off_sig = o;
        o += fused_round16(N * N) * 8;
        off_st = o;
        o += fused_round16(N) * 8;
        off_k2 = o;
        o += fused_round16(2 * N) * 8;
        off_w2 = o;
        o += fused_round16(2 * N) * 8;
        off_z = o;
        o += fused_round16(2 * n) * 8;
        off_mail = o;
        o += 128;
        off_bar = o;
        o += 16;
        total = o;
    }
};

template <int NL>
__global__ void __launch_bounds__(64, 12) ekf_fused2_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int n = NL;
    constexpr int N = 3 + 2 * NL;
    const Fused2Smem L(n);
    double* sig = reinterpret_cast<double*>(smem_raw + L.off_sig);
    double* st = reinterpret_cast<double*>(smem_raw + L.off_st);
    double2* K2 = reinterpret_cast<double2*>(smem_raw + L.off_k2);
    double2* W2 = reinterpret_cast<double2*>(smem_raw + L.off_w2);
    double* zbuf = reinterpret_cast<double*>(smem_raw + L.off_z);
    Mailbox* mail = reinterpret_cast<Mailbox*>(smem_raw + L.off_mail);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);

    const int lane = threadIdx.x & 31;
    const bool is_m = threadIdx.x < 32;  // warp 0 = matrix role, warp 1 = scalar role
    const long long b = blockIdx.x;
    if (b >= p.B) return;
    double* g_sig = p.sigma + b * (long long)p.sig_stride;
    double* g_st = p.state + b * (long long)p.st_stride;
    const uint32_t sig_bytes = (uint32_t)p.sig_stride * 8u, st_bytes = (uint32_t)p.st_stride * 8u;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, sig_bytes + st_bytes);
        bulk_g2s(sig, g_sig, sig_bytes, bar);
        bulk_g2s(st, g_st, st_bytes, bar);
    }
    __syncthreads();  // barrier initialised before anyone polls it
    const double* xy = p.xy + b * 2 * n;
    int init_flag = p.init_flag[b];
    // visible set (both warps walk the same mask)
    unsigned vis_mask = __ballot_sync(0xffffffffu, lane < n && p.vis[b * n + lane] != 0);

    if (!is_m) {
        // z = (range, bearing) of every slot's reading, lane-parallel (ekf_slam.cpp:140-146), while the copy is in flight
        if (lane < n) {
            double r, phi;
            range_bearing(xy[2 * lane], xy[2 * lane + 1], r, phi);
            zbuf[2 * lane] = r;
            zbuf[2 * lane + 1] = phi;
        }
    }
    mbar_wait(bar, 0);

    // ---- prediction (ekf_slam.cpp:55-106), matrix warp; the scalar warp only needs the new pose afterwards
    if (is_m && (p.mode & kDoPredict)) {
        const Motion mo = motion_model(st[0], p.twists[2 * b], p.twists[2 * b + 1]);
        __syncwarp();
        for (int c = lane; c < N; c += 32) {
            const double r0 = sig[c];
            sig[N + c] = fma(mo.a1, r0, sig[N + c]);
            sig[2 * N + c] = fma(mo.a2, r0, sig[2 * N + c]);
        }
        __syncwarp();
        for (int r = lane; r < N; r += 32) {
            const double c0 = sig[r * N];
            sig[r * N + 1] = fma(c0, mo.a1, sig[r * N + 1]);
            sig[r * N + 2] = fma(c0, mo.a2, sig[r * N + 2]);
        }
        __syncwarp();
        if (lane == 0) {
            sig[0] += kQ;
            sig[N + 1] += kQ;
            sig[2 * N + 2] += kQ;
            st[0] = st[0] + mo.u0;  // theta is not wrapped here (:99)
            st[1] = st[1] + mo.u1;
            st[2] = st[2] + mo.u2;
        }
    }
    __syncthreads();
    const double theta = st[0], x = st[1], y = st[2];  // entry-time pose, stale for later landmarks (:109-111)
    if (!init_flag) {  // first call: every slot from its reading (:113-128), scalar warp
        if (!is_m && lane < n) {
            double mx, my;
            landmark_from_reading(xy[2 * lane], xy[2 * lane + 1], theta, x, y, mx, my);
            st[3 + 2 * lane] = mx;
            st[4 + 2 * lane] = my;
        }
        init_flag = 1;
        __syncthreads();
    }

    unsigned long long n_corr = 0;
    unsigned rem = vis_mask;
    Hj h;
    if (!is_m && rem) {
        const int i0 = __ffs(rem) - 1;
        h = make_hj(st[3 + 2 * i0], st[4 + 2 * i0], theta, x, y);
    }
    while (rem) {
        const int i = __ffs(rem) - 1;
        rem &= rem - 1;
        const int i3 = 3 + 2 * i, i4 = i3 + 1;
        ++n_corr;
        if (!is_m) {
            // ------------------------------------------------------------ scalar role
            if (lane == 0) mail->h = h;
            pair_sync(1);  // A: H_j published
            pair_sync(2);  // B: W = H Sigma and P = Sigma H^T are in shared memory
            const double2 w0 = W2[0], w1 = W2[1], w2 = W2[2], w3 = W2[i3], w4 = W2[i4];
            const double s00 = h_row0(h, w1.x, w2.x, w3.x, w4.x) + kR;
            const double s01 = h_row1(h, w0.x, w1.x, w2.x, w3.x, w4.x);
            const double s10 = h_row0(h, w1.y, w2.y, w3.y, w4.y);
            const double s11 = h_row1(h, w0.y, w1.y, w2.y, w3.y, w4.y) + kR;
            const Sym2 si = inv2x2(s00, s01, s10, s11);
            const double nu0 = __dsub_rn(zbuf[2 * i], h.zr);
            const double nu1 = normalize_angle(__dsub_rn(zbuf[2 * i + 1], h.zphi));  // :182-183
            if (lane == 0) {
                mail->si = si;
                mail->nu0 = nu0;
                mail->nu1 = nu1;
            }
            pair_sync(3);  // C: S^-1 and the innovation published
            pair_sync(4);  // D: state updated
            if (rem) {     // H_j of the next landmark from the updated state: overlaps the rank-2 update of this one
                const int inext = __ffs(rem) - 1;
                h = make_hj(st[3 + 2 * inext], st[4 + 2 * inext], theta, x, y);
            }
        } else {
            // ------------------------------------------------------------ matrix role
            pair_sync(1);  // A
            const Hj hh = mail->h;
            for (int c = lane; c < N; c += 32) {
                const double s0 = sig[c], s1 = sig[N + c], s2 = sig[2 * N + c];
                const double s3 = sig[i3 * N + c], s4 = sig[i4 * N + c];
                W2[c] = make_double2(h_row0(hh, s1, s2, s3, s4), h_row1(hh, s0, s1, s2, s3, s4));
                const double* row = sig + c * N;
                const double r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[i3], r4 = row[i4];
                K2[c] = make_double2(h_row0(hh, r1, r2, r3, r4), h_row1(hh, r0, r1, r2, r3, r4));
            }
            pair_sync(2);  // B
            pair_sync(3);  // C
            const Sym2 si = mail->si;
            const double nu0 = mail->nu0, nu1 = mail->nu1;
            for (int r = lane; r < N; r += 32) {
                const double2 pp = K2[r];
                const double k0 = fma(pp.y, si.i10, pp.x * si.i00);
                const double k1 = fma(pp.y, si.i11, pp.x * si.i01);
                K2[r] = make_double2(k0, k1);
                st[r] = st[r] + fma(k1, nu1, k0 * nu0);
            }
            __syncwarp();
            if (lane == 0) st[0] = normalize_angle(st[0]);  // :187
            pair_sync(4);  // D
            // Sigma <- Sigma - K W (:191-192); 2 x 16 lane tiling, rows in batches (see ekf_fused.cuh)
            {
                const int g = lane >> 4, q = lane & 15;
                constexpr int CB = (N + 15) / 16;
                constexpr int RA = (N + 1) / 2;
                constexpr int RB = 6;
                double2 w[CB];
#pragma unroll
                for (int bb = 0; bb < CB; ++bb) w[bb] = (q + 16 * bb < N) ? W2[q + 16 * bb] : make_double2(0.0, 0.0);
#pragma unroll
                for (int a0 = 0; a0 < RA; a0 += RB) {
                    double2 k[RB];
                    double v[RB][CB];
#pragma unroll
                    for (int u = 0; u < RB; ++u) {
                        const int r = g + 2 * (a0 + u);
                        if (a0 + u < RA && r < N) {
                            k[u] = K2[r];
#pragma unroll
                            for (int bb = 0; bb < CB; ++bb)
                                if (q + 16 * bb < N) v[u][bb] = sig[r * N + q + 16 * bb];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < RB; ++u) {
                        const int r = g + 2 * (a0 + u);
                        if (a0 + u < RA && r < N) {
#pragma unroll
                            for (int bb = 0; bb < CB; ++bb)
                                if (q + 16 * bb < N)
                                    sig[r * N + q + 16 * bb] = fma(-k[u].y, w[bb].y, fma(-k[u].x, w[bb].x, v[u][bb]));
                        }
                    }
                }
            }
            __syncwarp();
        }
    }

    // ---- write back: smem -> HBM with the bulk copy engine
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_s2g(g_sig, sig, sig_bytes);
        bulk_s2g(g_st, st, st_bytes);
        bulk_commit();
        p.init_flag[b] = init_flag;
        if (p.n_updates && n_corr) atomicAdd(p.n_updates, n_corr);
        bulk_wait_all();
    }
}

}  // namespace ekf
