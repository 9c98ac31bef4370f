// Engine 1b: the batched fused step with Sigma kept SYMMETRIC (packed upper triangle) and up to FOUR corrections
// applied per pass over it.
//
// Same step as ekf_fused.cuh (prediction + measurement() or data_association() of one filter per warp, Sigma
// resident in shared memory, staged by the bulk copy engine), but
//   * only the upper triangle of Sigma exists, in HBM and on chip: row r keeps the columns [r, N), rows packed back
//     to back (N = 43: 946 of 1,849 entries, 7.6 KB instead of 14.8 KB).  The gain needs five ROWS only
//     (K = (H Sigma)^T S^-1, no column reads), the rank-2 pass touches the stored entries only, and HBM moves half
//     the bytes per filter-step;
//   * the shared memory that frees holds four (K, W) factor pairs, so the (I - K H) Sigma of ekf_slam.cpp:191-192 is
//     applied for up to four corrections in ONE pass (delayed application: the five rows the next correction reads
//     are rebuilt from the stored Sigma and the pending factors with the FMAs the pass will apply).  The kernel is
//     bound by shared-memory bandwidth, and the pass over Sigma is most of that traffic;
//   * 13.4 KB of shared memory per filter -> 16 resident filters per SM (the register file's limit at 128/thread).
// Mathematically Sigma stays symmetric under both the prediction and the correction; the reference's dense products
// leave rounding-level asymmetry (1e-16 relative) that this layout does not carry, far inside the 1e-9
// parity tolerance.
//
// Restates rigid2d/src/ekf_slam.cpp:55-106, :108-197, :200-214, :217-276, :278-402 (as ekf_fused.cuh).
#pragma once
#include "ekf_fused.cuh"

// row pairs per register batch of the rank-2 pass.  Measured on B200 (65,536 x n = 20): anything that makes the kernel
// spill under the 128-register cap of 16 resident filters per SM costs 15-20 %.
#ifndef EKF_SYM_RB
#define EKF_SYM_RB 1
#endif
#ifndef EKF_SYM_MAXP
#define EKF_SYM_MAXP 4  // corrections applied per pass over Sigma (1..4)
#endif
#ifndef EKF_SYM_STASH
#define EKF_SYM_STASH 1
#endif
#ifndef EKF_SYM_FENCE
#define EKF_SYM_FENCE 1
#endif
#ifndef EKF_SYM_MINB
#define EKF_SYM_MINB 16  // resident filters per SM the n = 20 kernel is compiled for
#endif

namespace ekf {

constexpr int kSymMaxPending = EKF_SYM_MAXP;

// ---- packed upper triangle ----------------------------------------------------------------------------------------
// offset of the diagonal element (r, r) = start of the stored part of row r
__host__ __device__ constexpr int tri_row(int r, int N) { return r * N - (r * (r - 1)) / 2; }
__host__ __device__ constexpr int tri_size(int N) { return tri_row(N, N); }
// offset of element (r, c) for any r, c (the mirror entry below the diagonal)
__host__ __device__ __forceinline__ int tri_at(int r, int c, int N) {
    return c >= r ? tri_row(r, N) + (c - r) : tri_row(c, N) + (r - c);
}

// per-filter strides (doubles) of the batch's Sigma and state arrays: multiples of 16 B for the bulk copies
__host__ __device__ constexpr int sym_sig_stride(int N) { return (tri_size(N) + 1) & ~1; }
__host__ __device__ constexpr int sym_st_stride(int N) { return (N + 1) & ~1; }

// shared memory carve-up (bytes), identical on host and device
struct SymSmem {
    int off_sig, off_st, off_k, off_w, off_z, off_bar, off_stash, total;
    __host__ __device__ SymSmem(int n, int m_max) {
        const int N = 3 + 2 * n;
        int o = 0;
        off_sig = o;
        o += sym_sig_stride(N) * 8;
        off_st = o;
        o += sym_st_stride(N) * 8;
        off_k = o;  // K_j (N x 2) as double2 per row, j = 0..3
        o += kSymMaxPending * N * 16;
        off_w = o;  // W_j (2 x N) as double2 per column
        o += kSymMaxPending * N * 16;
        // data_association() uses one factor at a time: its (range, bearing) list lives in the unused factor slots
        // when it fits; the marker-list scratch of measurement() (3 doubles per slot) always fits in slots 0..1 of K
        off_z = off_k + N * 16;
        const int z_room = (kSymMaxPending - 1) * N * 16;
        if (2 * m_max * 8 > z_room) {
            off_z = o;
            o += 2 * m_max * 8;
        }
        off_bar = o;
        o += 16;
        off_stash = o;  // 11 doubles: the next correction's H_j / nu and the entry pose while the pass runs
        o += 96;        // n = 20: 13,536 B per CTA: 16 CTAs of 13,568 B (+1 KB reserve each) fit one SM's 228 KB
        total = o;
    }
};

// Gain of one correction from five ROWS of the symmetric Sigma: W = Hj Sigma -> W[npend], K = W^T S^-1 -> K[npend],
// state += K nu.  The stored Sigma still misses the factors 0..npend-1; the entries read here are rebuilt with the
// FMAs the pass will apply to them, in their stored orientation (row index <= column index).
// cmv[s] = tri_row(c) - c for the lane's column slot s (loop invariant, computed once per launch).
template <int NL, class H>
__device__ __forceinline__ void sym_warp_gain(const double* __restrict__ sig, double* __restrict__ st,
                                              double2* __restrict__ Kbuf, double2* __restrict__ Wbuf, const int npend,
                                              const int N, const int lane, const int* __restrict__ cmv, const int i,
                                              const H h, const double nu0, const double nu1) {
    constexpr int NS = NL ? (3 + 2 * NL + 31) / 32 : 5;  // column slots per lane (n <= 64 -> N <= 131)
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const int row1 = tri_row(1, N) - 1, row2 = tri_row(2, N) - 2;
    const int row3 = tri_row(i3, N) - i3, row4 = tri_row(i4, N) - i4;
    double2 wreg[NS];  // this lane's W columns stay in registers for the K loop
    double s[NS][5];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        if (c < N) {
            const int cm = cmv[sl];
            s[sl][0] = sig[c];
            s[sl][1] = sig[c < 1 ? cm + 1 : row1 + c];
            s[sl][2] = sig[c < 2 ? cm + 2 : row2 + c];
            s[sl][3] = sig[c < i3 ? cm + i3 : row3 + c];
            s[sl][4] = sig[c < i4 ? cm + i4 : row4 + c];
        }
    }
    for (int j = 0; j < npend; ++j) {
        const double2* __restrict__ Kj = Kbuf + j * N;
        const double2* __restrict__ Wj = Wbuf + j * N;
        const double2 k0 = Kj[0], k1 = Kj[1], k2 = Kj[2], k3 = Kj[i3], k4 = Kj[i4];
        const double2 w1 = Wj[1], w2 = Wj[2], w3 = Wj[i3], w4 = Wj[i4];
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const int c = lane + 32 * sl;
            if (c < N) {
                const double2 wc = Wj[c], kc = Kj[c];
                s[sl][0] = apply_pair(s[sl][0], k0, wc);
                s[sl][1] = apply_pair(s[sl][1], c < 1 ? kc : k1, c < 1 ? w1 : wc);
                s[sl][2] = apply_pair(s[sl][2], c < 2 ? kc : k2, c < 2 ? w2 : wc);
                s[sl][3] = apply_pair(s[sl][3], c < i3 ? kc : k3, c < i3 ? w3 : wc);
                s[sl][4] = apply_pair(s[sl][4], c < i4 ? kc : k4, c < i4 ? w4 : wc);
            }
        }
    }
    double2* __restrict__ Kout = Kbuf + npend * N;
    double2* __restrict__ Wout = Wbuf + npend * N;
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        if (c < N) {
            wreg[sl] = make_double2(h_row0(h, s[sl][1], s[sl][2], s[sl][3], s[sl][4]),
                                    h_row1(h, s[sl][0], s[sl][1], s[sl][2], s[sl][3], s[sl][4]));
            Wout[c] = wreg[sl];
        }
    }
    __syncwarp();
    // S = (Hj Sigma) Hj^T + R from W at the five columns; closed-form inverse
    const double2 w0 = Wout[0], w1 = Wout[1], w2 = Wout[2], w3 = Wout[i3], w4 = Wout[i4];
    const double s00 = h_row0(h, w1.x, w2.x, w3.x, w4.x) + kR;
    const double s01 = h_row1(h, w0.x, w1.x, w2.x, w3.x, w4.x);
    const double s10 = h_row0(h, w1.y, w2.y, w3.y, w4.y);
    const double s11 = h_row1(h, w0.y, w1.y, w2.y, w3.y, w4.y) + kR;
    const Sym2 si = inv2x2(s00, s01, s10, s11);
    // K = W^T S^-1; state += K nu
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int r = lane + 32 * sl;
        if (r < N) {
            const double2 p = wreg[sl];
            const double k0 = fma(p.y, si.i10, p.x * si.i00);
            const double k1 = fma(p.y, si.i11, p.x * si.i01);
            Kout[r] = make_double2(k0, k1);
            double ns = st[r] + fma(k1, nu1, k0 * nu0);
            if (r == 0) ns = normalize_angle(ns);  // theta is wrapped after every correction (:187)
            st[r] = ns;
        }
    }
    __syncwarp();
}

// Sigma <- Sigma - sum_{j < NF} K_j W_j on the stored triangle, ONE pass over shared memory; the factors are applied
// in order, two FMAs each, per element.  Lane (g, q) = (lane / 16, lane % 16) owns the rows of parity g and the
// columns q + 16 b: its W pairs stay in registers for the whole pass, each K pair is fetched once per row, and a
// half-warp always touches 16 consecutive doubles (conflict-free).
template <int NL, int NF>
__device__ __forceinline__ void sym_warp_rank2(double* __restrict__ sig, const double2* __restrict__ Kbuf,
                                               const double2* __restrict__ Wbuf, const int N, const int lane) {
    if (EKF_DEBUG_SKIP_RANK2) return;
    constexpr int NC = NL ? 3 + 2 * NL : 0;
    const int g = lane >> 4, q = lane & 15;
    if constexpr (NC != 0) {
        constexpr int NBK = (NC + 15) / 16;  // column chunks
        constexpr int RP = (NC + 1) / 2;     // row pairs
        constexpr int RB = EKF_SYM_RB;
        double2 w[NF][NBK];
#pragma unroll
        for (int j = 0; j < NF; ++j)
#pragma unroll
            for (int b = 0; b < NBK; ++b)
                w[j][b] = (q + 16 * b < NC) ? Wbuf[j * NC + q + 16 * b] : make_double2(0.0, 0.0);
#pragma unroll
        for (int a0 = 0; a0 < RP; a0 += RB) {
            double2 k[RB][NF];
            double v[RB][NBK];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int a = a0 + u;
                if (a < RP) {
                    const int r = 2 * a + g;
                    // element (r, c) sits at tri_row(r) + c - r = [tri_row(2a) - 2a] + g (NC - 2a - 1) + c
                    const int base = tri_row(2 * a, NC) - 2 * a + g * (NC - 2 * a - 1) + q;
                    if (r < NC) {
#pragma unroll
                        for (int j = 0; j < NF; ++j) k[u][j] = Kbuf[j * NC + r];
#pragma unroll
                        for (int b = (2 * a) >> 4; b < NBK; ++b) {
                            const int c = q + 16 * b;
                            if (c >= r && c < NC) v[u][b] = sig[base + 16 * b];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int a = a0 + u;
                if (a < RP) {
                    const int r = 2 * a + g;
                    const int base = tri_row(2 * a, NC) - 2 * a + g * (NC - 2 * a - 1) + q;
                    if (r < NC) {
#pragma unroll
                        for (int b = (2 * a) >> 4; b < NBK; ++b) {
                            const int c = q + 16 * b;
                            if (c >= r && c < NC) {
                                double t = v[u][b];
#pragma unroll
                                for (int j = 0; j < NF; ++j) t = apply_pair(t, k[u][j], w[j][b]);
                                sig[base + 16 * b] = t;
                            }
                        }
                    }
                }
            }
#if EKF_SYM_FENCE
            asm volatile("" ::: "memory");  // keep the scheduler from hoisting later batches' loads (register cap)
#endif
        }
    } else {
        for (int c0 = 0; c0 < N; c0 += 32) {  // two column chunks per sweep keep the generic path in registers
            double2 w[NF][2];
#pragma unroll
            for (int j = 0; j < NF; ++j)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int c = c0 + q + 16 * b;
                    w[j][b] = c < N ? Wbuf[j * N + c] : make_double2(0.0, 0.0);
                }
            const int r_end = (c0 + 32 < N) ? c0 + 32 : N;  // rows beyond the chunk's last column store nothing of it
            for (int r = g; r < r_end; r += 2) {
                double2 k[NF];
#pragma unroll
                for (int j = 0; j < NF; ++j) k[j] = Kbuf[j * N + r];
                double* row = sig + tri_row(r, N) - r;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int c = c0 + q + 16 * b;
                    if (c >= r && c < N) {
                        double t = row[c];
#pragma unroll
                        for (int j = 0; j < NF; ++j) t = apply_pair(t, k[j], w[j][b]);
                        row[c] = t;
                    }
                }
            }
        }
    }
    __syncwarp();
}

template <int NL>
__device__ __forceinline__ void sym_apply_pending(double* __restrict__ sig, const double2* __restrict__ Kbuf,
                                                  const double2* __restrict__ Wbuf, const int npend, const int N,
                                                  const int lane) {
    switch (npend) {
        case 1: sym_warp_rank2<NL, 1>(sig, Kbuf, Wbuf, N, lane); break;
#if EKF_SYM_MAXP >= 2
        case 2: sym_warp_rank2<NL, 2>(sig, Kbuf, Wbuf, N, lane); break;
#endif
#if EKF_SYM_MAXP >= 3
        case 3: sym_warp_rank2<NL, 3>(sig, Kbuf, Wbuf, N, lane); break;
#endif
#if EKF_SYM_MAXP >= 4
        case 4: sym_warp_rank2<NL, 4>(sig, Kbuf, Wbuf, N, lane); break;
#endif
        default: break;
    }
}

// Mahalanobis distance from the 5 x 5 block of the packed Sigma (ekf_slam.cpp:217-276; see maha_distance_rows).
__device__ __forceinline__ double sym_maha_distance(const double* __restrict__ sig, int N, int i, double mx, double my,
                                                    double zr, double zphi, double theta, double x, double y) {
    const Hj h = make_hj(mx, my, theta, x, y);
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    double wl0[5], wl1[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        const int c = l < 3 ? l : i3 + (l - 3);
        const double s0 = sig[tri_at(0, c, N)];
        const double s1 = sig[tri_at(1, c, N)];
        const double s2 = sig[tri_at(2, c, N)];
        const double s3 = sig[tri_at(i3, c, N)];
        const double s4 = sig[tri_at(i4, c, N)];
        wl0[l] = h_row0(h, s1, s2, s3, s4);
        wl1[l] = h_row1(h, s0, s1, s2, s3, s4);
    }
    const double p00 = h_row0(h, wl0[1], wl0[2], wl0[3], wl0[4]) + kR;
    const double p01 = h_row1(h, wl0[0], wl0[1], wl0[2], wl0[3], wl0[4]);
    const double p10 = h_row0(h, wl1[1], wl1[2], wl1[3], wl1[4]);
    const double p11 = h_row1(h, wl1[0], wl1[1], wl1[2], wl1[3], wl1[4]) + kR;
    const Sym2 pi = inv2x2(p00, p01, p10, p11);
    const double v0 = __dsub_rn(zr, h.zr), v1 = __dsub_rn(zphi, h.zphi);
    const double t0 = __dadd_rn(__dmul_rn(v0, pi.i00), __dmul_rn(v1, pi.i10));
    const double t1 = __dadd_rn(__dmul_rn(v0, pi.i01), __dmul_rn(v1, pi.i11));
    return __dadd_rn(__dmul_rn(t0, v0), __dmul_rn(t1, v1));
}

// first-call / new-landmark initialisation happens once per landmark: keep its atan2 + sincos out of the hot code
static __device__ __noinline__ void landmark_from_reading_cold(double sx, double sy, double theta, double x, double y,
                                                               double& mx, double& my) {
    landmark_from_reading(sx, sy, theta, x, y, mx, my);
}

// Reading of landmark slot i from the lane that holds it (slot i lives in lane i % 32, register group i / 32).
template <int NG>
__device__ __forceinline__ Reading sym_fetch_reading(const Reading (&zreg)[NG], const int i) {
    Reading z;
    const int src = i & 31;
    z.zr = __shfl_sync(0xffffffffu, zreg[0].zr, src);
    z.ux = __shfl_sync(0xffffffffu, zreg[0].ux, src);
    z.uy = __shfl_sync(0xffffffffu, zreg[0].uy, src);
#pragma unroll
    for (int gq = 1; gq < NG; ++gq) {
        const double za = __shfl_sync(0xffffffffu, zreg[gq].zr, src);
        const double zb = __shfl_sync(0xffffffffu, zreg[gq].ux, src);
        const double zc = __shfl_sync(0xffffffffu, zreg[gq].uy, src);
        if ((i >> 5) == gq) z.zr = za, z.ux = zb, z.uy = zc;
    }
    return z;
}

template <int NL>
__global__ void __launch_bounds__(32, NL == 20 ? EKF_SYM_MINB : 1) ekf_fused_sym_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = NL ? NL : p.n;
    const int N = 3 + 2 * n;
    const SymSmem L(n, p.m_max);
    double* sig = reinterpret_cast<double*>(smem_raw + L.off_sig);
    double* st = reinterpret_cast<double*>(smem_raw + L.off_st);
    double2* Kbuf = reinterpret_cast<double2*>(smem_raw + L.off_k);
    double2* Wbuf = reinterpret_cast<double2*>(smem_raw + L.off_w);
    double* zbuf = reinterpret_cast<double*>(smem_raw + L.off_z);     // association: (range, bearing) per measurement
    double* scratch = reinterpret_cast<double*>(smem_raw + L.off_k);  // marker list: Reading per slot, before any gain
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
    double* stash = reinterpret_cast<double*>(smem_raw + L.off_stash);

    const int lane = threadIdx.x;
    const long long b = blockIdx.x;
    if (b >= p.B) return;
    double* g_sig = p.sigma + b * (long long)p.sig_stride;
    double* g_st = p.state + b * (long long)p.st_stride;
    const uint32_t sig_bytes = (uint32_t)p.sig_stride * 8u, st_bytes = (uint32_t)p.st_stride * 8u;

    constexpr int NS = NL ? (3 + 2 * NL + 31) / 32 : 5;
    constexpr int NG = NL ? (NL + 31) / 32 : 2;  // groups of 32 landmark slots (n <= 64)
    int cmv[NS];  // tri_row(c) - c of this lane's column slots (see sym_warp_gain)
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        cmv[sl] = tri_row(c, N) - c;
    }

    // ---- stage Sigma (packed triangle) and the state into shared memory with the bulk copy engine
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, sig_bytes + st_bytes);
        bulk_g2s(sig, g_sig, sig_bytes, bar);
        bulk_g2s(st, g_st, st_bytes, bar);
    }
    // inputs that do not depend on the filter state are fetched while the copy is in flight
    double dtheta = 0.0, dxv = 0.0;
    if (p.mode & kDoPredict) {
        dtheta = p.twists[2 * b];
        dxv = p.twists[2 * b + 1];
    }
    int init_flag = p.init_flag[b];
    int m = 0;
    // measurement(): visible landmarks as warp-uniform bit masks (32 slots each); the reading of slot lane + 32 g
    // as {zr, ux, uy} in this lane's registers
    unsigned vismask[NG];
    Reading zreg[NG];
#pragma unroll
    for (int gq = 0; gq < NG; ++gq) {
        vismask[gq] = 0u;
        zreg[gq] = Reading{0.0, 1.0, 0.0};
    }
    int sp_begin = 0, sp_count = 0;
    if ((p.mode & kDoMeasurement) && (p.mode & kSparseReadings)) {
        // marker list: p.mcount = CSR offsets [B + 1], p.vis = landmark ids, p.xy = (x, y) per listed marker.
        // Listed readings are routed to their slot's lane through the (still unused) factor buffers.
        sp_begin = p.mcount[b];
        sp_count = p.mcount[b + 1] - sp_begin;
        for (int k0 = 0; k0 < sp_count; k0 += 32) {
            const int k = k0 + lane;
            int id = n;
            if (k < sp_count) {
                id = p.vis[sp_begin + k];
                if (id < n) {
                    const Reading z =
                        make_reading(p.xy[2 * (long long)(sp_begin + k)], p.xy[2 * (long long)(sp_begin + k) + 1]);
                    scratch[3 * id] = z.zr;
                    scratch[3 * id + 1] = z.ux;
                    scratch[3 * id + 2] = z.uy;
                }
            }
#pragma unroll
            for (int gq = 0; gq < NG; ++gq)
                vismask[gq] |= __reduce_or_sync(0xffffffffu, (id < n && (id >> 5) == gq) ? 1u << (id & 31) : 0u);
        }
        __syncwarp();
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
            const int i = lane + 32 * gq;
            if ((vismask[gq] >> lane) & 1u) zreg[gq] = Reading{scratch[3 * i], scratch[3 * i + 1], scratch[3 * i + 2]};
        }
        __syncwarp();
    } else if (p.mode & kDoMeasurement) {
        // range and unit direction of every slot's reading, lane-parallel (ekf_slam.cpp:140-146)
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
            const int i = lane + 32 * gq;
            const bool in = i < n;
            vismask[gq] = __ballot_sync(0xffffffffu, in && p.vis[b * n + i] != 0);
            if (in) zreg[gq] = make_reading(p.xy[b * 2 * n + 2 * i], p.xy[b * 2 * n + 2 * i + 1]);
        }
    } else if (p.mode & kDoAssociation) {
        m = p.mcount ? p.mcount[b] : p.m_max;
        m = m < p.m_max ? m : p.m_max;
        for (int j = lane; j < m; j += 32) {
            const double sx = p.xy[(b * p.m_max + j) * 2], sy = p.xy[(b * p.m_max + j) * 2 + 1];
            double r, phi;
            range_bearing(sx, sy, r, phi);
            zbuf[2 * j] = r;
            zbuf[2 * j + 1] = phi;
        }
    }
    __syncwarp();
    mbar_wait(bar, 0);

    // ---- prediction (ekf_slam.cpp:55-106): Sigma <- A Sigma A^T + Q with A = I + a1 e1 e0^T + a2 e2 e0^T.
    // Upper triangle: (1, c) += a1 (0, c) and (2, c) += a2 (0, c) for c >= 3; the 3 x 3 corner in closed form.
    double sth = 0.0, cth = 1.0;
    bool have_sincos = false;
    if (p.mode & kDoPredict) {
        const Motion mo = motion_model(st[0], dtheta, dxv);
        sth = mo.s_new, cth = mo.c_new, have_sincos = true;
        __syncwarp();
        const int row1 = tri_row(1, N) - 1, row2 = tri_row(2, N) - 2;
        for (int c = 3 + lane; c < N; c += 32) {
            const double r0 = sig[c];
            sig[row1 + c] = fma(mo.a1, r0, sig[row1 + c]);
            sig[row2 + c] = fma(mo.a2, r0, sig[row2 + c]);
        }
        if (lane == 0) {
            const double s00 = sig[0], s01 = sig[1], s02 = sig[2];
            const double s11 = sig[row1 + 1], s12 = sig[row1 + 2], s22 = sig[row2 + 2];
            // T = A Sigma (rows 1, 2), then Sigma' = T A^T (columns 1, 2), same FMA order as the dense strip update
            const double t10 = fma(mo.a1, s00, s01), t11 = fma(mo.a1, s01, s11), t12 = fma(mo.a1, s02, s12);
            const double t20 = fma(mo.a2, s00, s02), t22 = fma(mo.a2, s02, s22);
            sig[0] = s00 + kQ;
            sig[1] = fma(s00, mo.a1, s01);
            sig[2] = fma(s00, mo.a2, s02);
            sig[row1 + 1] = fma(t10, mo.a1, t11) + kQ;
            sig[row1 + 2] = fma(t10, mo.a2, t12);
            sig[row2 + 2] = fma(t20, mo.a2, t22) + kQ;
            st[0] = st[0] + mo.u0;  // theta is not wrapped here (:99)
            st[1] = st[1] + mo.u1;
            st[2] = st[2] + mo.u2;
        }
        __syncwarp();
    }

    unsigned long long n_corr = 0;

    // ---- measurement(): known association (ekf_slam.cpp:108-197)
    if (p.mode & kDoMeasurement) {
        double theta = st[0], x = st[1], y = st[2];  // read once; stale for later i (:109-111)
        if (!init_flag) {
            if (p.mode & kSparseReadings) {
                // unlisted slots read (0, 0): the landmark starts at the robot's position, as with dense zeros
                for (int i = lane; i < n; i += 32) {
                    st[3 + 2 * i] = x;
                    st[4 + 2 * i] = y;
                }
                __syncwarp();
                for (int k = lane; k < sp_count; k += 32) {
                    const int id = p.vis[sp_begin + k];
                    if (id < n) {
                        double mx, my;
                        landmark_from_reading_cold(p.xy[2 * (long long)(sp_begin + k)],
                                              p.xy[2 * (long long)(sp_begin + k) + 1], theta, x, y, mx, my);
                        st[3 + 2 * id] = mx;
                        st[4 + 2 * id] = my;
                    }
                }
            } else {
                for (int i = lane; i < n; i += 32) {
                    double mx, my;
                    landmark_from_reading_cold(p.xy[b * 2 * n + 2 * i], p.xy[b * 2 * n + 2 * i + 1], theta, x, y, mx, my);
                    st[3 + 2 * i] = mx;
                    st[4 + 2 * i] = my;
                }
            }
            init_flag = 1;
            __syncwarp();
        }
        if (!have_sincos) sincos(theta, &sth, &cth);
        // Visible landmarks go through in ascending order (:132-136).  Each gain sees the factors of the corrections
        // before it as pending; after four of them (or the last one) ONE pass over Sigma applies them all.  H_j / nu of
        // the next landmark are evaluated right after the state update they depend on and before the pass, so the
        // scalar chain overlaps the pass's shared-memory latency.
        unsigned long long rem = vismask[0];
        if (NG > 1) rem |= (unsigned long long)vismask[NG - 1] << 32;
        int npend = 0;
        int cur = -1;
        if (rem) {
            cur = __ffsll((long long)rem) - 1;
            rem &= rem - 1;
        }
        Innov h;
        if (cur >= 0)
            h = make_innov(st[3 + 2 * cur], st[4 + 2 * cur], theta, sth, cth, x, y, sym_fetch_reading<NG>(zreg, cur));
        while (cur >= 0) {
            sym_warp_gain<NL>(sig, st, Kbuf, Wbuf, npend, N, lane, cmv, cur, h, h.nu0, h.nu1);
            ++npend;
            ++n_corr;
            int nxt = -1;
            if (rem) {
                nxt = __ffsll((long long)rem) - 1;
                rem &= rem - 1;
                h = make_innov(st[3 + 2 * nxt], st[4 + 2 * nxt], theta, sth, cth, x, y, sym_fetch_reading<NG>(zreg, nxt));
            }
            if (npend == kSymMaxPending || nxt < 0) {
#if EKF_SYM_STASH
                // The pass wants every register for its W / K pairs: park the scalars that live across it.
                if (nxt >= 0) {
                    if (lane == 0) {
                        stash[0] = h.a, stash[1] = h.b, stash[2] = h.e, stash[3] = h.f, stash[4] = h.nu0, stash[5] = h.nu1;
                        stash[6] = theta, stash[7] = x, stash[8] = y, stash[9] = sth, stash[10] = cth;
                    }
                    __syncwarp();
                }
#endif
                sym_apply_pending<NL>(sig, Kbuf, Wbuf, npend, N, lane);
                npend = 0;
#if EKF_SYM_STASH
                if (nxt >= 0) {
                    h.a = stash[0], h.b = stash[1], h.e = stash[2], h.f = stash[3], h.nu0 = stash[4], h.nu1 = stash[5];
                    theta = stash[6], x = stash[7], y = stash[8], sth = stash[9], cth = stash[10];
                }
#endif
            }
            cur = nxt;
        }
    }

    // ---- data_association(): Mahalanobis nearest neighbour + landmark initialisation (ekf_slam.cpp:278-402)
    if (p.mode & kDoAssociation) {
        uint8_t* known = p.known + b * n;
        int known_count = 0;  // leading-true prefix (:281-288)
        for (int base = 0; base < n; base += 32) {
            const int i_l = base + lane;
            const unsigned ones = __ballot_sync(0xffffffffu, i_l < n && known[i_l] != 0);
            const int lead = __ffs(~ones) - 1;  // number of leading ones in this group of 32 (32 -> -1)
            if (ones == 0xffffffffu) {
                known_count += 32;
                continue;
            }
            known_count += lead;
            break;
        }
        if (known_count > n) known_count = n;
        const int known_count0 = known_count;
        for (int j = 0; j < m; ++j) {
            const double zr = zbuf[2 * j], zphi = zbuf[2 * j + 1];
            const double theta = st[0], x = st[1], y = st[2];  // live pose (:219-221)
            double best = INFINITY, second = INFINITY;
            int best_i = 0x7fffffff;
            for (int i = lane; i < known_count; i += 32) {
                double d = sym_maha_distance(sig, N, i, st[3 + 2 * i], st[4 + 2 * i], zr, zphi, theta, x, y);
                if (!(d == d)) d = INFINITY;  // NaN never wins
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
                if (better(ob, oi, best, best_i)) {
                    second = fmin(best, os);
                    best = ob;
                    best_i = oi;
                } else {
                    second = fmin(second, ob);
                }
            }
            double min_d = kGateNew;
            int min_idx = known_count;
            if (best < kGateNew) {  // d < min_maha_dis, :305
                min_d = best;
                min_idx = best_i;
                second = fmin(second, kGateNew);
            } else {
                second = best;
            }
            const long long o = b * p.m_max + j;
            if (lane == 0) {
                if (p.dmin_out) p.dmin_out[o] = min_d;
                if (p.second_out) p.second_out[o] = second;
            }
            int created = 0;
            if (min_idx == known_count && min_idx < n) {  // :318-327
                if (lane == 0) {
                    double mx, my;
                    landmark_from_reading_cold(p.xy[o * 2], p.xy[o * 2 + 1], theta, x, y, mx, my);
                    st[3 + 2 * min_idx] = mx;
                    st[4 + 2 * min_idx] = my;
                }
                __syncwarp();
                known_count++;
                min_d = 0.0;
                created = 1;
            }
            int assoc = -1;
            if (min_d < kGateUpdate) {  // :330
                const double th_l = st[0], x_l = st[1], y_l = st[2];  // live pose (:331-333)
                const Hj h = make_hj(st[3 + 2 * min_idx], st[4 + 2 * min_idx], th_l, x_l, y_l);
                sym_warp_gain<NL>(sig, st, Kbuf, Wbuf, 0, N, lane, cmv, min_idx, h, __dsub_rn(zr, h.zr),
                                  normalize_angle(__dsub_rn(zphi, h.zphi)));  // :182-183
                sym_warp_rank2<NL, 1>(sig, Kbuf, Wbuf, N, lane);  // the next distances need the new Sigma
                ++n_corr;
                assoc = min_idx;
            }
            if (lane == 0) {
                if (p.assoc_out) p.assoc_out[o] = assoc;
                if (p.created_out) p.created_out[o] = (uint8_t)created;
            }
        }
        for (int i = known_count0 + lane; i < known_count; i += 32) known[i] = 1;
        if (lane == 0) {  // outputs beyond the valid count are defined too
            for (int j = m; j < p.m_max; ++j) {
                const long long o = b * p.m_max + j;
                if (p.assoc_out) p.assoc_out[o] = -1;
                if (p.created_out) p.created_out[o] = 0;
                if (p.dmin_out) p.dmin_out[o] = kGateNew;
                if (p.second_out) p.second_out[o] = INFINITY;
            }
        }
    }

    // ---- write back: smem -> HBM with the bulk copy engine
    __syncwarp();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        bulk_s2g(g_sig, sig, sig_bytes);
        bulk_s2g(g_st, st, st_bytes);
        bulk_commit();
        p.init_flag[b] = init_flag;
        if (p.n_updates && n_corr) atomicAdd(p.n_updates, n_corr);
        bulk_wait_read();  // shared memory must outlive the copy engine's reads; the writes drain with the grid
    }
}

// Sigma_0 = blockdiag(0_3, 100 I) in the packed layout (ekf_slam.cpp:36-47), zero state, init flag cleared.
__global__ void k_fused_sym_init(double* sigma, double* state, int32_t* init_flag, long long B, int N, int sig_stride,
                                 int st_stride) {
    const long long b = blockIdx.x;
    if (b >= B) return;
    double* s = sigma + b * (long long)sig_stride;
    for (int e = threadIdx.x; e < sig_stride; e += blockDim.x) s[e] = 0.0;
    __syncthreads();
    for (int r = 3 + threadIdx.x; r < N; r += blockDim.x) s[tri_row(r, N)] = kSigma0;
    for (int e = threadIdx.x; e < st_stride; e += blockDim.x) state[b * (long long)st_stride + e] = 0.0;
    if (threadIdx.x == 0) init_flag[b] = 0;
}

// One filter's packed Sigma -> dense row-major N x N (ld = N).
__global__ void k_fused_sym_unpack(const double* __restrict__ s, double* __restrict__ out, int N) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N * N; e += gridDim.x * blockDim.x) {
        const int r = e / N, c = e - r * N;
        out[e] = s[tri_at(r, c, N)];
    }
}

}  // namespace ekf
