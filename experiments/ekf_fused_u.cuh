// Engine 1c (experimental A/B of ekf_fused_sym.cuh): the same batched step with the covariance update in its SYMMETRIC
// FACTOR form.  K W = W^T S^-1 W = U U^T with U = W^T L, S^-1 = L L^T (Cholesky of the 2 x 2 inverse, from two rsqrt
// and no division: for S = [a b; b c], L = [sqrt(c / det) 0; -b / sqrt(c det)  1 / sqrt(c)]).  One N x 2 array per
// correction instead of two (K and W): no K store, half the pending-factor loads, the orientation of a mirrored
// entry no longer matters, and the freed shared memory holds a third pending factor, so that up to THREE corrections
// share one pass over Sigma.  state += K nu = U (L^T nu).
// Layout, staging, innovation and association as in ekf_fused_sym.cuh.
#pragma once
#include "ekf_fused_sym.cuh"

#ifndef EKF_U_MAXP
#define EKF_U_MAXP 3  // corrections applied per pass (1..3)
#endif
#ifndef EKF_U_RB
#define EKF_U_RB 1  // row pairs per register batch of the pass
#endif
#ifndef EKF_U_MINB
#define EKF_U_MINB 16
#endif

namespace ekf {

struct USmem {
    int off_sig, off_st, off_u, off_z, off_bar, total;
    __host__ __device__ USmem(int n, int m_max) {
        const int N = 3 + 2 * n;
        int o = 0;
        off_sig = o;
        o += sym_sig_stride(N) * 8;
        off_st = o;
        o += sym_st_stride(N) * 8;
        off_u = o;  // U_j (N x 2) as double2 per row, j = 0 .. EKF_U_MAXP - 1
        o += EKF_U_MAXP * N * 16;
        o = (o + 15) & ~15;
        off_z = o;
        const int zc = 3 * (n > m_max ? n : m_max);
        o += zc * 8;
        off_bar = o;
        o += 16;
        total = o;
    }
};

// Gain of one correction: W = Hj Sigma from five rows (pending factors 0 .. npend-1 applied on the fly), S, then
// U = W^T L -> Ubuf[npend] and state += U (L^T nu).
template <int NL, class H>
__device__ __forceinline__ void u_warp_gain(const double* __restrict__ sig, double* __restrict__ st,
                                            double2* __restrict__ Ubuf, const int npend, const int N, const int lane,
                                            const int* __restrict__ cmv, const int i, const H h, const double nu0,
                                            const double nu1) {
    constexpr int NS = NL ? (3 + 2 * NL + 31) / 32 : 5;
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const int b3 = i3 & ~15, b4 = i4 & ~15;
    int row3 = __shfl_sync(0xffffffffu, cmv[0], i3 & 31), row4 = __shfl_sync(0xffffffffu, cmv[0], i4 & 31);
#pragma unroll
    for (int sl = 1; sl < NS; ++sl) {
        const int r3 = __shfl_sync(0xffffffffu, cmv[sl], i3 & 31), r4 = __shfl_sync(0xffffffffu, cmv[sl], i4 & 31);
        if ((i3 >> 5) == sl) row3 = r3;
        if ((i4 >> 5) == sl) row4 = r4;
    }
    double s[NS][5];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        if (c < N) {
            const int cm = cmv[sl];
            s[sl][0] = sig[c];
            s[sl][1] = sig[N + c];
            s[sl][2] = sig[2 * N + c];
            s[sl][3] = sig[c < b3 ? cm + i3 : row3 + c];
            s[sl][4] = sig[c < b4 ? cm + i4 : row4 + c];
        }
    }
#pragma unroll
    for (int j = 0; j < EKF_U_MAXP - 1; ++j) {
        if (j < npend) {
            const double2* __restrict__ Uj = Ubuf + j * N;
            const double2 u0 = Uj[0], u1 = Uj[1], u2 = Uj[2], u3 = Uj[i3], u4 = Uj[i4];
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                const int c = lane + 32 * sl;
                if (c < N) {
                    const double2 uc = Uj[c];
                    s[sl][0] = apply_pair(s[sl][0], u0, uc);
                    s[sl][1] = apply_pair(s[sl][1], u1, uc);
                    s[sl][2] = apply_pair(s[sl][2], u2, uc);
                    s[sl][3] = apply_pair(s[sl][3], u3, uc);
                    s[sl][4] = apply_pair(s[sl][4], u4, uc);
                }
            }
        }
    }
    double2* __restrict__ Uout = Ubuf + npend * N;
    double2 wreg[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        if (c < N) {
            h_rows(h, s[sl][0], s[sl][1], s[sl][2], s[sl][3], s[sl][4], wreg[sl].x, wreg[sl].y);
            Uout[c] = wreg[sl];  // W for now: the five columns S needs are read back below
        }
    }
    __syncwarp();
    const double2 w0 = Uout[0], w1 = Uout[1], w2 = Uout[2], w3 = Uout[i3], w4 = Uout[i4];
    __syncwarp();  // every lane has its copies before the slots are overwritten with U
    double s00, s01, s10, s11;
    h_rows(h, w0.x, w1.x, w2.x, w3.x, w4.x, s00, s01);
    h_rows(h, w0.y, w1.y, w2.y, w3.y, w4.y, s10, s11);
    // S = [a b; b c] (+R on the diagonal, :172-178); S^-1 = L L^T
    const double a = s00 + kR, bb = s10, c = s11 + kR;
    (void)s01;
    const double det = fma(a, c, -(bb * bb));
    const double rc = rsqrt(c), rd = rsqrt(det);
    const double l11 = rc, l00 = c * rc * rd, l10 = -(bb * rc) * rd;
    const double mu0 = fma(l10, nu1, l00 * nu0), mu1 = l11 * nu1;  // L^T nu
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int r = lane + 32 * sl;
        if (r < N) {
            const double2 w = wreg[sl];
            const double ux = fma(w.y, l10, w.x * l00), uy = w.y * l11;  // U = L^T w
            Uout[r] = make_double2(ux, uy);
            double ns = st[r] + fma(uy, mu1, ux * mu0);
            if (r == 0) ns = normalize_angle(ns);  // theta is wrapped after every correction (:187)
            st[r] = ns;
        }
    }
    __syncwarp();
}

// Sigma <- Sigma - sum_{j < NF} U_j U_j^T on the stored staircase, ONE pass; block row RBK.
template <int NC, int NF, int RBK>
__device__ __forceinline__ void u_pass_blockrow(double* __restrict__ sig, const double2* __restrict__ Ubuf,
                                                const double2 (*uc)[(NC + 15) / 16], const int g, const int q) {
    constexpr int NBK = (NC + 15) / 16;
    constexpr int CH = NBK - RBK;
    constexpr int ROWS = (NC - 16 * RBK) < 16 ? (NC - 16 * RBK) : 16;
    constexpr int SL = (ROWS + 1) / 2;
    constexpr int LROW = NC - 16 * RBK;
    constexpr int BASE = stair_row_base(16 * RBK, NC);
    constexpr int RB = EKF_U_RB;
#pragma unroll
    for (int a0 = 0; a0 < SL; a0 += RB) {
        double2 ur[RB][NF];
        double v[RB][CH];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int lr = g + 2 * (a0 + u);
            if (a0 + u < SL && lr < ROWS) {
#pragma unroll
                for (int j = 0; j < NF; ++j) ur[u][j] = Ubuf[j * NC + 16 * RBK + lr];
#pragma unroll
                for (int b = 0; b < CH; ++b)
                    if (16 * (RBK + b) + q < NC) v[u][b] = sig[BASE + lr * LROW + q + 16 * b];
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int lr = g + 2 * (a0 + u);
            if (a0 + u < SL && lr < ROWS) {
#pragma unroll
                for (int b = 0; b < CH; ++b)
                    if (16 * (RBK + b) + q < NC) {
                        double t = v[u][b];
#pragma unroll
                        for (int j = 0; j < NF; ++j) t = apply_pair(t, ur[u][j], uc[j][RBK + b]);
                        sig[BASE + lr * LROW + q + 16 * b] = t;
                    }
            }
        }
    }
}

template <int NC, int NF, int... RBK>
__device__ __forceinline__ void u_pass_rows(double* __restrict__ sig, const double2* __restrict__ Ubuf,
                                            const double2 (*uc)[(NC + 15) / 16], const int g, const int q,
                                            std::integer_sequence<int, RBK...>) {
    (u_pass_blockrow<NC, NF, RBK>(sig, Ubuf, uc, g, q), ...);
}

template <int NL, int NF>
__device__ __forceinline__ void u_warp_pass(double* __restrict__ sig, const double2* __restrict__ Ubuf, const int N,
                                            const int lane) {
    if (EKF_DEBUG_SKIP_RANK2) return;
    constexpr int NC = NL ? 3 + 2 * NL : 0;
    const int g = lane >> 4, q = lane & 15;
    if constexpr (NC != 0) {
        constexpr int NBK = (NC + 15) / 16;
        double2 uc[NF][NBK];
#pragma unroll
        for (int j = 0; j < NF; ++j)
#pragma unroll
            for (int b = 0; b < NBK; ++b)
                uc[j][b] = (q + 16 * b < NC) ? Ubuf[j * NC + q + 16 * b] : make_double2(0.0, 0.0);
        u_pass_rows<NC, NF>(sig, Ubuf, uc, g, q, std::make_integer_sequence<int, NBK>{});
    } else {
        const int nbk = (N + 15) >> 4;
        for (int rb = 0; rb < nbk; ++rb) {
            const int L = N - 16 * rb, rows = L < 16 ? L : 16, base = stair_row_base(16 * rb, N);
            for (int c0 = 0; c0 < L; c0 += 32) {
                double2 uc[NF][2];
#pragma unroll
                for (int j = 0; j < NF; ++j)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int c = 16 * rb + c0 + q + 16 * b;
                        uc[j][b] = c < N ? Ubuf[j * N + c] : make_double2(0.0, 0.0);
                    }
                for (int lr = g; lr < rows; lr += 2) {
                    double2 ur[NF];
#pragma unroll
                    for (int j = 0; j < NF; ++j) ur[j] = Ubuf[j * N + 16 * rb + lr];
                    double* row = sig + base + lr * L + c0 + q;
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        if (c0 + q + 16 * b < L) {
                            double t = row[16 * b];
#pragma unroll
                            for (int j = 0; j < NF; ++j) t = apply_pair(t, ur[j], uc[j][b]);
                            row[16 * b] = t;
                        }
                }
            }
        }
    }
    __syncwarp();
}

template <int NL>
__device__ __forceinline__ void u_apply_pending(double* __restrict__ sig, const double2* __restrict__ Ubuf,
                                                const int npend, const int N, const int lane) {
    switch (npend) {
        case 1: u_warp_pass<NL, 1>(sig, Ubuf, N, lane); break;
#if EKF_U_MAXP >= 2
        case 2: u_warp_pass<NL, 2>(sig, Ubuf, N, lane); break;
#endif
#if EKF_U_MAXP >= 3
        case 3: u_warp_pass<NL, 3>(sig, Ubuf, N, lane); break;
#endif
        default: break;
    }
}

template <int NL>
__global__ void __launch_bounds__(32, NL == 20 ? EKF_U_MINB : 1) ekf_fused_u_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = NL ? NL : p.n;
    const int N = 3 + 2 * n;
    const USmem L(n, p.m_max);
    double* sig = reinterpret_cast<double*>(smem_raw + L.off_sig);
    double* st = reinterpret_cast<double*>(smem_raw + L.off_st);
    double2* Ubuf = reinterpret_cast<double2*>(smem_raw + L.off_u);
    double* zbuf = reinterpret_cast<double*>(smem_raw + L.off_z);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);

    const int lane = threadIdx.x;
    const long long b = blockIdx.x;
    if (b >= p.B) return;
    double* g_sig = p.sigma + b * (long long)p.sig_stride;
    double* g_st = p.state + b * (long long)p.st_stride;
    const uint32_t sig_bytes = (uint32_t)p.sig_stride * 8u, st_bytes = (uint32_t)p.st_stride * 8u;

    constexpr int NS = NL ? (3 + 2 * NL + 31) / 32 : 5;
    int cmv[NS];  // mirror-row offsets of this lane's column slots (see sym_warp_gain)
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        cmv[sl] = stair_row_base(c, N) - (c & ~15);
    }

    // ---- stage Sigma (staircase) and the state into shared memory with the bulk copy engine
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
    // visible landmarks as warp-uniform bit masks (slots 0..31 and 32..63), readings as {zr, ux, uy} per slot
    unsigned vismask[2] = {0u, 0u};
    int sp_begin = 0, sp_count = 0;
    if ((p.mode & kDoMeasurement) && (p.mode & kSparseReadings)) {
        // marker list: p.mcount = CSR offsets [B + 1], p.vis = landmark ids, p.xy = (x, y) per listed marker
        sp_begin = p.mcount[b];
        sp_count = p.mcount[b + 1] - sp_begin;
        for (int k0 = 0; k0 < sp_count; k0 += 32) {
            const int k = k0 + lane;
            const bool on = k < sp_count;
            int id = 0;
            if (on) {
                id = p.vis[sp_begin + k];
                const Reading z = make_reading(p.xy[2 * (long long)(sp_begin + k)], p.xy[2 * (long long)(sp_begin + k) + 1]);
                if (id < n) {
                    zbuf[3 * id] = z.zr;
                    zbuf[3 * id + 1] = z.ux;
                    zbuf[3 * id + 2] = z.uy;
                }
            }
            vismask[0] |= __reduce_or_sync(0xffffffffu, (on && id < 32 && id < n) ? 1u << id : 0u);
            vismask[1] |= __reduce_or_sync(0xffffffffu, (on && id >= 32 && id < n) ? 1u << (id - 32) : 0u);
        }
    } else if (p.mode & kDoMeasurement) {
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            vismask[base >> 5] = __ballot_sync(0xffffffffu, i < n && p.vis[b * n + i] != 0);
        }
        // range and unit direction of every slot's reading, lane-parallel (ekf_slam.cpp:140-146)
        for (int i = lane; i < n; i += 32) {
            const Reading z = make_reading(p.xy[b * 2 * n + 2 * i], p.xy[b * 2 * n + 2 * i + 1]);
            zbuf[3 * i] = z.zr;
            zbuf[3 * i + 1] = z.ux;
            zbuf[3 * i + 2] = z.uy;
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
    // Rows 1, 2 are stored in full (block row 0) and take a * row 0; columns 1, 2 exist as such only inside the
    // first diagonal block (r < 16) -- beyond it they ARE rows 1, 2.
    double sth = 0.0, cth = 1.0;
    bool have_sincos = false;
    if (p.mode & kDoPredict) {
        const Motion mo = motion_model(st[0], dtheta, dxv);
        sth = mo.s_new, cth = mo.c_new, have_sincos = true;
        __syncwarp();
        for (int c = lane; c < N; c += 32) {
            const double r0 = sig[c];
            sig[N + c] = fma(mo.a1, r0, sig[N + c]);
            sig[2 * N + c] = fma(mo.a2, r0, sig[2 * N + c]);
        }
        __syncwarp();
        if (lane < 16 && lane < N) {
            const double c0 = sig[lane * N];
            sig[lane * N + 1] = fma(c0, mo.a1, sig[lane * N + 1]);
            sig[lane * N + 2] = fma(c0, mo.a2, sig[lane * N + 2]);
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
        __syncwarp();
    }

    unsigned long long n_corr = 0;

    // ---- measurement(): known association (ekf_slam.cpp:108-197)
    if (p.mode & kDoMeasurement) {
        const double theta = st[0], x = st[1], y = st[2];  // read once; stale for later i (:109-111)
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
                        landmark_from_reading_cold(p.xy[2 * (long long)(sp_begin + k)], p.xy[2 * (long long)(sp_begin + k) + 1],
                                              theta, x, y, mx, my);
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
        // Visible landmarks in ascending order (:132-136).  Each gain sees the factors before it as pending; after
        // EKF_U_MAXP of them (or the last one) ONE pass over Sigma applies them all.  H_j / nu of the next landmark are
        // evaluated right after the state update they depend on and before the pass.
        unsigned long long rem = vismask[0];
        if (NL == 0 || NL > 32) rem |= (unsigned long long)vismask[1] << 32;
        int npend = 0, cur = -1;
        if (rem) {
            cur = __ffsll((long long)rem) - 1;
            rem &= rem - 1;
        }
        Innov h;
        if (cur >= 0)
            h = make_innov(st[3 + 2 * cur], st[4 + 2 * cur], theta, sth, cth, x, y,
                           Reading{zbuf[3 * cur], zbuf[3 * cur + 1], zbuf[3 * cur + 2]});
        while (cur >= 0) {
            u_warp_gain<NL>(sig, st, Ubuf, npend, N, lane, cmv, cur, h, h.nu0, h.nu1);
            ++npend;
            ++n_corr;
            int nxt = -1;
            if (rem) {
                nxt = __ffsll((long long)rem) - 1;
                rem &= rem - 1;
                h = make_innov(st[3 + 2 * nxt], st[4 + 2 * nxt], theta, sth, cth, x, y,
                               Reading{zbuf[3 * nxt], zbuf[3 * nxt + 1], zbuf[3 * nxt + 2]});
            }
            if (npend == EKF_U_MAXP || nxt < 0) {
                u_apply_pending<NL>(sig, Ubuf, npend, N, lane);
                npend = 0;
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
            const double rob[6] = {sig[0], sig[1], sig[2], sig[N + 1], sig[N + 2], sig[2 * N + 2]};
            for (int i = lane; i < known_count; i += 32) {
                double d = sym_maha_distance(sig, N, i, rob, st[3 + 2 * i], st[4 + 2 * i], zr, zphi, theta, x, y);
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
                u_warp_gain<NL>(sig, st, Ubuf, 0, N, lane, cmv, min_idx, h, __dsub_rn(zr, h.zr),
                                normalize_angle(__dsub_rn(zphi, h.zphi)));  // :182-183
                u_warp_pass<NL, 1>(sig, Ubuf, N, lane);  // the next distances need the new Sigma
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

}  // namespace ekf
