// Engine 1b: the batched fused step with Sigma kept SYMMETRIC in a block-staircase layout.
//
// Same step as ekf_fused.cuh (prediction + measurement() or data_association() of one filter per warp, Sigma
// resident in shared memory, staged by the bulk copy engine), but only the block-upper part of Sigma exists, in HBM
// and on chip: row r keeps the columns [16*floor(r/16), N).  For N = 43 that is 1,241 of 1,849 entries, which
//   * cuts the shared memory per filter from 18.3 KB to 13.5 KB -> 16 resident filters per SM instead of 12,
//   * cuts the rank-2 pass (the (I - K H) Sigma of ekf_slam.cpp:191-192) to the stored entries (46 instead of 66
//     half-warp tiles per factor) and the gain to five row reads (K = (H Sigma)^T S^-1, no column reads),
//   * cuts the HBM traffic per filter-step by the same third.
// Every row stride (N - 16 rb) is odd, so row reads, transposed reads and the 2 x 16 tiles are bank-conflict free.
// Mathematically Sigma stays symmetric under both the prediction and the correction; the reference's dense
// products leave rounding-level asymmetry (1e-16 relative) that this layout does not carry, far inside the 1e-9
// parity tolerance.  Entries below the diagonal inside a diagonal block are stored and updated like any other.
//
// Restates rigid2d/src/ekf_slam.cpp:55-106, :108-197, :200-214, :217-276, :278-402 (as ekf_fused.cuh).
#pragma once
#include <utility>

#include "ekf_fused.cuh"

// rows per register batch of the rank-2 pass, by column chunks of the block row (3 / 2 / 1) and factors per pass.
// Measured on B200 (65,536 x n = 20): anything that makes the kernel spill under the 128-register cap of 16 resident
// filters per SM costs 15-20 %; batches of 2 rows (126 registers, no spills) are the fastest setting.
#ifndef EKF_SYM_RB3_1
#define EKF_SYM_RB3_1 2
#endif
#ifndef EKF_SYM_RB3_2
#define EKF_SYM_RB3_2 2
#endif
#ifndef EKF_SYM_RB2_1
#define EKF_SYM_RB2_1 2
#endif
#ifndef EKF_SYM_RB2_2
#define EKF_SYM_RB2_2 2
#endif
#ifndef EKF_SYM_RB1
#define EKF_SYM_RB1 2
#endif
#ifndef EKF_SYM_MINB
#define EKF_SYM_MINB 16  // resident filters per SM the n = 20 kernel is compiled for
#endif

namespace ekf {

// ---- staircase layout -------------------------------------------------------------------------------------------
// offset of element (r, 16*floor(r/16)) = start of the stored part of row r
__host__ __device__ constexpr int stair_row_base(int r, int N) {
    const int rb = r >> 4;
    return 16 * rb * N - 128 * rb * (rb - 1) + (r - 16 * rb) * (N - 16 * rb);
}
__host__ __device__ constexpr int stair_size(int N) { return stair_row_base(N, N); }
// offset of element (r, c) for any r, c (the mirror entry is used below the stored part)
__host__ __device__ __forceinline__ int stair_at(int r, int c, int N) {
    const bool t = c < (r & ~15);
    const int rr = t ? c : r, cc = t ? r : c;
    return stair_row_base(rr, N) + cc - (rr & ~15);
}

// per-filter strides (doubles) of the batch's Sigma and state arrays: multiples of 16 B for the bulk copies
__host__ __device__ constexpr int sym_sig_stride(int N) { return (stair_size(N) + 1) & ~1; }
__host__ __device__ constexpr int sym_st_stride(int N) { return (N + 1) & ~1; }

struct SymSmem {
    int off_sig, off_st, off_k2, off_w2, off_k2b, off_w2b, off_z, off_bar, total;
    __host__ __device__ SymSmem(int n, int m_max) {
        const int N = 3 + 2 * n;
        int o = 0;
        off_sig = o;
        o += sym_sig_stride(N) * 8;
        off_st = o;
        o += sym_st_stride(N) * 8;
        off_k2 = o;
        o += N * 16;
        off_w2 = o;
        o += N * 16;
        off_k2b = o;
        o += N * 16;
        off_w2b = o;
        o += N * 16;
        o = (o + 15) & ~15;
        off_z = o;
        const int zc = 3 * (n > m_max ? n : m_max);  // Reading {zr, ux, uy} per slot
        o += zc * 8;  // n = 20: 13,536 B per CTA: 16 CTAs of 13,568 B (+1 KB reserve each) fit one SM's 228 KB
        off_bar = o;
        o += 16;
        total = o;
    }
};

// Gain of one correction from five ROWS of the symmetric Sigma: W = Hj Sigma -> Wout, K = W^T S^-1 -> Kout,
// state += K nu.  With PEND the stored Sigma still misses the previous correction's factor (Kpend, Wpend); the
// entries read here are rebuilt with the FMAs the rank-2 pass will apply to them (in their stored orientation).
// cmv[s] = offset of the mirror row of the lane's column slot s (loop invariant, computed once per launch).
template <bool PEND, int NL, class H>
__device__ __forceinline__ void sym_warp_gain(const double* __restrict__ sig, double* __restrict__ st,
                                              const double2* __restrict__ Kpend, const double2* __restrict__ Wpend,
                                              double2* __restrict__ Kout, double2* __restrict__ Wout, const int N,
                                              const int lane, const int* __restrict__ cmv, const int i, const H h,
                                              const double nu0, const double nu1) {
    constexpr int NS = NL ? (3 + 2 * NL + 31) / 32 : 5;  // column slots per lane (n <= 64 -> N <= 131)
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const int b3 = i3 & ~15, b4 = i4 & ~15;
    // row offsets of the landmark's two rows: lane r % 32 already holds stair_row_base(r) - (r & ~15) in cmv[r / 32]
    int row3 = __shfl_sync(0xffffffffu, cmv[0], i3 & 31), row4 = __shfl_sync(0xffffffffu, cmv[0], i4 & 31);
#pragma unroll
    for (int sl = 1; sl < NS; ++sl) {
        const int r3 = __shfl_sync(0xffffffffu, cmv[sl], i3 & 31), r4 = __shfl_sync(0xffffffffu, cmv[sl], i4 & 31);
        if ((i3 >> 5) == sl) row3 = r3;
        if ((i4 >> 5) == sl) row4 = r4;
    }
    double2 ka[5], wa3, wa4;
    if (PEND) {
        ka[0] = Kpend[0], ka[1] = Kpend[1], ka[2] = Kpend[2], ka[3] = Kpend[i3], ka[4] = Kpend[i4];
        wa3 = Wpend[i3], wa4 = Wpend[i4];
    }
    double2 wreg[NS];  // this lane's W columns stay in registers for the K loop
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int c = lane + 32 * sl;
        // compile-time N: a lane beyond the last column reads neighbouring shared memory and stores nothing (no branch)
        if (NL != 0 || c < N) {
            const int cm = cmv[sl];
            const bool t3 = c < b3, t4 = c < b4;
            double s0 = sig[c], s1 = sig[N + c], s2 = sig[2 * N + c];
            double s3 = sig[t3 ? cm + i3 : row3 + c];
            double s4 = sig[t4 ? cm + i4 : row4 + c];
            if (PEND) {
                const double2 wc = Wpend[c], kc = Kpend[c];
                s0 = apply_pair(s0, ka[0], wc);
                s1 = apply_pair(s1, ka[1], wc);
                s2 = apply_pair(s2, ka[2], wc);
                s3 = apply_pair(s3, t3 ? kc : ka[3], t3 ? wa3 : wc);
                s4 = apply_pair(s4, t4 ? kc : ka[4], t4 ? wa4 : wc);
            }
            h_rows(h, s0, s1, s2, s3, s4, wreg[sl].x, wreg[sl].y);
            if (c < N) Wout[c] = wreg[sl];
        }
    }
    __syncwarp();
    // S = (Hj Sigma) Hj^T + R from W at the five columns; closed-form inverse
    const double2 w0 = Wout[0], w1 = Wout[1], w2 = Wout[2], w3 = Wout[i3], w4 = Wout[i4];
    double s00, s01, s10, s11;
    h_rows(h, w0.x, w1.x, w2.x, w3.x, w4.x, s00, s01);
    h_rows(h, w0.y, w1.y, w2.y, w3.y, w4.y, s10, s11);
    const Sym2 si = inv2x2(s00 + kR, s01, s10, s11 + kR);
    // K = W^T S^-1 folded with nu: state += K nu
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int r = lane + 32 * sl;
        if (NL != 0 || r < N) {
            const double2 p = wreg[sl];
            const double k0 = fma(p.y, si.i10, p.x * si.i00);
            const double k1 = fma(p.y, si.i11, p.x * si.i01);
            double ns = st[r] + fma(k1, nu1, k0 * nu0);
            if (r == 0) ns = normalize_angle(ns);  // theta is wrapped after every correction (:187)
            if (r < N) {
                Kout[r] = make_double2(k0, k1);
                st[r] = ns;
            }
        }
    }
    __syncwarp();
}

// One block row (rows 16 RBK .. 16 RBK + 15, columns from 16 RBK on) of Sigma <- Sigma - Ka Wa [- Kb Wb].
// Lane (g, q) = (lane / 16, lane % 16) owns rows of parity g and the columns q + 16 b.
template <int NC, int NF, int RBK>
__device__ __forceinline__ void sym_rank2_blockrow(double* __restrict__ sig, const double2* __restrict__ Ka,
                                                   const double2* __restrict__ Kb, const double2* wa,
                                                   const double2* wb, const int g, const int q) {
    constexpr int NBK = (NC + 15) / 16;
    constexpr int CH = NBK - RBK;                                    // column chunks of this block row
    constexpr int ROWS = (NC - 16 * RBK) < 16 ? (NC - 16 * RBK) : 16;  // rows of this block row
    constexpr int SL = (ROWS + 1) / 2;                               // row slots per lane group
    constexpr int LROW = NC - 16 * RBK;                              // row stride (odd)
    constexpr int BASE = stair_row_base(16 * RBK, NC);
    constexpr int RB = (CH >= 3) ? (NF == 1 ? EKF_SYM_RB3_1 : EKF_SYM_RB3_2)
                                 : (CH == 2 ? (NF == 1 ? EKF_SYM_RB2_1 : EKF_SYM_RB2_2) : EKF_SYM_RB1);
    // Loads are unconditional: a lane beyond the row's last column or in the missing partner row of an odd block row
    // reads a neighbouring entry of the same shared-memory allocation and its result is simply not stored.  That
    // keeps the ragged edges (11 of 16 lanes in the last chunk, one row short in the last block row) free of
    // branches; only the stores are predicated.
#pragma unroll
    for (int a0 = 0; a0 < SL; a0 += RB) {
        double2 ka[RB], kb[RB];
        double v[RB][CH];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int lr = g + 2 * (a0 + u);
            if (a0 + u < SL) {
                ka[u] = Ka[16 * RBK + lr];
                if (NF == 2) kb[u] = Kb[16 * RBK + lr];
#pragma unroll
                for (int b = 0; b < CH; ++b) v[u][b] = sig[BASE + lr * LROW + q + 16 * b];
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int lr = g + 2 * (a0 + u);
            if (a0 + u < SL) {
#pragma unroll
                for (int b = 0; b < CH; ++b) {
                    double t = apply_pair(v[u][b], ka[u], wa[RBK + b]);
                    if (NF == 2) t = apply_pair(t, kb[u], wb[RBK + b]);
                    if (lr < ROWS && 16 * (RBK + b) + q < NC) sig[BASE + lr * LROW + q + 16 * b] = t;
                }
            }
        }
    }
}

template <int NC, int NF, int... RBK>
__device__ __forceinline__ void sym_rank2_rows(double* __restrict__ sig, const double2* __restrict__ Ka,
                                               const double2* __restrict__ Kb, const double2* wa, const double2* wb,
                                               const int g, const int q, std::integer_sequence<int, RBK...>) {
    (sym_rank2_blockrow<NC, NF, RBK>(sig, Ka, Kb, wa, wb, g, q), ...);
}

template <int NL, int NF>
__device__ __forceinline__ void sym_warp_rank2(double* __restrict__ sig, const double2* __restrict__ Ka,
                                               const double2* __restrict__ Wa, const double2* __restrict__ Kb,
                                               const double2* __restrict__ Wb, const int N, const int lane) {
    if (EKF_DEBUG_SKIP_RANK2) return;
    constexpr int NC = NL ? 3 + 2 * NL : 0;
    const int g = lane >> 4, q = lane & 15;
    if constexpr (NC != 0) {
        constexpr int NBK = (NC + 15) / 16;
        double2 wa[NBK], wb[NBK];
#pragma unroll
        for (int b = 0; b < NBK; ++b) {
            wa[b] = (q + 16 * b < NC) ? Wa[q + 16 * b] : make_double2(0.0, 0.0);
            wb[b] = (NF == 2 && q + 16 * b < NC) ? Wb[q + 16 * b] : make_double2(0.0, 0.0);
        }
        sym_rank2_rows<NC, NF>(sig, Ka, Kb, wa, wb, g, q, std::make_integer_sequence<int, NBK>{});
    } else {
        const int nbk = (N + 15) >> 4;
        for (int rb = 0; rb < nbk; ++rb) {
            const int L = N - 16 * rb, rows = L < 16 ? L : 16, base = stair_row_base(16 * rb, N);
            for (int c0 = 0; c0 < L; c0 += 64) {  // four column slots per pass keep the generic path in registers
                double2 wa[4], wb[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int c = 16 * rb + c0 + q + 16 * b;
                    wa[b] = c < N ? Wa[c] : make_double2(0.0, 0.0);
                    wb[b] = (NF == 2 && c < N) ? Wb[c] : make_double2(0.0, 0.0);
                }
                for (int lr = g; lr < rows; lr += 2) {
                    const double2 ka = Ka[16 * rb + lr];
                    double2 kb = make_double2(0.0, 0.0);
                    if (NF == 2) kb = Kb[16 * rb + lr];
                    double* row = sig + base + lr * L + c0 + q;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (c0 + q + 16 * b < L) {
                            double t = apply_pair(row[16 * b], ka, wa[b]);
                            if (NF == 2) t = apply_pair(t, kb, wb[b]);
                            row[16 * b] = t;
                        }
                }
            }
        }
    }
    __syncwarp();
}

// Mahalanobis distance from the 5 x 5 block Sigma[idx, idx], idx = {0, 1, 2, 3+2i, 4+2i} (ekf_slam.cpp:217-276; see
// maha_distance_rows).  The block is symmetric: 15 distinct entries, of which the robot's 3 x 3 corner `rob`
// (00 01 02 11 12 22) is the same for every landmark and is fetched once per measurement by the caller.
__device__ __forceinline__ double sym_maha_distance(const double* __restrict__ sig, int N, int i, const double* rob,
                                                    double mx, double my, double zr, double zphi, double theta,
                                                    double x, double y) {
    const Hj h = make_hj(mx, my, theta, x, y);
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const int rb3 = stair_row_base(i3, N) - (i3 & ~15), rb4 = stair_row_base(i4, N) - (i4 & ~15);
    const double c03 = sig[i3], c04 = sig[i4], c13 = sig[N + i3], c14 = sig[N + i4], c23 = sig[2 * N + i3],
                 c24 = sig[2 * N + i4];
    const double l33 = sig[rb3 + i3], l34 = sig[rb3 + i4], l44 = sig[rb4 + i4];
    const double blk[5][5] = {{rob[0], rob[1], rob[2], c03, c04},
                              {rob[1], rob[3], rob[4], c13, c14},
                              {rob[2], rob[4], rob[5], c23, c24},
                              {c03, c13, c23, l33, l34},
                              {c04, c14, c24, l34, l44}};
    double wl0[5], wl1[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) h_rows(h, blk[0][l], blk[1][l], blk[2][l], blk[3][l], blk[4][l], wl0[l], wl1[l]);
    double p00, p01, p10, p11;
    h_rows(h, wl0[0], wl0[1], wl0[2], wl0[3], wl0[4], p00, p01);
    h_rows(h, wl1[0], wl1[1], wl1[2], wl1[3], wl1[4], p10, p11);
    const Sym2 pi = inv2x2(p00 + kR, p01, p10, p11 + kR);
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

template <int NL>
__global__ void __launch_bounds__(32, NL == 20 ? EKF_SYM_MINB : 1) ekf_fused_sym_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = NL ? NL : p.n;
    const int N = 3 + 2 * n;
    const SymSmem L(n, p.m_max);
    double* sig = reinterpret_cast<double*>(smem_raw + L.off_sig);
    double* st = reinterpret_cast<double*>(smem_raw + L.off_st);
    double2* K2 = reinterpret_cast<double2*>(smem_raw + L.off_k2);
    double2* W2 = reinterpret_cast<double2*>(smem_raw + L.off_w2);
    double2* K2b = reinterpret_cast<double2*>(smem_raw + L.off_k2b);
    double2* W2b = reinterpret_cast<double2*>(smem_raw + L.off_w2b);
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
        // malformed offsets must not turn into out-of-bounds reads: such a filter gets an empty list
        if (sp_begin < 0 || sp_count < 0 || (long long)sp_begin + sp_count > p.sparse_total) sp_count = 0;
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
        for (int base = 0; base < n; base += 32) {
            unsigned rem = vismask[base >> 5];
            // Visible landmarks go through in PAIRS: both gains first (the second sees the first one's factor as
            // pending), then ONE pass over Sigma applies both rank-2 updates.  H_j / nu of the next landmark are
            // evaluated right after the state update they depend on and before the pass, so the scalar chain
            // overlaps the pass's shared-memory latency.
            Innov h;
            if (rem) {
                const int i0 = base + __ffs(rem) - 1;
                h = make_innov(st[3 + 2 * i0], st[4 + 2 * i0], theta, sth, cth, x, y,
                               Reading{zbuf[3 * i0], zbuf[3 * i0 + 1], zbuf[3 * i0 + 2]});
            }
            bool second = false;
            while (rem) {
                const int ic = base + __ffs(rem) - 1;
                rem &= rem - 1;
                if (!second)
                    sym_warp_gain<false, NL>(sig, st, nullptr, nullptr, K2, W2, N, lane, cmv, ic, h, h.nu0, h.nu1);
                else
                    sym_warp_gain<true, NL>(sig, st, K2, W2, K2b, W2b, N, lane, cmv, ic, h, h.nu0, h.nu1);
                ++n_corr;
                if (rem) {
                    const int in = base + __ffs(rem) - 1;
                    h = make_innov(st[3 + 2 * in], st[4 + 2 * in], theta, sth, cth, x, y,
                                   Reading{zbuf[3 * in], zbuf[3 * in + 1], zbuf[3 * in + 2]});
                }
                if (second)
                    sym_warp_rank2<NL, 2>(sig, K2, W2, K2b, W2b, N, lane);
                else if (!rem)
                    sym_warp_rank2<NL, 1>(sig, K2, W2, nullptr, nullptr, N, lane);
                second = !second;
            }
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
                sym_warp_gain<false, NL>(sig, st, nullptr, nullptr, K2, W2, N, lane, cmv, min_idx, h,
                                         __dsub_rn(zr, h.zr), normalize_angle(__dsub_rn(zphi, h.zphi)));  // :182-183
                sym_warp_rank2<NL, 1>(sig, K2, W2, nullptr, nullptr, N, lane);  // the next distances need the new Sigma
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

// Sigma_0 = blockdiag(0_3, 100 I) in the staircase layout (ekf_slam.cpp:36-47), zero state, init flag cleared.
__global__ void k_fused_sym_init(double* sigma, double* state, int32_t* init_flag, long long B, int N, int sig_stride,
                                 int st_stride) {
    const long long b = blockIdx.x;
    if (b >= B) return;
    double* s = sigma + b * (long long)sig_stride;
    for (int e = threadIdx.x; e < sig_stride; e += blockDim.x) s[e] = 0.0;
    __syncthreads();
    for (int r = 3 + threadIdx.x; r < N; r += blockDim.x) s[stair_at(r, r, N)] = kSigma0;
    for (int e = threadIdx.x; e < st_stride; e += blockDim.x) state[b * (long long)st_stride + e] = 0.0;
    if (threadIdx.x == 0) init_flag[b] = 0;
}

// One filter's staircase Sigma -> dense row-major N x N (ld = N).
__global__ void k_fused_sym_unpack(const double* __restrict__ s, double* __restrict__ out, int N) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N * N; e += gridDim.x * blockDim.x) {
        const int r = e / N, c = e - r * N;
        out[e] = s[stair_at(r, c, N)];
    }
}

}  // namespace ekf
