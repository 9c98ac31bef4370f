// Engine 1c: the batched fused step at the reference's map size (n = 20 slots, N = 43; nuslam/src/slam.cpp:250)
// with Sigma resident in REGISTERS and the rank-2 update on the FP64 tensor-core path (DMMA).
//
// Same step as ekf_fused_sym.cuh (prediction + measurement() or data_association() of one filter per warp, Sigma
// symmetric), but the warp no longer keeps Sigma in shared memory.  ekf_fused_sym_kernel<20> was bound by the
// shared-memory pipe (81 % of the LSU wavefront peak: 190 wavefronts per correction, of which 75 were the read +
// write of the stored Sigma entries by the rank-2 pass and 35 the per-row K fetches).  Here
//   * the 40 x 40 landmark block of Sigma is cut into 8 x 8 blocks (four landmarks by four landmarks).  The 15 blocks
//     of the upper block triangle live in the accumulator-fragment layout of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4):
//     lane (g, t) = (lane / 4, lane % 4) keeps row g, columns 2t, 2t + 1 of every block, 30 doubles, for the whole
//     step;
//   * the rank-2 update of ekf_slam.cpp:191-192, Sigma <- Sigma - K W with K (N x 2) and W = H_j Sigma (2 x N), is a
//     k = 2 outer product; two consecutive corrections make k = 4, i.e. exactly one DMMA per block:
//     D = C + [-K_a | -K_b] [W_a ; W_b].  A pass for TWO corrections is 15 DMMA and ten 8-byte shared-memory loads
//     per lane (one A fragment per block row, one B fragment per block column) instead of 2 x 72 DFMA and 2 x 12
//     16-byte loads with a scalar tiling.  DMMA accumulates k = 0..3 as a chain of FMAs (checked bit for bit in
//     experiments/dmma_rate.cu), so the result equals two sequential rank-2 updates exactly.  The second gain of a
//     pair sees the first one's factor as pending (its two gathered rows take the two FMAs the pass will apply);
//   * the three robot rows of Sigma (3 x 43) live column-distributed: lane c keeps Sigma[0..2][c] and
//     Sigma[0..2][c + 32].  W = H_j Sigma is computed column-parallel, so the robot rows never move; they take their
//     part of the update right in the gain, as Sigma[r][c] -= K[c] . W[r] (the transposed product: K W = W^T S^-1 W
//     is symmetric), which needs only values the lane already holds;
//   * only the two landmark rows 3+2i, 4+2i of a correction are gathered through shared memory; the landmark index
//     picks the registers through a warp-uniform five-way branch (i / 4), never through dynamic register indexing;
//   * HBM holds each filter in exactly that order ([30 slots][32 lanes] fragments + [3][44] robot rows + [44] state
//     = 1,136 doubles = 9,088 B, read and written once per step).  Warps are persistent; while a filter is being
//     corrected the bulk copy engine (cp.async.bulk, SASS UBLKCP) lands the next one in shared memory.
//
// Sigma stays exactly symmetric between blocks (mirror blocks are not stored); inside a diagonal block both (r, c)
// and (c, r) are kept and updated like any other entry.
//
// Restates rigid2d/src/ekf_slam.cpp:55-106, :108-197, :200-214, :217-276, :278-402 (as ekf_fused.cuh).
#pragma once
#include "ekf_fused.cuh"

#ifndef EKF_TILE_DBG
#define EKF_TILE_DBG 0  // timing experiments only (results are wrong): 1 no pass, 2 no gather, 4 no H_j, 32 no corrections
#endif
#ifndef EKF_TILE_MINB
#define EKF_TILE_MINB 12  // resident filters per SM the kernel is compiled for
#endif

#ifndef EKF_TILE_PROF
#define EKF_TILE_PROF 0  // 1: per-phase clock64() totals in g_tile_prof (development builds only)
#endif

namespace ekf {
namespace tile {

#if EKF_TILE_PROF
__device__ unsigned long long g_tile_prof[16];
#define TILE_PROF_MARK(k)                                 \
    do {                                                  \
        const long long now_ = clock64();                 \
        prof_acc[k] += (unsigned long long)(now_ - prof_t); \
        prof_t = now_;                                    \
    } while (0)
#else
#define TILE_PROF_MARK(k) \
    do {                  \
    } while (0)
#endif

constexpr int kNL = 20, kN = 43;
constexpr int kBS = 8;                        // block edge: four landmarks
constexpr int kNB = 5;                        // 40 landmark rows / 8
constexpr int kBlocks = kNB * (kNB + 1) / 2;  // 15 blocks of the upper block triangle
constexpr int kRobOff = 2 * kBlocks * 32;     // 960: robot rows follow the fragments
constexpr int kRobLd = 44;
constexpr int kStOff = kRobOff + 3 * kRobLd;  // 1,092: the state follows the robot rows
constexpr int kStLen = 44;
constexpr int kStride = kStOff + kStLen;      // 1,136 doubles per filter (Sigma and state)
constexpr int kPad = 64;      // padded length of the per-column arrays in shared memory (two slots per lane)
constexpr int kBufStride = 68;  // pairs between the two corrections' K / W buffers: 64 B off a multiple of 128 B, so
                                // that the fragment loads of a half-warp (both buffers) hit distinct banks
constexpr int kGStride = 72;    // doubles between the two gathered rows (likewise)

__host__ __device__ constexpr int block_index(int I, int J) { return I * kNB - I * (I - 1) / 2 + (J - I); }

// offset of Sigma(r, c) inside one filter's block, any 0 <= r, c < 43
__host__ __device__ inline int tile_at(int r, int c) {
    if (r < 3) return kRobOff + kRobLd * r + c;
    if (c < 3) return kRobOff + kRobLd * c + r;  // mirror of a robot-row entry
    int q = r - 3, p = c - 3;
    if (q / kBS > p / kBS) {  // below the block diagonal: the mirror block holds it
        const int t = q;
        q = p;
        p = t;
    }
    const int g = q % kBS, col = p % kBS;
    return (2 * block_index(q / kBS, p / kBS) + (col & 1)) * 32 + 4 * g + (col >> 1);
}

struct TileSmem {
    int off_land, off_bar, off_g, off_w, off_k, off_st, off_cst, off_robm, off_dgm, off_z, total;
    __host__ __device__ TileSmem(int m_max, bool assoc) {
        int o = 0;
        off_land = o, o += kStride * 8;       // landing buffer of the NEXT filter (bulk copy), 9,088 B
        off_bar = o, o += 16;                 // mbarrier of that copy
        off_g = o, o += 4 * kGStride * 8;     // rows 3+2i / 4+2i of a pair's two corrections (index c at [c + 1]:
                                              // column 3 is 16-B aligned for the two-column stores of the gather)
        off_w = o, o += 2 * kBufStride * 16;  // W = H_j Sigma, one pair per column; first / second correction of a pair
        off_k = o, o += 2 * kBufStride * 16;  // K, one pair per row; likewise
        off_st = o, o += kPad * 8;            // mirror of the state (landmark positions for the next H_j)
        off_cst = o, o += 8 * 8;              // measurement(): entry-time pose and its sine / cosine
        off_robm = o, o += assoc ? 3 * kPad * 8 : 0;      // association only: mirror of the robot rows
        off_dgm = o, o += assoc ? 3 * (kNL + 4) * 8 : 0;  // association only: every landmark's 2 x 2 diagonal block
        o = (o + 15) & ~15;
        off_z = o;
        o += 3 * (kNL > m_max ? kNL : m_max) * 8;
        total = (o + 15) & ~15;
    }
};

// ---- gather: rows 3+2i, 4+2i of Sigma -> G3[c], G4[c] for all c ---------------------------------------------------
// Landmark i sits in block row I = i / 4 at rows 2o, 2o + 1 (o = i % 4).  From the diagonal block rightwards the rows
// are held by the lanes g = 2o (row 3+2i) and g = 2o + 1 (row 4+2i), two columns per block each; left of the diagonal
// block they are the mirror of columns 2o, 2o + 1 of the blocks (I', I), held by the lanes t = o.
template <int I>
__device__ __forceinline__ void gather_block_row(double* __restrict__ G3, double* __restrict__ G4,
                                                 const double (&C)[kBlocks][2], const int g, const int t, const int o) {
    const bool r3 = g == 2 * o, r4 = g == 2 * o + 1;
    if (r3 || r4) {
        double* dst = (r3 ? G3 : G4) + 3 + 2 * t;
#pragma unroll
        for (int J = I; J < kNB; ++J)
            *reinterpret_cast<double2*>(dst + kBS * J) = make_double2(C[block_index(I, J)][0], C[block_index(I, J)][1]);
    }
    if (I > 0 && t == o) {
#pragma unroll
        for (int Ip = 0; Ip < I; ++Ip) {
            G3[3 + kBS * Ip + g] = C[block_index(Ip, I)][0];
            G4[3 + kBS * Ip + g] = C[block_index(Ip, I)][1];
        }
    }
}

// columns 0..2 of the two rows: the mirror of the robot rows at columns 3+2i, 4+2i (always up to date)
__device__ __forceinline__ void gather_robot_cols(double* __restrict__ G3, double* __restrict__ G4,
                                                  const double (&rob)[3][2], const int lane, const int i) {
    if (EKF_TILE_DBG & 2) return;
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const bool m3 = lane == (i3 & 31), m4 = lane == (i4 & 31);
    if (m3 || m4) {
        const bool hi = (m3 ? i3 : i4) >= 32;
        double* dst = m3 ? G3 : G4;
#pragma unroll
        for (int r = 0; r < 3; ++r) dst[r] = hi ? rob[r][1] : rob[r][0];
    }
}
// columns 3.. of the two rows, from the landmark blocks
__device__ __forceinline__ void gather_blocks(double* __restrict__ G3, double* __restrict__ G4,
                                              const double (&C)[kBlocks][2], const int lane, const int i) {
    if (EKF_TILE_DBG & 2) return;
    const int I = i >> 2, o = i & 3;  // warp-uniform
    const int g = lane >> 2, t = lane & 3;
    // compare-and-branch chain: a switch becomes a jump table, i.e. a constant load plus an indirect branch
    if (I == 0)
        gather_block_row<0>(G3, G4, C, g, t, o);
    else if (I == 1)
        gather_block_row<1>(G3, G4, C, g, t, o);
    else if (I == 2)
        gather_block_row<2>(G3, G4, C, g, t, o);
    else if (I == 3)
        gather_block_row<3>(G3, G4, C, g, t, o);
    else
        gather_block_row<4>(G3, G4, C, g, t, o);
}
__device__ __forceinline__ void gather_rows(double* __restrict__ G3, double* __restrict__ G4,
                                            const double (&C)[kBlocks][2], const double (&rob)[3][2], const int lane,
                                            const int i) {
    gather_robot_cols(G3, G4, rob, lane, i);
    gather_blocks(G3, G4, C, lane, i);
}

// ---- one landmark correction (ekf_slam.cpp:138-192) around the warp's shared-memory exchanges ---------------------
// phase 1: W = H_j Sigma, column-parallel (robot rows from registers, the two landmark rows from G3 / G4).  With PEND
// the landmark blocks still lack the factor of the pair's first correction (Kpend, this lane's W pairs in wpend): the
// gathered entries take the two FMAs the pass will apply to them.  Columns 0..2 of the gathered rows come from the
// robot rows, which are always up to date.
template <bool PEND, class H>
__device__ __forceinline__ void gain_w(const double* __restrict__ G3, const double* __restrict__ G4,
                                       double2* __restrict__ Wcur, const double2* __restrict__ Kpend,
                                       const double (&rob)[3][2], const double2 (&wpend)[2], double2 (&wown)[2],
                                       const int lane, const int i, const H h) {
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    __syncwarp();  // gathered rows (and the pending K) are in place
    double2 kp3 = make_double2(0.0, 0.0), kp4 = kp3;
    if (PEND) kp3 = Kpend[i3], kp4 = Kpend[i4];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = lane + 32 * s;
        double s3 = G3[c], s4 = G4[c];
        if (PEND) {
            const double f3 = apply_pair(s3, kp3, wpend[s]), f4 = apply_pair(s4, kp4, wpend[s]);
            const bool robot_col = s == 0 && lane < 3;
            s3 = robot_col ? s3 : f3;
            s4 = robot_col ? s4 : f4;
        }
        double2 w;
        h_rows(h, rob[0][s], rob[1][s], rob[2][s], s3, s4, w.x, w.y);
        wown[s] = w;
        Wcur[c] = w;
    }
    __syncwarp();
}

// phase 2: S = W H_j^T + R, K = W^T S^-1, state += K nu, and the robot rows' part of Sigma <- Sigma - K W.
// Leaves K in Kcur and the new state in st.
template <class H>
__device__ __forceinline__ void gain_k(const double2* __restrict__ Wcur, double2* __restrict__ Kcur,
                                       double* __restrict__ st, double (&rob)[3][2], double (&stl)[2],
                                       const double2 (&wown)[2], const int lane, const int i, const H h, const double nu0,
                                       const double nu1) {
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const double2 w0 = Wcur[0], w1 = Wcur[1], w2 = Wcur[2], w3 = Wcur[i3], w4 = Wcur[i4];
    double s00, s01, s10, s11;
    h_rows(h, w0.x, w1.x, w2.x, w3.x, w4.x, s00, s01);
    h_rows(h, w0.y, w1.y, w2.y, w3.y, w4.y, s10, s11);
    const Sym2 si = inv2x2(s00 + kR, s01, s10, s11 + kR);
    double raw0 = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = lane + 32 * s;
        const double2 pw = wown[s];
        const double2 k = make_double2(fma(pw.y, si.i10, pw.x * si.i00), fma(pw.y, si.i11, pw.x * si.i01));
        double ns = stl[s] + fma(k.y, nu1, k.x * nu0);
        if (s == 0) {
            // theta (lane 0) is wrapped after every correction (:187): normalize_angle's |rad| < 2 pi path as selects
            raw0 = ns;
            const double t = __dadd_rn(ns, kTwoPi);
            double ang = (t >= kTwoPi) ? __dsub_rn(t, kTwoPi) : t;
            ang = (ang > kPi) ? __dsub_rn(ang, kTwoPi) : ang;
            ns = (lane == 0) ? ang : ns;
        }
        stl[s] = ns;
        Kcur[c] = k;
        st[c] = ns;
        // Sigma[r][c] -= K[c] . W[r], r = 0..2: the transpose of the reference's K[r] . W[c] (same value: K W is
        // symmetric), from this lane's own K pair and the W pairs already fetched for S
        rob[0][s] = apply_pair(rob[0][s], k, w0);
        rob[1][s] = apply_pair(rob[1][s], k, w1);
        rob[2][s] = apply_pair(rob[2][s], k, w2);
    }
    if (lane == 0 && !(fabs(raw0) < kTwoPi)) {  // never on a sane filter
        stl[0] = normalize_angle_large(raw0);
        st[0] = stl[0];
    }
    __syncwarp();
}

// Sigma <- Sigma - K_a W_a [- K_b W_b] on the landmark blocks: one DMMA.8x8x4 per block with A = [-K_a | -K_b] rows of
// the block row (k = 0, 1: first correction; k = 2, 3: second) and B = [W_a ; W_b] columns of the block column.
// Lane (g, t) supplies A[g][t] and B[t][g].  Kab / Wab: the two corrections' buffers, kBufStride pairs apart.
template <bool PAIR>
__device__ __forceinline__ void pass_blocks(double (&C)[kBlocks][2], const double2* __restrict__ Kab,
                                            const double2* __restrict__ Wab, const int lane) {
    if (EKF_TILE_DBG & 1) return;
    const int g = lane >> 2, t = lane & 3;
    // element (3 + 8 I + g) of correction t / 2, component t % 2, as a double index into the pair arrays
    const int off = 2 * (kBufStride * (t >> 1) + 3 + g) + (t & 1);
    const double* ka = reinterpret_cast<const double*>(Kab) + off;
    const double* wb = reinterpret_cast<const double*>(Wab) + off;
    double a[kNB], b[kNB];
#pragma unroll
    for (int I = 0; I < kNB; ++I) {
        a[I] = (PAIR || t < 2) ? -ka[2 * kBS * I] : 0.0;
        b[I] = (PAIR || t < 2) ? wb[2 * kBS * I] : 0.0;
    }
#pragma unroll
    for (int I = 0; I < kNB; ++I)
#pragma unroll
        for (int J = I; J < kNB; ++J)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(C[block_index(I, J)][0]), "+d"(C[block_index(I, J)][1])
                         : "d"(a[I]), "d"(b[J]));
}

// ---- H_j / nu of a known landmark without a branch (so that ptxas can interleave it with the pass) ----------------
// rsqrt(d) as the CUDA library computes it in the normal range (MUFU.RSQ64H seed + one third-order step), without the
// library's branch for zero / subnormal / infinite / NaN arguments: those set `special` and the caller redoes the
// whole thing out of line with the library routines.
__device__ __forceinline__ double rsqrt_normal_range(double d, bool& special) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    special = (unsigned)(__double2hiint(d) - 0x00100000) >= 0x7fe00000u;
    const double t = __dmul_rn(y, y);
    const double e = fma(d, -t, 1.0);
    const double c = fma(e, 0.375, 0.5);
    const double ye = __dmul_rn(y, e);
    return fma(c, ye, y);
}

// asin(x) = x + x^3 (1/6 + 3/40 x^2 + ... + 231/13312 x^10) for |x| < 1/16: the first dropped term is 6e-17 x
static __constant__ double kAsinCoef[6] = {231.0 / 13312.0, 63.0 / 2816.0, 35.0 / 1152.0,
                                           5.0 / 112.0,     3.0 / 40.0,    1.0 / 6.0};

// same arithmetic as make_innov (ekf_math.cuh) on its fast path; `slow` is set where make_innov would leave it
__device__ __forceinline__ Innov make_innov_nobranch(double mx, double my, double sth, double cth, double x, double y,
                                                     const Reading z, bool& slow) {
    Innov h;
    const double dx = __dsub_rn(mx, x), dy = __dsub_rn(my, y);
    const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    bool special;
    const double isq = rsqrt_normal_range(d, special);
    const double sq = __dmul_rn(d, isq), id = __dmul_rn(isq, isq);
    h.a = -__dmul_rn(dx, isq);
    h.b = -__dmul_rn(dy, isq);
    h.e = __dmul_rn(dy, id);
    h.f = -__dmul_rn(dx, id);
    h.nu0 = __dsub_rn(z.zr, sq);
    const double px = fma(sth, dy, cth * dx), py = fma(cth, dy, -(sth * dx));
    const double sn = fma(px, z.uy, -(py * z.ux)) * isq;
    const double cs = fma(px, z.ux, py * z.uy);
    slow = special || !(cs > 0.0 && fabs(sn) < 0.0625);
    const double x2 = sn * sn;
    double pl = kAsinCoef[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) pl = fma(pl, x2, kAsinCoef[k]);
    h.nu1 = fma(sn * x2, pl, sn);
    return h;
}
// Range and unit direction of a reading without the library's sqrt and division (each with a slow-path branch): the
// rsqrt of the normal range, one correction step for the range, products for the direction; zero / subnormal /
// infinite / NaN arguments take make_reading() out of line.
static __device__ __noinline__ Reading make_reading_cold(double sx, double sy) { return make_reading(sx, sy); }
__device__ __forceinline__ Reading make_reading_fast(double sx, double sy) {
    const double d = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    bool special;
    const double isq = rsqrt_normal_range(d, special);
    Reading z;
    const double r0 = __dmul_rn(d, isq);
    z.zr = fma(fma(-r0, r0, d), 0.5 * isq, r0);
    z.ux = __dmul_rn(sx, isq);
    z.uy = __dmul_rn(sy, isq);
    if (special) z = make_reading_cold(sx, sy);
    return z;
}

static __device__ __noinline__ Innov make_innov_cold(double mx, double my, double theta, double sth, double cth, double x,
                                                     double y, double zr, double ux, double uy) {
    return make_innov(mx, my, theta, sth, cth, x, y, Reading{zr, ux, uy});
}

// Mahalanobis distance of measurement (zr, zphi) to landmark i from the 5 x 5 block Sigma[idx, idx]
// (ekf_slam.cpp:217-276), read from the mirrors of the robot rows and of the landmarks' diagonal blocks.
__device__ __forceinline__ double maha_distance_mirror(const double* __restrict__ robm, const double* __restrict__ dgm,
                                                       const int i, const double* rob6, double mx, double my, double zr,
                                                       double zphi, double theta, double x, double y, Hj& h_out) {
    const Hj h = make_hj(mx, my, theta, x, y);
    h_out = h;
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    const double c03 = robm[i3], c04 = robm[i4], c13 = robm[kPad + i3], c14 = robm[kPad + i4],
                 c23 = robm[2 * kPad + i3], c24 = robm[2 * kPad + i4];
    const double l33 = dgm[3 * i], l34 = dgm[3 * i + 1], l44 = dgm[3 * i + 2];
    const double blk[5][5] = {{rob6[0], rob6[1], rob6[2], c03, c04},
                              {rob6[1], rob6[3], rob6[4], c13, c14},
                              {rob6[2], rob6[4], rob6[5], c23, c24},
                              {c03, c13, c23, l33, l34},
                              {c04, c14, c24, l34, l44}};
    double wl0[5], wl1[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) h_rows(h, blk[0][l], blk[1][l], blk[2][l], blk[3][l], blk[4][l], wl0[l], wl1[l]);
    double p00, p01, p10, p11;
    h_rows(h, wl0[0], wl0[1], wl0[2], wl0[3], wl0[4], p00, p01);
    h_rows(h, wl1[0], wl1[1], wl1[2], wl1[3], wl1[4], p10, p11);
    const Sym2 pi = inv2x2(p00 + kR, p01, p10, p11 + kR);
    const double v0 = __dsub_rn(zr, h.zr), v1 = __dsub_rn(zphi, h.zphi);  // bearing difference NOT wrapped (:269)
    const double t0 = __dadd_rn(__dmul_rn(v0, pi.i00), __dmul_rn(v1, pi.i10));
    const double t1 = __dadd_rn(__dmul_rn(v0, pi.i01), __dmul_rn(v1, pi.i11));
    return __dadd_rn(__dmul_rn(t0, v0), __dmul_rn(t1, v1));
}

static __device__ __noinline__ Hj make_hj_cold(double mx, double my, double theta, double x, double y) {
    return make_hj(mx, my, theta, x, y);
}

// Order-preserving map of a double onto an unsigned 64-bit key (and back): the association's minimum and runner-up
// are then four warp-wide REDUX.MIN on the key halves instead of a five-round shuffle tree.
__device__ __forceinline__ unsigned long long ordered_key(double d) {
    const long long b = __double_as_longlong(d);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double key_value(unsigned hi, unsigned lo) {
    const long long k = (long long)(((unsigned long long)hi << 32) | lo);
    return __longlong_as_double(k ^ ((~k >> 63) | (long long)0x8000000000000000ull));
}

static __device__ __noinline__ void landmark_from_reading_cold(double sx, double sy, double theta, double x, double y,
                                                               double& mx, double& my) {
    landmark_from_reading(sx, sy, theta, x, y, mx, my);
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// Input loads with the PTX .volatile qualifier: ptxas may neither sink them to their first use nor predicate them
// away, so all of a filter's inputs are in flight together instead of one L2 round trip after the other.
__device__ __forceinline__ double ldg_f64_early(const double* ptr) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
    return v;
}
__device__ __forceinline__ int ldg_s32_early(const int32_t* ptr) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr));
    return v;
}
__device__ __forceinline__ unsigned ldg_u8_early(const uint8_t* ptr) {
    unsigned v;
    asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(v) : "l"(ptr));
    return v;
}

// Persistent kernel: one warp per CTA, 148 x EKF_TILE_MINB CTAs, each walking filters b, b + grid, ...  While a filter
// is being corrected in registers, the bulk copy engine lands the next one's Sigma and state in shared memory and its
// inputs are pulled into L2, so a filter's HBM latency is paid behind the previous filter's arithmetic.
template <bool ASSOC>
__global__ void __launch_bounds__(32, EKF_TILE_MINB) ekf_fused_tile_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem S(p.m_max, ASSOC);
    double* land = reinterpret_cast<double*>(smem_raw + S.off_land);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + S.off_bar);
    double* G3 = reinterpret_cast<double*>(smem_raw + S.off_g) + 1;
    double* G4 = G3 + kGStride;
    double *G3b = G3 + 2 * kGStride, *G4b = G3b + kGStride;  // second correction of a pair
    double2* Wab = reinterpret_cast<double2*>(smem_raw + S.off_w);
    double2* Kab = reinterpret_cast<double2*>(smem_raw + S.off_k);
    double* st = reinterpret_cast<double*>(smem_raw + S.off_st);
    double* cst = reinterpret_cast<double*>(smem_raw + S.off_cst);
    double* robm = reinterpret_cast<double*>(smem_raw + S.off_robm);
    double* dgm = reinterpret_cast<double*>(smem_raw + S.off_dgm);
    double* zbuf = reinterpret_cast<double*>(smem_raw + S.off_z);

    constexpr unsigned kFull = 0xffffffffu;
    constexpr int n = kNL;
    constexpr uint32_t kStageBytes = kStride * 8;
    const int lane = threadIdx.x;
    const long long stride_b = gridDim.x;
    long long b = blockIdx.x;
    if (b >= p.B) return;

    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, kStageBytes);
        bulk_g2s(land, p.sigma + b * (long long)kStride, kStageBytes, bar);
    }
    // padded tails of the per-column arrays stay zero for the whole launch
    G3[lane] = G3[lane + 32] = G4[lane] = G4[lane + 32] = 0.0;
    G3b[lane] = G3b[lane + 32] = G4b[lane] = G4b[lane + 32] = 0.0;
    uint32_t phase = 0;
    int n_corr = 0;
#if EKF_TILE_PROF
    unsigned long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif

    for (; b < p.B; b += stride_b) {
        const long long bn = b + stride_b;
        const bool more = bn < p.B;
        TILE_PROF_MARK(7);

        // ---- this filter's inputs (in L2 already, thanks to the previous iteration's prefetch); every load is issued
        // before the first use
        double dtheta = 0.0, dxv = 0.0;
        if (p.mode & kDoPredict) {
            dtheta = ldg_f64_early(p.twists + 2 * b);
            dxv = ldg_f64_early(p.twists + 2 * b + 1);
        }
        int init_flag = p.init_flag[b];  // plain load: written by this kernel
        int m = 0;
        unsigned vismask = 0u;
        int sp_begin = 0, sp_count = 0, sp_next = 0;
        if (!ASSOC && (p.mode & kDoMeasurement) && (p.mode & kSparseReadings)) {
            // marker list: p.mcount = CSR offsets [B + 1], p.vis = landmark ids, p.xy = (x, y) per listed marker
            sp_begin = ldg_s32_early(p.mcount + b);
            sp_count = ldg_s32_early(p.mcount + b + 1) - sp_begin;
            if (more) sp_next = ldg_s32_early(p.mcount + bn);
            if (sp_begin < 0 || sp_count < 0 || (long long)sp_begin + sp_count > p.sparse_total) sp_count = 0;
            for (int k0 = 0; k0 < sp_count; k0 += 32) {
                const int k = k0 + lane;
                const bool on = k < sp_count;
                int id = 0;
                if (on) {
                    const double sx = ldg_f64_early(p.xy + 2 * (long long)(sp_begin + k)),
                                 sy = ldg_f64_early(p.xy + 2 * (long long)(sp_begin + k) + 1);
                    id = (int)ldg_u8_early(p.vis + sp_begin + k);
                    const Reading z = make_reading_fast(sx, sy);
                    if (id < n) {
                        zbuf[3 * id] = z.zr;
                        zbuf[3 * id + 1] = z.ux;
                        zbuf[3 * id + 2] = z.uy;
                    }
                }
                vismask |= __reduce_or_sync(kFull, (on && id < n) ? 1u << id : 0u);
            }
        } else if (!ASSOC && (p.mode & kDoMeasurement)) {
            const int ls = lane < n ? lane : 0;
            const double sx = ldg_f64_early(p.xy + b * 2 * n + 2 * ls), sy = ldg_f64_early(p.xy + b * 2 * n + 2 * ls + 1);
            const unsigned vb = ldg_u8_early(p.vis + b * n + ls);
            vismask = __ballot_sync(kFull, lane < n && vb != 0);
            if (lane < n) {  // range and unit direction of every slot's reading (ekf_slam.cpp:140-146)
                const Reading z = make_reading_fast(sx, sy);
                zbuf[3 * lane] = z.zr;
                zbuf[3 * lane + 1] = z.ux;
                zbuf[3 * lane + 2] = z.uy;
            }
        } else if (ASSOC && (p.mode & kDoAssociation)) {
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

        TILE_PROF_MARK(0);  // inputs issued / reduced
        // ---- this filter's Sigma and state: landing buffer -> registers (lane-major, conflict-free)
        mbar_wait(bar, phase);
        TILE_PROF_MARK(1);  // wait for the landing buffer
        phase ^= 1u;
        double C[kBlocks][2], rob[3][2], stl[2];
#pragma unroll
        for (int k = 0; k < kBlocks; ++k) {
            C[k][0] = land[(2 * k) * 32 + lane];
            C[k][1] = land[(2 * k + 1) * 32 + lane];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            rob[r][0] = land[kRobOff + kRobLd * r + lane];
            rob[r][1] = lane < kN - 32 ? land[kRobOff + kRobLd * r + 32 + lane] : 0.0;
        }
        stl[0] = land[kStOff + lane];
        stl[1] = lane < kN - 32 ? land[kStOff + 32 + lane] : 0.0;
        bool staged_next = false;
        // Issued once every register loaded above has been consumed (after the first pass, or after the write-back
        // of a filter without corrections): the copy engine may then overwrite the landing buffer.
        auto stage_next = [&]() {
            __syncwarp();
            if (more) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, kStageBytes);
                    bulk_g2s(land, p.sigma + bn * (long long)kStride, kStageBytes, bar);
                }
                // the next filter's inputs -> L2
                if (ASSOC) {
                    const char* q = reinterpret_cast<const char*>(p.xy + bn * p.m_max * 2);
                    if (lane * 128 < p.m_max * 16) prefetch_l2(q + lane * 128);
                    if (lane == 29 && p.mcount) prefetch_l2(p.mcount + bn);
                    if (lane == 30) prefetch_l2(p.known + bn * n);
                } else if (p.mode & kSparseReadings) {
                    if (lane < 2) prefetch_l2(reinterpret_cast<const char*>(p.xy + 2 * (long long)sp_next) + lane * 128);
                    if (lane == 2) prefetch_l2(p.vis + sp_next);
                } else {
                    if (lane < 3) prefetch_l2(reinterpret_cast<const char*>(p.xy + bn * 2 * n) + lane * 128);
                    if (lane == 3) prefetch_l2(p.vis + bn * n);
                }
                if (lane == 31) prefetch_l2(p.twists + 2 * bn);
                if (lane == 28) prefetch_l2(p.init_flag + bn);
            }
            staged_next = true;
        };

        // ---- prediction (ekf_slam.cpp:55-106): Sigma <- A Sigma A^T + Q with A = I + a1 e1 e0^T + a2 e2 e0^T.
        // Rows 1, 2 take a * row 0 lane-locally.  Columns 1, 2 exist as such only inside the robot block; below it
        // they are the mirror of rows 1, 2.
        double sth = 0.0, cth = 1.0;
        bool have_sincos = false;
        if (p.mode & kDoPredict) {
            const Motion mo = motion_model(__shfl_sync(kFull, stl[0], 0), dtheta, dxv);
            sth = mo.s_new, cth = mo.c_new, have_sincos = true;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                rob[1][s] = fma(mo.a1, rob[0][s], rob[1][s]);
                rob[2][s] = fma(mo.a2, rob[0][s], rob[2][s]);
            }
            double c0[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) c0[r] = __shfl_sync(kFull, rob[r][0], 0);
            if (lane == 1 || lane == 2) {
                const double a = lane == 1 ? mo.a1 : mo.a2;
#pragma unroll
                for (int r = 0; r < 3; ++r) rob[r][0] = fma(c0[r], a, rob[r][0]);
            }
            if (lane == 0) rob[0][0] += kQ, stl[0] = stl[0] + mo.u0;  // theta is not wrapped here (:99)
            if (lane == 1) rob[1][0] += kQ, stl[0] = stl[0] + mo.u1;
            if (lane == 2) rob[2][0] += kQ, stl[0] = stl[0] + mo.u2;
        }
        st[lane] = stl[0];
        st[lane + 32] = stl[1];
        __syncwarp();
        TILE_PROF_MARK(2);  // landing buffer -> registers, prediction

        double2 wa[2], wb[2];  // this lane's W pairs of the pair's first / second correction

        // ---- measurement(): known association (ekf_slam.cpp:108-197)
        if (!ASSOC && (p.mode & kDoMeasurement)) {
            const double theta = st[0], x = st[1], y = st[2];  // read once; stale for later i (:109-111)
            if (!init_flag) {
                if (p.mode & kSparseReadings) {
                    // unlisted slots read (0, 0): the landmark starts at the robot's position, as with dense zeros
                    if (lane < n) {
                        st[3 + 2 * lane] = x;
                        st[4 + 2 * lane] = y;
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
                } else if (lane < n) {
                    double mx, my;
                    landmark_from_reading_cold(p.xy[b * 2 * n + 2 * lane], p.xy[b * 2 * n + 2 * lane + 1], theta, x, y, mx,
                                               my);
                    st[3 + 2 * lane] = mx;
                    st[4 + 2 * lane] = my;
                }
                init_flag = 1;
                __syncwarp();
                stl[0] = st[lane];
                stl[1] = st[lane + 32];
            }
            if (!have_sincos) sincos(theta, &sth, &cth);
            if (lane == 0) {
                cst[0] = theta, cst[1] = x, cst[2] = y, cst[3] = sth, cst[4] = cth;
            }
            unsigned rem = (EKF_TILE_DBG & 32) ? 0u : vismask;
            if (rem) {
                // Visible landmarks go through in PAIRS: both gains first (the second sees the first one's factor as
                // pending), then ONE DMMA pass applies both rank-2 updates; H_j / nu of the next pair's first
                // landmark are evaluated in the same basic block as the pass, without a branch, so that the scalar
                // chain and the tensor pipe overlap.
                int ia = __ffs(rem) - 1;
                rem &= rem - 1;
                int ib = rem ? __ffs(rem) - 1 : -1;  // the pair's second landmark, if any
                bool slow0 = false;
                Innov h = make_innov_nobranch(st[3 + 2 * ia], st[4 + 2 * ia], sth, cth, x, y,
                                              Reading{zbuf[3 * ia], zbuf[3 * ia + 1], zbuf[3 * ia + 2]}, slow0);
                if (slow0)
                    h = make_innov_cold(st[3 + 2 * ia], st[4 + 2 * ia], theta, sth, cth, x, y, zbuf[3 * ia], zbuf[3 * ia + 1],
                                        zbuf[3 * ia + 2]);
                // both landmarks' rows come from the same landmark blocks (no pass in between); only the robot
                // columns of the second one's rows wait for the first correction
                gather_rows(G3, G4, C, rob, lane, ia);
                if (ib >= 0) gather_blocks(G3b, G4b, C, lane, ib);
                TILE_PROF_MARK(3);  // measurement() entry: init check, first H_j, first gather
                for (;;) {
                    gain_w<false>(G3, G4, Wab, nullptr, rob, wa, wa, lane, ia, h);
                    TILE_PROF_MARK(8);
                    gain_k(Wab, Kab, st, rob, stl, wa, lane, ia, h, h.nu0, h.nu1);
                    TILE_PROF_MARK(9);
                    ++n_corr;
                    if (ib < 0) {  // odd count: the last correction goes through alone
                        pass_blocks<false>(C, Kab, Wab, lane);
                        TILE_PROF_MARK(13);
                        break;
                    }
                    rem &= rem - 1;
                    gather_robot_cols(G3b, G4b, rob, lane, ib);
                    bool slow = false;
                    if (!(EKF_TILE_DBG & 4))
                        h = make_innov_nobranch(st[3 + 2 * ib], st[4 + 2 * ib], cst[3], cst[4], cst[1], cst[2],
                                                Reading{zbuf[3 * ib], zbuf[3 * ib + 1], zbuf[3 * ib + 2]}, slow);
                    if (slow)
                        h = make_innov_cold(st[3 + 2 * ib], st[4 + 2 * ib], cst[0], cst[3], cst[4], cst[1], cst[2],
                                            zbuf[3 * ib], zbuf[3 * ib + 1], zbuf[3 * ib + 2]);
                    TILE_PROF_MARK(10);
                    gain_w<true>(G3b, G4b, Wab + kBufStride, Kab, rob, wa, wb, lane, ib, h);
                    TILE_PROF_MARK(11);
                    gain_k(Wab + kBufStride, Kab + kBufStride, st, rob, stl, wb, lane, ib, h, h.nu0, h.nu1);
                    TILE_PROF_MARK(12);
                    ++n_corr;
                    if (!rem) {
                        pass_blocks<true>(C, Kab, Wab, lane);
                        TILE_PROF_MARK(13);
                        break;
                    }
                    // next pair's first landmark: H_j / nu in one basic block with the pass
                    ia = __ffs(rem) - 1;
                    rem &= rem - 1;
                    ib = rem ? __ffs(rem) - 1 : -1;
                    slow = false;
                    if (!(EKF_TILE_DBG & 4))
                        h = make_innov_nobranch(st[3 + 2 * ia], st[4 + 2 * ia], cst[3], cst[4], cst[1], cst[2],
                                                Reading{zbuf[3 * ia], zbuf[3 * ia + 1], zbuf[3 * ia + 2]}, slow);
                    pass_blocks<true>(C, Kab, Wab, lane);
                    if (slow)
                        h = make_innov_cold(st[3 + 2 * ia], st[4 + 2 * ia], cst[0], cst[3], cst[4], cst[1], cst[2],
                                            zbuf[3 * ia], zbuf[3 * ia + 1], zbuf[3 * ia + 2]);
                    TILE_PROF_MARK(14);
                    if (!staged_next) stage_next();  // every register loaded from the landing buffer has been used
                    gather_rows(G3, G4, C, rob, lane, ia);
                    if (ib >= 0) gather_blocks(G3b, G4b, C, lane, ib);
                    TILE_PROF_MARK(15);
                }
                if (!staged_next) stage_next();
            }
        }

        // ---- data_association(): Mahalanobis nearest neighbour + landmark initialisation (ekf_slam.cpp:278-402)
        if (ASSOC && (p.mode & kDoAssociation)) {
            uint8_t* known = p.known + b * n;
            int known_count;  // leading-true prefix (:281-288)
            {
                const unsigned ones = __ballot_sync(kFull, lane < n && known[lane < n ? lane : 0] != 0);
                known_count = __ffs(~ones) - 1;
                if (known_count > n || known_count < 0) known_count = n;
            }
            const int known_count0 = known_count;
            constexpr int kHjLd = 24;      // six fields x 24 lanes in the rows of the (unused here) second correction
            double* const hjm = G3b - 1;
            bool mirrors_stale = true;
            const int g = lane >> 2, t = lane & 3;
            for (int j = 0; j < m; ++j) {
                if (mirrors_stale) {  // the distances read the robot rows and the landmarks' diagonal blocks
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        robm[r * kPad + lane] = rob[r][0];
                        robm[r * kPad + lane + 32] = rob[r][1];
                    }
                    if ((g >> 1) == t) {  // landmark 4 I + t of diagonal block I: rows 2t, 2t + 1, columns 2t, 2t + 1
#pragma unroll
                        for (int I = 0; I < kNB; ++I) {
                            double* d = dgm + 3 * (4 * I + t);
                            if (g & 1) {
                                d[2] = C[block_index(I, I)][1];
                            } else {
                                d[0] = C[block_index(I, I)][0];
                                d[1] = C[block_index(I, I)][1];
                            }
                        }
                    }
                    __syncwarp();
                    mirrors_stale = false;
                }
                TILE_PROF_MARK(8);
                const double zr = zbuf[2 * j], zphi = zbuf[2 * j + 1];
                const double theta = st[0], x = st[1], y = st[2];  // live pose (:219-221)
                // every known landmark's distance, one lane each; the lane's H_j stays in shared memory because the
                // update of the chosen landmark needs exactly that one again (same pose, same landmark position)
                double dist = INFINITY;
                if (lane < known_count) {
                    const double rob6[6] = {robm[0], robm[1], robm[2], robm[kPad + 1], robm[kPad + 2], robm[2 * kPad + 2]};
                    Hj hl;
                    dist = maha_distance_mirror(robm, dgm, lane, rob6, st[3 + 2 * lane], st[4 + 2 * lane], zr, zphi, theta, x,
                                                y, hl);
                    if (!(dist == dist)) dist = INFINITY;  // NaN never wins
                    hjm[lane] = hl.a;
                    hjm[kHjLd + lane] = hl.b;
                    hjm[2 * kHjLd + lane] = hl.e;
                    hjm[3 * kHjLd + lane] = hl.f;
                    hjm[4 * kHjLd + lane] = hl.zr;
                    hjm[5 * kHjLd + lane] = hl.zphi;
                }
                __syncwarp();
                TILE_PROF_MARK(9);
                // minimum with the lowest index on ties (:300-309) and the runner-up
                double best, second;
                int best_i;
                {
                    const unsigned long long key = ordered_key(dist);
                    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                    const unsigned mhi = __reduce_min_sync(kFull, hi);
                    const unsigned mlo = __reduce_min_sync(kFull, hi == mhi ? lo : 0xffffffffu);
                    best_i = __ffs(__ballot_sync(kFull, hi == mhi && lo == mlo)) - 1;
                    best = key_value(mhi, mlo);
                    const bool other = lane != best_i;
                    const unsigned shi = __reduce_min_sync(kFull, other ? hi : 0xffffffffu);
                    const unsigned slo = __reduce_min_sync(kFull, other && hi == shi ? lo : 0xffffffffu);
                    second = key_value(shi, slo);
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
                TILE_PROF_MARK(10);
                if (min_idx == known_count && min_idx < n) {  // :318-327
                    if (lane == 0) {
                        double mx, my;
                        landmark_from_reading_cold(p.xy[o * 2], p.xy[o * 2 + 1], theta, x, y, mx, my);
                        st[3 + 2 * min_idx] = mx;
                        st[4 + 2 * min_idx] = my;
                    }
                    __syncwarp();
                    stl[0] = st[lane];
                    stl[1] = st[lane + 32];
                    known_count++;
                    min_d = 0.0;
                    created = 1;
                }
                int assoc = -1;
                TILE_PROF_MARK(11);
                if (min_d < kGateUpdate) {  // :330
                    // H_j at the live pose (:331-346): what the chosen landmark's lane already evaluated for its
                    // distance, unless the landmark was created just now
                    Hj h;
                    if (created) {
                        h = make_hj_cold(st[3 + 2 * min_idx], st[4 + 2 * min_idx], st[0], st[1], st[2]);
                    } else {
                        h.a = hjm[min_idx];
                        h.b = hjm[kHjLd + min_idx];
                        h.e = hjm[2 * kHjLd + min_idx];
                        h.f = hjm[3 * kHjLd + min_idx];
                        h.zr = hjm[4 * kHjLd + min_idx];
                        h.zphi = hjm[5 * kHjLd + min_idx];
                    }
                    gather_rows(G3, G4, C, rob, lane, min_idx);
                    TILE_PROF_MARK(12);
                    gain_w<false>(G3, G4, Wab, nullptr, rob, wa, wa, lane, min_idx, h);
                    TILE_PROF_MARK(13);
                    gain_k(Wab, Kab, st, rob, stl, wa, lane, min_idx, h, __dsub_rn(zr, h.zr),
                           normalize_angle(__dsub_rn(zphi, h.zphi)));  // :182-183
                    TILE_PROF_MARK(14);
                    pass_blocks<false>(C, Kab, Wab, lane);  // the next distances need the new Sigma
                    __syncwarp();
                    if (!staged_next) stage_next();
                    mirrors_stale = true;
                    ++n_corr;
                    assoc = min_idx;
                    TILE_PROF_MARK(15);
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

        TILE_PROF_MARK(4);  // corrections
        // ---- write back: registers -> HBM, same lane-major order
        double* go = p.sigma + b * (long long)kStride;
#pragma unroll
        for (int k = 0; k < kBlocks; ++k) {
            __stcs(go + (2 * k) * 32 + lane, C[k][0]);
            __stcs(go + (2 * k + 1) * 32 + lane, C[k][1]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            __stcs(go + kRobOff + kRobLd * r + lane, rob[r][0]);
            if (lane < kN - 32) __stcs(go + kRobOff + kRobLd * r + 32 + lane, rob[r][1]);
        }
        __stcs(go + kStOff + lane, stl[0]);
        if (lane < kN - 32) __stcs(go + kStOff + 32 + lane, stl[1]);
        if (lane == 0) p.init_flag[b] = init_flag;
        if (!staged_next) stage_next();
        TILE_PROF_MARK(5);  // write-back
    }
#if EKF_TILE_PROF
    if (lane == 0)
        for (int k = 0; k < 16; ++k) atomicAdd(&g_tile_prof[k], prof_acc[k]);
#endif
    if (lane == 0 && p.n_updates && n_corr) atomicAdd(p.n_updates, (unsigned long long)n_corr);
}

// Sigma_0 = blockdiag(0_3, 100 I) in the fragment layout (ekf_slam.cpp:36-47), zero state, init flag cleared.
__global__ void k_fused_tile_init(double* sigma, int32_t* init_flag, long long B) {
    const long long b = blockIdx.x;
    if (b >= B) return;
    double* s = sigma + b * (long long)kStride;
    for (int e = threadIdx.x; e < kStride; e += blockDim.x) s[e] = 0.0;
    __syncthreads();
    for (int r = 3 + threadIdx.x; r < kN; r += blockDim.x) s[tile_at(r, r)] = kSigma0;
    if (threadIdx.x == 0) init_flag[b] = 0;
}

// One filter's Sigma -> dense row-major N x N (ld = N).
__global__ void k_fused_tile_unpack(const double* __restrict__ s, double* __restrict__ out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kN * kN; e += gridDim.x * blockDim.x) {
        const int r = e / kN, c = e - r * kN;
        out[e] = s[tile_at(r, c)];
    }
}

}  // namespace tile
}  // namespace ekf
