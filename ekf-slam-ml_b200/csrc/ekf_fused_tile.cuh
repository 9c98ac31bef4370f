// Engine 1c: the batched fused step at the reference's map size (n = 20 slots, N = 43; nuslam/src/slam.cpp:250)
// with Sigma resident in REGISTERS.
//
// Same step as ekf_fused_sym.cuh (prediction + measurement() or data_association() of one filter per warp, Sigma
// symmetric), but the warp no longer keeps Sigma in shared memory.  ekf_fused_sym_kernel<20> was bound by the
// shared-memory pipe (81 % of the LSU wavefront peak: 190 wavefronts per correction, of which 75 were the read +
// write of the stored Sigma entries by the rank-2 pass and 35 the per-row K fetches).  Here
//   * the 40 x 40 landmark block of Sigma is cut into 6 x 6 tiles (three landmarks by three landmarks); the 28 tiles
//     of the upper block triangle live one per lane, 36 doubles in registers, for the whole step.  The rank-2 pass
//     (the (I - K H) Sigma of ekf_slam.cpp:191-192) is 72 FMAs per lane on registers and costs 12 shared-memory
//     loads (six K pairs by tile row, six W pairs by tile column) instead of a read-modify-write of Sigma;
//   * the three robot rows of Sigma (3 x 43) live column-distributed: lane c keeps Sigma[0..2][c] and
//     Sigma[0..2][c + 32].  W = H_j Sigma is computed column-parallel, so the robot rows never move; only the two
//     landmark rows 3+2i, 4+2i of the correction are gathered through shared memory (seven lanes hold them: the
//     tile row of block i/3 and, mirrored, the tile column above it), and the landmark index picks the registers
//     through a warp-uniform three-way branch (i % 3), never through dynamic register indexing;
//   * HBM holds each filter in exactly that order ([36 tile slots][28 lanes] + [3][44] robot rows = 1,140 doubles =
//     9,120 B, all of it used), so staging is 42 fully coalesced 8-byte loads and stores per lane and no copy through
//     shared memory.
// Shared memory per filter drops from 13.5 KB to ~6 KB (K, W, the two gathered rows, a state mirror), so residency
// is set by the register file alone.  Per correction the warp issues ~25 shared-memory instructions instead of ~110.
//
// Sigma stays exactly symmetric above / below the tiles (mirror entries are not stored); inside a diagonal tile both
// (r, c) and (c, r) are kept and updated like any other entry, as in the staircase layout.
//
// Restates rigid2d/src/ekf_slam.cpp:55-106, :108-197, :200-214, :217-276, :278-402 (as ekf_fused.cuh).
#pragma once
#include "ekf_fused.cuh"

#ifndef EKF_TILE_MINB
#define EKF_TILE_MINB 12  // resident filters per SM the kernel is compiled for (168 registers)
#endif

namespace ekf {
namespace tile {

constexpr int kNL = 20, kN = 43;
constexpr int kNB = 7;                       // landmark blocks of three (six rows / columns); the last one holds two
constexpr int kTiles = kNB * (kNB + 1) / 2;  // 28 tiles of the upper block triangle, one per lane
constexpr int kTS = 6;                       // tile edge
constexpr int kRobOff = kTS * kTS * kTiles;  // 1,008: robot rows follow the tile slots
constexpr int kRobLd = 44;
constexpr int kSigStride = kRobOff + 3 * kRobLd;  // 1,140 doubles per filter
constexpr int kStStride = 44;
constexpr int kPad = 64;  // padded length of the per-column arrays in shared memory (two slots per lane)

__host__ __device__ constexpr int tile_index(int rb, int cb) { return rb * kNB - rb * (rb - 1) / 2 + (cb - rb); }

// offset of Sigma(r, c) inside one filter's block, any 0 <= r, c < 43
__host__ __device__ inline int tile_at(int r, int c) {
    if (r < 3) return kRobOff + kRobLd * r + c;
    if (c < 3) return kRobOff + kRobLd * c + r;  // mirror of a robot-row entry
    int lr = r - 3, lc = c - 3;
    if (lr / kTS > lc / kTS) {  // below the block diagonal: the mirror tile holds it
        const int t = lr;
        lr = lc;
        lc = t;
    }
    return ((lr % kTS) * kTS + (lc % kTS)) * kTiles + tile_index(lr / kTS, lc / kTS);
}

struct TileSmem {
    int off_sig, off_stg, off_bar, off_g, off_w, off_k, off_st, off_cst, off_robm, off_dgm, off_z, total;
    __host__ __device__ TileSmem(int m_max, bool assoc) {
        int o = 0;
        off_sig = o, o += kSigStride * 8;  // landing buffer of the NEXT filter's Sigma (bulk copy), 9,120 B
        off_stg = o, o += kStStride * 8;   // ... and of its state
        off_bar = o, o += 16;              // mbarrier of that copy
        off_g = o, o += kPad * 16;         // the correction's two landmark rows, as {Sigma[i3][c], Sigma[i4][c]}
        off_w = o, o += kPad * 16;         // W = H_j Sigma, one pair per column
        off_k = o, o += kPad * 16;         // K, one pair per row
        off_st = o, o += kPad * 8;         // mirror of the state (landmark positions for the next H_j)
        off_cst = o, o += 8 * 8;           // measurement(): entry-time pose and its sine / cosine
        off_robm = o, o += assoc ? 3 * kPad * 8 : 0;        // association only: mirror of the robot rows
        off_dgm = o, o += assoc ? 3 * (kNL + 1) * 8 : 0;    // association only: every landmark's 2 x 2 diagonal block
        o = (o + 15) & ~15;
        off_z = o;
        o += 3 * (kNL > m_max ? kNL : m_max) * 8;
        total = (o + 15) & ~15;
    }
};

struct LaneTile {
    int lane;
    int rb, cb;        // tile coordinates; 15 on the four lanes without a tile
    int row0, col0;    // first matrix row / column of the tile (0 on the lanes without a tile: loads stay in range)
    bool has, diag;
};

// the rows 3+2i, 4+2i of Sigma -> G[c] = {Sigma[i3][c], Sigma[i4][c]} for all c
template <int O>
__device__ __forceinline__ void gather_tile_rows(double2* __restrict__ G, const double (&T)[kTS][kTS], const LaneTile& L,
                                                 const bool isrow, const bool iscol) {
    if (isrow) {
#pragma unroll
        for (int b = 0; b < kTS; ++b) G[L.col0 + b] = make_double2(T[2 * O][b], T[2 * O + 1][b]);
    }
    if (iscol) {
#pragma unroll
        for (int a = 0; a < kTS; ++a) G[L.row0 + a] = make_double2(T[a][2 * O], T[a][2 * O + 1]);
    }
}

__device__ __forceinline__ void gather_rows(double2* __restrict__ G, const double (&T)[kTS][kTS],
                                            const double (&rob)[3][2], const LaneTile& L, const int i) {
    const int bi = i / 3, off = i - 3 * bi;  // warp-uniform
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    // columns 0..2 of the two rows are the mirror of the robot rows at columns i3, i4
    const bool m3 = L.lane == (i3 & 31), m4 = L.lane == (i4 & 31);
    if (m3 || m4) {
        const bool hi = (m3 ? i3 : i4) >= 32;
        double* dst = reinterpret_cast<double*>(G) + (m3 ? 0 : 1);
#pragma unroll
        for (int r = 0; r < 3; ++r) dst[2 * r] = hi ? rob[r][1] : rob[r][0];
    }
    const bool isrow = L.rb == bi, iscol = (L.cb == bi) && !isrow;
    if (off == 0)
        gather_tile_rows<0>(G, T, L, isrow, iscol);
    else if (off == 1)
        gather_tile_rows<1>(G, T, L, isrow, iscol);
    else
        gather_tile_rows<2>(G, T, L, isrow, iscol);
}

// Gain of one correction: W = H_j Sigma (column-parallel: robot rows from registers, landmark rows from G),
// S = W H_j^T + R, K = W^T S^-1, state += K nu.  Leaves K and W in shared memory and this lane's W pairs in wown.
template <class H>
__device__ __forceinline__ void gain(const double2* __restrict__ G, double2* __restrict__ W2, double2* __restrict__ K2,
                                     double* __restrict__ st, const double (&rob)[3][2], double (&stl)[2],
                                     double2 (&wown)[2], const int lane, const int i, const H h, const double nu0,
                                     const double nu1) {
    const int i3 = 3 + 2 * i, i4 = i3 + 1;
    __syncwarp();  // gathered rows are in place
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = lane + 32 * s;
        const double2 g = G[c];
        h_rows(h, rob[0][s], rob[1][s], rob[2][s], g.x, g.y, wown[s].x, wown[s].y);
        W2[c] = wown[s];
    }
    __syncwarp();
    const double2 w0 = W2[0], w1 = W2[1], w2 = W2[2], w3 = W2[i3], w4 = W2[i4];
    double s00, s01, s10, s11;
    h_rows(h, w0.x, w1.x, w2.x, w3.x, w4.x, s00, s01);
    h_rows(h, w0.y, w1.y, w2.y, w3.y, w4.y, s10, s11);
    const Sym2 si = inv2x2(s00 + kR, s01, s10, s11 + kR);
    double raw0 = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = lane + 32 * s;
        const double2 pw = wown[s];
        const double k0 = fma(pw.y, si.i10, pw.x * si.i00);
        const double k1 = fma(pw.y, si.i11, pw.x * si.i01);
        double ns = stl[s] + fma(k1, nu1, k0 * nu0);
        if (s == 0) {
            // theta (lane 0) is wrapped after every correction (:187): normalize_angle's |rad| < 2 pi path as selects
            raw0 = ns;
            const double t = __dadd_rn(ns, kTwoPi);
            double ang = (t >= kTwoPi) ? __dsub_rn(t, kTwoPi) : t;
            ang = (ang > kPi) ? __dsub_rn(ang, kTwoPi) : ang;
            ns = (lane == 0) ? ang : ns;
        }
        stl[s] = ns;
        K2[c] = make_double2(k0, k1);
        st[c] = ns;
    }
    if (lane == 0 && !(fabs(raw0) < kTwoPi)) {  // never on a sane filter
        stl[0] = normalize_angle_large(raw0);
        st[0] = stl[0];
    }
    __syncwarp();
}

// Sigma <- Sigma - K W on this lane's registers ((I - K H_j) Sigma, ekf_slam.cpp:191-192)
__device__ __forceinline__ void rank2(double (&T)[kTS][kTS], double (&rob)[3][2], const double2* __restrict__ K2,
                                      const double2* __restrict__ W2, const double2 (&wown)[2], const LaneTile& L) {
    double2 w[kTS];
#pragma unroll
    for (int b = 0; b < kTS; ++b) w[b] = W2[L.col0 + b];
#pragma unroll
    for (int a = 0; a < kTS; ++a) {
        const double2 k = K2[L.row0 + a];
#pragma unroll
        for (int b = 0; b < kTS; ++b) T[a][b] = apply_pair(T[a][b], k, w[b]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double2 k = K2[r];
#pragma unroll
        for (int s = 0; s < 2; ++s) rob[r][s] = apply_pair(rob[r][s], k, wown[s]);
    }
}

// ---- H_j / nu of a known landmark without a branch (so that ptxas can interleave it with the rank-2 pass) --------
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

static __constant__ double kAsinCoef[9] = {12155.0 / 2490368.0, 6435.0 / 557056.0, 143.0 / 10240.0,
                                           231.0 / 13312.0,     63.0 / 2816.0,      35.0 / 1152.0,
                                           5.0 / 112.0,         3.0 / 40.0,         1.0 / 6.0};

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
    slow = special || !(cs > 0.0 && fabs(sn) < 0.125);
    const double x2 = sn * sn;
    double pl = kAsinCoef[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) pl = fma(pl, x2, kAsinCoef[k]);
    h.nu1 = fma(sn * x2, pl, sn);
    return h;
}
static __device__ __noinline__ Innov make_innov_cold(double mx, double my, double theta, double sth, double cth, double x,
                                                     double y, double zr, double ux, double uy) {
    return make_innov(mx, my, theta, sth, cth, x, y, Reading{zr, ux, uy});
}

// Mahalanobis distance of measurement (zr, zphi) to landmark i from the 5 x 5 block Sigma[idx, idx]
// (ekf_slam.cpp:217-276), read from the mirrors of the robot rows and of the landmarks' diagonal blocks.
__device__ __forceinline__ double maha_distance_mirror(const double* __restrict__ robm, const double* __restrict__ dgm,
                                                       const int i, const double* rob6, double mx, double my, double zr,
                                                       double zphi, double theta, double x, double y) {
    const Hj h = make_hj(mx, my, theta, x, y);
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

static __device__ __noinline__ void landmark_from_reading_cold(double sx, double sy, double theta, double x, double y,
                                                               double& mx, double& my) {
    landmark_from_reading(sx, sy, theta, x, y, mx, my);
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// Persistent kernel: one warp per CTA, 148 x EKF_TILE_MINB CTAs, each walking filters b, b + grid, ...  While a filter
// is being corrected in registers, the bulk copy engine lands the next one's Sigma and state in shared memory and its
// inputs are pulled into L2, so a filter's HBM latency is paid behind the previous filter's arithmetic.
template <bool ASSOC>
__global__ void __launch_bounds__(32, EKF_TILE_MINB) ekf_fused_tile_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem S(p.m_max, ASSOC);
    double* buf_sig = reinterpret_cast<double*>(smem_raw + S.off_sig);
    double* buf_st = reinterpret_cast<double*>(smem_raw + S.off_stg);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + S.off_bar);
    double2* G = reinterpret_cast<double2*>(smem_raw + S.off_g);
    double2* W2 = reinterpret_cast<double2*>(smem_raw + S.off_w);
    double2* K2 = reinterpret_cast<double2*>(smem_raw + S.off_k);
    double* st = reinterpret_cast<double*>(smem_raw + S.off_st);
    double* cst = reinterpret_cast<double*>(smem_raw + S.off_cst);
    double* robm = reinterpret_cast<double*>(smem_raw + S.off_robm);
    double* dgm = reinterpret_cast<double*>(smem_raw + S.off_dgm);
    double* zbuf = reinterpret_cast<double*>(smem_raw + S.off_z);

    constexpr unsigned kFull = 0xffffffffu;
    constexpr int n = kNL;
    constexpr uint32_t kStageBytes = (kSigStride + kStStride) * 8;
    const int lane = threadIdx.x;
    const long long stride_b = gridDim.x;
    long long b = blockIdx.x;
    if (b >= p.B) return;

    LaneTile L;
    L.lane = lane;
    L.has = lane < kTiles;
    {
        int rb = 0, rem = lane;
        while (rb < kNB && rem >= kNB - rb) {
            rem -= kNB - rb;
            ++rb;
        }
        L.rb = L.has ? rb : 15;
        L.cb = L.has ? rb + rem : 15;
        L.row0 = L.has ? 3 + kTS * L.rb : 0;
        L.col0 = L.has ? 3 + kTS * L.cb : 0;
        L.diag = L.has && L.rb == L.cb;
    }

    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, kStageBytes);
        bulk_g2s(buf_sig, p.sigma + b * (long long)kSigStride, kSigStride * 8, bar);
        bulk_g2s(buf_st, p.state + b * (long long)kStStride, kStStride * 8, bar);
    }
    // padded tails of the per-column arrays stay zero for the whole launch
    G[lane] = G[lane + 32] = make_double2(0.0, 0.0);
    uint32_t phase = 0;
    int n_corr = 0;

    for (; b < p.B; b += stride_b) {
        const long long bn = b + stride_b;
        const bool more = bn < p.B;

        // ---- this filter's inputs (in L2 already, thanks to the previous iteration's prefetch)
        double dtheta = 0.0, dxv = 0.0;
        if (p.mode & kDoPredict) {
            dtheta = p.twists[2 * b];
            dxv = p.twists[2 * b + 1];
        }
        int init_flag = p.init_flag[b];
        int m = 0;
        unsigned vismask = 0u;
        int sp_begin = 0, sp_count = 0, sp_next = 0;
        if (!ASSOC && (p.mode & kDoMeasurement) && (p.mode & kSparseReadings)) {
            // marker list: p.mcount = CSR offsets [B + 1], p.vis = landmark ids, p.xy = (x, y) per listed marker
            sp_begin = p.mcount[b];
            sp_count = p.mcount[b + 1] - sp_begin;
            if (more) sp_next = p.mcount[bn];
            if (sp_begin < 0 || sp_count < 0 || (long long)sp_begin + sp_count > p.sparse_total) sp_count = 0;
            for (int k0 = 0; k0 < sp_count; k0 += 32) {
                const int k = k0 + lane;
                const bool on = k < sp_count;
                int id = 0;
                if (on) {
                    id = p.vis[sp_begin + k];
                    const Reading z =
                        make_reading(p.xy[2 * (long long)(sp_begin + k)], p.xy[2 * (long long)(sp_begin + k) + 1]);
                    if (id < n) {
                        zbuf[3 * id] = z.zr;
                        zbuf[3 * id + 1] = z.ux;
                        zbuf[3 * id + 2] = z.uy;
                    }
                }
                vismask |= __reduce_or_sync(kFull, (on && id < n) ? 1u << id : 0u);
            }
        } else if (!ASSOC && (p.mode & kDoMeasurement)) {
            vismask = __ballot_sync(kFull, lane < n && p.vis[b * n + (lane < n ? lane : 0)] != 0);
            if (lane < n) {  // range and unit direction of every slot's reading (ekf_slam.cpp:140-146)
                const Reading z = make_reading(p.xy[b * 2 * n + 2 * lane], p.xy[b * 2 * n + 2 * lane + 1]);
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

        // ---- this filter's Sigma and state: landing buffer -> registers (lane-major, conflict-free)
        mbar_wait(bar, phase);
        phase ^= 1u;
        double T[kTS][kTS], rob[3][2], stl[2];
#pragma unroll
        for (int a = 0; a < kTS; ++a)
#pragma unroll
            for (int c = 0; c < kTS; ++c) T[a][c] = L.has ? buf_sig[(a * kTS + c) * kTiles + lane] : 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            rob[r][0] = buf_sig[kRobOff + kRobLd * r + lane];
            rob[r][1] = lane < kN - 32 ? buf_sig[kRobOff + kRobLd * r + 32 + lane] : 0.0;
        }
        stl[0] = buf_st[lane];
        stl[1] = lane < kN - 32 ? buf_st[32 + lane] : 0.0;
        bool staged_next = false;
        // Issued once every register loaded above has been consumed (after the first rank-2 pass, or after the
        // write-back of a filter without corrections): the copy engine may then overwrite the landing buffer.
        auto stage_next = [&]() {
            __syncwarp();
            if (more) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, kStageBytes);
                    bulk_g2s(buf_sig, p.sigma + bn * (long long)kSigStride, kSigStride * 8, bar);
                    bulk_g2s(buf_st, p.state + bn * (long long)kStStride, kStStride * 8, bar);
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

        double2 wown[2];

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
            unsigned rem = vismask;
            Innov h;
            if (rem) {
                const int i0 = __ffs(rem) - 1;
                h = make_innov(st[3 + 2 * i0], st[4 + 2 * i0], theta, sth, cth, x, y,
                               Reading{zbuf[3 * i0], zbuf[3 * i0 + 1], zbuf[3 * i0 + 2]});
                gather_rows(G, T, rob, L, i0);
            }
            while (rem) {
                const int ic = __ffs(rem) - 1;
                rem &= rem - 1;
                gain(G, W2, K2, st, rob, stl, wown, lane, ic, h, h.nu0, h.nu1);
                ++n_corr;
                // H_j / nu of the next landmark depend on the state just written.  They are evaluated without a
                // branch, in one basic block with the pass, so that the scalar chain overlaps the pass's FMAs (after
                // the last correction the values are simply not used).
                const int in = rem ? __ffs(rem) - 1 : ic;
                bool slow;
                h = make_innov_nobranch(st[3 + 2 * in], st[4 + 2 * in], cst[3], cst[4], cst[1], cst[2],
                                        Reading{zbuf[3 * in], zbuf[3 * in + 1], zbuf[3 * in + 2]}, slow);
                rank2(T, rob, K2, W2, wown, L);
                if (slow)
                    h = make_innov_cold(st[3 + 2 * in], st[4 + 2 * in], cst[0], cst[3], cst[4], cst[1], cst[2], zbuf[3 * in],
                                        zbuf[3 * in + 1], zbuf[3 * in + 2]);
                if (!staged_next) stage_next();
                if (rem) gather_rows(G, T, rob, L, in);
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
            bool mirrors_stale = true;
            for (int j = 0; j < m; ++j) {
                if (mirrors_stale) {  // the distances read the robot rows and the landmarks' diagonal blocks
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        robm[r * kPad + lane] = rob[r][0];
                        robm[r * kPad + lane + 32] = rob[r][1];
                    }
                    if (L.diag) {
#pragma unroll
                        for (int o = 0; o < 3; ++o) {
                            dgm[3 * (3 * L.rb + o)] = T[2 * o][2 * o];
                            dgm[3 * (3 * L.rb + o) + 1] = T[2 * o][2 * o + 1];
                            dgm[3 * (3 * L.rb + o) + 2] = T[2 * o + 1][2 * o + 1];
                        }
                    }
                    __syncwarp();
                    mirrors_stale = false;
                }
                const double zr = zbuf[2 * j], zphi = zbuf[2 * j + 1];
                const double theta = st[0], x = st[1], y = st[2];  // live pose (:219-221)
                double best = INFINITY, second = INFINITY;
                int best_i = 0x7fffffff;
                if (lane < known_count) {
                    const double rob6[6] = {robm[0], robm[1], robm[2], robm[kPad + 1], robm[kPad + 2], robm[2 * kPad + 2]};
                    double d = maha_distance_mirror(robm, dgm, lane, rob6, st[3 + 2 * lane], st[4 + 2 * lane], zr, zphi,
                                                    theta, x, y);
                    if (!(d == d)) d = INFINITY;  // NaN never wins
                    best = d;
                    best_i = lane;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ob = __shfl_xor_sync(kFull, best, off);
                    const double os = __shfl_xor_sync(kFull, second, off);
                    const int oi = __shfl_xor_sync(kFull, best_i, off);
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
                    stl[0] = st[lane];
                    stl[1] = st[lane + 32];
                    known_count++;
                    min_d = 0.0;
                    created = 1;
                }
                int assoc = -1;
                if (min_d < kGateUpdate) {  // :330
                    const double th_l = st[0], x_l = st[1], y_l = st[2];  // live pose (:331-333)
                    const Hj h = make_hj(st[3 + 2 * min_idx], st[4 + 2 * min_idx], th_l, x_l, y_l);
                    gather_rows(G, T, rob, L, min_idx);
                    gain(G, W2, K2, st, rob, stl, wown, lane, min_idx, h, __dsub_rn(zr, h.zr),
                         normalize_angle(__dsub_rn(zphi, h.zphi)));  // :182-183
                    rank2(T, rob, K2, W2, wown, L);  // the next distances need the new Sigma
                    __syncwarp();
                    if (!staged_next) stage_next();
                    mirrors_stale = true;
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

        // ---- write back: registers -> HBM, same lane-major order
        double* go = p.sigma + b * (long long)kSigStride;
        double* g_st = p.state + b * (long long)kStStride;
        if (L.has) {
#pragma unroll
            for (int a = 0; a < kTS; ++a)
#pragma unroll
                for (int c = 0; c < kTS; ++c) __stcs(go + (a * kTS + c) * kTiles + lane, T[a][c]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            __stcs(go + kRobOff + kRobLd * r + lane, rob[r][0]);
            if (lane < kN - 32) __stcs(go + kRobOff + kRobLd * r + 32 + lane, rob[r][1]);
        }
        __stcs(g_st + lane, stl[0]);
        if (lane < kN - 32) __stcs(g_st + 32 + lane, stl[1]);
        if (lane == 0) p.init_flag[b] = init_flag;
        if (!staged_next) stage_next();
    }
    if (lane == 0 && p.n_updates && n_corr) atomicAdd(p.n_updates, (unsigned long long)n_corr);
}

// Sigma_0 = blockdiag(0_3, 100 I) in the tile layout (ekf_slam.cpp:36-47), zero state, init flag cleared.
__global__ void k_fused_tile_init(double* sigma, double* state, int32_t* init_flag, long long B) {
    const long long b = blockIdx.x;
    if (b >= B) return;
    double* s = sigma + b * (long long)kSigStride;
    for (int e = threadIdx.x; e < kSigStride; e += blockDim.x) s[e] = 0.0;
    __syncthreads();
    for (int r = 3 + threadIdx.x; r < kN; r += blockDim.x) s[tile_at(r, r)] = kSigma0;
    for (int e = threadIdx.x; e < kStStride; e += blockDim.x) state[b * (long long)kStStride + e] = 0.0;
    if (threadIdx.x == 0) init_flag[b] = 0;
}

// One filter's tiled Sigma -> dense row-major N x N (ld = N).
__global__ void k_fused_tile_unpack(const double* __restrict__ s, double* __restrict__ out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kN * kN; e += gridDim.x * blockDim.x) {
        const int r = e / kN, c = e - r * kN;
        out[e] = s[tile_at(r, c)];
    }
}

}  // namespace tile
}  // namespace ekf
