/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C CPU restatement of the reference's EKF-SLAM hot path,
 * in the O(N^2) sparse form (the reference's dense N^3 GEMMs are intractable beyond N ~ 1e3).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this; the product path (ekf-slam-ml_b200/) never does.
 *
 * PARITY PINNING: the reference ships no EKF test, so this file is pinned against
 * the reference's OWN source run on identical inputs: oracle/_ref/libekf_ref.so (reference
 * rigid2d/src/ekf_slam.cpp compiled unmodified, see oracle/Makefile) in tests/test_oracle_vs_ref.py,
 * and against golden vectors generated from that build (tests/golden/, scripts/make_golden.py).
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 * Layout: state = [theta, x, y, m1x, m1y, ...] (ekf_slam.cpp:72-74,15-21); sigma row-major, ld = N.
 * Built with -ffp-contract=off so a*b+c rounds twice, as the reference's x86-64 -O3 build does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_PI 3.14159265358979323846 /* rigid2d.hpp:13 */

typedef struct {
    int n;          /* landmark slots */
    int N;          /* 3 + 2n */
    double* state;  /* N */
    double* sigma;  /* N*N row-major */
    int init_flag;  /* landmark_init_flag, ekf_slam.cpp:50 */
    double* K;      /* scratch N*2 */
    double* W;      /* scratch 2*N */
} ekf_oracle_t;

/* rigid2d.cpp:336-345 */
double oracle_normalize_angle(double rad) {
    double reduced = fmod(rad, 2 * ORACLE_PI);
    double ang = fmod(reduced + 2 * ORACLE_PI, 2 * ORACLE_PI);
    if (ang > ORACLE_PI) ang = ang - 2 * ORACLE_PI;
    return ang;
}

/* diff_drive.cpp:38-47; out = {angular, linear_x} */
void oracle_body_twist(double wheel_base, double wheel_radius, double left, double right, double* out2) {
    double D = wheel_base * 0.5;
    double r = wheel_radius;
    out2[0] = (r / (2.0 * D)) * (right - left);
    out2[1] = (r / 2.0) * (right + left);
}

/* Transform2D::operator*=, rigid2d.cpp:222-245: the composed rotation goes through acos / asin (kept as written:
 * a negative rotation beyond -pi/2 comes out wrong in the reference, and small ones carry acos' cancellation). */
static void compose_tf(double ax, double ay, double ath, double bx, double by, double bth, double* ox, double* oy,
                       double* oth) {
    double r11 = (cos(ath) * cos(bth)) - (sin(ath) * sin(bth));
    double r21 = (sin(ath) * cos(bth)) + (cos(ath) * sin(bth));
    double rc = acos(r11), rs = asin(r21);
    double rot = rc;
    if ((rc * rs) < 0) rot = rs;
    double cx = bx * cos(ath) - by * sin(ath) + ax;
    double cy = bx * sin(ath) + by * cos(ath) + ay;
    *ox = cx;
    *oy = cy;
    *oth = rot;
}

/* integrateTwist, rigid2d.cpp:304-333; out = {x, y, theta} of T_bb' */
void oracle_integrate_twist(double w, double vx, double vy, double* out3) {
    if (fabs(w) > 0.0001) {
        double xs = vy / w, ys = -vx / w;
        /* t_sb = (xs, ys, 0); t_bs = t_sb.inv() (rigid2d.cpp:208-217) */
        double bsx = -xs * cos(0.0) - ys * sin(0.0), bsy = xs * sin(0.0) - ys * cos(0.0), bsth = -0.0;
        double t1x, t1y, t1th;
        compose_tf(0.0, 0.0, w, xs, ys, 0.0, &t1x, &t1y, &t1th); /* t_ssq * t_sqbq */
        compose_tf(bsx, bsy, bsth, t1x, t1y, t1th, &out3[0], &out3[1], &out3[2]);
    } else {
        out3[0] = vx;
        out3[1] = vy;
        out3[2] = 0.0;
    }
}

/* DiffDrive::updatePose, diff_drive.cpp:50-67; pose = {x, y, theta} in place (theta is not wrapped) */
void oracle_update_pose(double wheel_base, double wheel_radius, double* pose3, double left, double right) {
    double tw[2], t[3];
    oracle_body_twist(wheel_base, wheel_radius, left, right, tw);
    oracle_integrate_twist(tw[0], tw[1], 0.0, t);
    /* q = Transform2D(theta).adjConvert(Twist2D(t.theta, (t.x, t.y))), rigid2d.cpp:262-274 with zero translation */
    double th = pose3[2];
    double qa = t[2];
    double qx = qa * 0.0 + t[0] * cos(th) - t[1] * sin(th);
    double qy = -qa * 0.0 + t[0] * sin(th) + t[1] * cos(th);
    pose3[0] += qx;
    pose3[1] += qy;
    pose3[2] += qa;
}

/* ekf_slam.cpp:27-53 */
ekf_oracle_t* oracle_create(int n) {
    ekf_oracle_t* o = (ekf_oracle_t*)calloc(1, sizeof(*o));
    o->n = n;
    o->N = 3 + 2 * n;
    const size_t N = (size_t)o->N;
    o->state = (double*)calloc(N, sizeof(double));
    o->sigma = (double*)calloc(N * N, sizeof(double));
    o->K = (double*)calloc(2 * N, sizeof(double));
    o->W = (double*)calloc(2 * N, sizeof(double));
    for (size_t k = 3; k < N; ++k) o->sigma[k * N + k] = 100.0;
    o->init_flag = 0;
    return o;
}
void oracle_destroy(ekf_oracle_t* o) {
    if (!o) return;
    free(o->state);
    free(o->sigma);
    free(o->K);
    free(o->W);
    free(o);
}
int oracle_dim(const ekf_oracle_t* o) { return o->N; }
void oracle_get_state(const ekf_oracle_t* o, double* out) { memcpy(out, o->state, sizeof(double) * o->N); }
void oracle_set_state(ekf_oracle_t* o, const double* in) { memcpy(o->state, in, sizeof(double) * o->N); }
void oracle_get_sigma(const ekf_oracle_t* o, double* out) {
    memcpy(out, o->sigma, sizeof(double) * (size_t)o->N * o->N);
}
void oracle_set_sigma(ekf_oracle_t* o, const double* in) {
    memcpy(o->sigma, in, sizeof(double) * (size_t)o->N * o->N);
}
int oracle_get_init_flag(const ekf_oracle_t* o) { return o->init_flag; }
void oracle_set_init_flag(ekf_oracle_t* o, int v) { o->init_flag = v != 0; }

/* ekf_slam.cpp:55-106.  Sigma = (I+A) Sigma (I+A)^T + Q with A non-zero only at (1,0),(2,0):
 * rows 1,2 += a*row0 (At*sigma), then cols 1,2 += a*col0 of that product (*At.t()), then Q. */
void oracle_prediction(ekf_oracle_t* o, double dtheta, double dx) {
    const size_t N = (size_t)o->N;
    double* s = o->sigma;
    const double theta = o->state[0];
    double u0, u1, u2, a1, a2;
    if (fabs(dtheta) < 0.000001) { /* :79 */
        u0 = 0;
        u1 = dx * cos(theta);
        u2 = dx * sin(theta);
        a1 = -dx * sin(theta);
        a2 = dx * cos(theta);
    } else {
        u0 = dtheta;
        u1 = -(dx / dtheta) * sin(theta) + (dx / dtheta) * sin(theta + dtheta);
        u2 = (dx / dtheta) * cos(theta) - (dx / dtheta) * cos(theta + dtheta);
        a1 = -(dx / dtheta) * cos(theta) + (dx / dtheta) * cos(theta + dtheta);
        a2 = -(dx / dtheta) * sin(theta) + (dx / dtheta) * sin(theta + dtheta);
    }
    o->state[0] += u0; /* :99, theta not wrapped here */
    o->state[1] += u1;
    o->state[2] += u2;
    for (size_t c = 0; c < N; ++c) {
        const double r0 = s[c];
        s[N + c] = s[N + c] + a1 * r0;
        s[2 * N + c] = s[2 * N + c] + a2 * r0;
    }
    for (size_t r = 0; r < N; ++r) {
        const double c0 = s[r * N];
        s[r * N + 1] = s[r * N + 1] + c0 * a1;
        s[r * N + 2] = s[r * N + 2] + c0 * a2;
    }
    s[0] += 0.0001; /* Q, :41-43 */
    s[N + 1] += 0.0001;
    s[2 * N + 2] += 0.0001;
}

/* H_j's five non-zero columns (ekf_slam.cpp:158-170 / 247-259 / 356-368) and z_hat (:152-155). */
static void h_and_zhat(const ekf_oracle_t* o, int i, double theta, double x, double y, double H[2][5],
                       double zhat[2]) {
    const double mx = o->state[3 + 2 * i], my = o->state[4 + 2 * i];
    const double esti_r = sqrt((mx - x) * (mx - x) + (my - y) * (my - y));
    const double esti_phi = atan2(my - y, mx - x) - theta;
    zhat[0] = esti_r;
    zhat[1] = oracle_normalize_angle(esti_phi);
    const double dx = mx - x, dy = my - y;
    const double d = dx * dx + dy * dy;
    H[0][0] = 0;
    H[0][1] = -dx / sqrt(d);
    H[0][2] = -dy / sqrt(d);
    H[0][3] = dx / sqrt(d);
    H[0][4] = dy / sqrt(d);
    H[1][0] = -1;
    H[1][1] = dy / d;
    H[1][2] = -dx / d;
    H[1][3] = -dy / d;
    H[1][4] = dx / d;
}

/* One landmark correction: ekf_slam.cpp:138-192 (== :335-390).  Pose (theta,x,y) is supplied by the
 * caller because measurement() uses the entry-time pose (:109-111) and data_association() the live one
 * (:331-333).  Returns nothing; mutates state and sigma. */
static void correct_landmark(ekf_oracle_t* o, int i, double sensor_x, double sensor_y, double theta,
                             double x, double y) {
    const size_t N = (size_t)o->N;
    double* s = o->sigma;
    double H[2][5], zhat[2];
    const size_t idx[5] = {0, 1, 2, (size_t)(3 + 2 * i), (size_t)(4 + 2 * i)};
    const double zr = sqrt(sensor_x * sensor_x + sensor_y * sensor_y);
    const double zphi = atan2(sensor_y, sensor_x);
    h_and_zhat(o, i, theta, x, y, H, zhat);

    double* W = o->W; /* W[a*N + c] = (Hj*sigma)(a,c) */
    double* K = o->K; /* first P = sigma*Hj^T in K[r*2+a], then the gain */
    for (size_t c = 0; c < N; ++c)
        for (int a = 0; a < 2; ++a) {
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) acc += H[a][k] * s[idx[k] * N + c];
            W[a * N + c] = acc;
        }
    for (size_t r = 0; r < N; ++r)
        for (int a = 0; a < 2; ++a) {
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) acc += s[r * N + idx[k]] * H[a][k];
            K[r * 2 + a] = acc;
        }
    double S[2][2];
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) acc += W[a * N + idx[k]] * H[b][k];
            S[a][b] = acc + (a == b ? 0.01 : 0.0); /* R, :172-175 */
        }
    const double det = S[0][0] * S[1][1] - S[0][1] * S[1][0];
    const double Si[2][2] = {{S[1][1] / det, -S[0][1] / det}, {-S[1][0] / det, S[0][0] / det}};
    for (size_t r = 0; r < N; ++r) {
        const double p0 = K[r * 2], p1 = K[r * 2 + 1];
        K[r * 2] = p0 * Si[0][0] + p1 * Si[1][0];
        K[r * 2 + 1] = p0 * Si[0][1] + p1 * Si[1][1];
    }
    double nu0 = zr - zhat[0];
    double nu1 = oracle_normalize_angle(zphi - zhat[1]); /* :182-183 */
    for (size_t r = 0; r < N; ++r) o->state[r] = o->state[r] + (K[r * 2] * nu0 + K[r * 2 + 1] * nu1);
    o->state[0] = oracle_normalize_angle(o->state[0]); /* :187 */
    for (size_t r = 0; r < N; ++r) {
        const double k0 = K[r * 2], k1 = K[r * 2 + 1];
        double* row = s + r * N;
        for (size_t c = 0; c < N; ++c) row[c] = row[c] - (k0 * W[c] + k1 * W[N + c]); /* :191-192 */
    }
}

/* ekf_slam.cpp:200-214 (and the first-call loop :113-128) */
static void init_landmark(ekf_oracle_t* o, int i, double sensor_x, double sensor_y, double theta,
                          double x, double y) {
    const double ri = sqrt(sensor_x * sensor_x + sensor_y * sensor_y);
    const double phii = atan2(sensor_y, sensor_x);
    o->state[3 + 2 * i] = x + ri * cos(phii + theta);
    o->state[4 + 2 * i] = y + ri * sin(phii + theta);
}

/* ekf_slam.cpp:108-197 */
void oracle_measurement(ekf_oracle_t* o, const double* xy, const uint8_t* visible) {
    const double theta = o->state[0], x = o->state[1], y = o->state[2]; /* read once, :109-111 */
    if (!o->init_flag) {
        for (int i = 0; i < o->n; ++i) init_landmark(o, i, xy[2 * i], xy[2 * i + 1], theta, x, y);
        o->init_flag = 1;
    }
    for (int i = 0; i < o->n; ++i) {
        if (!visible[i]) continue;
        correct_landmark(o, i, xy[2 * i], xy[2 * i + 1], theta, x, y);
    }
}

/* ekf_slam.cpp:217-276: needs only the 5x5 block sigma[idx,idx].  Bearing innovation NOT wrapped (:269). */
double oracle_maha(const ekf_oracle_t* o, double sensor_x, double sensor_y, int i) {
    const size_t N = (size_t)o->N;
    const double* s = o->sigma;
    const double theta = o->state[0], x = o->state[1], y = o->state[2];
    double H[2][5], zhat[2];
    const size_t idx[5] = {0, 1, 2, (size_t)(3 + 2 * i), (size_t)(4 + 2 * i)};
    const double zr = sqrt(sensor_x * sensor_x + sensor_y * sensor_y);
    const double zphi = atan2(sensor_y, sensor_x);
    h_and_zhat(o, i, theta, x, y, H, zhat);
    double Wl[2][5];
    for (int a = 0; a < 2; ++a)
        for (int l = 0; l < 5; ++l) {
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) acc += H[a][k] * s[idx[k] * N + idx[l]];
            Wl[a][l] = acc;
        }
    double psi[2][2];
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            double acc = 0.0;
            for (int l = 0; l < 5; ++l) acc += Wl[a][l] * H[b][l];
            psi[a][b] = acc + (a == b ? 0.01 : 0.0);
        }
    const double det = psi[0][0] * psi[1][1] - psi[0][1] * psi[1][0];
    const double pi[2][2] = {{psi[1][1] / det, -psi[0][1] / det}, {-psi[1][0] / det, psi[0][0] / det}};
    const double v0 = zr - zhat[0], v1 = zphi - zhat[1];
    const double t0 = v0 * pi[0][0] + v1 * pi[1][0];
    const double t1 = v0 * pi[0][1] + v1 * pi[1][1];
    return t0 * v0 + t1 * v1;
}

/* ekf_slam.cpp:278-402.  assoc_out[j] = landmark corrected by measurement j, or -1 if dropped;
 * dmin_out[j] = min_maha_dis before the new-landmark reset (:326); second_out[j] = runner-up distance
 * (tie-margin diagnostics, +inf if none); created_out[j] = 1 if a landmark was initialised (:318-327).
 * Output pointers may be NULL.  Returns the final known_count. */
int oracle_data_association(ekf_oracle_t* o, const double* xy, int m, uint8_t* known, int32_t* assoc_out,
                            double* dmin_out, double* second_out, uint8_t* created_out) {
    int known_count = 0;
    for (int i = 0; i < o->n; ++i) { /* leading-true prefix, :281-288 */
        if (known[i])
            known_count++;
        else
            break;
    }
    for (int j = 0; j < m; ++j) {
        double min_d = 10.0; /* :293 */
        double second = INFINITY;
        int min_idx = known_count;
        const double sx = xy[2 * j], sy = xy[2 * j + 1];
        for (int i = 0; i < known_count; ++i) {
            const double d = oracle_maha(o, sx, sy, i);
            if (d < min_d) { /* strict, ascending i: lowest index wins ties; NaN never wins */
                second = min_d;
                min_d = d;
                min_idx = i;
            } else if (d < second) {
                second = d;
            }
        }
        if (dmin_out) dmin_out[j] = min_d;
        if (second_out) second_out[j] = second;
        int created = 0;
        if (min_idx == known_count && min_idx < o->n) { /* :318 */
            init_landmark(o, min_idx, sx, sy, o->state[0], o->state[1], o->state[2]);
            known[known_count] = 1;
            known_count++;
            min_d = 0.0;
            created = 1;
        }
        if (created_out) created_out[j] = (uint8_t)created;
        if (min_d < 1.0) { /* :330 */
            correct_landmark(o, min_idx, sx, sy, o->state[0], o->state[1], o->state[2]);
            if (assoc_out) assoc_out[j] = min_idx;
        } else {
            if (assoc_out) assoc_out[j] = -1;
        }
    }
    return known_count;
}

/* Convenience for bench.py's "port" CPU baseline: T steps of prediction + measurement on one filter.
 * twists[t][2] = {dtheta, dx}; xy[t][2n]; vis[t][n].  Returns the number of corrections done. */
int64_t oracle_run_known(ekf_oracle_t* o, int steps, const double* twists, const double* xy,
                         const uint8_t* vis) {
    int64_t updates = 0;
    for (int t = 0; t < steps; ++t) {
        oracle_prediction(o, twists[2 * t], twists[2 * t + 1]);
        const uint8_t* v = vis + (size_t)t * o->n;
        for (int i = 0; i < o->n; ++i) updates += v[i] ? 1 : 0;
        oracle_measurement(o, xy + (size_t)t * 2 * o->n, v);
    }
    return updates;
}
