// TEST INFRASTRUCTURE ONLY (oracle/).  extern "C" driver around the reference's OWN
// rigid2d::EKF_SLAM / rigid2d::CircleFitting / rigid2d::DiffDrive, compiled in place from
// /root/reference (see oracle/Makefile) against oracle/armadillo_standin/armadillo.
// Output: oracle/_ref/libekf_ref.so.  Used by tests/ (golden-vector generation, oracle pinning)
// and by bench.py's cpu_baseline / --impl reference legs.  Never linked into the product.
//
// The reference keeps `state` and `sigma` private (rigid2d/include/rigid2d/ekf_slam.hpp:61-65);
// parity needs them, so this translation unit (only) sees the class with `private` opened up.
#include <armadillo>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <random>
#include <sstream>
#include <pthread.h>
#include <sched.h>
#include <thread>
#include <vector>

#include "rigid2d/rigid2d.hpp"
#define private public
#include "rigid2d/ekf_slam.hpp"
#include "rigid2d/circle_fitting.hpp"
#undef private
#include "rigid2d/diff_drive.hpp"

using rigid2d::EKF_SLAM;

namespace {
std::vector<bool> to_bools(const uint8_t* p, int n) {
    std::vector<bool> v(n);
    for (int i = 0; i < n; ++i) v[i] = p[i] != 0;
    return v;
}
}  // namespace

extern "C" void scipy_openblas_set_num_threads(int);

extern "C" {

// One filter per host core is the CPU-baseline protocol: keep OpenBLAS itself single-threaded.
void ref_blas_threads(int n) { scipy_openblas_set_num_threads(n); }

// The reference prints inside its hot loops (ekf_slam.cpp:213,268,290-329,399; circle_fitting.cpp:72).
void ref_silence_stdout(int on) {
    if (on)
        std::cout.setstate(std::ios::failbit);
    else
        std::cout.clear();
}

void* ref_ekf_create(int n) { return new EKF_SLAM(n); }
void ref_ekf_destroy(void* h) { delete static_cast<EKF_SLAM*>(h); }

void ref_ekf_prediction(void* h, double dtheta, double dx) {
    rigid2d::Twist2D tw(dtheta, rigid2d::Vector2D(dx, 0.0));
    static_cast<EKF_SLAM*>(h)->prediction(tw);
}

void ref_ekf_measurement(void* h, const double* xy, const uint8_t* visible) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    const int n = f->n;
    mat z = zeros<mat>(2 * n, 1);
    for (int i = 0; i < 2 * n; ++i) z(i, 0) = xy[i];
    f->measurement(z, to_bools(visible, n), std::vector<bool>(n, false));
}

void ref_ekf_data_association(void* h, const double* xy, int m, uint8_t* known) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    const int n = f->n;
    std::vector<rigid2d::Vector2D> meas;
    for (int j = 0; j < m; ++j) meas.emplace_back(xy[2 * j], xy[2 * j + 1]);
    std::vector<bool> kl = to_bools(known, n);
    f->data_association(meas, kl);
    for (int i = 0; i < n; ++i) known[i] = kl[i] ? 1 : 0;
}

double ref_ekf_maha(void* h, double mx, double my, int i) {
    return static_cast<EKF_SLAM*>(h)->calculate_maha_dis(rigid2d::Vector2D(mx, my), i);
}

int ref_ekf_dim(void* h) { return 3 + 2 * static_cast<EKF_SLAM*>(h)->n; }

void ref_ekf_get_state(void* h, double* out) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    const int N = 3 + 2 * f->n;
    for (int i = 0; i < N; ++i) out[i] = f->state(i, 0);
}
void ref_ekf_set_state(void* h, const double* in) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    const int N = 3 + 2 * f->n;
    for (int i = 0; i < N; ++i) f->state(i, 0) = in[i];
}
// row-major N x N
void ref_ekf_get_sigma(void* h, double* out) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    const int N = 3 + 2 * f->n;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) out[(size_t)i * N + j] = f->sigma(i, j);
}
void ref_ekf_set_sigma(void* h, const double* in) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    const int N = 3 + 2 * f->n;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) f->sigma(i, j) = in[(size_t)i * N + j];
}
int ref_ekf_get_init_flag(void* h) { return static_cast<EKF_SLAM*>(h)->landmark_init_flag ? 1 : 0; }
void ref_ekf_set_init_flag(void* h, int v) { static_cast<EKF_SLAM*>(h)->landmark_init_flag = v != 0; }
void ref_ekf_getters(void* h, double* out3) {
    EKF_SLAM* f = static_cast<EKF_SLAM*>(h);
    out3[0] = f->getStateTheta();
    out3[1] = f->getStateX();
    out3[2] = f->getStateY();
}

double ref_normalize_angle(double a) { return rigid2d::normalize_angle(a); }

void ref_body_twist(double wheel_base, double wheel_radius, double dl, double dr, double* out2) {
    rigid2d::DiffDrive dd(wheel_base, wheel_radius);
    rigid2d::Twist2D t = dd.getBodyTwistForUpdate(dl, dr);
    out2[0] = t.angular();
    out2[1] = t.linearX();
}

void ref_integrate_twist(double w, double vx, double vy, double* out3) {
    rigid2d::Transform2D t = rigid2d::integrateTwist(rigid2d::Twist2D(w, rigid2d::Vector2D(vx, vy)));
    out3[0] = t.x();
    out3[1] = t.y();
    out3[2] = t.theta();
}

void ref_wheel_velocity(double wheel_base, double wheel_radius, double ang, double vx, double* out2) {
    rigid2d::DiffDrive dd(wheel_base, wheel_radius);
    rigid2d::Vector2D w = dd.calculateWheelVelocity(rigid2d::Twist2D(ang, rigid2d::Vector2D(vx, 0.0)));
    out2[0] = w.x;
    out2[1] = w.y;
}

// One DiffDrive::updatePose from pose (x, y, theta); in/out pose3 = {x, y, theta}.
void ref_update_pose(double wheel_base, double wheel_radius, double* pose3, double dl, double dr) {
    rigid2d::Vector2D p(pose3[0], pose3[1]);
    rigid2d::DiffDrive dd(wheel_base, wheel_radius, p, pose3[2]);
    dd.updatePose(dl, dr);
    pose3[0] = dd.getPosition().x;
    pose3[1] = dd.getPosition().y;
    pose3[2] = dd.getTheta();
}

// ---------------------------------------------------------------- circle fitting
// clusteringRanges: returns number of clusters; sizes[k], and concatenated ranges / xy.
int ref_cluster(const double* ranges, int len, int* sizes, int max_clusters, double* flat_ranges,
                double* flat_xy, int max_points) {
    rigid2d::CircleFitting cf;
    cf.clusteringRanges(std::vector<double>(ranges, ranges + len));
    const auto& pc = cf.point_cluster;
    const auto& xc = cf.xy_cluster;
    int k = 0, p = 0;
    for (size_t c = 0; c < pc.size() && k < max_clusters; ++c, ++k) {
        sizes[k] = (int)pc[c].size();
        for (size_t j = 0; j < pc[c].size() && p < max_points; ++j, ++p) {
            flat_ranges[p] = pc[c][j];
            flat_xy[2 * p] = xc[c][j].x;
            flat_xy[2 * p + 1] = xc[c][j].y;
        }
    }
    return (int)pc.size();
}

// circleRegression (+ classifyCircle flag) on caller-supplied clusters (the set_xy_cluster seam,
// circle_fitting.cpp:96-98).  out_cxr: per cluster (cx, cy, r).  is_circle: classification.
void ref_fit_clusters(const double* flat_xy, const int* sizes, int n_clusters, double* out_cxr,
                      uint8_t* is_circle) {
    rigid2d::CircleFitting cf;
    std::vector<std::vector<rigid2d::Vector2D>> cl;
    int p = 0;
    for (int c = 0; c < n_clusters; ++c) {
        std::vector<rigid2d::Vector2D> pts;
        for (int j = 0; j < sizes[c]; ++j, ++p) pts.emplace_back(flat_xy[2 * p], flat_xy[2 * p + 1]);
        cl.push_back(pts);
    }
    cf.set_xy_cluster(cl);
    std::vector<rigid2d::Vector2D> pos = cf.circleRegression();
    std::vector<double> r = cf.get_r_cluster();
    for (int c = 0; c < n_clusters; ++c) {
        out_cxr[3 * c] = pos[c].x;
        out_cxr[3 * c + 1] = pos[c].y;
        out_cxr[3 * c + 2] = r[c];
    }
    if (is_circle) {
        // classifyCircle compacts; recover the flags by matching in order.
        std::vector<rigid2d::Vector2D> kept = cf.classifyCircle(pos);
        size_t q = 0;
        for (int c = 0; c < n_clusters; ++c) {
            is_circle[c] = 0;
            if (q < kept.size() && kept[q].x == pos[c].x && kept[q].y == pos[c].y) {
                is_circle[c] = 1;
                ++q;
            }
        }
    }
}

// approxCirclePositions on one scan.  Returns number of circles (<= max), or -1 when the
// reference would hit its empty-cluster UB (circle_fitting.cpp:54) — detected here first.
int ref_circles(const double* ranges, int len, double* centers, int max) {
    // pre-check the UB condition with the reference's own split rule (circle_fitting.cpp:30-40)
    int run = 1, kept = 0;
    for (int i = 1; i < len; ++i) {
        if ((std::fabs(ranges[i] - ranges[i - 1]) < 0.2) && (i != len - 1)) {
            ++run;
        } else {
            if (run > 6) ++kept;
            run = 1;
        }
    }
    if (kept == 0) return -1;
    rigid2d::CircleFitting cf;
    std::vector<rigid2d::Vector2D> c = cf.approxCirclePositions(std::vector<double>(ranges, ranges + len));
    int k = 0;
    for (; k < (int)c.size() && k < max; ++k) {
        centers[2 * k] = c[k].x;
        centers[2 * k + 1] = c[k].y;
    }
    return (int)c.size();
}

// ---------------------------------------------------------------- CPU baseline (bench.py only)
// Runs `n_filters` independent reference filters, one std::thread each, over the same per-filter
// traces: per step prediction(twist) then measurement(xy, visible) (nuslam/src/slam.cpp:433-434).
// Layout: twists[f][t][2], xy[f][t][2n], vis[f][t][n].  Returns wall seconds; *updates_out gets the
// number of landmark corrections executed (visible flags after the init-only first call).
double ref_bench_known(int n, int n_filters, int steps, const double* twists, const double* xy,
                       const uint8_t* vis, int64_t* updates_out, double* final_states /*may be null*/) {
    std::vector<EKF_SLAM> filters;
    for (int f = 0; f < n_filters; ++f) filters.emplace_back(n);
    std::atomic<int64_t> updates{0};
    const int N = 3 + 2 * n;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int f = 0; f < n_filters; ++f) {
        th.emplace_back([&, f]() {
            // one filter per host core, pinned: thread f on the f-th CPU of the allowed set
            cpu_set_t allowed;
            if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
                int seen = 0, target = f % CPU_COUNT(&allowed);
                for (int c = 0; c < CPU_SETSIZE; ++c)
                    if (CPU_ISSET(c, &allowed) && seen++ == target) {
                        cpu_set_t one;
                        CPU_ZERO(&one);
                        CPU_SET(c, &one);
                        pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
                        break;
                    }
            }
            EKF_SLAM& F = filters[f];
            int64_t u = 0;
            mat z = zeros<mat>(2 * n, 1);
            std::vector<bool> kl(n, false);
            for (int t = 0; t < steps; ++t) {
                const size_t s = (size_t)f * steps + t;
                rigid2d::Twist2D tw(twists[2 * s], rigid2d::Vector2D(twists[2 * s + 1], 0.0));
                F.prediction(tw);
                for (int i = 0; i < 2 * n; ++i) z(i, 0) = xy[s * 2 * n + i];
                std::vector<bool> v(n);
                for (int i = 0; i < n; ++i) {
                    v[i] = vis[s * n + i] != 0;
                    u += v[i] ? 1 : 0;
                }
                F.measurement(z, v, kl);
            }
            updates += u;
        });
    }
    for (auto& t : th) t.join();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *updates_out = updates.load();
    if (final_states)
        for (int f = 0; f < n_filters; ++f)
            for (int i = 0; i < N; ++i) final_states[(size_t)f * N + i] = filters[f].state(i, 0);
    return sec;
}

}  // extern "C"
