// Per-call latency of the C++ facade on the reference's node cadence (nuslam/src/slam.cpp:433-434 and
// unknown_data_assoc.cpp:414-415): prediction + measurement / data_association + the three pose getters the node
// reads afterwards (slam.cpp:278-279,343).  Prints one JSON object; used by bench.py's single_filter leg.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rigid2d/diff_drive.hpp"
#include "rigid2d/ekf_slam.hpp"

using namespace rigid2d;
using Clock = std::chrono::steady_clock;

int main(int argc, char** argv) {
    const int steps = argc > 1 ? std::atoi(argv[1]) : 400;
    const double lx[6] = {0.5, 0.2, -0.3, 0.6, -0.3, 0.0}, ly[6] = {0.1, -0.3, 0.5, -0.8, -0.6, 1.1};
    double us_known = 0.0, us_unknown = 0.0;
    int upd_known = 0, upd_unknown = 0;
    for (int mode = 0; mode < 2; ++mode) {
        EKF_SLAM f;
        f = EKF_SLAM(20);
        if (f.last_status() != 0) return 1;
        std::vector<bool> visible(20, false), known(20, false);
        mat reading = zeros<mat>(40, 1);
        double th = 0, x = 0, y = 0;
        for (int step = -20; step < steps; ++step) {  // 20 untimed warm-up steps
            const auto t0 = Clock::now();
            f.prediction(Twist2D(0.05, Vector2D(0.01, 0.0)));
            if (mode == 0) {
                for (int i = 0; i < 6; ++i) {
                    const double dx = lx[i] - x, dy = ly[i] - y;
                    reading(2 * i, 0) = std::cos(th) * dx + std::sin(th) * dy;
                    reading(2 * i + 1, 0) = -std::sin(th) * dx + std::cos(th) * dy;
                    visible[i] = step > -20;
                }
                f.measurement(reading, visible, known);
            } else {
                std::vector<Vector2D> m;
                for (int i = 0; i < 6; ++i) {
                    const double dx = lx[i] - x, dy = ly[i] - y;
                    m.push_back(Vector2D(std::cos(th) * dx + std::sin(th) * dy, -std::sin(th) * dx + std::cos(th) * dy));
                }
                f.data_association(m, known);
            }
            th = f.getStateTheta(), x = f.getStateX(), y = f.getStateY();
            const double us = std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
            if (step >= 0) {
                (mode == 0 ? us_known : us_unknown) += us;
                (mode == 0 ? upd_known : upd_unknown) += 6;
            }
        }
    }
    std::printf("{\"steps\": %d, \"us_per_step_known\": %.3f, \"us_per_step_unknown\": %.3f, \"updates_per_step\": 6}\n", steps,
                us_known / steps, us_unknown / steps);
    return 0;
}
