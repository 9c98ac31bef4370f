// Exercises the C++ facade exactly the way the reference's callers do:
//   nuslam/src/slam.cpp:428,433-434 (copy-assign a fresh filter, prediction + measurement, getters),
//   nuslam/src/unknown_data_assoc.cpp:414-415 (prediction + data_association),
//   nuslam/tests/circle_tests.cpp (the four known-answer cases).
// Prints "FACADE OK" and exits 0 when everything holds.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "rigid2d/circle_fitting.hpp"
#include "rigid2d/diff_drive.hpp"
#include "rigid2d/ekf_slam.hpp"

using namespace rigid2d;

#define REQUIRE(c)                                                       \
    do {                                                                 \
        if (!(c)) {                                                      \
            std::fprintf(stderr, "REQUIRE failed line %d: %s\n", __LINE__, #c); \
            return 1;                                                    \
        }                                                                \
    } while (0)

static bool approx(double a, double b) { return std::fabs(a - b) <= 1.2e-5 * std::fmax(std::fabs(a), std::fabs(b)); }

int main() {
    // --- SLAM node pattern: by-value member, fresh filter copy-assigned in INIT
    EKF_SLAM slam_agent;
    slam_agent = rigid2d::EKF_SLAM(20);
    REQUIRE(slam_agent.last_status() == 0);
    DiffDrive dd(0.16, 0.033);
    std::vector<bool> visible(20, false), known(20, false);
    mat reading = zeros<mat>(40, 1);
    const double lx[3] = {0.5, 0.2, -0.3}, ly[3] = {0.1, -0.3, 0.5};
    for (int step = 0; step < 30; ++step) {
        Twist2D tw = dd.getBodyTwistForUpdate(10.0 * 0.025, 10.0 * 0.035);
        slam_agent.prediction(tw);
        const double th = slam_agent.getStateTheta(), x = slam_agent.getStateX(), y = slam_agent.getStateY();
        for (int i = 0; i < 3; ++i) {
            const double dx = lx[i] - x, dy = ly[i] - y;
            reading(2 * i, 0) = std::cos(th) * dx + std::sin(th) * dy;
            reading(2 * i + 1, 0) = -std::sin(th) * dx + std::cos(th) * dy;
            visible[i] = step > 0;
        }
        slam_agent.measurement(reading, visible, known);
        REQUIRE(slam_agent.last_status() == 0);
    }
    mat lm = slam_agent.getStateLandmark();
    REQUIRE(lm.n_rows == 40 && lm.n_cols == 1);
    for (int i = 0; i < 3; ++i) REQUIRE(std::fabs(lm(2 * i, 0) - lx[i]) < 0.05 && std::fabs(lm(2 * i + 1, 0) - ly[i]) < 0.05);
    EKF_SLAM copy = slam_agent;  // deep copy
    copy.prediction(Twist2D(0.1, Vector2D(0.05, 0.0)));
    REQUIRE(std::fabs(copy.getStateTheta() - slam_agent.getStateTheta()) > 0.05);

    // --- the odometer the node runs beside the filter (slam.cpp:96), known answer of rigid2d/tests/tests.cpp:372-383
    DiffDrive odo(0.2, 0.05);
    odo.updatePose(0.0, 2 * 3.1415926);
    REQUIRE(std::fabs(odo.getTheta() - 1.5708) < 2e-5);
    REQUIRE(std::fabs(odo.getPosition().x - 0.1) < 2e-6 && std::fabs(odo.getPosition().y - 0.1) < 2e-6);

    // --- unknown data association pattern
    EKF_SLAM uda(20);
    std::vector<bool> known_list(20, false);
    for (int step = 0; step < 10; ++step) {
        uda.prediction(Twist2D(0.0, Vector2D(0.01, 0.0)));
        const double x = uda.getStateX(), y = uda.getStateY();
        std::vector<Vector2D> measures;
        for (int i = 0; i < 3; ++i) measures.push_back(Vector2D(lx[i] - x, ly[i] - y));
        uda.data_association(measures, known_list);
        REQUIRE(uda.last_status() == 0);
    }
    REQUIRE(known_list[0] && known_list[1] && known_list[2] && !known_list[3]);
    REQUIRE(uda.last_association().size() == 3 && uda.last_association()[0] == 0 && uda.last_association()[2] == 2);

    // --- the reference's four known-answer circle cases (values from nuslam/tests/circle_tests.cpp:8-76)
    const std::vector<double> scan21{0.713136, 0.682084, 0.668864, 0.660664, 0.65551,  0.652665, 0.651814,
                                     0.652875, 0.655952, 0.661391, 0.670004, 0.684042, 1.01247,  1.01543,
                                     1.01872,  1.02234,  1.0263,   1.03061,  1.04061,  1.05061,  1.06061};
    auto fit_one = [](std::initializer_list<Vector2D> pts, double& cx, double& cy, double& r) {
        CircleFitting fitter;
        fitter.set_xy_cluster({std::vector<Vector2D>(pts)});
        const std::vector<Vector2D> centre = fitter.circleRegression();
        cx = centre[0].x;
        cy = centre[0].y;
        r = fitter.get_r_cluster()[0];
    };
    {
        CircleFitting fitter;
        fitter.clusteringRanges(scan21);
        const auto clusters = fitter.get_point_cluster();
        REQUIRE(clusters.size() == 2 && approx(clusters[1][0], 1.01247));
        REQUIRE(fitter.classifyCircle(fitter.circleRegression()).empty());
        REQUIRE(fitter.approxCirclePositions(scan21).empty());
    }
    double cx, cy, r;
    fit_one({{1.0, 7.0}, {2.0, 6.0}, {5.0, 8.0}, {7.0, 7.0}, {9.0, 5.0}, {3.0, 7.0}}, cx, cy, r);
    REQUIRE(approx(cx, 4.615482) && approx(cy, 2.807354) && approx(r, 4.827575));
    fit_one({{-1.0, 0.0}, {-0.3, -0.06}, {0.3, 0.1}, {1.0, 0.0}}, cx, cy, r);
    REQUIRE(approx(cx, 0.4908357) && approx(cy, -22.15212) && approx(r, 22.17979));
    std::printf("FACADE OK\n");
    return 0;
}
