// Minimal host-side subset of the reference's rigid2d types that the two hot-path classes use
// (reference: rigid2d/include/rigid2d/rigid2d.hpp:13, 68-107, 162-190; rigid2d/src/rigid2d.cpp:100-131, 336-345).
// Only what EKF_SLAM / CircleFitting need is provided here; where the full reference library is present, include
// its own rigid2d.hpp first and define RIGID2D_INCLUDE_GUARD_HPP-compatible types — this header then stands down.
#ifndef RIGID2D_INCLUDE_GUARD_HPP
#define RIGID2D_INCLUDE_GUARD_HPP
#define EKF_B200_RIGID2D_SUBSET 1

#include <cmath>

namespace rigid2d {

constexpr double PI = 3.14159265358979323846;

struct Vector2D {
    double x = 0.0;
    double y = 0.0;
    Vector2D() {}
    Vector2D(double x_val, double y_val) : x(x_val), y(y_val) {}
};

class Twist2D {
  public:
    Twist2D() {}
    Twist2D(double angular, const Vector2D& linear) : ang(angular), lin(linear) {}
    double linearX() const { return lin.x; }
    double linearY() const { return lin.y; }
    double angular() const { return ang; }

  private:
    double ang = 0.0;
    Vector2D lin;
};

// (-pi, pi]; two exact fmod reductions, like the reference.
inline double normalize_angle(double rad) {
    double a = std::fmod(std::fmod(rad, 2 * PI) + 2 * PI, 2 * PI);
    return a > PI ? a - 2 * PI : a;
}

}  // namespace rigid2d
#endif
