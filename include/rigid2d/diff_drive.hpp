// The one DiffDrive member on the filter's input path: getBodyTwistForUpdate
// (reference: rigid2d/src/diff_drive.cpp:38-47, called as odometer twist x10 in nuslam/src/slam.cpp:173-176).
#ifndef DIFF_DRIVE_INCLUDE_GUARD_HPP
#define DIFF_DRIVE_INCLUDE_GUARD_HPP
#include "../ekf_slam_b200.h"
#include "rigid2d.hpp"

namespace rigid2d {
class DiffDrive {
  public:
    DiffDrive(double wheel_base, double wheel_radius) : wheel_b(wheel_base), wheel_r(wheel_radius) {}
    Twist2D getBodyTwistForUpdate(double left_angle, double right_angle) {
        double out[2] = {0.0, 0.0};
        ekf_body_twist(wheel_b, wheel_r, left_angle, right_angle, out);
        return Twist2D(out[0], Vector2D(out[1], 0.0));
    }

  private:
    double wheel_b, wheel_r;
};
}  // namespace rigid2d
#endif
