// DiffDrive members on the filter's input path: getBodyTwistForUpdate (reference: rigid2d/src/diff_drive.cpp:38-47,
// called as odometer twist x10 in nuslam/src/slam.cpp:173-176) and the dead-reckoning odometer updatePose /
// getPosition / getTheta (diff_drive.cpp:50-77, driven per joint_states message at slam.cpp:96).
#ifndef DIFF_DRIVE_INCLUDE_GUARD_HPP
#define DIFF_DRIVE_INCLUDE_GUARD_HPP
#include "../ekf_slam_b200.h"
#include "rigid2d.hpp"

namespace rigid2d {
class DiffDrive {
  public:
    DiffDrive(double wheel_base, double wheel_radius) : wheel_b(wheel_base), wheel_r(wheel_radius) {}
    DiffDrive(double wheel_base, double wheel_radius, Vector2D& init_pos, double init_theta)
        : wheel_b(wheel_base), wheel_r(wheel_radius), pose{init_pos.x, init_pos.y, init_theta} {}
    void updatePose(double left_angle, double right_angle) {
        ekf_update_pose(wheel_b, wheel_r, 1, pose, &left_angle, &right_angle);
    }
    Vector2D getPosition() { return Vector2D(pose[0], pose[1]); }
    double getTheta() { return pose[2]; }
    Twist2D getBodyTwistForUpdate(double left_angle, double right_angle) {
        double out[2] = {0.0, 0.0};
        ekf_body_twist(wheel_b, wheel_r, left_angle, right_angle, out);
        return Twist2D(out[0], Vector2D(out[1], 0.0));
    }

  private:
    double wheel_b, wheel_r;
    double pose[3] = {0.0, 0.0, 0.0};
};
}  // namespace rigid2d
#endif
