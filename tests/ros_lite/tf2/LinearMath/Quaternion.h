#pragma once
#include <cmath>
namespace tf2 {
// tf2::Quaternion::setRPY as in tf2/LinearMath/Quaternion.h (half-angle products)
class Quaternion {
  public:
    void setRPY(double roll, double pitch, double yaw) {
        const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
        const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
        x_ = sr * cp * cy - cr * sp * sy;
        y_ = cr * sp * cy + sr * cp * sy;
        z_ = cr * cp * sy - sr * sp * cy;
        w_ = cr * cp * cy + sr * sp * sy;
    }
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    double w() const { return w_; }

  private:
    double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
};
}  // namespace tf2
