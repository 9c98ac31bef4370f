#pragma once
#include <iomanip>
#include <ostream>

#include "std_msgs/Header.h"
namespace geometry_msgs {
struct Point {
    double x = 0, y = 0, z = 0;
};
struct Vector3 {
    double x = 0, y = 0, z = 0;
};
struct Quaternion {
    double x = 0, y = 0, z = 0, w = 0;
};
struct Pose {
    Point position;
    Quaternion orientation;
};
struct PoseStamped {
    std_msgs::Header header;
    Pose pose;
};
struct Twist {
    Vector3 linear, angular;
};
struct Transform {
    Vector3 translation;
    Quaternion rotation;
};
struct TransformStamped {
    std_msgs::Header header;
    std::string child_frame_id;
    Transform transform;
};
struct PoseWithCovariance {
    Pose pose;
};
struct TwistWithCovariance {
    Twist twist;
};
inline std::ostream& lite_prec(std::ostream& o) { return o << std::setprecision(17); }
inline void lite_dump(std::ostream& o, const TransformStamped& t) {
    lite_prec(o) << t.header.frame_id << ' ' << t.child_frame_id << ' ' << t.transform.translation.x << ' '
                 << t.transform.translation.y << ' ' << t.transform.rotation.x << ' ' << t.transform.rotation.y << ' '
                 << t.transform.rotation.z << ' ' << t.transform.rotation.w;
}
inline void lite_dump(std::ostream& o, const Twist& t) { lite_prec(o) << t.linear.x << ' ' << t.angular.z; }
}  // namespace geometry_msgs
