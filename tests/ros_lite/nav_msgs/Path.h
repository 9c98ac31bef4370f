#pragma once
#include <vector>

#include "geometry_msgs/Point.h"
namespace nav_msgs {
struct Path {
    std_msgs::Header header;
    std::vector<geometry_msgs::PoseStamped> poses;
};
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
    geometry_msgs::TwistWithCovariance twist;
};
// a Path grows by one pose per publication: the newest pose is the information
inline void lite_dump(std::ostream& o, const Path& p) {
    geometry_msgs::lite_prec(o) << p.poses.size();
    if (!p.poses.empty()) {
        const auto& q = p.poses.back().pose;
        o << ' ' << q.position.x << ' ' << q.position.y << ' ' << q.orientation.x << ' ' << q.orientation.y << ' '
          << q.orientation.z << ' ' << q.orientation.w;
    }
}
inline void lite_dump(std::ostream& o, const Odometry& d) {
    geometry_msgs::lite_prec(o) << d.pose.pose.position.x << ' ' << d.pose.pose.position.y;
}
}  // namespace nav_msgs
