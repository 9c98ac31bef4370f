#pragma once
#include <string>
#include <vector>

#include "geometry_msgs/Point.h"
namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4 };
    enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
    std_msgs::Header header;
    std::string ns;
    int32_t id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
    bool frame_locked = false;
    std::vector<geometry_msgs::Point> points;
};
struct MarkerArray {
    std::vector<Marker> markers;
};
inline void lite_dump(std::ostream& o, const Marker& m) {
    geometry_msgs::lite_prec(o) << m.id << ' ' << m.action << ' ' << m.pose.position.x << ' ' << m.pose.position.y;
}
inline void lite_dump(std::ostream& o, const MarkerArray& a) {
    o << a.markers.size();
    for (const auto& m : a.markers) {
        o << ' ';
        lite_dump(o, m);
    }
}
}  // namespace visualization_msgs
