#pragma once
#include <vector>

#include "geometry_msgs/Point.h"
namespace sensor_msgs {
struct LaserScan {
    std_msgs::Header header;
    float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0, range_min = 0, range_max = 0;
    std::vector<float> ranges, intensities;
};
inline void lite_dump(std::ostream& o, const LaserScan& s) {
    o << std::setprecision(9) << s.ranges.size();
    for (float v : s.ranges) o << ' ' << v;
}
}  // namespace sensor_msgs
