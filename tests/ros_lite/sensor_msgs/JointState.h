#pragma once
#include <string>
#include <vector>

#include "geometry_msgs/Point.h"
namespace sensor_msgs {
struct JointState {
    std_msgs::Header header;
    std::vector<std::string> name;
    std::vector<double> position, velocity, effort;
};
inline void lite_dump(std::ostream& o, const JointState& j) {
    geometry_msgs::lite_prec(o);
    for (double v : j.position) o << v << ' ';
    for (double v : j.velocity) o << v << ' ';
}
}  // namespace sensor_msgs
