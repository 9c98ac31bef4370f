#pragma once
#include "geometry_msgs/Point.h"
namespace tf2_ros {
class TransformBroadcaster {
  public:
    void sendTransform(const geometry_msgs::TransformStamped& t) { ros::Publisher("/tf").publish(t); }
};
}  // namespace tf2_ros
